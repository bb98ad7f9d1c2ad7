/* m3dreg_node.h — the callers and data formats either side of the registration path, as a C ABI
 * of libm3dreg.so: a ROS-free replay of the reference's per-scan driver `class gpu6DSLAM`
 *   ref: include/gpu6DSLAM.h:28-262, src/gpu6DSLAM.cpp:4-222 (registerSingleScan), :203-262 (getMetascan),
 *        :597-631 (registerAll service), :665-718 (loadmapfromfile), :720-750 (callbackInitialPose), :752-758 (downsample)
 * with its persistence formats:
 *   - binary PCD files of PointXYZIRNLRGB as pcl::io::savePCDFileBinary writes them for the point type registered in
 *     include/custom_point_types.h:22-32 (src/gpu6DSLAM.cpp:41, :88; read back by pcl::io::loadPCDFile, :693),
 *   - the XML pose model of `class data_model` (include/data_model.hpp, src/data_model.cpp:6-154: a boost::property_tree
 *     written with tab indentation; Model.Algorithms.name, Model.DatasetPath, Model.Transformations.<id>.Affine.{Type,Data}
 *     with the 4x4 matrix in COLUMN-major order, Model.Transformations.<id>.cloudname).
 * No ROS, PCL, Eigen or Boost: the formats are restated from their definitions; files written here load in the reference
 * and vice versa (tests/test_formats.py pins the byte layout against fixtures written in PCL's / property_tree's format).
 *
 * The node owns the scan store: a host copy of every processed scan (what getMetascan concatenates) and its resident copy
 * on the device (m3dreg_scan_upload slots), so that the 30+30+30 pair iterations and the 3 x 10 sweeps of the schedule
 * (src/gpu6DSLAM.cpp:159-187) run device-resident through m3dreg_icp_pair / m3dreg_slam_sweep.
 */
#ifndef M3DREG_NODE_H_
#define M3DREG_NODE_H_

#include "m3dreg.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- persistence formats (pure host functions: usable without a GPU) ---------------------------------------------- */

/* ref: pcl::io::savePCDFileBinary(path, cloud) as called at src/gpu6DSLAM.cpp:41,88 = PCL's typed writer
 * (pcl/io/impl/pcd_io.hpp: PCDWriter::generateHeader<PointT> + writeBinary<PointT>; PCL is not part of /root/reference,
 * version unpinned — 1.7.2 on the Ubuntu 16.04 / ROS kinetic the tree targets).  Header for this point type:
 *   FIELDS x y z intensity ring normal_x normal_y normal_z label rgb / SIZE 4 4 4 4 2 4 4 4 4 4 / TYPE F F F F U F F F I F /
 *   COUNT 1 x 10 / WIDTH n / HEIGHT 1 / VIEWPOINT 0 0 0 1 0 0 0 / POINTS n / DATA binary,
 * then the REGISTERED fields of every point packed back to back: 38 bytes per point (the struct's 2 padding bytes after
 * `ring` are not written).  0 or M3DREG_E_IO. */
int m3dreg_pcd_write_binary(const char *path, const m3dreg_point *cloud, int n);
/* ref: pcl::io::loadPCDFile (src/gpu6DSLAM.cpp:693).  Reads `DATA binary` and `DATA ascii` files whose fields include
 * x y z (the others are filled when present, by NAME, so files written by other PCL point types load too).
 * out == NULL: only the point count is returned in *n_out.  M3DREG_E_IO / M3DREG_E_SIZE_MISMATCH (cap too small). */
int m3dreg_pcd_read(const char *path, m3dreg_point *out, int cap, int *n_out);

typedef struct m3dreg_model m3dreg_model;       /* ref: class data_model (include/data_model.hpp) */
m3dreg_model *m3dreg_model_create(void);
void m3dreg_model_destroy(m3dreg_model *m);
int  m3dreg_model_load(m3dreg_model *m, const char *xml_path);                           /* loadFile, data_model.cpp:6-27 */
int  m3dreg_model_save(const m3dreg_model *m, const char *xml_path);                     /* saveFile, :29-48 */
void m3dreg_model_set_algorithm_name(m3dreg_model *m, const char *name);                 /* :198-201 */
void m3dreg_model_set_dataset_path(m3dreg_model *m, const char *path);                   /* :227-230 */
int  m3dreg_model_get_dataset_path(const m3dreg_model *m, char *out, int cap);           /* :231-241 */
void m3dreg_model_set_affine(m3dreg_model *m, const char *scan_id, const float *m4x4_rowmajor);      /* setAffine(Matrix4f), :137-152 */
int  m3dreg_model_get_affine(const m3dreg_model *m, const char *scan_id, float *m4x4_rowmajor);      /* getAffine, :50-91 (both Types) */
void m3dreg_model_set_cloud_name(m3dreg_model *m, const char *scan_id, const char *file_name);       /* :184-187 */
int  m3dreg_model_get_cloud_name(const m3dreg_model *m, const char *scan_id, char *out, int cap);    /* :93-98 */
int  m3dreg_model_scan_count(const m3dreg_model *m);                                                  /* getAllScansId, :189-196 */
int  m3dreg_model_scan_id(const m3dreg_model *m, int index, char *out, int cap);
/* <directory of the xml>/<DatasetPath>/<cloudname> (getFullPathOfPointcloud, :242-253) */
int  m3dreg_model_full_cloud_path(const m3dreg_model *m, const char *scan_id, char *out, int cap);

/* ---- the per-scan driver -------------------------------------------------------------------------------------------- */

/* ref: the public parameter members of class gpu6DSLAM with their defaults (include/gpu6DSLAM.h:41-106, :163-223) */
typedef struct m3dreg_node_params {
	float   noise_removal_resolution;                    /* 0.5 */
	int32_t noise_removal_number_of_points_in_bucket_threshold;   /* 3 */
	float   noise_removal_bounding_box_extension;        /* 1.0 */
	float   downsampling_resolution;                     /* 0.3 (also passed as the box extension, src/gpu6DSLAM.cpp:71) */
	float   semantic_classification_normal_vectors_search_radius;           /* 1.0 */
	float   semantic_classification_curvature_threshold;                    /* 10.0 */
	float   semantic_classification_ground_Z_coordinate_threshold;          /* 1.0 */
	int32_t semantic_classification_number_of_points_needed_for_plane_threshold;   /* 15 */
	int32_t semantic_classification_max_number_considered_in_INNER_bucket;  /* 100 */
	int32_t semantic_classification_max_number_considered_in_OUTER_bucket;  /* 100 */
	float   semantic_classification_bounding_box_extension;                 /* 1.0 */
	float   slam_registerLastArrivedScan_distance_threshold;                /* 100.0 */
	float   slam_registerAll_distance_threshold;                            /* 10.0 */
	int32_t slam_number_of_observations_threshold;                          /* 100 */
	float   slam_search_radius_step[3];                  /* 2.5, 2.0, 1.0 */
	float   slam_bucket_size_step[3];                    /* 2.5, 2.0, 1.0 */
	int32_t slam_registerLastArrivedScan_number_of_iterations_step[3];      /* 30, 30, 30 */
	int32_t slam_registerAll_number_of_iterations_step[3];                  /* 10, 10, 10 */
	float   slam_search_radius_register_all;             /* 0.5 */
	float   slam_bucket_size_step_register_all;          /* 0.5 */
	float   slam_bounding_box_extension;                 /* 1.0 */
	int32_t slam_max_number_considered_in_INNER_bucket;  /* 100 */
	int32_t slam_max_number_considered_in_OUTER_bucket;  /* 100 */
	float   slam_observation_weight[4];                  /* plane 10, edge 1, ceiling 10, ground 10 (indexed by label) */
	float   findBestYaw_start_angle, findBestYaw_finish_angle, findBestYaw_step_angle;   /* -30, 30, 0.5 */
	float   findBestYaw_bucket_size, findBestYaw_bounding_box_extension, findBestYaw_search_radius;   /* 1.0, 1.0, 0.3 */
	int32_t findBestYaw_max_number_considered_in_INNER_bucket, findBestYaw_max_number_considered_in_OUTER_bucket;   /* 50, 50 */
	float   viewpoint[3];                                /* 0, 0, 2 */
	/* what upstream hard-codes */
	float   cutoff_z_min, cutoff_z_max, cutoff_xy2_min;  /* -1, 15, 1.5: keep z in (min, max) and x^2 + y^2 > xy2_min (src/gpu6DSLAM.cpp:50-57) */
	int32_t number_of_last_scans_in_sweeps;              /* 3 (src/gpu6DSLAM.cpp:176,181,186) */
	int32_t dof;                                         /* 4: registerLS_4DOF is the live call (src/gpu6DSLAM.cpp:405-406, 574-575) */
	int32_t use_find_best_yaw;                           /* 0: the call is commented out upstream (src/gpu6DSLAM.cpp:135-155) */
	int32_t write_files;                                 /* 1: raw + processed PCD and the three XML models per scan */
} m3dreg_node_params;

void m3dreg_node_default_params(m3dreg_node_params *p);

typedef struct m3dreg_node m3dreg_node;

/* ref: gpu6DSLAM::gpu6DSLAM(root_folder_name) (include/gpu6DSLAM.h:109-161): creates <root>, <root>/rawData,
 * <root>/processedData (root_folder may be NULL or params->write_files 0: nothing is written) and names the three models.
 * The node uses `ctx` (not owned) for every device operation. */
int  m3dreg_node_create(m3dreg_node **out, m3dreg_ctx *ctx, const m3dreg_node_params *params, const char *root_folder);
void m3dreg_node_destroy(m3dreg_node *node);

typedef struct m3dreg_node_scan_stats {
	int32_t n_raw, n_after_cutoff, n_after_noise_removal, n_after_downsampling;
	int32_t pair_iterations, pair_last_status;   /* registerLastArrivedScan iterations run / status of the last solve */
	int32_t sweeps, sweep_solved_last;           /* registerAll sweeps run / scans solved in the last one */
	float   yaw_deg;                             /* findBestYaw result (0 unless enabled) */
	float   preprocess_ms, register_ms;          /* host wall clock */
} m3dreg_node_scan_stats;

/* ref: gpu6DSLAM::registerSingleScan(pc, mtf, iso_time_str) (src/gpu6DSLAM.cpp:4-222): raw PCD, cut-off, noise filter,
 * downsampling, classification, processed PCD, odometry-increment chaining of vmtf / vmregistered and the re-anchoring of
 * every registered pose on the new scan's tf pose (:100-119), then the schedule — 3 steps of registerLastArrivedScan
 * (the last scan against its predecessor) and 3 steps of registerAll over the last 3 scans (:159-187) — and the three
 * XML models.  mtf: row-major 4x4. */
int m3dreg_node_register_single_scan(m3dreg_node *node, const m3dreg_point *cloud, int n, const float *mtf, const char *iso_time_str,
		m3dreg_node_scan_stats *stats /* may be NULL */);

int m3dreg_node_scan_count(const m3dreg_node *node);
/* pose of scan i: registered (vmregistered) and odometry (vmtf), row-major 4x4, either may be NULL */
int m3dreg_node_get_pose(const m3dreg_node *node, int i, float *registered, float *tf);
int m3dreg_node_scan_size(const m3dreg_node *node, int i);
int m3dreg_node_get_scan(const m3dreg_node *node, int i, m3dreg_point *out, int cap);        /* the processed scan, local frame */
int m3dreg_node_scan_id(const m3dreg_node *node, int i, char *out, int cap);
/* ref: gpu6DSLAM::getMetascan() (src/gpu6DSLAM.cpp:224-262): every scan transformed by its registered pose, concatenated.
 * out == NULL: only the size. */
int m3dreg_node_metascan(m3dreg_node *node, m3dreg_point *out, int cap, int *n_out);
/* ref: gpu6DSLAM::registerAll() (src/gpu6DSLAM.cpp:597-631): one sweep over ALL scans at slam_search_radius_register_all */
int m3dreg_node_register_all(m3dreg_node *node, int *solved_out);
/* ref: gpu6DSLAM::loadmapfromfile (src/gpu6DSLAM.cpp:665-718): replaces the scan store by the model's scans and poses */
int m3dreg_node_load_map(m3dreg_node *node, const char *xml_path);
/* ref: gpu6DSLAM::callbackInitialPose (src/gpu6DSLAM.cpp:720-750): re-anchors the map on the registered pose closest to
 * `initial_pose` (every pose becomes pose * closest^-1 * initial_pose) */
int m3dreg_node_set_initial_pose(m3dreg_node *node, const float *initial_pose);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* M3DREG_NODE_H_ */
