/* oracle/m3d_oracle.h — CPU ORACLE: TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the algorithm of gpu_6dslam's registration hot path, written from the
 * mathematical specification with each function citing the reference
 * file:line it follows (paths relative to /root/reference/gpu_6dslam/gpu_6dslam/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (mandala-mapping_b200/) never does and has no CPU fallback.
 *
 * PARITY PINNING.  The reference ships no tests, golden vectors or CPU implementation of this path
 * (its kernels are CUDA-only), so this oracle is pinned by EXECUTING the reference's own kernels
 * (oracle/_ref/libm3dref.so, built from the sources under /root/reference) on a B200:
 * tests/test_gpu_oracle_vs_ref.py checks grid parameters, bucket keys, sorted table, bucket table
 * and NN indices bit-for-bit, and normal equations / solutions to fp64 round-off, on seeded
 * synthetic scans; small golden vectors produced by that run are committed under tests/golden/.
 * Stages with no pinnable reference arithmetic (Eigen float pose composition, the host-side Eigen
 * cloud transform, cuBLAS/cuSOLVER summation order, and NDT which the reference lacks) are
 * "parity unpinned" and compared by tolerance only; see DESIGN.md.
 */
#ifndef M3D_ORACLE_H_
#define M3D_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same layouts as include/m3dreg.h (= the reference's, include/lesson_16.h:15-57) */
typedef struct orc_point {
	float x, y, z, intensity;
	uint16_t ring;
	float normal_x, normal_y, normal_z;
	int32_t label;
	float rgb;
} orc_point;
typedef struct orc_hash_element { int32_t index_of_point, index_of_bucket; } orc_hash_element;
typedef struct orc_bucket { int32_t index_begin, index_end, number_of_points; } orc_bucket;
typedef struct orc_grid_params {
	float bounding_box_min_X, bounding_box_min_Y, bounding_box_min_Z;
	float bounding_box_max_X, bounding_box_max_Y, bounding_box_max_Z;
	int32_t number_of_buckets_X, number_of_buckets_Y, number_of_buckets_Z;
	int32_t _pad0;
	int64_t number_of_buckets;
	float resolution_X, resolution_Y, resolution_Z;
	int32_t _pad1;
} orc_grid_params;
typedef struct orc_obs_nn { float x_diff, y_diff, z_diff, x0, y0, z0, P; } orc_obs_nn;

typedef struct orc_reg_params {
	float search_radius, bucket_size, bbox_extension;
	int32_t max_inner, max_outer, obs_threshold;
	float weight[4];
	int32_t dof;
	int32_t mode;
} orc_reg_params;

int  orc_num_threads(void);
void orc_set_num_threads(int n);

/* grid ---------------------------------------------------------------------------------------- */
void orc_grid_params_compute(const orc_point *cloud, int n, float rx, float ry, float rz, float ext, orc_grid_params *out);
void orc_bucket_keys(const orc_point *cloud, int n, const orc_grid_params *p, int32_t *keys_out);
void orc_build_grid(const orc_point *cloud, int n, const orc_grid_params *p, orc_bucket *buckets, orc_hash_element *table);

/* semantic NN ----------------------------------------------------------------------------------- */
int  orc_angle_gate(float dot);
void orc_nn_search(const orc_point *first, int n_first, const orc_point *second, int n_second,
		const orc_hash_element *table, const orc_bucket *buckets, const orc_grid_params *p,
		float search_radius, int max_inner, int max_outer, int32_t *nn_out);
/* candidate evaluations the reference kernel would perform (sum over queries), for reporting E/query */
int64_t orc_nn_count_evaluations(const orc_point *second, int n_second, const orc_bucket *buckets,
		const orc_grid_params *p, int max_inner, int max_outer);

/* observations, normal equations, solve ----------------------------------------------------------- */
int  orc_build_observations(const orc_point *first_global, const orc_point *first_local,
		const orc_point *second_global, int n_second, const int32_t *nn, const float *weight4,
		orc_obs_nn *obs_out);
void orc_normal_equations(const orc_obs_nn *obs, int n_obs, const double *pose6, int dof,
		double *AtPA_out, double *AtPl_out);
int  orc_chol_solve(const double *A_colmajor, const double *b, int n, double *x_out);
int  orc_register_ls(const orc_obs_nn *obs, int n_obs, double *pose6, int dof, double *x_out);

/* pose helpers + transform ------------------------------------------------------------------------ */
void orc_matrix4_to_euler(const float *m4x4, float *omfika, float *xyz);
void orc_euler_to_matrix(const float *omfika, const float *xyz, float *m4x4);
void orc_transform_cloud(const orc_point *in, orc_point *out, int n, const float *m4x4);

/* NDT (point-to-distribution) — NOT IN THE REFERENCE: definition of record for this repository,
 * "parity unpinned".  neq28 = 21 upper-triangular AtPA + 6 AtPl + observation count.  Returns n_obs. */
int64_t orc_ndt_normal_equations(const orc_point *first_global, const orc_point *first_local, int n_first,
		const orc_point *second_global, int n_second, const orc_hash_element *table, const orc_bucket *buckets,
		const orc_grid_params *p, const double *pose6, double *neq28);
int  orc_solve_packed(const double *neq28, int dof, double *x_out);

/* loops ------------------------------------------------------------------------------------------- */
/* one registerLastArrivedScan iteration body on an (i=first, j=second) pair; second_global is
 * already transformed.  scratch_first (n_first points) receives the transformed first cloud.
 * nn_out (n_second) may be NULL.  Returns the status: 0 solved, -3 not SPD, -4 too few obs. */
int  orc_icp_iteration(const orc_point *first_local, int n_first, const orc_point *second_global, int n_second,
		float *pose_first4x4, const orc_reg_params *prm, orc_point *scratch_first, int32_t *nn_out,
		int64_t *n_obs_out, double *x_out);
/* one Jacobi registerAll sweep over n_scans scans stored back to back (offsets[n_scans+1]) */
int  orc_register_all_sweep(const orc_point *scans_local, const int64_t *offsets, int n_scans,
		float *poses4x4, const orc_reg_params *prm, float pair_distance_threshold,
		double *neq_out /* n_scans*28 or NULL */, int32_t *status_out /* n_scans or NULL */);
int  orc_register_all_sweep_last(const orc_point *scans_local, const int64_t *offsets, int n_scans,
		float *poses4x4, const orc_reg_params *prm, float pair_distance_threshold, int first_optimised,
		double *neq_out, int32_t *status_out);

/* pre-registration steps --------------------------------------------------------- */
void orc_remove_noise_markers(const orc_point *cloud, int n, float res, float ext, int threshold, uint8_t *markers);
void orc_downsample_markers(const orc_point *cloud, int n, float res, float ext, uint8_t *markers);
void orc_classify(orc_point *cloud, int n, float radius, float curvature_threshold, float ground_z, int plane_points, float ext,
		int max_in, int max_out, float vx, float vy, float vz, float *mean_out, orc_hash_element *table_out);
int  orc_find_best_yaw(const orc_point *first, int n1, const orc_point *second, int n2, const float *second_m, const float *first_inv_m,
		float bucket, float ext, float radius, int max_in, int max_out, const float *yaw_mats, int n_angles, int32_t *counts_out);

#ifdef __cplusplus
}
#endif
#endif
