/* m3dreg.h — C ABI of the B200-native registration hot path (drop-in boundary).
 *
 * This is the seam between the m3d / gpu_6dslam host code (C++, unchanged) and the sm_100a
 * kernels of this repository.  Every entry point is `extern "C"`, takes plain pointers and
 * sizes, never throws, and returns an `int` status:
 *      0                success
 *      > 0              a cudaError_t value (same convention as the reference's L0 functions,
 *                       lesson_16.h:59-143, which all return cudaError_t)
 *      < 0              an M3DREG_E_* library code (below)
 *
 * "ref:" citations are relative to /root/reference/gpu_6dslam/gpu_6dslam/ and name the reference
 * interface each declaration replaces.  POD structs have EXACTLY the reference's memory layout
 * so device/host buffers can be handed over without conversion.
 *
 * Threading: one context per device; a context is not thread-safe.  All work of a context is
 * issued on one CUDA stream (its own non-blocking stream, or the one given to
 * m3dreg_set_stream).  Functions taking HOST result pointers synchronise that stream before
 * returning; functions whose results stay on the device do not.
 */
#ifndef M3DREG_H_
#define M3DREG_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3DREG_VERSION 100

/* ---- status codes (negative = library, positive = cudaError_t) ------------------------- */
#define M3DREG_OK                  0
#define M3DREG_E_INVALID_ARG     (-1)   /* null pointer, negative size, dof not in {4,6}, ...        */
#define M3DREG_E_TOO_MANY_BUCKETS (-2)  /* nbX*nbY*nbZ does not fit int32 (ref silently overflows,
                                           lesson_16.cu:80) or exceeds the planned capacity         */
#define M3DREG_E_NOT_SPD         (-3)   /* Cholesky pivot <= 0: pose left unchanged
                                           (ref prints and carries on, CCUDAAXBSolverWrapper.cpp:520) */
#define M3DREG_E_TOO_FEW_OBS     (-4)   /* n_obs <= obs_threshold: no solve (ref: gpu6DSLAM.cpp:402,572) */
#define M3DREG_E_BAD_SLOT        (-5)   /* scan slot not uploaded / out of range                     */
#define M3DREG_E_NO_DEVICE       (-6)   /* no CUDA device / wrong architecture (needs sm_100)        */
#define M3DREG_E_SIZE_MISMATCH   (-7)   /* ref: cudaWrapper.cpp:354 silently returns; we report it    */
#define M3DREG_E_NO_NCCL         (-8)   /* multi-GPU sweep without NCCL in the process / without a communicator */
#define M3DREG_E_NCCL            (-9)   /* an NCCL call returned an error                            */
#define M3DREG_E_IO              (-10)  /* a file could not be read / written / parsed (m3dreg_node.h) */

/* ---- labels (ref: include/lesson_16.h:9-12) --------------------------------------------- */
#define M3DREG_LABEL_PLANE   0
#define M3DREG_LABEL_EDGE    1
#define M3DREG_LABEL_CEILING 2
#define M3DREG_LABEL_GROUND  3

/* ---- POD types with the reference's exact layouts ---------------------------------------- */

/* ref: include/custom_point_types.h:8-20  lidar_pointcloud::PointXYZIRNLRGB — 40 bytes, align 4 */
typedef struct m3dreg_point {
	float    x, y, z;          /*  0  4  8 */
	float    intensity;        /* 12 */
	uint16_t ring;             /* 16 (+2 padding) */
	float    normal_x;         /* 20 */
	float    normal_y;         /* 24 */
	float    normal_z;         /* 28 */
	int32_t  label;            /* 32 */
	float    rgb;              /* 36 */
} m3dreg_point;

/* ref: include/lesson_16.h:15-18  hashElement — 8 bytes */
typedef struct m3dreg_hash_element {
	int32_t index_of_point;
	int32_t index_of_bucket;
} m3dreg_hash_element;

/* ref: include/lesson_16.h:20-24  bucket — 12 bytes */
typedef struct m3dreg_bucket {
	int32_t index_begin;
	int32_t index_end;
	int32_t number_of_points;
} m3dreg_bucket;

/* ref: include/lesson_16.h:26-40  gridParameters — 64 bytes (number_of_buckets at offset 40) */
typedef struct m3dreg_grid_params {
	float   bounding_box_min_X, bounding_box_min_Y, bounding_box_min_Z;
	float   bounding_box_max_X, bounding_box_max_Y, bounding_box_max_Z;
	int32_t number_of_buckets_X, number_of_buckets_Y, number_of_buckets_Z;
	int32_t _pad0;
	int64_t number_of_buckets;
	float   resolution_X, resolution_Y, resolution_Z;
	int32_t _pad1;
} m3dreg_grid_params;

/* ref: include/lesson_16.h:48-57  obs_nn_t — 28 bytes */
typedef struct m3dreg_obs_nn {
	float x_diff, y_diff, z_diff;   /* p1(global) - p2(global), computed in float (gpu6DSLAM.cpp:370-372) */
	float x0, y0, z0;               /* matched point of the first cloud in ITS LOCAL frame (gpu6DSLAM.cpp:367-369) */
	float P;                        /* weight_label / count_label (gpu6DSLAM.cpp:373-395) */
} m3dreg_obs_nn;

/* Registration knobs = the hot-path subset of the reference's ROS params
 * (ref: include/gpu6DSLAM.h:179-210, src/main.cpp:271-330). */
typedef struct m3dreg_reg_params {
	float   search_radius;        /* slam_search_radius_*                       */
	float   bucket_size;          /* slam_bucket_size_* (cubic buckets)         */
	float   bbox_extension;       /* slam_bounding_box_extension (1.0)          */
	int32_t max_inner;            /* slam_max_number_considered_in_INNER_bucket */
	int32_t max_outer;            /* slam_max_number_considered_in_OUTER_bucket */
	int32_t obs_threshold;        /* slam_number_of_observations_threshold (100): solve iff n_obs > threshold */
	float   weight[4];            /* slam_observation_weight_{plane,edge,ceiling,ground} indexed by label */
	int32_t dof;                  /* 6 = registerLS, 4 = registerLS_4DOF (x,y,z,yaw) */
	int32_t mode;                 /* M3DREG_MODE_ICP | M3DREG_MODE_NDT */
} m3dreg_reg_params;

#define M3DREG_MODE_ICP 0   /* semantic point-to-point (the reference's only mode)            */
#define M3DREG_MODE_NDT 1   /* point-to-distribution on per-bucket mean/covariance (new; no reference) */

/* Per-call statistics of the fused entry points (all optional). */
typedef struct m3dreg_icp_stats {
	int32_t iterations_run;
	int32_t last_status;          /* status of the last solve (0, M3DREG_E_NOT_SPD, M3DREG_E_TOO_FEW_OBS) */
	int64_t n_obs_last;           /* matched correspondences in the last iteration                       */
	int64_t n_buckets_last;       /* dense bucket count of the last grid                                 */
	double  x_last[6];            /* last solution vector (dof entries)                                  */
	float   device_ms;            /* CUDA-event time of the whole call on the context's stream           */
} m3dreg_icp_stats;

typedef struct m3dreg_ctx m3dreg_ctx;

/* ---- lifecycle ---------------------------------------------------------------------------- */

int         m3dreg_version(void);
const char *m3dreg_status_string(int status);

/* ref: CCudaWrapper::CCudaWrapper + warmUpGPU(int cudaDevice) (src/cudaWrapper.cpp:4-18,36-44):
 * selects the device, creates the stream and the persistent arena (grown on demand, never on the
 * iteration path).  Fails with M3DREG_E_NO_DEVICE when no sm_100 device is visible — there is no
 * CPU fallback. */
int  m3dreg_create(m3dreg_ctx **out, int cuda_device);
void m3dreg_destroy(m3dreg_ctx *ctx);
/* ref: cudaWarmUpGPU (include/lesson_16.h:59) */
int  m3dreg_warm_up(m3dreg_ctx *ctx);
/* Use an externally owned cudaStream_t (passed as void*) for all work; NULL = context's own. */
int  m3dreg_set_stream(m3dreg_ctx *ctx, void *cuda_stream);
void *m3dreg_get_stream(m3dreg_ctx *ctx);
int  m3dreg_synchronize(m3dreg_ctx *ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches claim). */
int64_t m3dreg_launch_count(const m3dreg_ctx *ctx);

/* Diagnostic: switch the NN search's exact bucket pruning off (every query then walks all <= 27 buckets like the
 * reference kernel does).  Results are identical either way (tests/test_gpu_stages.py::test_pruning_is_exact);
 * only m3dreg_nn_search / m3dreg_semantic_nn_host honour it.  Default: on. */
int m3dreg_set_pruning(m3dreg_ctx *ctx, int enabled);

/* Diagnostic counter: number of candidate distance evaluations the NN search made (summed over queries) since the
 * last reset.  Only counted while m3dreg_set_profiling(ctx, 1) is in effect (the counter costs one atomic per warp).
 * Reported by bench.py as evaluations per query. */
int m3dreg_get_nn_evaluations(m3dreg_ctx *ctx, uint64_t *count_out, int reset);

/* Diagnostic counter (also only while profiling is on): number of queries the warp-shared search handed to the
 * per-thread search because their warp's queries were scattered.  0 for spatially sorted queries (the scan store's
 * order); tests use it to prove which code path produced a result. */
int m3dreg_get_nn_fallbacks(m3dreg_ctx *ctx, uint64_t *count_out, int reset);

/* Diagnostic (only while profiling is on): %globaltimer stamps (ns) of block 0 at the phase boundaries of the last
 * k_grid_build launch — 0 entry, 1 box done, 2 barrier, 3 keys done, 4 barrier, 5 first sort pass done, 6 barrier,
 * 7 bucket table + second pass done, 8 barrier, 9 (third pass), 10 candidate sets done; 15-21 inside the first sort pass
 * (count scan done, entry, digit bases, ranks, warp scan, scatter, next histogram), 22-27 one bucket's candidate set
 * (start, record read, binning start, binned, table written, placed), 29 its candidate / point counts.
 * stamps_out: 32 entries. */
int m3dreg_get_grid_phase_ns(m3dreg_ctx *ctx, uint64_t *stamps_out);

/* ---- stage-level entry points on DEVICE pointers (parity surface = reference L0) ---------- */

/* ref: cudaCalculateGridParams (include/lesson_16.h:61-62, src/lesson_16.cu:23-106).
 * One fused min/max pass instead of 3 thrust::minmax_element + 6 D2H copies. out is HOST memory. */
int m3dreg_calculate_grid_params(m3dreg_ctx *ctx, const m3dreg_point *d_cloud, int n,
		float resolution_X, float resolution_Y, float resolution_Z,
		float bounding_box_extension, m3dreg_grid_params *out);

/* ref: cudaCalculateGrid (include/lesson_16.h:64-65, src/lesson_16.cu:200-243).
 * Fills d_table (n entries, sorted by bucket then original index) and the dense d_buckets
 * (params->number_of_buckets entries, empty = {-1,-1,0}) bit-identically to the reference,
 * including the first-element quirk of kernel_updateBuckets (src/lesson_16.cu:148-158). */
int m3dreg_calculate_grid(m3dreg_ctx *ctx, const m3dreg_point *d_cloud, int n,
		const m3dreg_grid_params *params, m3dreg_bucket *d_buckets, m3dreg_hash_element *d_table);

/* ref: cudaSemanticNearestNeighborSearch (include/lesson_16.h:76-88, src/lesson_16.cu:531-738).
 * d_nn[q] = index into d_first (original order) or -1. */
int m3dreg_nn_search(m3dreg_ctx *ctx,
		const m3dreg_point *d_first, int n_first,
		const m3dreg_point *d_second, int n_second,
		const m3dreg_hash_element *d_table, const m3dreg_bucket *d_buckets,
		const m3dreg_grid_params *params, float search_radius,
		int max_inner, int max_outer, int *d_nn);

/* ref: cudaTransformPointCloud, out-of-place overload (include/lesson_16.h:137-143,
 * src/lesson_16.cu:1341-1384).  m = row-major 3x4 [R|t] (host).  d_in may equal d_out. */
int m3dreg_transform(m3dreg_ctx *ctx, const m3dreg_point *d_in, m3dreg_point *d_out, int n, const float *m3x4);

/* ref: fill_A_l_cuda / fill_A_l_4DOFcuda + cudaCompute_AtP + 2x cublasDgemm
 * (include/lesson_16.h:67-74, src/CCUDAAXBSolverWrapper.cpp:407-428).  A, P, AtP are never
 * materialised: one fused fp64 reduction.  pose6 = {tx,ty,tz,om,fi,ka} (host).
 * AtPA_out: dof*dof doubles column-major, AtPl_out: dof doubles (both HOST). */
int m3dreg_normal_equations(m3dreg_ctx *ctx, const m3dreg_obs_nn *d_obs, int n_obs,
		const double *pose6, int dof, double *AtPA_out, double *AtPl_out);

/* ref: CCUDA_AX_B_SolverWrapper::linearSolverCHOL (src/CCUDAAXBSolverWrapper.cpp:484-539):
 * lower Cholesky + two triangular solves of a dof x dof system, executed on the device.
 * All pointers HOST.  Returns M3DREG_E_NOT_SPD if a pivot is not positive. */
int m3dreg_solve_chol(m3dreg_ctx *ctx, const double *AtPA, const double *AtPl, int dof, double *x_out);

/* ref: CCUDA_AX_B_SolverWrapper::Solve_ATPA_ATPl_x_data_on_GPU
 * (include/CCUDAAXBSolverWrapper.h:58-59): observations on the device in, x (host) out. */
int m3dreg_solve_observations(m3dreg_ctx *ctx, const m3dreg_obs_nn *d_obs, int n_obs,
		const double *pose6, int dof, double *x_out);

/* ---- CCudaWrapper-level entry points on HOST buffers (what gpu6DSLAM.cpp calls today) ----- */

/* ref: CCudaWrapper::semanticNearestNeighbourhoodSearch (include/cudaWrapper.h:57-65,
 * src/cudaWrapper.cpp:344-424).  nn_out has n_second entries. */
int m3dreg_semantic_nn_host(m3dreg_ctx *ctx,
		const m3dreg_point *first, int n_first, const m3dreg_point *second, int n_second,
		float search_radius, float bucket_size, float bounding_box_extension,
		int max_inner, int max_outer, int *nn_out);

/* ref: CCudaWrapper::registerLS (dof 6, src/cudaWrapper.cpp:516-581) and registerLS_4DOF
 * (dof 4, src/cudaWrapper.cpp:583-648).  pose6 = {tx,ty,tz,om,fi,ka} updated in place. */
int m3dreg_register_ls_host(m3dreg_ctx *ctx, const m3dreg_obs_nn *obs, int n_obs,
		double *pose6, int dof, double *x_out);

/* ref: CCudaWrapper::Matrix4ToEuler / EulerToMatrix (src/cudaWrapper.cpp:470-514).
 * m = row-major 4x4 float; omfika, xyz = 3 floats.  Pure host helpers. */
void m3dreg_matrix4_to_euler(const float *m4x4, float *omfika, float *xyz);
void m3dreg_euler_to_matrix(const float *omfika, const float *xyz, float *m4x4);

/* ---- device-resident scan store + fused loops (the B200 path) ----------------------------- */

/* Upload (or replace) scan `slot` (0-based).  src is HOST memory unless src_on_device != 0.
 * The context keeps the scan in HBM as two float4 streams (xyz+label, normal). */
int m3dreg_scan_upload(m3dreg_ctx *ctx, int slot, const m3dreg_point *src, int n, int src_on_device);
int m3dreg_scan_size(const m3dreg_ctx *ctx, int slot);
int m3dreg_scan_clear(m3dreg_ctx *ctx);

/* ref: the iteration body of gpu6DSLAM::registerLastArrivedScan (src/gpu6DSLAM.cpp:264-422)
 * run `iterations` times without leaving the device: Euler round-trip, transform of the first
 * scan by its current pose, grid, semantic NN with the second scan (transformed once by
 * pose_second) as queries, per-label weights, normal equations, Cholesky, pose update.
 * pose_first (row-major 4x4 float, in/out) and pose_second (row-major 4x4 float) are HOST. */
int m3dreg_icp_pair(m3dreg_ctx *ctx, int first_slot, int second_slot,
		float *pose_first, const float *pose_second,
		const m3dreg_reg_params *params, int iterations, m3dreg_icp_stats *stats);

/* The same loop in three steps, for callers that want to enqueue iterations without any host round trip in
 * between (benchmarks, CUDA-graph style replays, multi-GPU drivers):
 *   icp_begin   uploads the poses, transforms the queries once and sizes the bucket table (one stream sync);
 *   icp_step    ENQUEUES `iterations` device-resident iterations on the context's stream and returns at once;
 *   icp_end     reads pose + statistics back (synchronises).
 * icp_copy_neq copies the last iteration's 28-double normal-equation block (21 upper-triangular AtPA, 6 AtPl,
 * count) device-to-device into d_dst, asynchronously on the context's stream (for NCCL all-reduce). */
int m3dreg_icp_begin(m3dreg_ctx *ctx, int first_slot, int second_slot, const float *pose_first,
		const float *pose_second, const m3dreg_reg_params *params);
int m3dreg_icp_step(m3dreg_ctx *ctx, int iterations);
int m3dreg_icp_end(m3dreg_ctx *ctx, float *pose_first_out, m3dreg_icp_stats *stats);
int m3dreg_icp_copy_neq(m3dreg_ctx *ctx, double *d_dst);
/* Alternative without the copy: every following iteration of the fused loop also WRITES its 28-double block to d_dst
 * (device memory, caller-owned; 0 switches it off) from the kernel that forms it — the buffer an NCCL all-reduce on a
 * second stream picks up.  With several iterations per icp_step call the last one wins. */
int m3dreg_icp_set_neq_out(m3dreg_ctx *ctx, double *d_dst);

/* Per-stage CUDA-event timing of the fused iteration (off by default).  Stages: 0 transform+bounds,
 * 1 grid (params, keys, radix sort, bucket table, gather), 2 semantic NN, 3 normal equations + solve.
 * get_stage_ms returns the accumulated milliseconds and launch counts since the last reset and resets them. */
#define M3DREG_STAGE_COUNT 4
int m3dreg_set_profiling(m3dreg_ctx *ctx, int enabled);
int m3dreg_get_stage_ms(m3dreg_ctx *ctx, float *ms_out, int *iterations_out);

/* Same iteration, but through HOST clouds every call the way the reference crosses the
 * boundary (both 40-B clouds H2D, nn + pose D2H): first is in its LOCAL frame, second already
 * in the global frame.  nn_out may be NULL.  The uploads run on an internal copy stream and the
 * first cloud's grid is built under the second upload; the call synchronises before returning
 * (the caller's buffers are free again, pose_first / stats / nn_out are final). */
int m3dreg_icp_iteration_host(m3dreg_ctx *ctx,
		const m3dreg_point *first_local, int n_first, const m3dreg_point *second_global, int n_second,
		float *pose_first, const m3dreg_reg_params *params, int *nn_out, m3dreg_icp_stats *stats);

/* Export hooks for parity checks: materialise what the last fused iteration built, in the
 * reference's layouts (HOST outputs; any may be NULL). buckets_cap in entries. */
int m3dreg_export_last_grid(m3dreg_ctx *ctx, m3dreg_grid_params *params_out,
		m3dreg_hash_element *table_out, int table_cap, m3dreg_bucket *buckets_out, int64_t buckets_cap);
int m3dreg_export_last_nn(m3dreg_ctx *ctx, int *nn_out, int nn_cap);

/* ref: gpu6DSLAM::registerAll (src/gpu6DSLAM.cpp:424-597), split so pairs can be sharded over
 * GPUs.  Normal equations per scan are 28 doubles: 21 upper-triangular AtPA entries (row-major
 * order of the 6x6 upper triangle), 6 AtPl entries, 1 observation count.
 *   sweep_accumulate: for the given (i,j) pairs (i = gridded/optimised scan, j = query scan)
 *       ADD each pair's weighted contribution into d_neq[i*28 ..].  poses = n_scans row-major
 *       4x4 float (HOST), the OLD poses of the Jacobi sweep.  d_neq is DEVICE memory
 *       (n_scans*28 doubles) owned by the caller so it can be all-reduced (NCCL) in place.
 *   sweep_solve: gate on obs_threshold, Cholesky, pose update, Euler round-trip for every
 *       scan in [scan_begin, scan_end); poses updated in place (HOST). status_out (n_scans
 *       ints, HOST, may be NULL) receives the per-scan solve status. */
int m3dreg_sweep_zero(m3dreg_ctx *ctx, double *d_neq, int n_scans);
int m3dreg_sweep_accumulate(m3dreg_ctx *ctx, int n_pairs, const int *pair_i, const int *pair_j,
		const float *poses, int n_scans, const m3dreg_reg_params *params, double *d_neq);
int m3dreg_sweep_solve(m3dreg_ctx *ctx, const double *d_neq, int n_scans, int scan_begin, int scan_end,
		float *poses, const m3dreg_reg_params *params, int *status_out);

/* ---- the whole sweep as ONE call, sharded over the GPUs of a box ---------------------------------------------
 * ref: gpu6DSLAM::registerAll(cudaWrapper, radius, bucket, number_of_last_EOZ) (src/gpu6DSLAM.cpp:424-597) and the
 * service loop that calls it (src/main.cpp:36-60).  One context per GPU / per rank, every scan uploaded to every
 * context (m3dreg_scan_upload, slots 0..n_scans-1).  Every rank calls m3dreg_slam_sweep with the SAME poses:
 *   1. pair gate: (i, j), i in [first_optimised, n_scans), j != i, |t_i - t_j| < distance_threshold
 *      (slam_registerAll_distance_threshold, include/gpu6DSLAM.h:180; src/gpu6DSLAM.cpp:464-469);
 *   2. deterministic partition of the pairs over the ranks (whole groups of equal i, heavy groups split), balanced by point
 *      counts in the first sweep and by the MEASURED device time of every group from then on (a second all-reduce of
 *      n_scans doubles per sweep carries the measurements);
 *   3. m3dreg_sweep_accumulate of the rank's pairs into n_scans x 28 doubles on the device;
 *   4. ONE ncclAllReduce (sum, double) of that block over NVLink / NVSwitch, on the context's stream;
 *   5. gate on obs_threshold, Cholesky, pose update, Euler round trip of scans [first_optimised, n_scans) on every rank
 *      (src/gpu6DSLAM.cpp:572-593); poses (n_scans row-major 4x4 float, HOST) are updated in place, identically on
 *      every rank; earlier scans keep their pose untouched (src/gpu6DSLAM.cpp:430).
 * NCCL is looked up at run time (the process' own copy first, then libnccl.so.2): no link-time dependency. */
typedef struct m3dreg_slam_params {
	m3dreg_reg_params reg;          /* radius / bucket of this sweep, caps, weights, dof, mode (ICP or NDT)              */
	float   distance_threshold;     /* slam_registerAll_distance_threshold (10 m)                                        */
	int32_t first_optimised;        /* n_scans - number_of_last_EOZ: scans before it are only neighbours (0 = all)       */
} m3dreg_slam_params;

typedef struct m3dreg_sweep_stats {
	int64_t n_pairs, n_pairs_mine;  /* gated pairs in the sweep / handled by this rank                                    */
	int64_t points_all, points_mine;/* sum over those pairs of (points of i + points of j)                                */
	float   accumulate_ms;          /* CUDA-event time of this rank's accumulation                                       */
	float   allreduce_ms;           /* CUDA-event time from the end of the accumulation to the end of the all-reduce
	                                   (includes waiting for the slowest rank)                                           */
	int32_t rank, world;
} m3dreg_sweep_stats;

/* Communicator of this context's rank: either created here from a 128-byte ncclUniqueId that rank 0 obtained with
 * m3dreg_nccl_get_unique_id and handed to the other ranks (any transport: MPI, a file, torch.distributed, ...), or an
 * existing ncclComm_t of the host application (attach: not destroyed with the context).  world == 1 detaches. */
int m3dreg_nccl_get_unique_id(void *id128);
int m3dreg_nccl_init(m3dreg_ctx *ctx, const void *id128, int rank, int world);
int m3dreg_nccl_attach(m3dreg_ctx *ctx, void *nccl_comm, int rank, int world);

int m3dreg_slam_sweep(m3dreg_ctx *ctx, int n_scans, float *poses, const m3dreg_slam_params *params,
		int *status_out /* n_scans, HOST, may be NULL */, m3dreg_sweep_stats *stats /* may be NULL */);
/* The (all-reduced) normal-equation blocks of the last m3dreg_slam_sweep, n_scans x 28 doubles to HOST memory (checks). */
int m3dreg_slam_copy_neq(m3dreg_ctx *ctx, double *neq_out, int n_scans);
/* The plan alone, without a device (pure host function; the library loads without a GPU): pairs in sweep order and the
 * rank that owns each.  Returns the number of pairs (or < 0); arrays may be NULL to only count.  sizes: points per scan. */
int m3dreg_slam_plan(const float *poses, int n_scans, const int *sizes, float distance_threshold, int first_optimised,
		int world, int *pair_i, int *pair_j, int *owner, int cap);
/* The plan m3dreg_slam_sweep uses from its SECOND sweep on (multi-GPU only): instead of point counts the groups are
 * balanced by the device time per pair the previous sweep measured for every scan's group (cost_per_pair, n_scans doubles,
 * ms; <= 0 = not measured: the mean of the measured ones).  The measurement is all-reduced, so every rank plans alike. */
int m3dreg_slam_plan_measured(const float *poses, int n_scans, const int *sizes, float distance_threshold, int first_optimised,
		int world, const double *cost_per_pair, int *pair_i, int *pair_j, int *owner, int cap);

/* ---- pre-registration steps on the registration path's own grid ---------------------
 * Every incoming scan passes through these before it is registered (src/gpu6DSLAM.cpp:63-85); they build the same
 * regular grid (bounds -> keys -> stable sort -> dense table) and produce the normals / labels the semantic search keys
 * on.  HOST clouds in, HOST clouds out, synchronous — as the reference's wrapper methods. */

/* ref: CCudaWrapper::removeNoiseNaive(cloud, resolution, bounding_box_extension, number_of_points_in_bucket_threshold)
 * (src/cudaWrapper.cpp:118-179) = cudaCalculateGridParams + cudaCalculateGrid + cudaRemoveNoiseNaive
 * (src/lesson_16.cu:740-787): a point stays iff its bucket holds MORE than `threshold` points.  The survivors are
 * written to `out` (capacity n) in their original order, their number to *n_out; markers_out (n bytes, 0/1, may be
 * NULL) receives the reference's d_markers.  `out` may alias `cloud`. */
int m3dreg_remove_noise_host(m3dreg_ctx *ctx, const m3dreg_point *cloud, int n, float resolution, float bounding_box_extension,
		int number_of_points_in_bucket_threshold, m3dreg_point *out, int *n_out, unsigned char *markers_out);

/* ref: CCudaWrapper::downsampling(cloud, resolution, bounding_box_extension) (src/cudaWrapper.cpp:181-262) =
 * grid + cudaDownSample (src/lesson_16.cu:789-815): the first point (smallest original index) of every bucket that
 * has an index_begin stays. */
int m3dreg_downsample_host(m3dreg_ctx *ctx, const m3dreg_point *cloud, int n, float resolution, float bounding_box_extension,
		m3dreg_point *out, int *n_out, unsigned char *markers_out);

/* ref: CCudaWrapper::classify(cloud, normal_vectors_search_radius, curvature_threshold, ground_Z_coordinate_threshold,
 * number_of_points_needed_for_plane_threshold, bounding_box_extension, max_INNER, max_OUTER, viewpoint)
 * (src/cudaWrapper.cpp:264-342) = grid with cubic buckets of the search radius + cudaSemanticLabelingPlaneEdges +
 * cudaSemanticLabelingFloorCeiling (src/lesson_16.cu:817-1239): normal_x/y/z and label of every point are rewritten
 * in place.  mean_out (3 floats per SORTED position, may be NULL) receives the reference's d_mean for parity checks,
 * table_out (n records, may be NULL) the sorted table the positions refer to. */
int m3dreg_classify_host(m3dreg_ctx *ctx, m3dreg_point *cloud, int n, float normal_vectors_search_radius, float curvature_threshold,
		float ground_Z_coordinate_threshold, int number_of_points_needed_for_plane_threshold, float bounding_box_extension,
		int max_number_considered_in_INNER_bucket, int max_number_considered_in_OUTER_bucket,
		float viewpointX, float viewpointY, float viewpointZ, float *mean_out, m3dreg_hash_element *table_out);

/* ref: CCudaWrapper::findBestYaw(first, m_first, second, m_second, bucket_size, bounding_box_extension, search_radius,
 * max_INNER, max_OUTER, angle_start, angle_finish, angle_step, myaw_out) (src/cudaWrapper.cpp:662-836;
 * src/lesson_16.cu:1241-1384): the second cloud is brought into the first one's frame (second_transform3x4, then
 * first_transform_inverse3x4: row-major 3x4, may be NULL = identity), and for every angle a = start, start + step, ...
 * <= finish (degrees, float accumulation as upstream) rotated about Z, searched against the grid of the first cloud
 * and the matched queries counted; the angle with the most matches wins (strictly more: the first maximum).
 * The grid and the candidate sets of the first cloud are built ONCE (upstream rebuilds nothing either, but copies the
 * second cloud and launches per angle with a device-wide sync each).  counts_out (may be NULL): matches per angle. */
int m3dreg_find_best_yaw_host(m3dreg_ctx *ctx, const m3dreg_point *first, int n1, const m3dreg_point *second, int n2,
		const float *second_transform3x4, const float *first_transform_inverse3x4,
		float bucket_size, float bounding_box_extension, float search_radius, int max_inner, int max_outer,
		float angle_start, float angle_finish, float angle_step, float *best_angle_out, int *best_count_out, int *counts_out, int counts_cap);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* M3DREG_H_ */
