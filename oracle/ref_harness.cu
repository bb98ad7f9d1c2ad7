/* oracle/ref_harness.cu — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A thin extern "C" driver around the UNMODIFIED reference sources
 *   /root/reference/gpu_6dslam/gpu_6dslam/src/lesson_16.cu
 *   /root/reference/gpu_6dslam/gpu_6dslam/src/CCUDAAXBSolverWrapper.cpp
 * which are compiled where they lie (see oracle/Makefile) and linked with this file into
 * oracle/_ref/libm3dref.so.  Nothing from the reference is copied into this repository:
 * this file only *calls* the reference's own free functions / solver class the way
 * CCudaWrapper does (cudaWrapper.cpp:344-424 for the NN search, cudaWrapper.cpp:516-648
 * for registerLS / registerLS_4DOF), because cudaWrapper.cpp itself needs PCL + Eigen,
 * which are not installed.
 *
 * Used by tests/ (to pin oracle/m3d_oracle.c and the CUDA product against the reference's
 * own kernels on a B200) and by `bench.py --impl reference`.  The product never links it.
 */
#include "lesson_16.h"
#include "CCUDAAXBSolverWrapper.h"

#include <cublas_v2.h>
#include <cstdio>
#include <cstring>
#include <chrono>

typedef lidar_pointcloud::PointXYZIRNLRGB ref_point_t;

/* Per-stage wall-clock of the reference's own call sequence (every reference function ends in cudaDeviceSynchronize or a
 * blocking copy, so host time around a call IS its device time plus its malloc / launch overhead — what a caller pays):
 *   0 malloc + H2D of both clouds        1 cudaCalculateGridParams      2 mallocs + cudaCalculateGrid
 *   3 cudaSemanticNearestNeighborSearch  4 D2H of nn + frees            5 malloc + H2D of the observations
 *   6 fill_A_l_cuda[_4DOF]               7 Solve_ATPA_ATPl_x (AtP, 2x DGEMM, potrf/potrs incl. handle creation) + frees */
static double g_stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
static int g_stage_calls[2] = {0, 0};
struct StageClock {
	std::chrono::steady_clock::time_point t;
	StageClock() : t(std::chrono::steady_clock::now()) {}
	void lap(int stage)
	{
		cudaDeviceSynchronize();
		auto n = std::chrono::steady_clock::now();
		g_stage_ms[stage] += std::chrono::duration<double, std::milli>(n - t).count();
		t = n;
	}
};

#define REF_CHECK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { rc = (int)e__; goto done; } } while (0)

static int ref_threads_for_device(int dev)
{
	/* cudaWrapper.cpp:57-91 picks prop.maxThreadsPerBlock for every compute capability. */
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 1024;
	return prop.maxThreadsPerBlock;
}

extern "C" {

int ref_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

int ref_sizeof_point(void) { return (int)sizeof(ref_point_t); }
int ref_sizeof_grid_params(void) { return (int)sizeof(gridParameters); }

/* accumulated stage times (ms) since the last reset; calls_out[0] = NN calls, [1] = registerLS calls */
void ref_get_stage_ms(double *ms_out, int *calls_out, int reset)
{
	for (int k = 0; k < 8; k++) { if (ms_out) ms_out[k] = g_stage_ms[k]; if (reset) g_stage_ms[k] = 0.0; }
	if (calls_out) { calls_out[0] = g_stage_calls[0]; calls_out[1] = g_stage_calls[1]; }
	if (reset) g_stage_calls[0] = g_stage_calls[1] = 0;
}

int ref_warm_up(int dev)
{
	cudaError_t e = cudaSetDevice(dev);
	if (e != cudaSuccess) return (int)e;
	return (int)cudaWarmUpGPU();
}

/* Replays CCudaWrapper::semanticNearestNeighbourhoodSearch (cudaWrapper.cpp:344-424) on HOST
 * buffers: malloc + H2D both clouds, cudaCalculateGridParams, malloc table/buckets/nn,
 * cudaCalculateGrid, cudaSemanticNearestNeighborSearch, D2H nn, free everything.
 * Optional exports (NULL to skip): grid parameters, sorted table (n1 entries), dense bucket
 * table (up to buckets_cap entries).  Returns the cudaError_t value (0 = success),
 * -2 if buckets_cap is too small for the export. */
int ref_nn_search_host(const void *first, int n1, const void *second, int n2,
		float search_radius, float bucket_size, float bbox_extension,
		int max_inner, int max_outer, int *nn_out,
		void *params_out, void *table_out, void *buckets_out, long long buckets_cap)
{
	int rc = 0;
	int dev = 0;
	ref_point_t *d_first = 0, *d_second = 0;
	hashElement *d_table = 0;
	bucket *d_buckets = 0;
	int *d_nn = 0;
	gridParameters p;
	memset(&p, 0, sizeof(p));
	cudaGetDevice(&dev);
	int threads = ref_threads_for_device(dev);
	StageClock clk;
	g_stage_calls[0]++;

	REF_CHECK(cudaMalloc((void **)&d_first, (size_t)n1 * sizeof(ref_point_t)));
	REF_CHECK(cudaMemcpy(d_first, first, (size_t)n1 * sizeof(ref_point_t), cudaMemcpyHostToDevice));
	REF_CHECK(cudaMalloc((void **)&d_second, (size_t)n2 * sizeof(ref_point_t)));
	REF_CHECK(cudaMemcpy(d_second, second, (size_t)n2 * sizeof(ref_point_t), cudaMemcpyHostToDevice));
	clk.lap(0);
	REF_CHECK(cudaCalculateGridParams(d_first, n1, bucket_size, bucket_size, bucket_size, bbox_extension, p));
	clk.lap(1);
	REF_CHECK(cudaMalloc((void **)&d_table, (size_t)n1 * sizeof(hashElement)));
	REF_CHECK(cudaMalloc((void **)&d_buckets, (size_t)p.number_of_buckets * sizeof(bucket)));
	REF_CHECK(cudaMalloc((void **)&d_nn, (size_t)n2 * sizeof(int)));
	REF_CHECK(cudaCalculateGrid(threads, d_first, d_buckets, d_table, n1, p));
	clk.lap(2);
	REF_CHECK(cudaSemanticNearestNeighborSearch(threads, d_first, n1, d_second, n2, d_table, d_buckets, p,
			search_radius, max_inner, max_outer, d_nn));
	clk.lap(3);
	REF_CHECK(cudaMemcpy(nn_out, d_nn, (size_t)n2 * sizeof(int), cudaMemcpyDeviceToHost));
	if (params_out) memcpy(params_out, &p, sizeof(p));
	if (table_out) REF_CHECK(cudaMemcpy(table_out, d_table, (size_t)n1 * sizeof(hashElement), cudaMemcpyDeviceToHost));
	if (buckets_out) {
		if (p.number_of_buckets > buckets_cap) { rc = -2; goto done; }
		REF_CHECK(cudaMemcpy(buckets_out, d_buckets, (size_t)p.number_of_buckets * sizeof(bucket), cudaMemcpyDeviceToHost));
	}
done:
	cudaFree(d_first); cudaFree(d_second); cudaFree(d_table); cudaFree(d_buckets); cudaFree(d_nn);
	clk.lap(4);
	return rc;
}

/* Grid parameters only (lesson_16.cu:23-106) on a host cloud. */
int ref_grid_params_host(const void *cloud, int n, float rx, float ry, float rz, float ext, void *params_out)
{
	int rc = 0;
	ref_point_t *d = 0;
	gridParameters p;
	memset(&p, 0, sizeof(p));
	REF_CHECK(cudaMalloc((void **)&d, (size_t)n * sizeof(ref_point_t)));
	REF_CHECK(cudaMemcpy(d, cloud, (size_t)n * sizeof(ref_point_t), cudaMemcpyHostToDevice));
	REF_CHECK(cudaCalculateGridParams(d, n, rx, ry, rz, ext, p));
	memcpy(params_out, &p, sizeof(p));
done:
	cudaFree(d);
	return rc;
}

/* Device-side rigid transform (lesson_16.cu:1300-1339), in place on a host cloud copy.
 * m is a row-major 3x4 [R|t]. */
int ref_transform_host(void *cloud, int n, const float *m)
{
	int rc = 0;
	int dev = 0;
	ref_point_t *d = 0;
	cudaGetDevice(&dev);
	int threads = ref_threads_for_device(dev);
	REF_CHECK(cudaMalloc((void **)&d, (size_t)n * sizeof(ref_point_t)));
	REF_CHECK(cudaMemcpy(d, cloud, (size_t)n * sizeof(ref_point_t), cudaMemcpyHostToDevice));
	REF_CHECK(cudaTransformPointCloud(threads, d, n,
			m[0], m[4], m[8], m[1], m[5], m[9], m[2], m[6], m[10], m[3], m[7], m[11]));
	REF_CHECK(cudaMemcpy(cloud, d, (size_t)n * sizeof(ref_point_t), cudaMemcpyDeviceToHost));
done:
	cudaFree(d);
	return rc;
}

/* Normal equations exactly as the reference forms them (fill_A_l_cuda[_4DOF] →
 * cudaCompute_AtP → cuBLAS DGEMM ×2; cudaWrapper.cpp:523-545 + CCUDAAXBSolverWrapper.cpp:407-428)
 * but stopping before the solve so AtPA (dof×dof, column-major) and AtPl (dof) can be exported.
 * pose6 = {tx,ty,tz,om,fi,ka}. */
int ref_normal_equations_host(const void *obs, int n_obs, const double *pose6, int dof,
		double *AtPA_out, double *AtPl_out)
{
	int rc = 0;
	int dev = 0;
	double *d_A = 0, *d_P = 0, *d_l = 0, *d_AtP = 0, *d_AtPA = 0, *d_AtPl = 0;
	obs_nn_t *d_obs = 0;
	cublasHandle_t h = 0;
	cudaGetDevice(&dev);
	int threads = ref_threads_for_device(dev);
	int rows = dof, cols = n_obs * 3;
	if (dof != 6 && dof != 4) return -3;
	{
		CCUDA_AX_B_SolverWrapper wr(false, dev);
		if (cublasCreate(&h) != CUBLAS_STATUS_SUCCESS) { rc = -4; goto done; }
		REF_CHECK(cudaMalloc((void **)&d_A, (size_t)n_obs * 3 * dof * sizeof(double)));
		REF_CHECK(cudaMalloc((void **)&d_P, (size_t)n_obs * 3 * sizeof(double)));
		REF_CHECK(cudaMalloc((void **)&d_l, (size_t)n_obs * 3 * sizeof(double)));
		REF_CHECK(cudaMalloc((void **)&d_obs, (size_t)n_obs * sizeof(obs_nn_t)));
		REF_CHECK(cudaMemcpy(d_obs, obs, (size_t)n_obs * sizeof(obs_nn_t), cudaMemcpyHostToDevice));
		if (dof == 6)
			REF_CHECK(fill_A_l_cuda(threads, d_A, pose6[0], pose6[1], pose6[2], pose6[3], pose6[4], pose6[5], d_obs, n_obs, d_P, d_l));
		else
			REF_CHECK(fill_A_l_4DOFcuda(threads, d_A, pose6[0], pose6[1], pose6[2], pose6[3], pose6[4], pose6[5], d_obs, n_obs, d_P, d_l));
		REF_CHECK(cudaMalloc((void **)&d_AtP, sizeof(double) * rows * cols));
		REF_CHECK(cudaCompute_AtP(threads, d_A, d_P, d_AtP, rows, cols));
		REF_CHECK(cudaMalloc((void **)&d_AtPA, sizeof(double) * rows * rows));
		REF_CHECK(cudaMalloc((void **)&d_AtPl, sizeof(double) * rows));
		if (wr.multiplyCUBLAS(h, d_AtP, d_A, d_AtPA, rows, cols, rows) != CUBLAS_STATUS_SUCCESS) { rc = -5; goto done; }
		if (wr.multiplyCUBLAS(h, d_AtP, d_l, d_AtPl, rows, cols, 1) != CUBLAS_STATUS_SUCCESS) { rc = -5; goto done; }
		REF_CHECK(cudaDeviceSynchronize());
		REF_CHECK(cudaMemcpy(AtPA_out, d_AtPA, sizeof(double) * rows * rows, cudaMemcpyDeviceToHost));
		REF_CHECK(cudaMemcpy(AtPl_out, d_AtPl, sizeof(double) * rows, cudaMemcpyDeviceToHost));
done:
		cudaFree(d_A); cudaFree(d_P); cudaFree(d_l); cudaFree(d_obs); cudaFree(d_AtP); cudaFree(d_AtPA); cudaFree(d_AtPl);
		if (h) cublasDestroy(h);
	}
	return rc;
}

/* Replays CCudaWrapper::registerLS (dof=6, cudaWrapper.cpp:516-581) or registerLS_4DOF
 * (dof=4, cudaWrapper.cpp:583-648) on a HOST observation vector, including the per-call
 * solver-wrapper construction the reference does.  pose6 = {tx,ty,tz,om,fi,ka} is updated
 * in place; x_out (dof doubles, may be NULL) receives the raw solution.
 * Returns 0 on success, 1 if the solver reported failure, else the cudaError_t value. */
int ref_register_ls_host(const void *obs, int n_obs, double *pose6, int dof, double *x_out)
{
	int rc = 0;
	int dev = 0;
	double *d_A = 0, *d_P = 0, *d_l = 0;
	obs_nn_t *d_obs = 0;
	double x[6] = {0, 0, 0, 0, 0, 0};
	cudaGetDevice(&dev);
	int threads = ref_threads_for_device(dev);
	if (dof != 6 && dof != 4) return -3;
	StageClock clk;
	g_stage_calls[1]++;
	REF_CHECK(cudaMalloc((void **)&d_A, (size_t)n_obs * 3 * dof * sizeof(double)));
	REF_CHECK(cudaMalloc((void **)&d_P, (size_t)n_obs * 3 * sizeof(double)));
	REF_CHECK(cudaMalloc((void **)&d_l, (size_t)n_obs * 3 * sizeof(double)));
	REF_CHECK(cudaMalloc((void **)&d_obs, (size_t)n_obs * sizeof(obs_nn_t)));
	REF_CHECK(cudaMemcpy(d_obs, obs, (size_t)n_obs * sizeof(obs_nn_t), cudaMemcpyHostToDevice));
	clk.lap(5);
	if (dof == 6)
		REF_CHECK(fill_A_l_cuda(threads, d_A, pose6[0], pose6[1], pose6[2], pose6[3], pose6[4], pose6[5], d_obs, n_obs, d_P, d_l));
	else
		REF_CHECK(fill_A_l_4DOFcuda(threads, d_A, pose6[0], pose6[1], pose6[2], pose6[3], pose6[4], pose6[5], d_obs, n_obs, d_P, d_l));
	clk.lap(6);
	{
		CCUDA_AX_B_SolverWrapper *wr = new CCUDA_AX_B_SolverWrapper(false, dev);
		CCUDA_AX_B_SolverWrapper::CCUDA_AX_B_SolverWrapper_error e =
			wr->Solve_ATPA_ATPl_x_data_on_GPU(threads, d_A, d_P, d_l, x, dof, n_obs * 3, CCUDA_AX_B_SolverWrapper::chol);
		delete wr;
		if (e != CCUDA_AX_B_SolverWrapper::success) { rc = 1; goto done; }
	}
	pose6[0] += x[0];
	pose6[1] += x[1];
	pose6[2] += x[2];
	if (dof == 6) { pose6[3] += x[3]; pose6[4] += x[4]; pose6[5] += x[5]; }
	else { pose6[5] += x[3]; }
	if (x_out) memcpy(x_out, x, sizeof(double) * dof);
done:
	cudaFree(d_A); cudaFree(d_P); cudaFree(d_l); cudaFree(d_obs);
	clk.lap(7);
	return rc;
}


/* ---- pre-registration steps: the reference's own kernels behind the call sequences of
 * CCudaWrapper::removeNoiseNaive / downsampling / classify / findBestYaw (cudaWrapper.cpp:118-342, 662-836) ---- */

/* mode 0: removeNoiseNaive (threshold used), mode 1: downsampling.  markers_out: n bytes (the reference's bool d_markers). */
static int ref_mark_host(int mode, const void *cloud, int n, float resolution, float ext, int threshold, unsigned char *markers_out)
{
	int rc = 0, dev = 0;
	ref_point_t *d_cloud = 0;
	hashElement *d_table = 0;
	bucket *d_buckets = 0;
	bool *d_markers = 0;
	gridParameters p;
	memset(&p, 0, sizeof(p));
	cudaGetDevice(&dev);
	int threads = ref_threads_for_device(dev);
	REF_CHECK(cudaMalloc((void **)&d_cloud, (size_t)n * sizeof(ref_point_t)));
	REF_CHECK(cudaMemcpy(d_cloud, cloud, (size_t)n * sizeof(ref_point_t), cudaMemcpyHostToDevice));
	REF_CHECK(cudaCalculateGridParams(d_cloud, n, resolution, resolution, resolution, ext, p));
	REF_CHECK(cudaMalloc((void **)&d_table, (size_t)n * sizeof(hashElement)));
	REF_CHECK(cudaMalloc((void **)&d_buckets, (size_t)p.number_of_buckets * sizeof(bucket)));
	REF_CHECK(cudaCalculateGrid(threads, d_cloud, d_buckets, d_table, n, p));
	REF_CHECK(cudaMalloc((void **)&d_markers, (size_t)n * sizeof(bool)));
	if (mode == 0) REF_CHECK(cudaRemoveNoiseNaive(threads, d_markers, d_cloud, d_table, d_buckets, p, n, threshold));
	else REF_CHECK(cudaDownSample(threads, d_markers, d_table, d_buckets, p, n));
	REF_CHECK(cudaMemcpy(markers_out, d_markers, (size_t)n * sizeof(bool), cudaMemcpyDeviceToHost));
done:
	cudaFree(d_cloud); cudaFree(d_table); cudaFree(d_buckets); cudaFree(d_markers);
	return rc;
}

int ref_remove_noise_host(const void *cloud, int n, float resolution, float ext, int threshold, unsigned char *markers_out)
{
	return ref_mark_host(0, cloud, n, resolution, ext, threshold, markers_out);
}

int ref_downsample_host(const void *cloud, int n, float resolution, float ext, unsigned char *markers_out)
{
	return ref_mark_host(1, cloud, n, resolution, ext, 0, markers_out);
}

/* CCudaWrapper::classify (cudaWrapper.cpp:264-342) in place on a host cloud.  The grid and the two labelling kernels run
 * with threadsNV = maxThreadsPerBlock / 4 (cudaWrapper.cpp:85), the floor/ceiling pass with threads.  Optional exports:
 * mean_out = d_mean (3 floats per SORTED position), table_out = the sorted table (n records). */
int ref_classify_host(void *cloud, int n, float radius, float curvature_threshold, float ground_z, int plane_points, float ext,
		int max_inner, int max_outer, float vx, float vy, float vz, float *mean_out, void *table_out)
{
	int rc = 0, dev = 0;
	ref_point_t *d_cloud = 0;
	hashElement *d_table = 0;
	bucket *d_buckets = 0;
	simple_point3D *d_mean = 0;
	gridParameters p;
	memset(&p, 0, sizeof(p));
	cudaGetDevice(&dev);
	int threads = ref_threads_for_device(dev), threadsNV = threads / 4;
	REF_CHECK(cudaMalloc((void **)&d_cloud, (size_t)n * sizeof(ref_point_t)));
	REF_CHECK(cudaMemcpy(d_cloud, cloud, (size_t)n * sizeof(ref_point_t), cudaMemcpyHostToDevice));
	REF_CHECK(cudaCalculateGridParams(d_cloud, n, radius, radius, radius, ext, p));
	REF_CHECK(cudaMalloc((void **)&d_table, (size_t)n * sizeof(hashElement)));
	REF_CHECK(cudaMalloc((void **)&d_buckets, (size_t)p.number_of_buckets * sizeof(bucket)));
	REF_CHECK(cudaCalculateGrid(threadsNV, d_cloud, d_buckets, d_table, n, p));
	REF_CHECK(cudaMalloc((void **)&d_mean, (size_t)n * sizeof(simple_point3D)));
	REF_CHECK(cudaMemset(d_mean, 0, (size_t)n * sizeof(simple_point3D)));      /* upstream leaves rows of invalid positions unset */
	REF_CHECK(cudaSemanticLabelingPlaneEdges(threadsNV, d_cloud, n, d_table, d_buckets, d_mean, p, radius, max_inner, max_outer,
			curvature_threshold, plane_points, vx, vy, vz));
	REF_CHECK(cudaSemanticLabelingFloorCeiling(threads, d_cloud, n, ground_z));
	REF_CHECK(cudaMemcpy(cloud, d_cloud, (size_t)n * sizeof(ref_point_t), cudaMemcpyDeviceToHost));
	if (mean_out) REF_CHECK(cudaMemcpy(mean_out, d_mean, (size_t)n * sizeof(simple_point3D), cudaMemcpyDeviceToHost));
	if (table_out) REF_CHECK(cudaMemcpy(table_out, d_table, (size_t)n * sizeof(hashElement), cudaMemcpyDeviceToHost));
done:
	cudaFree(d_cloud); cudaFree(d_table); cudaFree(d_buckets); cudaFree(d_mean);
	return rc;
}

/* CCudaWrapper::findBestYaw (cudaWrapper.cpp:662-836) without Eigen: the caller supplies the matrices upstream takes from
 * Eigen — second_transform and first_transform.inverse() as row-major 3x4 (NULL = skip that transform) and ONE row-major
 * 3x4 yaw matrix per angle (yaw_mats: n_angles x 12).  counts_out: matched queries per angle
 * (cudaCountNumberOfSemanticNearestNeighbours).  Returns the index of the winning angle in *best_out (strict >). */
int ref_find_best_yaw_host(const void *first, int n1, const void *second, int n2, const float *second_m, const float *first_inv_m,
		float bucket_size, float ext, float search_radius, int max_inner, int max_outer,
		const float *yaw_mats, int n_angles, int *counts_out, int *best_out)
{
	int rc = 0, dev = 0;
	ref_point_t *d_first = 0, *d_second = 0, *d_rot = 0;
	hashElement *d_table = 0;
	bucket *d_buckets = 0;
	int *d_nn = 0;
	gridParameters p;
	memset(&p, 0, sizeof(p));
	cudaGetDevice(&dev);
	int threads = ref_threads_for_device(dev);
	int best = -1, best_n = 0;
	REF_CHECK(cudaMalloc((void **)&d_first, (size_t)n1 * sizeof(ref_point_t)));
	REF_CHECK(cudaMemcpy(d_first, first, (size_t)n1 * sizeof(ref_point_t), cudaMemcpyHostToDevice));
	REF_CHECK(cudaMalloc((void **)&d_second, (size_t)n2 * sizeof(ref_point_t)));
	REF_CHECK(cudaMemcpy(d_second, second, (size_t)n2 * sizeof(ref_point_t), cudaMemcpyHostToDevice));
	for (int k = 0; k < 2; k++) {
		const float *m = k == 0 ? second_m : first_inv_m;
		if (m) REF_CHECK(cudaTransformPointCloud(threads, d_second, n2, m[0], m[4], m[8], m[1], m[5], m[9], m[2], m[6], m[10], m[3], m[7], m[11]));
	}
	REF_CHECK(cudaMalloc((void **)&d_rot, (size_t)n2 * sizeof(ref_point_t)));
	REF_CHECK(cudaCalculateGridParams(d_first, n1, bucket_size, bucket_size, bucket_size, ext, p));
	REF_CHECK(cudaMalloc((void **)&d_table, (size_t)n1 * sizeof(hashElement)));
	REF_CHECK(cudaMalloc((void **)&d_buckets, (size_t)p.number_of_buckets * sizeof(bucket)));
	REF_CHECK(cudaMalloc((void **)&d_nn, (size_t)n2 * sizeof(int)));
	REF_CHECK(cudaCalculateGrid(threads, d_first, d_buckets, d_table, n1, p));
	for (int a = 0; a < n_angles; a++) {
		const float *m = yaw_mats + 12 * (size_t)a;
		int number_of_nn = 0;
		REF_CHECK(cudaTransformPointCloud(threads, d_second, n2, d_rot, n2, m[0], m[4], m[8], m[1], m[5], m[9], m[2], m[6], m[10], m[3], m[7], m[11]));
		REF_CHECK(cudaCountNumberOfSemanticNearestNeighbours(threads, d_first, n1, d_rot, n2, d_table, d_buckets, p, search_radius,
				max_inner, max_outer, d_nn, number_of_nn));
		if (counts_out) counts_out[a] = number_of_nn;
		if (number_of_nn > best_n) { best_n = number_of_nn; best = a; }
	}
	if (best_out) *best_out = best;
done:
	cudaFree(d_first); cudaFree(d_second); cudaFree(d_rot); cudaFree(d_table); cudaFree(d_buckets); cudaFree(d_nn);
	return rc;
}

} /* extern "C" */
