/* cuda_wrapper_shim.hpp — header-only C++ mirror of the reference's L1 call surface, `class CCudaWrapper`
 * (ref: gpu_6dslam/gpu_6dslam/include/cudaWrapper.h:26-105, src/cudaWrapper.cpp), forwarding to the C ABI of
 * include/m3dreg.h.  Same method names, argument order/meaning and error behaviour as the reference, so
 * gpu6DSLAM.cpp's call sites (src/gpu6DSLAM.cpp:12-13, 313, 405-406, 478, 574-575) compile against it unchanged.
 *
 * The reference's signatures use pcl::PointCloud<lidar_pointcloud::PointXYZIRNLRGB> and Eigen::Affine3f.  Neither
 * PCL nor Eigen is needed here: the methods are templates over
 *     Cloud  — anything with `.points` (contiguous container of 40-byte points with `.data()` / `.size()`) and `.size()`
 *              (pcl::PointCloud<PointXYZIRNLRGB> qualifies: custom_point_types.h:8-20 is byte-identical to m3dreg_point);
 *     Affine — anything with `float operator()(int row, int col)` (read) and `float &operator()(int,int)` (write)
 *              (Eigen::Affine3f qualifies); Vec3 — anything with x() y() z() accessors (Eigen::Vector3f qualifies).
 * m3dreg::Affine3f / Vector3f below are minimal stand-ins used by this repository's own tests.
 *
 * Beyond the reference surface the shim exposes the device-resident fused loop (registerPair / scan store), which is
 * what a patched gpu6DSLAM::registerLastArrivedScan calls instead of one NN + one registerLS call per iteration, and the
 * multi-GPU Jacobi sweep (registerAll -> m3dreg_slam_sweep).
 *
 * Two ways to use it:
 *   - alone, as `class CCudaWrapper` (default), in a harness that only needs the registration surface;
 *   - NEXT TO the reference's own cudaWrapper.h (a staged migration: some call sites still on the reference's class) by
 *     compiling with -DM3DREG_SHIM_CLASS=CM3dRegWrapper -DM3DREG_SHIM_NO_OBSERVATIONS: the class then has its own
 *     name and the header does not redefine observations_t; its template methods accept the reference's observations_t
 *     and obs_nn_t as they are (same field names, 28-byte layout checked at compile time).
 */
#ifndef M3DREG_CUDA_WRAPPER_SHIM_HPP_
#define M3DREG_CUDA_WRAPPER_SHIM_HPP_

#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "m3dreg.h"

namespace m3dreg {

/* stand-ins for Eigen types (tests / ROS-free harnesses) */
struct Vector3f {
	float v[3] = {0, 0, 0};
	Vector3f() {}
	Vector3f(float a, float b, float c) { v[0] = a; v[1] = b; v[2] = c; }
	float &x() { return v[0]; } float &y() { return v[1]; } float &z() { return v[2]; }
	float x() const { return v[0]; } float y() const { return v[1]; } float z() const { return v[2]; }
};
struct Affine3f {
	float m[16];   /* row-major 4x4 */
	Affine3f() { setIdentity(); }
	void setIdentity() { for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0f : 0.0f; }
	float &operator()(int r, int c) { return m[r * 4 + c]; }
	float operator()(int r, int c) const { return m[r * 4 + c]; }
};

/* Stands in for thrust::system_error thrown by throw_on_cuda_error (ref: src/cudaWrapper.cpp:650-660): carries the
 * status (a cudaError_t value when > 0, an M3DREG_E_* code when < 0) and "file(line)". */
class system_error : public std::runtime_error {
public:
	system_error(int code, const std::string &where)
		: std::runtime_error(where + ": " + m3dreg_status_string(code)), code_(code) {}
	int code() const { return code_; }
private:
	int code_;
};

} /* namespace m3dreg */

/* ref: include/cudaWrapper.h:14-24 */
template <class Affine = m3dreg::Affine3f>
struct observations_tmpl {
	std::vector<m3dreg_obs_nn> vobs_nn;
	Affine m_pose;
	double om = 0, fi = 0, ka = 0, tx = 0, ty = 0, tz = 0;
};
#ifndef M3DREG_SHIM_NO_OBSERVATIONS
typedef observations_tmpl<> observations_t;
#endif

#ifndef M3DREG_SHIM_CLASS
#define M3DREG_SHIM_CLASS CCudaWrapper
#endif

class M3DREG_SHIM_CLASS {
public:
	/* ref: src/cudaWrapper.cpp:4-18 — the context is created lazily by warmUpGPU, like the reference picks its device there */
	M3DREG_SHIM_CLASS() : threads(0), threadsNV(0), cuda_device(0), ctx_(nullptr) {}
	~M3DREG_SHIM_CLASS() { if (ctx_) m3dreg_destroy(ctx_); }
	M3DREG_SHIM_CLASS(const M3DREG_SHIM_CLASS &) = delete;
	M3DREG_SHIM_CLASS &operator=(const M3DREG_SHIM_CLASS &) = delete;

	/* ref: src/cudaWrapper.cpp:36-44 */
	void warmUpGPU(int cudaDevice)
	{
		if (!ctx_ || cuda_device != cudaDevice) {
			if (ctx_) { m3dreg_destroy(ctx_); ctx_ = nullptr; }
			throw_on_cuda_error(m3dreg_create(&ctx_, cudaDevice), __FILE__, __LINE__);
			cuda_device = cudaDevice;
		}
		throw_on_cuda_error(m3dreg_warm_up(ctx_), __FILE__, __LINE__);
		getNumberOfAvailableThreads(cudaDevice, threads, threadsNV);
	}

	/* ref: src/cudaWrapper.cpp:46-91 — kept for source compatibility; launch shapes are internal to the kernels now */
	int getNumberOfAvailableThreads(int) { return 1024; }
	bool getNumberOfAvailableThreads(int, int &t, int &tNV) { t = 1024; tNV = 256; return true; }
	void coutMemoryStatus() {}

	/* ref: src/cudaWrapper.cpp:344-424 */
	template <class Cloud>
	void semanticNearestNeighbourhoodSearch(Cloud &first_point_cloud, Cloud &second_point_cloud, float search_radius,
			float bucket_size, float bounding_box_extension, int max_number_considered_in_INNER_bucket,
			int max_number_considered_in_OUTER_bucket, std::vector<int> &nearest_neighbour_indexes)
	{
		static_assert(sizeof(first_point_cloud.points[0]) == sizeof(m3dreg_point), "point type must be the 40-byte PointXYZIRNLRGB");
		if (nearest_neighbour_indexes.size() != second_point_cloud.size()) return;     /* ref: cudaWrapper.cpp:354 */
		int st = m3dreg_semantic_nn_host(context(),
				reinterpret_cast<const m3dreg_point *>(first_point_cloud.points.data()), (int)first_point_cloud.points.size(),
				reinterpret_cast<const m3dreg_point *>(second_point_cloud.points.data()), (int)second_point_cloud.points.size(),
				search_radius, bucket_size, bounding_box_extension, max_number_considered_in_INNER_bucket,
				max_number_considered_in_OUTER_bucket, nearest_neighbour_indexes.data());
		throw_on_cuda_error(st, __FILE__, __LINE__);
	}

	/* ---- pre-registration steps (every incoming scan, src/gpu6DSLAM.cpp:63-85) ------------------------------------
	 * ref: src/cudaWrapper.cpp:118-179 — the cloud is replaced by its survivors, original order kept */
	template <class Cloud>
	void removeNoiseNaive(Cloud &point_cloud, float resolution, float bounding_box_extension, int number_of_points_in_bucket_threshold)
	{
		static_assert(sizeof(point_cloud.points[0]) == sizeof(m3dreg_point), "point type must be the 40-byte PointXYZIRNLRGB");
		const int n = (int)point_cloud.points.size();
		if (n == 0) return;
		int kept = 0;
		m3dreg_point *p = reinterpret_cast<m3dreg_point *>(point_cloud.points.data());
		throw_on_cuda_error(m3dreg_remove_noise_host(context(), p, n, resolution, bounding_box_extension,
				number_of_points_in_bucket_threshold, p, &kept, nullptr), __FILE__, __LINE__);
		shrink(point_cloud, kept);
	}
	/* ref: src/cudaWrapper.cpp:181-262 */
	template <class Cloud>
	void downsampling(Cloud &point_cloud, float resolution, float bounding_box_extension)
	{
		static_assert(sizeof(point_cloud.points[0]) == sizeof(m3dreg_point), "point type must be the 40-byte PointXYZIRNLRGB");
		const int n = (int)point_cloud.points.size();
		if (n == 0) return;
		int kept = 0;
		m3dreg_point *p = reinterpret_cast<m3dreg_point *>(point_cloud.points.data());
		throw_on_cuda_error(m3dreg_downsample_host(context(), p, n, resolution, bounding_box_extension, p, &kept, nullptr), __FILE__, __LINE__);
		shrink(point_cloud, kept);
	}
	/* ref: src/cudaWrapper.cpp:264-342 — normals and labels rewritten in place */
	template <class Cloud>
	void classify(Cloud &point_cloud, float normal_vectors_search_radius, float curvature_threshold, float ground_Z_coordinate_threshold,
			int number_of_points_needed_for_plane_threshold, float bounding_box_extension, int max_number_considered_in_INNER_bucket,
			int max_number_considered_in_OUTER_bucket, float viewpointX, float viewpointY, float viewpointZ)
	{
		static_assert(sizeof(point_cloud.points[0]) == sizeof(m3dreg_point), "point type must be the 40-byte PointXYZIRNLRGB");
		const int n = (int)point_cloud.points.size();
		if (n == 0) return;
		throw_on_cuda_error(m3dreg_classify_host(context(), reinterpret_cast<m3dreg_point *>(point_cloud.points.data()), n,
				normal_vectors_search_radius, curvature_threshold, ground_Z_coordinate_threshold, number_of_points_needed_for_plane_threshold,
				bounding_box_extension, max_number_considered_in_INNER_bucket, max_number_considered_in_OUTER_bucket,
				viewpointX, viewpointY, viewpointZ, nullptr, nullptr), __FILE__, __LINE__);
	}
	/* ref: src/cudaWrapper.cpp:662-836 — myaw receives the rotation about Z with the most matches (untouched when no angle
	 * matches anything, as upstream).  `first_transform_inverse` is passed in by the caller: upstream inverts an
	 * Eigen::Affine3f, which this header does not depend on (Eigen callers pass first_transform.inverse()). */
	template <class Cloud, class Affine>
	void findBestYaw(Cloud &first_point_cloud, const Affine &first_transform_inverse, Cloud &second_point_cloud, const Affine &second_transform,
			float bucket_size, float bounding_box_extension, float search_radius, int max_number_considered_in_INNER_bucket,
			int max_number_considered_in_OUTER_bucket, float angle_start, float angle_finish, float angle_step, Affine &myaw,
			float *best_angle_deg = nullptr)
	{
		float m1[12], m2[12];
		for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) { m1[r * 4 + c] = first_transform_inverse(r, c); m2[r * 4 + c] = second_transform(r, c); }
		float best = angle_start;
		int best_n = 0;
		throw_on_cuda_error(m3dreg_find_best_yaw_host(context(),
				reinterpret_cast<const m3dreg_point *>(first_point_cloud.points.data()), (int)first_point_cloud.points.size(),
				reinterpret_cast<const m3dreg_point *>(second_point_cloud.points.data()), (int)second_point_cloud.points.size(),
				m2, m1, bucket_size, bounding_box_extension, search_radius, max_number_considered_in_INNER_bucket,
				max_number_considered_in_OUTER_bucket, angle_start, angle_finish, angle_step, &best, &best_n, nullptr, 0), __FILE__, __LINE__);
		if (best_angle_deg) *best_angle_deg = best;
		if (best_n > 0) {
			const float rad = (float)((double)best * 3.14159265358979323846 / 180.0);
			float of[3] = {0.0f, 0.0f, rad}, t[3] = {0.0f, 0.0f, 0.0f}, a[16];
			m3dreg_euler_to_matrix(of, t, a);
			for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) myaw(r, c) = a[r * 4 + c];
		}
	}

	/* ref: src/cudaWrapper.cpp:427-468 (double[16] column-major variant) */
	static void Matrix4ToEuler(const double *alignxf, double *rPosTheta, double *rPos)
	{
		float m[16], of[3], t[3];
		for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) m[r * 4 + c] = (float)alignxf[c * 4 + r];
		m3dreg_matrix4_to_euler(m, of, t);
		for (int k = 0; k < 3; k++) { rPosTheta[k] = of[k]; if (rPos) rPos[k] = alignxf[12 + k]; }
	}
	/* ref: src/cudaWrapper.cpp:470-504 */
	template <class Affine, class Vec3>
	static void Matrix4ToEuler(const Affine &m, Vec3 &omfika, Vec3 &xyz)
	{
		float a[16], of[3], t[3];
		for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) a[r * 4 + c] = (r < 3) ? m(r, c) : (c == 3 ? 1.0f : 0.0f);
		m3dreg_matrix4_to_euler(a, of, t);
		omfika.x() = of[0]; omfika.y() = of[1]; omfika.z() = of[2];
		xyz.x() = t[0]; xyz.y() = t[1]; xyz.z() = t[2];
	}
	/* ref: src/cudaWrapper.cpp:506-514 */
	template <class Affine, class Vec3>
	static void EulerToMatrix(const Vec3 &omfika, const Vec3 &xyz, Affine &m)
	{
		float of[3] = {omfika.x(), omfika.y(), omfika.z()}, t[3] = {xyz.x(), xyz.y(), xyz.z()}, a[16];
		m3dreg_euler_to_matrix(of, t, a);
		for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) m(r, c) = a[r * 4 + c];
	}

	/* ref: src/cudaWrapper.cpp:516-581 (6-DOF) and 583-648 (4-DOF: x, y, z, yaw).  Returns false when the system
	 * cannot be solved (the reference prints and carries on, CCUDAAXBSolverWrapper.cpp:520-522); obs is then unchanged. */
	template <class Obs>
	bool registerLS(Obs &obs) { return register_ls(obs, 6); }
	template <class Obs>
	bool registerLS_4DOF(Obs &obs) { return register_ls(obs, 4); }

	/* ref: src/cudaWrapper.cpp:650-660 */
	void throw_on_cuda_error(int code, const char *file, int line)
	{
		if (code != 0) {
			std::stringstream ss;
			ss << file << "(" << line << ")";
			throw m3dreg::system_error(code, ss.str());
		}
	}

	/* ---- beyond the reference surface: device-resident scans + fused loop --------------------------------------
	 * uploadScan keeps scan `slot` in HBM; registerPair runs `iterations` complete iterations of
	 * gpu6DSLAM::registerLastArrivedScan (src/gpu6DSLAM.cpp:264-422) on the device and updates pose_first. */
	template <class Cloud>
	void uploadScan(int slot, const Cloud &cloud)
	{
		throw_on_cuda_error(m3dreg_scan_upload(context(), slot, reinterpret_cast<const m3dreg_point *>(cloud.points.data()),
				(int)cloud.points.size(), 0), __FILE__, __LINE__);
	}
	template <class Affine>
	bool registerPair(int first_slot, int second_slot, Affine &pose_first, const Affine &pose_second,
			const m3dreg_reg_params &params, int iterations, m3dreg_icp_stats *stats = nullptr)
	{
		float a[16], b[16];
		for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) {
			a[r * 4 + c] = (r < 3) ? pose_first(r, c) : (c == 3 ? 1.0f : 0.0f);
			b[r * 4 + c] = (r < 3) ? pose_second(r, c) : (c == 3 ? 1.0f : 0.0f);
		}
		m3dreg_icp_stats local;
		int st = m3dreg_icp_pair(context(), first_slot, second_slot, a, b, &params, iterations, stats ? stats : &local);
		throw_on_cuda_error(st, __FILE__, __LINE__);
		for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) pose_first(r, c) = a[r * 4 + c];
		return (stats ? stats : &local)->last_status == 0;
	}

	/* ref: gpu6DSLAM::registerAll(cudaWrapper, radius, bucket, number_of_last_EOZ) (src/gpu6DSLAM.cpp:424-597) as ONE call:
	 * scans 0..poses.size()-1 must have been uploaded (uploadScan); the poses (any container of Affine) are updated in
	 * place.  With a communicator attached (attachNccl, one wrapper per GPU / rank, the same poses on every rank) the
	 * pairs are sharded over the ranks and the normal-equation blocks all-reduced inside the library. */
	template <class AffineVec>
	int registerAll(AffineVec &poses, const m3dreg_reg_params &params, float distance_threshold, size_t number_of_last_EOZ,
			std::vector<int> *status = nullptr, m3dreg_sweep_stats *stats = nullptr)
	{
		const int n = (int)poses.size();
		if (n == 0 || (size_t)n < number_of_last_EOZ) return 0;                    /* ref: src/gpu6DSLAM.cpp:428 */
		std::vector<float> flat((size_t)n * 16);
		for (int k = 0; k < n; k++)
			for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++)
				flat[(size_t)k * 16 + r * 4 + c] = (r < 3) ? poses[k](r, c) : (c == 3 ? 1.0f : 0.0f);
		m3dreg_slam_params sp;
		sp.reg = params;
		sp.distance_threshold = distance_threshold;
		sp.first_optimised = n - (int)number_of_last_EOZ;
		std::vector<int> st_local((size_t)n, 0);
		int st = m3dreg_slam_sweep(context(), n, flat.data(), &sp, st_local.data(), stats);
		throw_on_cuda_error(st, __FILE__, __LINE__);
		for (int k = 0; k < n; k++)
			for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) poses[k](r, c) = flat[(size_t)k * 16 + r * 4 + c];
		int solved = 0;
		for (int k = sp.first_optimised; k < n; k++) solved += st_local[(size_t)k] == 0 ? 1 : 0;
		if (status) *status = st_local;
		return solved;
	}
	/* existing ncclComm_t of the host application (or 0 / world 1 to detach); see also m3dreg_nccl_init */
	void attachNccl(void *nccl_comm, int rank, int world)
	{
		throw_on_cuda_error(m3dreg_nccl_attach(context(), nccl_comm, rank, world), __FILE__, __LINE__);
	}

	m3dreg_ctx *context()
	{
		if (!ctx_) warmUpGPU(cuda_device);
		return ctx_;
	}

	int threads;
	int threadsNV;
	int cuda_device;

private:
	/* keep the first `kept` points (pcl::PointCloud keeps width / height beside the vector: unorganised cloud of `kept` points) */
	template <class Cloud>
	static auto shrink(Cloud &c, int kept) -> decltype(c.width = 0u, void())
	{
		c.points.resize((size_t)kept);
		c.width = (unsigned)kept; c.height = 1;
	}
	template <class Cloud>
	static void shrink(Cloud &c, long kept) { c.points.resize((size_t)kept); }

	template <class Obs>
	bool register_ls(Obs &obs, int dof)
	{
		if (obs.vobs_nn.empty()) return false;
		static_assert(sizeof(obs.vobs_nn[0]) == sizeof(m3dreg_obs_nn), "obs_nn_t must be the 28-byte reference layout");
		double pose6[6] = {obs.tx, obs.ty, obs.tz, obs.om, obs.fi, obs.ka};
		int st = m3dreg_register_ls_host(context(), reinterpret_cast<const m3dreg_obs_nn *>(obs.vobs_nn.data()),
				(int)obs.vobs_nn.size(), pose6, dof, nullptr);
		if (st == M3DREG_E_NOT_SPD) {
			std::fprintf(stderr, "problem with solving Ax=B\n");
			return false;
		}
		throw_on_cuda_error(st, __FILE__, __LINE__);
		obs.tx = pose6[0]; obs.ty = pose6[1]; obs.tz = pose6[2];
		obs.om = pose6[3]; obs.fi = pose6[4]; obs.ka = pose6[5];
		return true;
	}

	m3dreg_ctx *ctx_;
};

#endif /* M3DREG_CUDA_WRAPPER_SHIM_HPP_ */
