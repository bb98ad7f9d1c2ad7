/* shim_host.cpp — TEST PROGRAM: a ROS/PCL/Eigen-free C++ host that drives include/cuda_wrapper_shim.hpp exactly the
 * way gpu6DSLAM::registerLastArrivedScan drives the reference's CCudaWrapper (ref: src/gpu6DSLAM.cpp:264-422):
 * Euler round trip, CPU transform of both clouds, semanticNearestNeighbourhoodSearch, per-label weights,
 * observation assembly, registerLS / registerLS_4DOF, EulerToMatrix — and then the same registration through the
 * device-resident fused loop (registerPair).  tests/test_gpu_shim.py compares both against the oracle.
 *
 *   shim_host <first.bin> <second.bin> <poses.bin> <radius> <iterations> <dof> <out.bin>
 * first/second: raw 40-byte points; poses.bin: 32 floats (pose_first, pose_second row-major 4x4);
 * out.bin: 16 floats (legacy-loop pose) + 16 floats (fused-loop pose) + int32 n2 + n2 int32 (last legacy nn).
 */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuda_wrapper_shim.hpp"

struct Cloud {   /* the part of pcl::PointCloud the wrapper touches */
	std::vector<m3dreg_point> points;
	size_t size() const { return points.size(); }
	m3dreg_point &operator[](size_t i) { return points[i]; }
};

static bool read_cloud(const char *path, Cloud &c)
{
	FILE *f = std::fopen(path, "rb");
	if (!f) return false;
	std::fseek(f, 0, SEEK_END);
	long bytes = std::ftell(f);
	std::fseek(f, 0, SEEK_SET);
	c.points.resize((size_t)bytes / sizeof(m3dreg_point));
	size_t got = std::fread(c.points.data(), sizeof(m3dreg_point), c.points.size(), f);
	std::fclose(f);
	return got == c.points.size();
}

/* gpu6DSLAM::transformPointCloud (ref: src/gpu6DSLAM.cpp:635-663) with the rounding pattern of the reference's
 * device transform (src/lesson_16.cu:1354-1366) so results are comparable bit for bit with the oracle. */
static void transformPointCloud(Cloud &c, const m3dreg::Affine3f &m)
{
	for (auto &p : c.points) {
		float x = p.x, y = p.y, z = p.z, nx = p.normal_x, ny = p.normal_y, nz = p.normal_z;
		p.x = m(0, 3) + std::fmaf(m(0, 2), z, std::fmaf(m(0, 0), x, m(0, 1) * y));
		p.y = m(1, 3) + std::fmaf(m(1, 2), z, std::fmaf(m(1, 0), x, m(1, 1) * y));
		p.z = m(2, 3) + std::fmaf(m(2, 2), z, std::fmaf(m(2, 0), x, m(2, 1) * y));
		p.normal_x = std::fmaf(m(0, 2), nz, std::fmaf(m(0, 0), nx, m(0, 1) * ny));
		p.normal_y = std::fmaf(m(1, 2), nz, std::fmaf(m(1, 0), nx, m(1, 1) * ny));
		p.normal_z = std::fmaf(m(2, 2), nz, std::fmaf(m(2, 0), nx, m(2, 1) * ny));
	}
}

int main(int argc, char **argv)
{
	if (argc != 8) { std::fprintf(stderr, "usage: shim_host first second poses radius iterations dof out\n"); return 2; }
	Cloud first, second;
	if (!read_cloud(argv[1], first) || !read_cloud(argv[2], second)) { std::fprintf(stderr, "cannot read clouds\n"); return 2; }
	float poses[32];
	FILE *pf = std::fopen(argv[3], "rb");
	if (!pf || std::fread(poses, sizeof(float), 32, pf) != 32) { std::fprintf(stderr, "cannot read poses\n"); return 2; }
	std::fclose(pf);
	float radius = (float)std::atof(argv[4]);
	int iterations = std::atoi(argv[5]), dof = std::atoi(argv[6]);
	const float weight[4] = {10.0f, 1.0f, 10.0f, 10.0f};      /* ref: include/gpu6DSLAM.h:207-210 */
	const size_t obs_threshold = 100;                           /* ref: include/gpu6DSLAM.h:206 */

	m3dreg::Affine3f vmregistered_i, vmregistered_j;
	std::memcpy(vmregistered_i.m, poses, 64);
	std::memcpy(vmregistered_j.m, poses + 16, 64);
	std::vector<int> nearest_neighbour_indexes;

	try {
		CCudaWrapper cudaWrapper;
		cudaWrapper.warmUpGPU(0);

		/* ---- the reference's call pattern, one NN + one registerLS per iteration ------------------------------ */
		for (int it = 0; it < iterations; it++) {
			m3dreg::Vector3f omfika1, xyz1, omfika2, xyz2;
			m3dreg::Affine3f pose1, pose2;
			cudaWrapper.Matrix4ToEuler(vmregistered_i, omfika1, xyz1);
			cudaWrapper.EulerToMatrix(omfika1, xyz1, pose1);
			Cloud pc1 = first;
			transformPointCloud(pc1, pose1);
			observations_t obs;
			obs.om = omfika1.x(); obs.fi = omfika1.y(); obs.ka = omfika1.z();
			obs.tx = xyz1.x(); obs.ty = xyz1.y(); obs.tz = xyz1.z();
			cudaWrapper.Matrix4ToEuler(vmregistered_j, omfika2, xyz2);
			cudaWrapper.EulerToMatrix(omfika2, xyz2, pose2);
			Cloud pc2 = second;
			transformPointCloud(pc2, pose2);
			nearest_neighbour_indexes.assign(pc2.size(), -1);
			cudaWrapper.semanticNearestNeighbourhoodSearch(pc1, pc2, radius, radius, 1.0f, 100, 100, nearest_neighbour_indexes);
			int count[4] = {0, 0, 0, 0};
			for (size_t ii = 0; ii < nearest_neighbour_indexes.size(); ii++)
				if (nearest_neighbour_indexes[ii] != -1 && pc2[ii].label >= 0 && pc2[ii].label < 4) count[pc2[ii].label]++;
			for (size_t ii = 0; ii < nearest_neighbour_indexes.size(); ii++) {
				int k = nearest_neighbour_indexes[ii];
				if (k == -1) continue;
				m3dreg_obs_nn o;
				o.x0 = first[(size_t)k].x; o.y0 = first[(size_t)k].y; o.z0 = first[(size_t)k].z;
				o.x_diff = pc1[(size_t)k].x - pc2[ii].x; o.y_diff = pc1[(size_t)k].y - pc2[ii].y; o.z_diff = pc1[(size_t)k].z - pc2[ii].z;
				int L = pc2[ii].label;
				o.P = (L >= 0 && L < 4) ? weight[L] / count[L] : 0.0f;
				obs.vobs_nn.push_back(o);
			}
			if (obs.vobs_nn.size() > obs_threshold) {
				bool ok = dof == 6 ? cudaWrapper.registerLS(obs) : cudaWrapper.registerLS_4DOF(obs);
				if (ok) {
					m3dreg::Vector3f of((float)obs.om, (float)obs.fi, (float)obs.ka), t((float)obs.tx, (float)obs.ty, (float)obs.tz);
					cudaWrapper.EulerToMatrix(of, t, vmregistered_i);
				}
			}
		}

		/* ---- the same registration through the device-resident fused loop ------------------------------------- */
		m3dreg::Affine3f fused_i, fused_j;
		std::memcpy(fused_i.m, poses, 64);
		std::memcpy(fused_j.m, poses + 16, 64);
		cudaWrapper.uploadScan(0, first);
		cudaWrapper.uploadScan(1, second);
		m3dreg_reg_params prm;
		std::memset(&prm, 0, sizeof(prm));
		prm.search_radius = radius; prm.bucket_size = radius; prm.bbox_extension = 1.0f;
		prm.max_inner = 100; prm.max_outer = 100; prm.obs_threshold = (int)obs_threshold;
		for (int k = 0; k < 4; k++) prm.weight[k] = weight[k];
		prm.dof = dof; prm.mode = M3DREG_MODE_ICP;
		cudaWrapper.registerPair(0, 1, fused_i, fused_j, prm, iterations);

		FILE *out = std::fopen(argv[7], "wb");
		if (!out) return 2;
		std::fwrite(vmregistered_i.m, sizeof(float), 16, out);
		std::fwrite(fused_i.m, sizeof(float), 16, out);
		int n2 = (int)nearest_neighbour_indexes.size();
		std::fwrite(&n2, sizeof(int), 1, out);
		std::fwrite(nearest_neighbour_indexes.data(), sizeof(int), (size_t)n2, out);
		std::fclose(out);
	} catch (const m3dreg::system_error &e) {
		std::fprintf(stderr, "m3dreg::system_error %d: %s\n", e.code(), e.what());
		return 3;
	}
	return 0;
}
