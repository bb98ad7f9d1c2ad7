/* shim_preproc.cpp — TEST PROGRAM: the pre-registration sequence of gpu6DSLAM::registerSingleScan
 * (ref: src/gpu6DSLAM.cpp:47-85: cut-off, removeNoiseNaive, downsampling, classify) and a findBestYaw call
 * (ref: src/gpu6DSLAM.cpp:135-155) driven through include/cuda_wrapper_shim.hpp used as `class CCudaWrapper` — the
 * "full swap" arrangement of INTEGRATION.md section 3 — in a ROS/PCL/Eigen-free host.
 *
 *   shim_preproc <scan.bin> <first.bin> <second.bin> <out_processed.bin> <out_yaw.bin>
 * scan: an unprocessed scan; first/second: a classified pair for the yaw sweep (raw 40-byte points);
 * out_processed: the processed scan; out_yaw: float best angle + 12 floats of myaw. */
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cuda_wrapper_shim.hpp"

struct Cloud {   /* the part of pcl::PointCloud the wrapper touches (width / height as PCL keeps them) */
	std::vector<m3dreg_point> points;
	unsigned width = 0, height = 1;
	size_t size() const { return points.size(); }
	m3dreg_point &operator[](size_t i) { return points[i]; }
	void push_back(const m3dreg_point &p) { points.push_back(p); width = (unsigned)points.size(); }
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
	c.width = (unsigned)c.points.size();
	return got == c.points.size();
}

int main(int argc, char **argv)
{
	if (argc != 6) { std::fprintf(stderr, "usage: shim_preproc scan first second out_processed out_yaw\n"); return 2; }
	Cloud pc_in, first, other;
	if (!read_cloud(argv[1], pc_in) || !read_cloud(argv[2], first) || !read_cloud(argv[3], other)) { std::fprintf(stderr, "cannot read clouds\n"); return 2; }
	try {
		CCudaWrapper cudaWrapper;
		cudaWrapper.warmUpGPU(0);
		/* cut off (src/gpu6DSLAM.cpp:47-58), z window widened to the synthetic scanner's frame */
		Cloud pc;
		for (size_t i = 0; i < pc_in.size(); i++)
			if ((pc_in[i].z < 15 && pc_in[i].z > -3) && (pc_in[i].x * pc_in[i].x + pc_in[i].y * pc_in[i].y > 1.5)) pc.push_back(pc_in[i]);
		/* defaults of include/gpu6DSLAM.h:163-176 */
		cudaWrapper.removeNoiseNaive(pc, 0.5f, 1.0f, 3);
		cudaWrapper.downsampling(pc, 0.3f, 0.3f);
		cudaWrapper.classify(pc, 1.0f, 10.0f, 1.0f, 15, 1.0f, 100, 100, 0.0f, 0.0f, 0.0f);
		if (pc.width != pc.points.size() || pc.height != 1) { std::fprintf(stderr, "width/height not maintained\n"); return 3; }
		FILE *f = std::fopen(argv[4], "wb");
		std::fwrite(pc.points.data(), sizeof(m3dreg_point), pc.points.size(), f);
		std::fclose(f);
		/* findBestYaw with the defaults of include/gpu6DSLAM.h:212-219 on a coarser angle grid */
		m3dreg::Affine3f first_inv, second_tf, myaw;
		float best = 0.0f;
		cudaWrapper.findBestYaw(first, first_inv, other, second_tf, 1.0f, 1.0f, 0.3f, 50, 50, -12.0f, 12.0f, 1.5f, myaw, &best);
		f = std::fopen(argv[5], "wb");
		std::fwrite(&best, sizeof(float), 1, f);
		std::fwrite(myaw.m, sizeof(float), 12, f);
		std::fclose(f);
	} catch (const m3dreg::system_error &e) {
		std::fprintf(stderr, "m3dreg error %d: %s\n", e.code(), e.what());
		return 1;
	}
	return 0;
}
