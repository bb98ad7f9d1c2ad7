/* m3dreg_kernels.cuh — sm_100a kernels of the registration hot path.
 *
 * Data layout in HBM (all device-resident, SoA, 16-byte records so every access is one LDG.128/STG.128):
 *   xyzl[i] = {x, y, z, label-bits}      nrm[i] = {nx, ny, nz, 0}
 * for (a) each stored scan in its local frame, (b) the first cloud transformed into the global frame in
 * original order (g_*), (c) the same in grid-sorted order (s_*), (d) the queries (q_*).
 * Grid: keys/vals u32 ping-pong buffers for the LSD radix sort, a dense bucket table in the reference's
 * 12-byte layout, nn[] in query order.
 *
 * Arithmetic contract: every float operation that decides an index is written with
 * explicit round-to-nearest intrinsics in the exact association nvcc 12.9 gives the reference's kernels
 * (src/lesson_16.cu of the reference, PTX inspected), so results are bit-identical:
 *   cell    = cvt.rzi( div.rn( sub(v, min), res ) )                       (lesson_16.cu:124-126, 578-580)
 *   dist    = fma(dz,dz, fma(dx,dx, dy*dy))                               (lesson_16.cu:658-660)
 *   dot     = fma(nz,nnz, fma(nx,nnx, ny*nny))                            (lesson_16.cu:662-664)
 *   v'      = t + fma(r02,z, fma(r00,x, r01*y))                           (lesson_16.cu:1354-1356)
 * No -use_fast_math, no reciprocal multiplication.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/m3dreg.h"
#include "nn_core.cuh"

namespace m3d {

/* Programmatic dependent launch (sm_90+): the host launches every kernel with programmatic stream serialisation, so
 * a kernel is set up and its blocks are scheduled as the blocks of the previous kernel of the stream exit, instead of
 * after the whole grid has drained and a fresh launch has been processed.  Nothing the previous kernel wrote may be
 * read, and nothing it reads may be overwritten, before griddepcontrol.wait returns — it is the FIRST statement of
 * every kernel in this file.  No kernel calls griddepcontrol.launch_dependents: triggering early leaves the next grid's
 * blocks resident and waiting, which measured SLOWER than plain stream order on small scans (C1: 134 vs 122 us per
 * iteration), while the implicit trigger at block exit measured 98 us (C1) and 189 vs 211 us (C2).
 * Without the launch attribute the instruction is a no-op. */
__device__ __forceinline__ void pdl_enter()
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

constexpr int kMomentCount = 24;   /* S, M1[3], M2[6], L1[3], L2[9], n_obs, (pad) */
constexpr int kPartialCols = 28;   /* row width of the per-block partial sums (ICP uses 24 of them, NDT all 28) */
constexpr int kNeqCount = 28;      /* 21 upper-tri AtPA + 6 AtPl + count */

/* flags[] slots */
enum { FLAG_ERROR = 0, FLAG_COUNT = 4 };

/* Device-resident pose / solve state of the fused loop. */
struct PoseState {
	float  m[16];        /* stored pose (row-major 4x4), the reference's vmregistered[i]          */
	float  pose1[16];    /* Euler round trip of m, used to transform the first cloud this iteration */
	double pose6[6];     /* tx,ty,tz,om,fi,ka = double(float Euler) that the solve updates          */
	double x[6];         /* last solution                                                          */
	double neq[kNeqCount];
	long long n_obs;
	int    status;       /* 0 | M3DREG_E_NOT_SPD | M3DREG_E_TOO_FEW_OBS                            */
	int    iterations;
};

/* ---- ordered-uint encoding of floats for atomicMin/Max -------------------------------------------- */
__device__ __forceinline__ uint32_t f2o(float f)
{
	uint32_t u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(uint32_t o)
{
	return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

/* Block-level min/max of xyz and one atomic per block per bound. blockDim.x multiple of 32, <= 1024. */
__device__ __forceinline__ void block_bounds_commit(float mnx, float mny, float mnz, float mxx, float mxy, float mxz,
		uint32_t *bounds)
{
	__shared__ float sm[6][32];
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	mnx = warp_min(mnx); mny = warp_min(mny); mnz = warp_min(mnz);
	mxx = warp_max(mxx); mxy = warp_max(mxy); mxz = warp_max(mxz);
	if (lane == 0) { sm[0][w] = mnx; sm[1][w] = mny; sm[2][w] = mnz; sm[3][w] = mxx; sm[4][w] = mxy; sm[5][w] = mxz; }
	__syncthreads();
	if (w == 0) {
		float a = lane < nw ? sm[0][lane] : INFINITY, b = lane < nw ? sm[1][lane] : INFINITY, c = lane < nw ? sm[2][lane] : INFINITY;
		float d = lane < nw ? sm[3][lane] : -INFINITY, e = lane < nw ? sm[4][lane] : -INFINITY, f = lane < nw ? sm[5][lane] : -INFINITY;
		a = warp_min(a); b = warp_min(b); c = warp_min(c);
		d = warp_max(d); e = warp_max(e); f = warp_max(f);
		if (lane == 0) {
			atomicMin(&bounds[0], f2o(a)); atomicMin(&bounds[1], f2o(b)); atomicMin(&bounds[2], f2o(c));
			atomicMax(&bounds[3], f2o(d)); atomicMax(&bounds[4], f2o(e)); atomicMax(&bounds[5], f2o(f));
		}
	}
}

/* rigid transform applied to a stored scan's points on the fly (the fused loop never materialises the transformed
 * cloud): same operation sequence as k_transform_soa, so the bits equal the ones the bounding box was made from */
struct PointXform { float r[12]; int on; };

__device__ __forceinline__ float4 xform_point(const PointXform &x, const float4 &p)
{
	if (!x.on) return p;
	return make_float4(__fadd_rn(x.r[3], __fmaf_rn(x.r[2], p.z, __fmaf_rn(x.r[0], p.x, __fmul_rn(x.r[1], p.y)))),
			__fadd_rn(x.r[7], __fmaf_rn(x.r[6], p.z, __fmaf_rn(x.r[4], p.x, __fmul_rn(x.r[5], p.y)))),
			__fadd_rn(x.r[11], __fmaf_rn(x.r[10], p.z, __fmaf_rn(x.r[8], p.x, __fmul_rn(x.r[9], p.y)))), p.w);
}

/* ---- AoS (reference 40-byte point) <-> SoA ----------------------------------------------------------- */
__global__ void k_unpack_points(const m3dreg_point *__restrict__ in, int n, float4 *__restrict__ xyzl, float4 *__restrict__ nrm)
{
	pdl_enter();
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint2 *p = reinterpret_cast<const uint2 *>(in + i);   /* 40-byte records are 8-byte aligned */
	uint2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4);
	/* a={x,y} b={z,intensity} c={ring,normal_x} d={normal_y,normal_z} e={label,rgb} */
	xyzl[i] = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(b.x), __uint_as_float(e.x));
	nrm[i] = make_float4(__uint_as_float(c.y), __uint_as_float(d.x), __uint_as_float(d.y), 0.0f);
}

/* min/max straight from an AoS cloud (stage-level m3dreg_grid_params). */
__global__ void k_bounds_aos(const m3dreg_point *__restrict__ in, int n, uint32_t *bounds)
{
	pdl_enter();
	float mnx = INFINITY, mny = INFINITY, mnz = INFINITY, mxx = -INFINITY, mxy = -INFINITY, mxz = -INFINITY;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint2 *p = reinterpret_cast<const uint2 *>(in + i);
		uint2 a = __ldg(p), b = __ldg(p + 1);
		float x = __uint_as_float(a.x), y = __uint_as_float(a.y), z = __uint_as_float(b.x);
		mnx = fminf(mnx, x); mny = fminf(mny, y); mnz = fminf(mnz, z);
		mxx = fmaxf(mxx, x); mxy = fmaxf(mxy, y); mxz = fmaxf(mxz, z);
	}
	block_bounds_commit(mnx, mny, mnz, mxx, mxy, mxz, bounds);
}

__global__ void k_reset_bounds(uint32_t *bounds)
{
	pdl_enter();
	if (threadIdx.x < 3) bounds[threadIdx.x] = 0xFFFFFFFFu;
	else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;
}

/* Rigid transform of an AoS cloud, bit-compatible with the reference's device kernel (lesson_16.cu:1341-1367). */
__global__ void k_transform_aos(const m3dreg_point *__restrict__ in, m3dreg_point *__restrict__ out, int n,
		float r00, float r01, float r02, float t0, float r10, float r11, float r12, float t1,
		float r20, float r21, float r22, float t2)
{
	pdl_enter();
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	m3dreg_point p = in[i];
	float x = p.x, y = p.y, z = p.z, nx = p.normal_x, ny = p.normal_y, nz = p.normal_z;
	p.x = __fadd_rn(t0, __fmaf_rn(r02, z, __fmaf_rn(r00, x, __fmul_rn(r01, y))));
	p.y = __fadd_rn(t1, __fmaf_rn(r12, z, __fmaf_rn(r10, x, __fmul_rn(r11, y))));
	p.z = __fadd_rn(t2, __fmaf_rn(r22, z, __fmaf_rn(r20, x, __fmul_rn(r21, y))));
	p.normal_x = __fmaf_rn(r02, nz, __fmaf_rn(r00, nx, __fmul_rn(r01, ny)));
	p.normal_y = __fmaf_rn(r12, nz, __fmaf_rn(r10, nx, __fmul_rn(r11, ny)));
	p.normal_z = __fmaf_rn(r22, nz, __fmaf_rn(r20, nx, __fmul_rn(r21, ny)));
	out[i] = p;
}

/* Transform a stored scan (SoA) by the 3x4 matrix at `m` (DEVICE memory, row-major 4x4) and, when bounds != 0,
 * reduce the bounding box of the result in the same pass (replaces 3x thrust::minmax_element, lesson_16.cu:34-45). */
template <bool WITH_BOUNDS>
__global__ void k_transform_soa(const float4 *__restrict__ in_xyzl, const float4 *__restrict__ in_nrm, int n,
		const float *__restrict__ m, float4 *__restrict__ out_xyzl, float4 *__restrict__ out_nrm, uint32_t *bounds)
{
	pdl_enter();
	float r00 = __ldg(m + 0), r01 = __ldg(m + 1), r02 = __ldg(m + 2), t0 = __ldg(m + 3);
	float r10 = __ldg(m + 4), r11 = __ldg(m + 5), r12 = __ldg(m + 6), t1 = __ldg(m + 7);
	float r20 = __ldg(m + 8), r21 = __ldg(m + 9), r22 = __ldg(m + 10), t2 = __ldg(m + 11);
	float mnx = INFINITY, mny = INFINITY, mnz = INFINITY, mxx = -INFINITY, mxy = -INFINITY, mxz = -INFINITY;
	/* four independent loads per thread in flight (one per trip measured 16 us for 16 MB: latency, not bandwidth) */
	const int stride = gridDim.x * blockDim.x;
	for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
		float4 pv[4], qv[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int i = i0 + k * stride;
			pv[k] = __ldg(in_xyzl + (i < n ? i : i0));
			if (out_nrm) qv[k] = __ldg(in_nrm + (i < n ? i : i0));
		}
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int i = i0 + k * stride;
			if (i >= n) continue;
			const float4 p = pv[k];
			float x = __fadd_rn(t0, __fmaf_rn(r02, p.z, __fmaf_rn(r00, p.x, __fmul_rn(r01, p.y))));
			float y = __fadd_rn(t1, __fmaf_rn(r12, p.z, __fmaf_rn(r10, p.x, __fmul_rn(r11, p.y))));
			float z = __fadd_rn(t2, __fmaf_rn(r22, p.z, __fmaf_rn(r20, p.x, __fmul_rn(r21, p.y))));
			if (out_xyzl) out_xyzl[i] = make_float4(x, y, z, p.w);      /* 0: bounding box only (the fused loop never stores the transformed cloud) */
			if (out_nrm) {      /* the fused loop rotates only the candidates' normals (k_build_candidates), not all of them */
				const float4 q = qv[k];
				float nx = __fmaf_rn(r02, q.z, __fmaf_rn(r00, q.x, __fmul_rn(r01, q.y)));
				float ny = __fmaf_rn(r12, q.z, __fmaf_rn(r10, q.x, __fmul_rn(r11, q.y)));
				float nz = __fmaf_rn(r22, q.z, __fmaf_rn(r20, q.x, __fmul_rn(r21, q.y)));
				out_nrm[i] = make_float4(nx, ny, nz, 0.0f);
			}
			if (WITH_BOUNDS) {
				mnx = fminf(mnx, x); mny = fminf(mny, y); mnz = fminf(mnz, z);
				mxx = fmaxf(mxx, x); mxy = fmaxf(mxy, y); mxz = fmaxf(mxz, z);
			}
		}
	}
	if (WITH_BOUNDS) block_bounds_commit(mnx, mny, mnz, mxx, mxy, mxz, bounds);
}

/* ---- query segments of a batched sweep step ----------------------------------------------------------------------
 * registerAll matches every neighbour j of a scan i against the SAME grid of i (gpu6DSLAM.cpp:469-565).  The sweep
 * therefore concatenates the (transformed) clouds of several neighbours into one query array — segment s occupies
 * [off_s, off_s + n_s), off_s a multiple of kSegChunk, padding queries sit at +inf (outside every box: no match) — and
 * runs ONE search and ONE moment reduction over it.  seg_of_chunk[q / kSegChunk] names the segment of query q; the
 * per-label match counts (the observation weights are weight_label / count_label PER PAIR, gpu6DSLAM.cpp:490-562)
 * are kept per segment. */
constexpr int kSegChunk = 256;
constexpr int kMaxSegs = 64;
struct SweepSeg {
	const float4 *sx, *sn;      /* the neighbour's stored scan in query order (local frame) */
	int n, off, pose, pad;      /* points, first query, index of its pose in the pose array */
};

/* One block per chunk.  The chunk -> segment map the search and the moment reduction use afterwards is produced here
 * (binary search over the segments' first queries) instead of on the host. */
__global__ void __launch_bounds__(kSegChunk) k_transform_segments(const SweepSeg *__restrict__ segs, int n_segs, int *__restrict__ seg_of_chunk,
		const float *__restrict__ poses, float4 *__restrict__ out_xyzl, float4 *__restrict__ out_nrm)
{
	pdl_enter();
	__shared__ int s_off[kMaxSegs];
	if ((int)threadIdx.x < n_segs) s_off[threadIdx.x] = segs[threadIdx.x].off;
	__syncthreads();
	int lo = 0, hi = n_segs - 1;      /* last segment whose first query is <= this chunk's first query */
	const int q0 = blockIdx.x * kSegChunk;
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (s_off[mid] <= q0) lo = mid; else hi = mid - 1;
	}
	if (threadIdx.x == 0) seg_of_chunk[blockIdx.x] = lo;
	const SweepSeg sg = segs[lo];
	const float *m = poses + 16 * (size_t)sg.pose;
	const int q = q0 + threadIdx.x, i = q - sg.off;
	if (i >= sg.n) {
		out_xyzl[q] = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(-1));
		out_nrm[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		return;
	}
	/* same operation sequence as k_transform_soa */
	const float r00 = __ldg(m + 0), r01 = __ldg(m + 1), r02 = __ldg(m + 2), t0 = __ldg(m + 3);
	const float r10 = __ldg(m + 4), r11 = __ldg(m + 5), r12 = __ldg(m + 6), t1 = __ldg(m + 7);
	const float r20 = __ldg(m + 8), r21 = __ldg(m + 9), r22 = __ldg(m + 10), t2 = __ldg(m + 11);
	const float4 p = __ldg(sg.sx + i), nq = __ldg(sg.sn + i);
	const float x = __fadd_rn(t0, __fmaf_rn(r02, p.z, __fmaf_rn(r00, p.x, __fmul_rn(r01, p.y))));
	const float y = __fadd_rn(t1, __fmaf_rn(r12, p.z, __fmaf_rn(r10, p.x, __fmul_rn(r11, p.y))));
	const float z = __fadd_rn(t2, __fmaf_rn(r22, p.z, __fmaf_rn(r20, p.x, __fmul_rn(r21, p.y))));
	out_xyzl[q] = make_float4(x, y, z, p.w);
	const float nx = __fmaf_rn(r02, nq.z, __fmaf_rn(r00, nq.x, __fmul_rn(r01, nq.y)));
	const float ny = __fmaf_rn(r12, nq.z, __fmaf_rn(r10, nq.x, __fmul_rn(r11, nq.y)));
	const float nz = __fmaf_rn(r22, nq.z, __fmaf_rn(r20, nq.x, __fmul_rn(r21, nq.y)));
	out_nrm[q] = make_float4(nx, ny, nz, 0.0f);
}

/* Grid parameters from the reduced bounds, on the device (host part of cudaCalculateGridParams, lesson_16.cu:64-91):
 * max += ext; min -= ext; nb = int((max-min)/res + 1); B = nbX*nbY*nbZ.
 * Returns false (and number_of_buckets = 0) when B overflows int32 or the planned capacity. */
__device__ __forceinline__ bool grid_params_from_bounds_dev(const uint32_t *__restrict__ bounds, float rx, float ry, float rz, float ext,
		long long bucket_cap, m3dreg_grid_params &g)
{
	float mnx = o2f(bounds[0]), mny = o2f(bounds[1]), mnz = o2f(bounds[2]);
	float mxx = o2f(bounds[3]), mxy = o2f(bounds[4]), mxz = o2f(bounds[5]);
	mxx = __fadd_rn(mxx, ext); mnx = __fsub_rn(mnx, ext);
	mxy = __fadd_rn(mxy, ext); mny = __fsub_rn(mny, ext);
	mxz = __fadd_rn(mxz, ext); mnz = __fsub_rn(mnz, ext);
	int nbx = (int)__fadd_rn(__fdiv_rn(__fsub_rn(mxx, mnx), rx), 1.0f);
	int nby = (int)__fadd_rn(__fdiv_rn(__fsub_rn(mxy, mny), ry), 1.0f);
	int nbz = (int)__fadd_rn(__fdiv_rn(__fsub_rn(mxz, mnz), rz), 1.0f);
	long long nb = (long long)nbx * (long long)nby * (long long)nbz;
	g.bounding_box_min_X = mnx; g.bounding_box_min_Y = mny; g.bounding_box_min_Z = mnz;
	g.bounding_box_max_X = mxx; g.bounding_box_max_Y = mxy; g.bounding_box_max_Z = mxz;
	g.number_of_buckets_X = nbx; g.number_of_buckets_Y = nby; g.number_of_buckets_Z = nbz;
	g.resolution_X = rx; g.resolution_Y = ry; g.resolution_Z = rz;
	g._pad0 = 0; g._pad1 = 0;
	bool ok = !(nbx <= 0 || nby <= 0 || nbz <= 0 || nb > 2147483647LL || nb > bucket_cap) && nn_columns_usable(nbx, nby, nbz);
	g.number_of_buckets = ok ? nb : 0;
	return ok;
}

/* Warp-level histogram update for one item per lane: lanes holding the same bin as their left neighbour form a run,
 * and only the first lane of a run touches shared memory (bucket keys of a scan are strongly coherent, so a warp
 * usually holds one or two runs).  bin < 0 marks an idle lane. */
__device__ __forceinline__ void warp_run_hist(int bin, uint32_t *sh, int lane)
{
	const unsigned full = 0xffffffffu;
	int prev = __shfl_up_sync(full, bin, 1);
	bool head = (lane == 0) || (prev != bin);
	unsigned heads = __ballot_sync(full, head);
	if (head && bin >= 0) {
		unsigned later = heads & ~((2u << lane) - 1u);      /* heads strictly above this lane */
		int len = (later ? __ffs(later) - 1 : 32) - lane;
		atomicAdd(&sh[bin], (uint32_t)len);
	}
}

constexpr int kRadixBits = 8;
constexpr int kRadixSize = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;

/* Head of the grid build of the fused loop, one block per sort tile:
 *   - grid parameters from the reduced bounds (every block recomputes them; block 0 publishes them, raises the error
 *     flag and resets the occupied-bucket counter),
 *   - dense bucket table reset to {-1,-1,0} (kernel_initializeBuckets, lesson_16.cu:131-140),
 *   - bucket key of every point (kernel_initializeIndByKey + kernel_getIndexOfBucketForPoints, lesson_16.cu:109-129;
 *     values are the implicit original indices and are not stored),
 *   - the first radix pass's per-tile digit histogram, and zeroing of the later passes' histograms. */
template <int ITEMS>
__global__ void __launch_bounds__(kSortThreads) k_grid_head(const float4 *__restrict__ xyzl, const float *__restrict__ pose, int n, const uint32_t *__restrict__ bounds,
		float rx, float ry, float rz, float ext, long long bucket_cap, m3dreg_grid_params *__restrict__ gp_out, int *__restrict__ flags,
		unsigned int *__restrict__ cell_count, m3dreg_bucket *__restrict__ buckets, uint32_t *__restrict__ keys,
		int tiles, int passes, uint32_t *__restrict__ hist)
{
	pdl_enter();
	__shared__ uint32_t sh[kRadixSize];
	PointXform xf;      /* pose != 0: xyzl is the stored scan (local frame), transformed here exactly as the box pass did */
	xf.on = pose != nullptr;
#pragma unroll
	for (int k = 0; k < 12; k++) xf.r[k] = xf.on ? __ldg(pose + k) : 0.0f;
	m3dreg_grid_params g;
	bool ok = grid_params_from_bounds_dev(bounds, rx, ry, rz, ext, bucket_cap, g);
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		*gp_out = g;
		*cell_count = 0u;
		if (!ok) atomicExch(&flags[FLAG_ERROR], M3DREG_E_TOO_MANY_BUCKETS);
	}
	if (!ok) return;
	sh[threadIdx.x] = 0;
	{   /* bucket table reset, 12-byte records written as a flat int stream: -1,-1,0,-1,-1,0,... */
		long long total = g.number_of_buckets * 3;
		int *flat = reinterpret_cast<int *>(buckets);
		for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
			flat[i] = (i % 3 == 2) ? 0 : -1;
		/* histograms of passes 1.. are accumulated by the scatter kernels */
		long long hz = (long long)(passes - 1) * kRadixSize * tiles;
		for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hz; i += (long long)gridDim.x * blockDim.x)
			hist[(size_t)kRadixSize * tiles + i] = 0u;
	}
	__syncthreads();
	const int nby = g.number_of_buckets_Y, nbz = g.number_of_buckets_Z;
	const int lane = threadIdx.x & 31;
	const int base = blockIdx.x * (kSortThreads * ITEMS);
	/* all of the tile's loads first (ITEMS independent 16-byte loads per thread in flight), then the keys */
	float4 pt[ITEMS];
#pragma unroll
	for (int j = 0; j < ITEMS; j++) {
		const int i = base + j * kSortThreads + threadIdx.x;
		pt[j] = __ldg(xyzl + (i < n ? i : 0));
	}
#pragma unroll
	for (int j = 0; j < ITEMS; j++) {
		int i = base + j * kSortThreads + threadIdx.x;
		int bin = -1;
		if (i < n) {
			float4 p = xform_point(xf, pt[j]);
			int ix = cell_of(p.x, g.bounding_box_min_X, rx), iy = cell_of(p.y, g.bounding_box_min_Y, ry), iz = cell_of(p.z, g.bounding_box_min_Z, rz);
			uint32_t k = (uint32_t)(ix * nby * nbz + iy * nbz + iz);
			keys[i] = k;
			bin = (int)(k & (kRadixSize - 1));
		}
		warp_run_hist(bin, sh, lane);
	}
	__syncthreads();
	hist[(size_t)threadIdx.x * tiles + blockIdx.x] = sh[threadIdx.x];
}

__global__ void k_keys_aos(const m3dreg_point *__restrict__ in, int n, const m3dreg_grid_params *__restrict__ gp,
		uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
	pdl_enter();
	float mnx = gp->bounding_box_min_X, mny = gp->bounding_box_min_Y, mnz = gp->bounding_box_min_Z;
	float rx = gp->resolution_X, ry = gp->resolution_Y, rz = gp->resolution_Z;
	int nby = gp->number_of_buckets_Y, nbz = gp->number_of_buckets_Z;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint2 *p = reinterpret_cast<const uint2 *>(in + i);
		uint2 a = __ldg(p), b = __ldg(p + 1);
		int ix = cell_of(__uint_as_float(a.x), mnx, rx), iy = cell_of(__uint_as_float(a.y), mny, ry), iz = cell_of(__uint_as_float(b.x), mnz, rz);
		keys[i] = (uint32_t)(ix * nby * nbz + iy * nbz + iz);
		vals[i] = (uint32_t)i;
	}
}

/* Query-role ordering of a stored scan: key = label (2 bits) | 27-bit Morton code of a fine local grid, so that 32
 * consecutive queries are one small patch of one semantic surface.  Any order is correct (the NN result does not
 * depend on query order); this one makes the warp-cooperative search cheap. */
__device__ __forceinline__ uint32_t spread3(uint32_t v)   /* 9 bits -> every third bit */
{
	v &= 0x1FFu;
	v = (v | (v << 16)) & 0x030000FFu;
	v = (v | (v << 8)) & 0x0300F00Fu;
	v = (v | (v << 4)) & 0x030C30C3u;
	v = (v | (v << 2)) & 0x09249249u;
	return v;
}

__global__ void k_keys_presort(const m3dreg_point *__restrict__ in, int n, float mnx, float mny, float mnz, float inv_res,
		uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
	pdl_enter();
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint2 *p = reinterpret_cast<const uint2 *>(in + i);
		uint2 a = __ldg(p), b = __ldg(p + 1), e = __ldg(p + 4);
		int ix = min(511, max(0, (int)((__uint_as_float(a.x) - mnx) * inv_res)));
		int iy = min(511, max(0, (int)((__uint_as_float(a.y) - mny) * inv_res)));
		int iz = min(511, max(0, (int)((__uint_as_float(b.x) - mnz) * inv_res)));
		keys[i] = ((e.x & 3u) << 27) | spread3((uint32_t)ix) | (spread3((uint32_t)iy) << 1) | (spread3((uint32_t)iz) << 2);
		vals[i] = (uint32_t)i;
	}
}

/* ---- stable LSD radix sort by bucket key (replaces thrust::sort + compareHashElements, lesson_16.cu:213-227) ----
 * 8-bit digits, ceil(bits/8) passes where bits = bit width of the bucket capacity.  Each pass:
 *   k_radix_hist    per-tile digit histogram  -> hist[digit * tiles + tile]
 *   k_radix_scan    exclusive scan of that digit-major matrix (single block)
 *   k_radix_scatter stable in-tile ranking (warp match-any multisplit) + scatter
 * Stability gives ties in ascending original index, i.e. exactly the permutation of the reference's stable merge sort.
 * Keys are non-negative ints (valid cells), so unsigned order == signed order. */
template <int ITEMS>
__global__ void __launch_bounds__(kSortThreads) k_radix_hist(const uint32_t *__restrict__ keys, int n, int shift,
		int tiles, uint32_t *__restrict__ hist, const m3dreg_grid_params *__restrict__ gp)
{
	pdl_enter();
	__shared__ uint32_t sh[kRadixSize];
	if (gp && gp->number_of_buckets <= 0) return;
	sh[threadIdx.x] = 0;
	__syncthreads();
	int base = blockIdx.x * (kSortThreads * ITEMS);
	int lane = threadIdx.x & 31;
#pragma unroll
	for (int j = 0; j < ITEMS; j++) {
		int i = base + j * kSortThreads + threadIdx.x;
		int bin = i < n ? (int)((__ldg(keys + i) >> shift) & (kRadixSize - 1)) : -1;
		warp_run_hist(bin, sh, lane);
	}
	__syncthreads();
	hist[(size_t)threadIdx.x * tiles + blockIdx.x] = sh[threadIdx.x];
}

/* One block per digit: exclusive scan of that digit's row of per-tile counts (in place) and the row total. */
__global__ void __launch_bounds__(256) k_radix_scan(uint32_t *__restrict__ hist, int tiles, uint32_t *__restrict__ digit_tot,
		const m3dreg_grid_params *__restrict__ gp)
{
	pdl_enter();
	__shared__ uint32_t warp_tot[8];
	__shared__ uint32_t carry;
	if (gp && gp->number_of_buckets <= 0) return;
	uint32_t *row = hist + (size_t)blockIdx.x * tiles;
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < tiles; base += 256) {
		int i = base + threadIdx.x;
		uint32_t v = i < tiles ? row[i] : 0u, incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		if (lane == 31) warp_tot[w] = incl;
		__syncthreads();
		uint32_t off = carry;
		for (int k = 0; k < w; k++) off += warp_tot[k];
		if (i < tiles) row[i] = off + incl - v;
		__syncthreads();
		if (threadIdx.x == 255) carry = off + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) digit_tot[blockIdx.x] = carry;
}

template <int ITEMS>
__global__ void __launch_bounds__(kSortThreads) k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
		uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int n, int shift, int tiles,
		const uint32_t *__restrict__ hist, const uint32_t *__restrict__ digit_tot, const m3dreg_grid_params *__restrict__ gp,
		uint32_t *__restrict__ next_hist)
{
	pdl_enter();
	__shared__ uint32_t wcnt[kSortWarps][kRadixSize];   /* per-warp digit counts, then per-warp exclusive offsets */
	__shared__ uint32_t gbase[kRadixSize];
	__shared__ uint32_t wtot[kSortWarps];
	if (gp && gp->number_of_buckets <= 0) return;
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < kSortWarps; k++) wcnt[k][threadIdx.x] = 0;
	{   /* exclusive scan of the 256 digit totals (thread = digit) + this tile's prefix inside the digit */
		uint32_t v = __ldg(digit_tot + threadIdx.x), incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		if (lane == 31) wtot[w] = incl;
		__syncthreads();
		uint32_t off = 0;
		for (int k = 0; k < w; k++) off += wtot[k];
		gbase[threadIdx.x] = off + incl - v + __ldg(hist + (size_t)threadIdx.x * tiles + blockIdx.x);
	}
	__syncthreads();

	int wbase = blockIdx.x * (kSortThreads * ITEMS) + w * (32 * ITEMS);
	uint32_t key[ITEMS], val[ITEMS], rank[ITEMS];
	const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
	for (int j = 0; j < ITEMS; j++) {
		int i = wbase + j * 32 + lane;
		bool valid = i < n;
		key[j] = valid ? __ldg(keys_in + i) : 0xFFFFFFFFu;
		val[j] = valid ? (vals_in ? __ldg(vals_in + i) : (uint32_t)i) : 0u;      /* vals_in == 0: implicit original indices */
		uint32_t d = (key[j] >> shift) & (kRadixSize - 1);
		/* invalid lanes must not disturb the counts of digit 0xFF: give them their own match group */
		uint32_t mk = valid ? d : (0x100u + lane);
		uint32_t peers = __match_any_sync(0xffffffffu, mk);
		int leader = __ffs(peers) - 1;
		uint32_t old = 0;
		if (lane == leader && valid) {
			old = wcnt[w][d];
			wcnt[w][d] = old + __popc(peers);
		}
		old = __shfl_sync(0xffffffffu, old, leader);
		rank[j] = old + __popc(peers & lt);
		__syncwarp();
	}
	__syncthreads();
	{   /* exclusive scan over warps for digit = threadIdx.x */
		uint32_t run = 0;
#pragma unroll
		for (int k = 0; k < kSortWarps; k++) {
			uint32_t c = wcnt[k][threadIdx.x];
			wcnt[k][threadIdx.x] = run;
			run += c;
		}
	}
	__syncthreads();
#pragma unroll
	for (int j = 0; j < ITEMS; j++) {
		int i = wbase + j * 32 + lane;
		if (i < n) {
			uint32_t d = (key[j] >> shift) & (kRadixSize - 1);
			uint32_t pos = gbase[d] + wcnt[w][d] + rank[j];
			keys_out[pos] = key[j];
			vals_out[pos] = val[j];
			rank[j] = pos;
		}
	}
	if (next_hist) {
		/* the next pass's per-tile digit histogram, accumulated where the items land: one atomic per group of equal
		 * (next digit, destination tile) inside the warp */
#pragma unroll
		for (int j = 0; j < ITEMS; j++) {
			int i = wbase + j * 32 + lane;
			bool valid = i < n;
			uint32_t slot = valid ? ((key[j] >> (shift + kRadixBits)) & (kRadixSize - 1)) * (uint32_t)tiles + rank[j] / (uint32_t)(kSortThreads * ITEMS)
					: (0xFFFFFF00u + lane);
			uint32_t peers = __match_any_sync(0xffffffffu, slot);
			if (valid && lane == __ffs(peers) - 1) atomicAdd(next_hist + slot, (uint32_t)__popc(peers));
		}
	}
}

/* ---- dense bucket table (kernel_initializeBuckets / updateBuckets / countNumberOfPointsForBuckets,
 *      lesson_16.cu:131-189) + gather of the first cloud into sorted order ---------------------------------- */
__global__ void k_init_buckets(m3dreg_bucket *__restrict__ buckets, const m3dreg_grid_params *__restrict__ gp, long long nb_host)
{
	pdl_enter();
	long long nb = gp ? gp->number_of_buckets : nb_host;
	/* 12-byte records written as a flat int stream: -1,-1,0,-1,-1,0,... */
	long long total = nb * 3;
	int *flat = reinterpret_cast<int *>(buckets);
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
		flat[i] = (i % 3 == 2) ? 0 : -1;
}

__device__ __forceinline__ int lower_bound_u32(const uint32_t *__restrict__ a, int n, uint32_t key)
{
	int lo = 0, hi = n;
	while (lo < hi) {
		int mid = (lo + hi) >> 1;
		if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
	}
	return lo;
}

/* One thread per sorted position p (kernel_updateBuckets + kernel_countNumberOfPointsForBuckets, lesson_16.cu:142-189).
 * The thread at the START of a run writes index_begin, the one at its END index_end, and every warp adds the length of
 * its piece of the run to number_of_points (integer atomics: deterministic) — no search, every step is one memory
 * latency.  Reference quirk reproduced (lesson_16.cu:148-158): when element 0 is alone in its bucket, the run that
 * starts at position 1 never gets index_begin (it receives index_end = 1 instead), so that bucket reads {-1, end, 0}.
 * Its index_end is a write race upstream (1 vs run end); we store the run end.  The table must have been reset to
 * {-1,-1,0}.  Optionally materialises the reference's hashElement table and the compact list of searchable buckets. */
__global__ void k_finalize_grid(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int n,
		const m3dreg_grid_params *__restrict__ gp, m3dreg_bucket *__restrict__ buckets, m3dreg_hash_element *__restrict__ table_out,
		uint32_t *__restrict__ cell_list, unsigned int *__restrict__ cell_count)
{
	pdl_enter();
	if (gp && gp->number_of_buckets <= 0) return;
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int nround = (n + 31) & ~31;      /* whole warps stay together for the ballots below */
	const uint32_t k0 = __ldg(keys), k1 = n > 1 ? __ldg(keys + 1) : k0;
	const bool has_quirk = n > 1 && k0 != k1;
	const int stride = gridDim.x * blockDim.x;
	for (int p0 = blockIdx.x * blockDim.x + threadIdx.x; p0 < nround; p0 += 4 * stride) {
		/* four positions per trip: their key loads (and the neighbours') go out together */
		uint32_t kc[4], kp[4], kn[4], vv[4];
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const int p = p0 + u * stride;
			const bool valid = p < n;
			kc[u] = valid ? __ldg(keys + p) : 0xFFFFFFFFu;
			kp[u] = (valid && p > 0) ? __ldg(keys + p - 1) : 0xFFFFFFFFu;
			kn[u] = (valid && p < n - 1) ? __ldg(keys + p + 1) : 0xFFFFFFFFu;
			vv[u] = (valid && table_out) ? __ldg(vals + p) : 0u;
		}
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const int p = p0 + u * stride;
			if (p >= nround) break;                    /* warp-uniform: nround and the strides are multiples of 32 */
			bool listed = false, valid = p < n;
			const uint32_t k = kc[u];
			bool quirk = false;
			if (valid) {
				if (table_out) {
					m3dreg_hash_element h;
					h.index_of_point = (int)vv[u];
					h.index_of_bucket = (int)k;
					table_out[p] = h;
				}
				quirk = has_quirk && k == k1;
				bool run_start = (p == 0) || (kp[u] != k);
				bool run_end = (p == n - 1) || (kn[u] != k);
				int *bp = reinterpret_cast<int *>(buckets + k);
				if (run_start && !quirk) bp[0] = p;
				if (run_end) { bp[1] = p + 1; listed = !quirk; }
			}
			{   /* number_of_points: one atomic per piece of a run inside this warp */
				uint32_t prev = __shfl_up_sync(full, k, 1);
				bool head = (lane == 0) || (prev != k);
				unsigned heads = __ballot_sync(full, head);
				if (head && valid && !quirk) {
					unsigned later = heads & ~((2u << lane) - 1u);
					int len = (later ? __ffs(later) - 1 : 32) - lane;
					if (p + len > n) len = n - p;
					atomicAdd(reinterpret_cast<int *>(buckets + k) + 2, len);
				}
			}
			if (cell_list) {   /* compact list of the searchable buckets (order irrelevant): one atomic per warp */
				unsigned m = __ballot_sync(full, listed);
				if (m) {
					unsigned int base = 0;
					if (lane == 0) base = atomicAdd(cell_count, (unsigned int)__popc(m));
					base = __shfl_sync(full, base, 0);
					if (listed) cell_list[base + __popc(m & ((1u << lane) - 1u))] = k;
				}
			}
		}
	}
}

/* Same list from an externally supplied bucket table (stage-level m3dreg_nn_search). */
__global__ void k_list_cells(const uint32_t *__restrict__ keys, int n, const m3dreg_bucket *__restrict__ buckets,
		uint32_t *__restrict__ cell_list, unsigned int *__restrict__ cell_count)
{
	pdl_enter();
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int nround = (n + 31) & ~31;
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nround; p += gridDim.x * blockDim.x) {
		bool listed = false;
		uint32_t k = 0;
		if (p < n) {
			k = __ldg(keys + p);
			bool run_end = (p == n - 1) || (__ldg(keys + p + 1) != k);
			listed = run_end && __ldg(reinterpret_cast<const int *>(buckets + k) + 2) > 0;
		}
		unsigned m = __ballot_sync(full, listed);
		if (m) {
			unsigned int base = 0;
			if (lane == 0) base = atomicAdd(cell_count, (unsigned int)__popc(m));
			base = __shfl_sync(full, base, 0);
			if (listed) cell_list[base + __popc(m & ((1u << lane) - 1u))] = k;
		}
	}
}

/* ---- candidate sets -------------------------------------------------------------------------------------------
 * The reference never looks at every point of a bucket: it walks sorted positions begin, begin+s, begin+2s, ... with
 * s = n / cap (lesson_16.cu:628-640), i.e. at most 2*cap-1 CANDIDATES per bucket, and only those can ever be a
 * result.  k_build_candidates gathers exactly those, per bucket, into the bucket's own [begin, begin+ncand) range of
 * the compact arrays (a bucket's candidates always fit inside its own range, so no prefix sum is needed), grouped by
 * (label & 3, sub-cell) bin with a u16 table of bin offsets — layout and exactness argument in nn_core.cuh.  The order
 * inside a bin is arbitrary (counting sort with shared-memory atomics); each record carries its sorted position l for
 * the tie-break and the result, which is all the search needs.  Two sets exist when the INNER and OUTER caps differ (different strides). */
constexpr int kBuildWarps = 4;
constexpr int kBuildChunk = 4;            /* gathers per lane in flight while binning */
constexpr int kBuildTabMax = 4 * 64 + 1;   /* bins + 1 at the finest level */

struct NormalRotation { float r[9]; int on; };   /* rotation applied to the candidates' normals (same rounding as the transform kernels) */

struct CellGeom { float mnx, mny, mnz, rx, ry, rz; int cx, cy, cz; };

__device__ __forceinline__ void prefetch_l2(const void *p)
{
	asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

/* One warp builds one bucket's candidate set.  Two sweeps over the bucket's candidates (walk positions begin + k * iter):
 * sweep 1 bins every candidate (histogram with shared-memory atomics; the normals the second sweep will want are
 * prefetched into L2), the warp scans the bin counts into the bucket's offset table, sweep 2 bins them again (table value
 * and point are L1/L2 hits by then) and places them.  Nothing per candidate is kept in registers between the sweeps:
 * the kernel is a chain of dependent gathers, resident warps are what hides it, so registers matter more than the
 * ~60 repeated instructions per candidate.
 * VALS_CG: the sorted table was written earlier in the SAME launch by other blocks (grid megakernel): read it past L1. */
template <bool VALS_CG = false>
__device__ __forceinline__ void build_cell_candidates(const uint32_t *vals, const float4 *__restrict__ src_xyzl,
		const float4 *__restrict__ src_nrm, const float4 *__restrict__ loc_src, const NormalRotation &rot, const PointXform &xf, int begin, int npts, int cap, int tables, const CellGeom &g,
		const CandSet &set, uint32_t *hist, int lane, unsigned long long *dbg = nullptr)
{
	const unsigned full = 0xffffffffu;
	if (npts <= 0 || cap <= 0 || begin < 0) return;
	const int iter = candidate_stride(npts, cap);
	const int ncand = (npts + iter - 1) / iter;
	const int level = tables ? nn_level(npts) : -1;
	const int nbins = level >= 0 ? (4 << (3 * level)) : 0;
	const int lv = level < 0 ? 0 : level;
	auto stamp = [&](int slot) {
		if (dbg && lane == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); dbg[slot] = t_; }
	};
	stamp(24);
	if (dbg && lane == 0) dbg[29] = (unsigned long long)ncand | ((unsigned long long)npts << 32);
	const float wx = nn_subcell_width(g.rx, lv), wy = nn_subcell_width(g.ry, lv), wz = nn_subcell_width(g.rz, lv);
	auto table_value = [&](int k) { return VALS_CG ? __ldcg(vals + begin + k * iter) : __ldg(vals + begin + k * iter); };
	auto bin_of = [&](const float4 &p) {
		const int ux = nn_col(p.x, g.mnx, wx, g.cx, lv), uy = nn_col(p.y, g.mny, wy, g.cy, lv), uz = nn_col(p.z, g.mnz, wz, g.cz, lv);
		return (uint32_t)nn_bin(__float_as_int(p.w), ux, uy, uz, lv);
	};
	auto store = [&](int pos, int k, uint32_t vi, const float4 &c0, const float4 &c) {
		const float4 n = __ldg(src_nrm + vi);
		{	/* the point in the frame the moment reduction wants (loc_src == 0: the source cloud itself) + its original index */
			const float4 l0 = loc_src ? __ldg(loc_src + vi) : c0;
			set.loc[pos] = make_float4(l0.x, l0.y, l0.z, __uint_as_float(vi));
		}
		float4 nn = n;
		if (rot.on) {
			nn.x = __fmaf_rn(rot.r[2], n.z, __fmaf_rn(rot.r[0], n.x, __fmul_rn(rot.r[1], n.y)));
			nn.y = __fmaf_rn(rot.r[5], n.z, __fmaf_rn(rot.r[3], n.x, __fmul_rn(rot.r[4], n.y)));
			nn.z = __fmaf_rn(rot.r[8], n.z, __fmaf_rn(rot.r[6], n.x, __fmul_rn(rot.r[7], n.y)));
		}
		set.xyzl[pos] = make_float4(c.x, c.y, c.z, __int_as_float(begin + k * iter));
		set.nrm[pos] = make_float4(nn.x, nn.y, nn.z, c.w);
	};
	if (level < 0) {     /* no table: candidates in walk order */
		for (int k = lane; k < ncand; k += 32) {
			const uint32_t vi = table_value(k);
			const float4 c0 = __ldg(src_xyzl + vi);
			store(begin + k, k, vi, c0, xform_point(xf, c0));
		}
		return;
	}
	for (int k = lane; k <= nbins; k += 32) hist[k] = 0;
	__syncwarp();
	/* sweep 1: bin histogram (shared-memory atomics: independent of each other, no warp-wide step per candidate),
	 * kBuildChunk gathers per lane in flight */
	for (int k0 = 0; k0 < ncand; k0 += 32 * kBuildChunk) {
		uint32_t v[kBuildChunk];
		float4 p[kBuildChunk];
#pragma unroll
		for (int j = 0; j < kBuildChunk; j++) { const int k = k0 + j * 32 + lane; v[j] = k < ncand ? table_value(k) : 0u; }      /* 0: a valid address */
#pragma unroll
		for (int j = 0; j < kBuildChunk; j++) { p[j] = __ldg(src_xyzl + v[j]); prefetch_l2(src_nrm + v[j]); }
#pragma unroll
		for (int j = 0; j < kBuildChunk; j++) {
			const int k = k0 + j * 32 + lane;
			if (k < ncand) atomicAdd(&hist[bin_of(xform_point(xf, p[j]))], 1u);
		}
	}
	__syncwarp();
	stamp(25);
	/* exclusive scan of the nbins + 1 table entries (9 consecutive entries per lane cover 288 >= 257), table out */
	{
		uint32_t h[9], sum = 0;
#pragma unroll
		for (int k = 0; k < 9; k++) { int e = lane * 9 + k; h[k] = e < nbins ? hist[e] : 0u; sum += h[k]; }
		uint32_t incl = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(full, incl, o);
			if (lane >= o) incl += t;
		}
		uint32_t run = incl - sum;
		__syncwarp();
		unsigned short *tab = set.tab + 2 * (size_t)begin;
#pragma unroll
		for (int k = 0; k < 9; k++) {
			int e = lane * 9 + k;
			if (e <= nbins) { hist[e] = run; tab[e] = (unsigned short)run; }
			run += h[k];
		}
		__syncwarp();
	}
	stamp(26);
	/* sweep 2: placement.  The order of the candidates INSIDE a bin is whatever the atomics hand out: the search takes
	 * the lexicographic minimum of (dist, l) with l carried in the record, so it does not depend on it */
	for (int k0 = 0; k0 < ncand; k0 += 32 * kBuildChunk) {
		uint32_t v[kBuildChunk];
		float4 p[kBuildChunk];
#pragma unroll
		for (int j = 0; j < kBuildChunk; j++) { const int k = k0 + j * 32 + lane; v[j] = k < ncand ? table_value(k) : 0u; }
#pragma unroll
		for (int j = 0; j < kBuildChunk; j++) p[j] = __ldg(src_xyzl + v[j]);
#pragma unroll
		for (int j = 0; j < kBuildChunk; j++) {
			const int k = k0 + j * 32 + lane;
			if (k < ncand) {
				const float4 c = xform_point(xf, p[j]);
				store(begin + (int)atomicAdd(&hist[bin_of(c)], 1u), k, v[j], p[j], c);
			}
		}
	}
	__syncwarp();
	stamp(27);
}

/* One warp per searchable bucket of the compact list k_finalize_grid / k_list_cells left behind. */
__global__ void __launch_bounds__(kBuildWarps * 32, 7) k_build_candidates(const uint32_t *__restrict__ vals,
		const m3dreg_grid_params *__restrict__ gp, const m3dreg_bucket *__restrict__ buckets,
		const uint32_t *__restrict__ cell_list, const unsigned int *__restrict__ cell_count,
		const float4 *__restrict__ src_xyzl, const float4 *__restrict__ src_nrm, const float4 *__restrict__ loc_src, const float *__restrict__ nrm_m,
		int xform_points, int max_inner, int max_outer, CandSet ci, CandSet co, int two_sets)
{
	pdl_enter();
	__shared__ uint32_t s_hist[kBuildWarps][kBuildTabMax + 7];
	NormalRotation rot;
	rot.on = nrm_m != nullptr;
#pragma unroll
	for (int k = 0; k < 9; k++) rot.r[k] = rot.on ? __ldg(nrm_m + (k / 3) * 4 + (k % 3)) : 0.0f;    /* row-major 4x4 */
	if (gp->number_of_buckets <= 0) return;
	PointXform xf;      /* xform_points: src_xyzl is the stored scan (local frame) and nrm_m the pose: points are transformed on the fly */
	xf.on = (xform_points && nrm_m) ? 1 : 0;
#pragma unroll
	for (int k = 0; k < 12; k++) xf.r[k] = xf.on ? __ldg(nrm_m + k) : 0.0f;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int nby = gp->number_of_buckets_Y, nbz = gp->number_of_buckets_Z;
	CellGeom g;
	g.mnx = gp->bounding_box_min_X; g.mny = gp->bounding_box_min_Y; g.mnz = gp->bounding_box_min_Z;
	g.rx = gp->resolution_X; g.ry = gp->resolution_Y; g.rz = gp->resolution_Z;
	const int tables = nn_tables_usable(max_inner, max_outer) ? 1 : 0;
	const unsigned int ncells = *cell_count;
	const unsigned int nwarps = gridDim.x * kBuildWarps;
	for (unsigned int t = blockIdx.x * kBuildWarps + w; t < ncells; t += nwarps) {
		int c = (int)__ldg(cell_list + t);
		const int *bp = reinterpret_cast<const int *>(buckets + c);
		int c_begin = __ldg(bp), c_n = __ldg(bp + 2);
		g.cx = c / (nby * nbz); g.cy = (c / nbz) % nby; g.cz = c % nbz;
		build_cell_candidates(vals, src_xyzl, src_nrm, loc_src, rot, xf, c_begin, c_n, max_inner, tables, g, ci, s_hist[w], lane);
		if (two_sets) build_cell_candidates(vals, src_xyzl, src_nrm, loc_src, rot, xf, c_begin, c_n, max_outer, tables, g, co, s_hist[w], lane);
	}
}

/* split an externally supplied reference-layout table into key / value streams (stage-level m3dreg_nn_search) */
__global__ void k_split_table(const m3dreg_hash_element *__restrict__ table, int n, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
	pdl_enter();
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
		m3dreg_hash_element h = table[p];
		keys[p] = (uint32_t)h.index_of_bucket;
		vals[p] = (uint32_t)h.index_of_point;
	}
}

/* ---- semantic nearest neighbour (kernel_semanticNearestNeighborSearch, lesson_16.cu:531-703) -------------
 * One query per thread; the search itself is nn_query() in nn_core.cuh.  Queries are expected in a spatially coherent
 * order (the scan store keeps a (label, Morton)-sorted copy of every scan, and a rigid transform preserves coherence),
 * so the 32 queries of a warp read the same few bins and their loads coalesce in L1.
 * q_perm (may be null = identity) maps the query's position to its index in the caller's order: nn_out is written
 * in the caller's order (the reference's layout); obs_rec (may be null), in query-array order, is what the moment
 * reduction streams next: {matched point of the first cloud as stored in src_xyzl (its LOCAL frame in the fused loops:
 * x0,y0,z0 of obs_nn_t, gpu6DSLAM.cpp:367-369), bits of its original index or -1}. */
constexpr int kNNThreads = 128;

__global__ void __launch_bounds__(kNNThreads) k_nn_search(const float4 *__restrict__ q_xyzl, const float4 *__restrict__ q_nrm,
		const uint32_t *__restrict__ q_perm, int n_second, CandSet ci, CandSet co,
		const uint32_t *__restrict__ s_vals, int n_first,
		const m3dreg_bucket *__restrict__ buckets, const m3dreg_grid_params *__restrict__ gp,
		float search_radius, int max_inner, int max_outer, int prune,
		int *__restrict__ nn_out, float4 *__restrict__ obs_rec, const float4 *__restrict__ src_xyzl, unsigned long long *__restrict__ label_counts,
		unsigned long long *__restrict__ eval_counter, const int *__restrict__ seg_of_chunk)
{
	pdl_enter();
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int qi = blockIdx.x * blockDim.x + threadIdx.x;
	if (seg_of_chunk && label_counts) label_counts += 4 * __ldg(seg_of_chunk + (blockIdx.x * blockDim.x) / kSegChunk);   /* a block never straddles segments */
	NNParams P;
	P.mnx = gp->bounding_box_min_X; P.mny = gp->bounding_box_min_Y; P.mnz = gp->bounding_box_min_Z;
	P.mxx = gp->bounding_box_max_X; P.mxy = gp->bounding_box_max_Y; P.mxz = gp->bounding_box_max_Z;
	P.rx = gp->resolution_X; P.ry = gp->resolution_Y; P.rz = gp->resolution_Z;
	P.nbx = gp->number_of_buckets_X; P.nby = gp->number_of_buckets_Y; P.nbz = gp->number_of_buckets_Z;
	P.nb = gp->number_of_buckets;
	P.buckets = buckets; P.ci = ci; P.co = co;
	P.cap_in = max_inner; P.cap_out = max_outer;
	nn_params_finish(P, search_radius, prune);
	unsigned int evals = 0;
	int best_l = kNNNone, label = -1;
	if (qi < n_second && P.nb > 0) {
		float4 p = __ldg(q_xyzl + qi), pn = __ldg(q_nrm + qi);
		label = __float_as_int(p.w);
		best_l = nn_query(P, p, pn, evals);
	}
	int result = -1;
	if (best_l != kNNNone && best_l >= 0 && best_l < n_first) result = (int)__ldg(s_vals + best_l);
	if (qi < n_second) {
		if (obs_rec) {
			float4 rec = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
			if (result >= 0) { rec = __ldg(src_xyzl + result); rec.w = __int_as_float(result); }
			obs_rec[qi] = rec;
		}
		if (nn_out) nn_out[q_perm ? __ldg(q_perm + qi) : (uint32_t)qi] = result;
	}
	if (eval_counter) {
		unsigned int tot = __reduce_add_sync(full, evals);
		if (lane == 0 && tot) atomicAdd(eval_counter, (unsigned long long)tot);
	}
	if (label_counts) {   /* per-label match counts (gpu6DSLAM.cpp:323-357): warp ballots, one atomic per label per warp */
		bool hit = qi < n_second && result >= 0;
#pragma unroll
		for (int L = 0; L < 4; L++) {
			unsigned m = __ballot_sync(full, hit && label == L);
			if (m && lane == L) atomicAdd(&label_counts[L], (unsigned long long)__popc(m));
		}
	}
}

/* ---- semantic NN, warp-shared search (k_nn_search_grid) ----------------------------------------------------------
 * Same answer as k_nn_search / nn_query() (lexicographic minimum of (dist, l) over the admissible candidates, rounds of
 * growing radius, conservative fine-column boxes — nn_core.cuh), organised so that the 32 queries of a warp (neighbours
 * on one surface: the scan store keeps queries (label, Morton)-sorted) stay in lock step:
 *   1. every unsettled lane computes its fine-column box for the round; the warp takes the HULL of the boxes (REDUX);
 *   2. the lanes look up, side by side, the fine cells of the hull (bucket record + bin table, once per cell and warp)
 *      and compact the non-empty bins of the warp's label into a segment list {first candidate, count, cell};
 *   3. the segments' candidates are staged in shared memory with coalesced loads (groups of four, padded with +inf)
 *      and ALL lanes evaluate ALL of them with broadcast LDS.128: a branch-free loop keeps, per lane, the minimum
 *      distance and the group that produced it, plus a flag for an exact tie between groups;
 *   4. one cold step per lane and batch applies the full predicate (label, angle gate, radius, (dist, l) order) to the
 *      winning group; a tie or an inadmissible winner triggers a re-scan of the batch with the full predicate.
 * Looking at more candidates than a lane's own box is harmless for a minimum as long as they belong to buckets of the
 * lane's own 27-neighbourhood; when the hull leaves some lane's neighbourhood (search radius > bucket size) this is
 * tested per group.  A bin of a coarser bucket (S = 2 or 1) covers several fine cells and is listed from one
 * representative cell per axis (the hull's first cell or a cell aligned to the bin); cells inside the previous round's
 * hull are skipped: a skipped representative means the bin touched that hull, where — by induction over the rounds —
 * every touching bin was evaluated by all lanes.  A bin may be listed twice (harmless).  Hulls larger than the list
 * are processed in chunks of whole (y,z) rows; warps whose queries are scattered (hull much larger than the lanes' own
 * boxes) and grids beyond the column arithmetic fall back to nn_query().
 * Requires one candidate set (INNER cap == OUTER cap, the reference's default). */
#ifndef M3D_NNG_THREADS
#define M3D_NNG_THREADS 64
#endif
#ifndef M3D_NNG_MINBLOCKS
#define M3D_NNG_MINBLOCKS 12
#endif
/* heuristics only (the answer is exact for any values): first-round radius = min(res) / rho_div; a warp whose hull has
 * more than hull_min cells AND more than hull_ratio times its largest own box searches per lane instead */
struct NNTuning { int rho_div, hull_min, hull_ratio; };
#ifndef M3D_NNG_CELLS
#define M3D_NNG_CELLS 128
#endif
#ifndef M3D_NNG_STAGE
#define M3D_NNG_STAGE 192
#endif
constexpr int kNNCells = M3D_NNG_CELLS;       /* hull cells looked up per chunk (segment list capacity) */
constexpr int kNNStage = M3D_NNG_STAGE;       /* candidates staged per batch (multiple of 4) */
constexpr int kNNGThreads = M3D_NNG_THREADS;
constexpr int kNNWarps = kNNGThreads / 32;

/* Packed f32x2 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2): two IEEE round-to-nearest operations per instruction, each
 * component rounded exactly like the scalar sub.rn / mul.rn / fma.rn the reference's distance is made of. */
__device__ __forceinline__ unsigned long long f2_pack(float a, float b)
{
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
	return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float &a, float &b)
{
	asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
/* {dist(q, c_a), dist(q, c_b)} with dist = fma(dz,dz, fma(dx,dx, dy*dy)) (lesson_16.cu:658-660) */
__device__ __forceinline__ unsigned long long nn_dist2(unsigned long long qx2, unsigned long long qy2, unsigned long long qz2,
		unsigned long long x2, unsigned long long y2, unsigned long long z2)
{
	unsigned long long dx, dy, dz, t;
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(qx2), "l"(x2));
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(qy2), "l"(y2));
	asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(qz2), "l"(z2));
	asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(t) : "l"(dy));
	asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(t) : "l"(dx), "l"(t));
	asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(t) : "l"(dz), "l"(t));
	return t;
}

__device__ __noinline__ int nn_query_fallback(const m3dreg_grid_params *__restrict__ gp, const m3dreg_bucket *__restrict__ buckets,
		const float4 *cxyzl, const float4 *cnrm, const unsigned short *ctab, float search_radius, int cap, int prune, float4 p, float4 pn,
		unsigned int *evals)
{
	NNParams P;
	P.mnx = gp->bounding_box_min_X; P.mny = gp->bounding_box_min_Y; P.mnz = gp->bounding_box_min_Z;
	P.mxx = gp->bounding_box_max_X; P.mxy = gp->bounding_box_max_Y; P.mxz = gp->bounding_box_max_Z;
	P.rx = gp->resolution_X; P.ry = gp->resolution_Y; P.rz = gp->resolution_Z;
	P.nbx = gp->number_of_buckets_X; P.nby = gp->number_of_buckets_Y; P.nbz = gp->number_of_buckets_Z;
	P.nb = gp->number_of_buckets;
	P.buckets = buckets;
	P.ci.xyzl = const_cast<float4 *>(cxyzl); P.ci.nrm = const_cast<float4 *>(cnrm); P.ci.tab = const_cast<unsigned short *>(ctab);
	P.co = P.ci;
	P.cap_in = cap; P.cap_out = cap;
	nn_params_finish(P, search_radius, prune);
	unsigned int ev = 0;
	const int l = nn_query(P, p, pn, ev);
	*evals += ev;
	return l;
}

/* full predicate of the reference on one candidate (lesson_16.cu:658-686) with the (dist, l) order made explicit */
#define M3D_NN_CONSIDER(D, C, J)                                                                                       \
	if ((D) <= lim) {                                                                                                  \
		const int l_ = __float_as_int((C).w);                                                                           \
		if ((D) < best_d || l_ < best_l) {                                                                             \
			const float4 n_ = __ldg(cn + (J));                                                                          \
			if (__float_as_int(n_.w) == label) {                                                                        \
				const float dot_ = f_fma(pn.z, n_.z, f_fma(pn.x, n_.x, f_mul(pn.y, n_.y)));                             \
				if (angle_gate(dot_)) { best_d = (D); best_l = l_; best_j = (J); lim = (D); }                          \
			}                                                                                                           \
		}                                                                                                               \
	}

__global__ void __launch_bounds__(kNNGThreads, M3D_NNG_MINBLOCKS) k_nn_search_grid(const float4 *__restrict__ q_xyzl, const float4 *__restrict__ q_nrm,
		const uint32_t *__restrict__ q_perm, int n_second, CandSet cs,
		const uint32_t *__restrict__ s_vals, int n_first,
		const m3dreg_bucket *__restrict__ buckets, const m3dreg_grid_params *__restrict__ gp,
		float search_radius, int cap, int prune, NNTuning tune,
		int *__restrict__ nn_out, float4 *__restrict__ obs_rec, const float4 *__restrict__ src_xyzl, unsigned long long *__restrict__ label_counts,
		unsigned long long *__restrict__ eval_counter, const int *__restrict__ seg_of_chunk)
{
	pdl_enter();
	if (seg_of_chunk && label_counts) label_counts += 4 * __ldg(seg_of_chunk + (blockIdx.x * blockDim.x) / kSegChunk);   /* a block never straddles segments */
	__shared__ int4 s_segs[kNNWarps][kNNCells];                /* {first candidate, count, hull cell x | y << 16, hull cell z} */
	__shared__ float4 s_cand[kNNWarps][kNNStage];              /* staged candidates, per group of four: {x0..x3}, {y0..y3}, {z0..z3}, {l0..l3} */
	__shared__ int4 s_grp[kNNWarps][kNNStage / 4];             /* per group: {index of its first candidate, -, hull cell x | y << 16, z} */
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const unsigned lt_mask = (1u << lane) - 1u;
	int4 *segs = s_segs[threadIdx.x >> 5];
	float4 *stage = s_cand[threadIdx.x >> 5];
	float *stagef = reinterpret_cast<float *>(stage);
	int4 *grp = s_grp[threadIdx.x >> 5];
	const int qi = blockIdx.x * blockDim.x + threadIdx.x;
	/* grid and search parameters (same expressions as nn_params_finish) */
	const float mnx = gp->bounding_box_min_X, mny = gp->bounding_box_min_Y, mnz = gp->bounding_box_min_Z;
	const float res_x = gp->resolution_X, res_y = gp->resolution_Y, res_z = gp->resolution_Z;
	const int nbx = gp->number_of_buckets_X, nby = gp->number_of_buckets_Y, nbz = gp->number_of_buckets_Z;
	const long long nb = gp->number_of_buckets;
	const int tables = (nn_tables_usable(cap, cap) && nn_columns_usable(nbx, nby, nbz)) ? 1 : 0;
	const float r2 = f_mul(search_radius, search_radius);                /* lesson_16.cu:553 */
	const float iwx = f_div(4.0f, res_x), iwy = f_div(4.0f, res_y), iwz = f_div(4.0f, res_z);
	float rho2_first;
	{
		const float rmin = f_div(fminf(res_x, fminf(res_y, res_z)), (float)(tune.rho_div > 0 ? tune.rho_div : 16));
		rho2_first = fmaxf(f_mul(rmin, rmin), 1.0e-30f);
	}
	const float4 *__restrict__ cx = cs.xyzl;
	const float4 *__restrict__ cn = cs.nrm;

	unsigned int evals = 0;
	int best_l = kNNNone, best_j = -1, label = -1;                     /* best_j: the winner's slot in the candidate set (-1: found by nn_query()) */
	float best_d = 100000000.0f;                                        /* lesson_16.cu:597 */
	float lim = fminf(r2, 99999992.0f);
	float qx = 0.0f, qy = 0.0f, qz = 0.0f;
	float4 pn = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	int ix = 0, iy = 0, iz = 0;
	bool active = false;
	if (qi < n_second && nb > 0 && cap > 0) {
		const float4 p = __ldg(q_xyzl + qi);
		pn = __ldg(q_nrm + qi);
		qx = p.x; qy = p.y; qz = p.z;
		label = __float_as_int(p.w);
		/* lesson_16.cu:562-583 */
		if (!(qx < mnx || qx > gp->bounding_box_max_X || qy < mny || qy > gp->bounding_box_max_Y || qz < mnz || qz > gp->bounding_box_max_Z)) {
			ix = cell_of(qx, mnx, res_x); iy = cell_of(qy, mny, res_y); iz = cell_of(qz, mnz, res_z);
			const int home = ix * nby * nbz + iy * nbz + iz;
			active = home >= 0 && (long long)home < nb && lim >= 0.0f;
		}
	}
	if (!nn_columns_usable(nbx, nby, nbz)) {                            /* warp-uniform */
		if (active) best_l = nn_query_fallback(gp, buckets, cs.xyzl, cs.nrm, cs.tab, search_radius, cap, prune ? 1 : 0,
				make_float4(qx, qy, qz, __int_as_float(label)), pn, &evals);
		active = false;
	}
	/* the 27-neighbourhood with edge clamping (lesson_16.cu:588-608), as a box of fine columns */
	const int cx0 = (ix > 0 ? ix - 1 : ix) << 2, cx1 = ((ix != nbx - 1 ? ix + 1 : ix) << 2) + 3;
	const int cy0 = (iy > 0 ? iy - 1 : iy) << 2, cy1 = ((iy != nby - 1 ? iy + 1 : iy) << 2) + 3;
	const int cz0 = (iz > 0 ? iz - 1 : iz) << 2, cz1 = ((iz != nbz - 1 ? iz + 1 : iz) << 2) + 3;
	const float mgx = f_fma(f_mul(fabsf(qx) + fabsf(mnx), iwx), 3.814697265625e-06f, 9.765625e-04f);
	const float mgy = f_fma(f_mul(fabsf(qy) + fabsf(mny), iwy), 3.814697265625e-06f, 9.765625e-04f);
	const float mgz = f_fma(f_mul(fabsf(qz) + fabsf(mnz), iwz), 3.814697265625e-06f, 9.765625e-04f);

	/* every candidate with dist <= tau lies in the box of fine columns this returns (nn_query()'s box: conservative
	 * column bounds of [q - R, q + R], R >= sqrt(tau) rounded outwards, clamped to the lane's 27-neighbourhood) */
	auto fine_box = [&](float tau, int &xl, int &xh, int &yl, int &yh, int &zl, int &zh) {
		/* R >= sqrt(tau) * (1 + 2^-20) is all the proof needs: tau * rsqrt(tau) is within 2^-21 of sqrt(tau) (MUFU.RSQ: 2 ulp),
		 * the factor 1 + 2^-13 covers that with three orders of magnitude to spare and costs a box 0.01 % wider */
		const float R = !prune ? INFINITY : (tau > 1.0e-30f ? f_fma(f_mul(tau, rsqrtf(tau)), 1.0001220703125f, 1.0e-18f) : 1.1e-15f);
		xl = col_floor(f_sub(qx, R), mnx, iwx, mgx); xh = col_ceil(f_add(qx, R), mnx, iwx, mgx);
		yl = col_floor(f_sub(qy, R), mny, iwy, mgy); yh = col_ceil(f_add(qy, R), mny, iwy, mgy);
		zl = col_floor(f_sub(qz, R), mnz, iwz, mgz); zh = col_ceil(f_add(qz, R), mnz, iwz, mgz);
		xl = xl > cx0 ? xl : cx0; xh = xh < cx1 ? xh : cx1;
		yl = yl > cy0 ? yl : cy0; yh = yh < cy1 ? yh : cy1;
		zl = zl > cz0 ? zl : cz0; zh = zh < cz1 ? zh : cz1;
	};

	unsigned todo = __ballot_sync(full, active);
	while (todo) {                                                      /* one pass per label present in the warp (almost always one) */
		const int L = __shfl_sync(full, label, __ffs(todo) - 1);
		const bool mine = active && label == L;
		bool unsettled = mine;
		todo &= ~__ballot_sync(full, mine);
		int hxl = 1, hxh = 0, hyl = 1, hyh = 0, hzl = 1, hzh = 0;      /* previous round's hull (empty) */
		float rho2 = rho2_first;
		for (int round = 0; round < 80; round++) {
			if (!__any_sync(full, unsettled)) break;
			int xl = 0x7fffffff, xh = -0x7fffffff, yl = 0x7fffffff, yh = -0x7fffffff, zl = 0x7fffffff, zh = -0x7fffffff;
			if (unsettled) {
				fine_box(prune ? fminf(lim, rho2) : lim, xl, xh, yl, yh, zl, zh);
				if (xl > xh || yl > yh || zl > zh) { xl = yl = zl = 0x7fffffff; xh = yh = zh = -0x7fffffff; }
			}
			const int uxl = __reduce_min_sync(full, xl), uxh = __reduce_max_sync(full, xh);
			const int uyl = __reduce_min_sync(full, yl), uyh = __reduce_max_sync(full, yh);
			const int uzl = __reduce_min_sync(full, zl), uzh = __reduce_max_sync(full, zh);
			if (uxl <= uxh) {
				const int dx = uxh - uxl + 1, dy = uyh - uyl + 1, dz = uzh - uzl + 1;
				const long long ncell = (long long)dx * dy * dz;
				const int own = xl <= xh ? (xh - xl + 1) * (yh - yl + 1) * (zh - zl + 1) : 0;
				const int own_max = __reduce_max_sync(full, own);
				if (dx > kNNCells || dy > 32767 || dz > 32767 || (ncell > tune.hull_min && ncell > (long long)tune.hull_ratio * own_max)) {   /* scattered warp: per-lane search from scratch */
					if (eval_counter) {
						const unsigned fb = __ballot_sync(full, unsettled);
						if (lane == 0) atomicAdd(eval_counter + 1, (unsigned long long)__popc(fb));
					}
					if (unsettled) {
						best_l = nn_query_fallback(gp, buckets, cs.xyzl, cs.nrm, cs.tab, search_radius, cap, prune ? 1 : 0,
								make_float4(qx, qy, qz, __int_as_float(label)), pn, &evals);
						best_j = -1;
					}
					unsettled = false;
					break;
				}
				/* may every lane look at every bucket the hull touches? (always, unless the radius exceeds the bucket size) */
				const bool nb_all = __all_sync(full, !mine || (uxl >= cx0 && uxh <= cx1 && uyl >= cy0 && uyh <= cy1 && uzl >= cz0 && uzh <= cz1));
				const int nrows = dy * dz, rpc = __float2int_rz(__fdividef((float)kNNCells + 0.5f, (float)dx));   /* = kNNCells / dx for 1 <= dx <= kNNCells */
				const float inv_dx = __frcp_rn((float)dx), inv_dy = __frcp_rn((float)dy);
				for (int row0 = 0; row0 < nrows; row0 += rpc) {
					const int nr = nrows - row0 < rpc ? nrows - row0 : rpc;
					const int ncc = nr * dx;
					/* 2. look up the hull cells of rows [row0, row0 + nr), row = (z - uzl) * dy + (y - uyl); list the non-empty bins */
					int nseg = 0;
					for (int c0 = 0; c0 < ncc; c0 += 32) {
						const int c = c0 + lane;
						int start = 0, cnt = 0, rel_xy = 0, rel_z = 0;
						if (c < ncc) {
							const int rr = __float2int_rz(((float)c + 0.5f) * inv_dx);
							const int ax = c - rr * dx;
							const int r = row0 + rr;
							const int az = __float2int_rz(((float)r + 0.5f) * inv_dy);
							const int ay = r - az * dy;
							const int gx = uxl + ax, gy = uyl + ay, gz = uzl + az;
							const bool in_old = gx >= hxl && gx <= hxh && gy >= hyl && gy <= hyh && gz >= hzl && gz <= hzh;
							if (!in_old) {
								const int cell = ((gx >> 2) * nby + (gy >> 2)) * nbz + (gz >> 2);
								const int *rec = reinterpret_cast<const int *>(buckets + cell);
								const int npts = __ldg(rec + 2), begin = __ldg(rec);
								if (npts > 0 && begin >= 0) {           /* lesson_16.cu:615-616 (also the quirk bucket) */
									const int level = tables ? nn_level(npts) : -1;
									const int sh = level < 0 ? 2 : 2 - level;
									const int am = (1 << sh) - 1;
									const bool rep = (ax == 0 || !(gx & am)) && (ay == 0 || !(gy & am)) && (az == 0 || !(gz & am));
									if (rep) {
										if (level < 0) {                /* no table: the whole bucket, in walk order */
											const int iter = candidate_stride(npts, cap);
											start = begin; cnt = (npts + iter - 1) / iter;
										} else {
											const int bin = nn_bin(L, (gx & 3) >> sh, (gy & 3) >> sh, (gz & 3) >> sh, level);
											const unsigned short *tab = cs.tab + 2 * (size_t)begin + bin;
											const int s = __ldg(tab), e = __ldg(tab + 1);
											start = begin + s; cnt = e - s;
										}
										rel_xy = ax | (ay << 16); rel_z = az;
									}
								}
							}
						}
						const unsigned m = __ballot_sync(full, cnt > 0);
						if (cnt > 0) segs[nseg + __popc(m & lt_mask)] = make_int4(start, cnt, rel_xy, rel_z);
						nseg += __popc(m);
					}
					__syncwarp();
					/* 3./4. batches of at most kNNStage staged candidates */
					int k0 = 0;
					while (k0 < nseg) {
						/* lane k owns segment k0 + k: padded sizes, inclusive scan, how many segments fit */
						int4 sg = make_int4(0, 0, 0, 0);
						if (k0 + lane < nseg) sg = segs[k0 + lane];
						const int padded = (sg.y + 3) & ~3;
						int incl = padded;
#pragma unroll
						for (int o = 1; o < 32; o <<= 1) {
							const int t = __shfl_up_sync(full, incl, o);
							if (lane >= o) incl += t;
						}
						const unsigned fit = __ballot_sync(full, k0 + lane < nseg && incl <= kNNStage);
						int ntake = __popc(fit);                        /* segments are taken in order: fit is a prefix mask */
						int ncand;
						if (ntake == 0) {                               /* the first segment alone exceeds a batch: take a part of it */
							const int4 s0 = segs[k0];
							__syncwarp();
							if (lane == 0) segs[k0] = make_int4(s0.x + kNNStage, s0.y - kNNStage, s0.z, s0.w);
							if (lane < kNNStage / 4) grp[lane] = make_int4(s0.x + 4 * lane, 4, s0.z, s0.w);
							if (lane + 32 < kNNStage / 4) grp[lane + 32] = make_int4(s0.x + 4 * (lane + 32), 4, s0.z, s0.w);
							ncand = kNNStage;
						} else {
							ncand = __shfl_sync(full, incl, ntake - 1);
							if (lane < ntake) {                         /* group records of the lane's own segment */
								const int g0 = (incl - padded) >> 2, ng = padded >> 2;
								for (int g = 0; g < ng; g++) {
									const int left = sg.y - 4 * g;
									grp[g0 + g] = make_int4(sg.x + 4 * g, left < 4 ? left : 4, sg.z, sg.w);
								}
							}
							k0 += ntake;
						}
						__syncwarp();
						/* flat, coalesced copy: slot t belongs to group t / 4 (all loads independent, in flight together);
						 * a group is stored transposed so that the hot loop reads coordinate pairs as 64-bit registers */
#pragma unroll 2
						for (int t = lane; t < ncand; t += 32) {
							const int4 gi = grp[t >> 2];
							const float4 cv = (t & 3) < gi.y ? __ldg(cx + gi.x + (t & 3)) : make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(kNNNone));
							float *dst = stagef + ((t >> 2) << 4) + (t & 3);
							dst[0] = cv.x; dst[4] = cv.y; dst[8] = cv.z; dst[12] = cv.w;
						}
						__syncwarp();
						const int ngrp = ncand >> 2;
						/* branch-free minimum over every staged candidate */
						float rb = mine ? lim : -INFINITY;
						int bg = -1;
						bool flag = false;
						const unsigned long long qx2 = f2_pack(qx, qx), qy2 = f2_pack(qy, qy), qz2 = f2_pack(qz, qz);
						if (nb_all) {
							if (mine) evals += (unsigned int)ncand;
#pragma unroll 2
							for (int g = 0; g < ngrp; g++) {
								const float4 X = stage[4 * g], Y = stage[4 * g + 1], Z = stage[4 * g + 2];
								float d0, d1, d2, d3;
								f2_unpack(nn_dist2(qx2, qy2, qz2, f2_pack(X.x, X.y), f2_pack(Y.x, Y.y), f2_pack(Z.x, Z.y)), d0, d1);
								f2_unpack(nn_dist2(qx2, qy2, qz2, f2_pack(X.z, X.w), f2_pack(Y.z, Y.w), f2_pack(Z.z, Z.w)), d2, d3);
								const float m4 = fminf(fminf(d0, d1), fminf(d2, d3));
								const bool lt = m4 < rb;
								flag = flag || (m4 == rb);
								rb = lt ? m4 : rb; bg = lt ? g : bg;
							}
						} else {
							for (int g = 0; g < ngrp; g++) {
								const int4 gi = grp[g];
								const int gx = uxl + (gi.z & 0xffff), gy = uyl + (gi.z >> 16), gz = uzl + gi.w;
								const bool use = gx >= cx0 && gx <= cx1 && gy >= cy0 && gy <= cy1 && gz >= cz0 && gz <= cz1;
								if (mine && use) evals += 4u;
								const float4 X = stage[4 * g], Y = stage[4 * g + 1], Z = stage[4 * g + 2];
								float d0, d1, d2, d3;
								f2_unpack(nn_dist2(qx2, qy2, qz2, f2_pack(X.x, X.y), f2_pack(Y.x, Y.y), f2_pack(Z.x, Z.y)), d0, d1);
								f2_unpack(nn_dist2(qx2, qy2, qz2, f2_pack(X.z, X.w), f2_pack(Y.z, Y.w), f2_pack(Z.z, Z.w)), d2, d3);
								float m4 = fminf(fminf(d0, d1), fminf(d2, d3));
								m4 = use ? m4 : INFINITY;
								const bool lt = m4 < rb;
								flag = flag || (m4 == rb);
								rb = lt ? m4 : rb; bg = lt ? g : bg;
							}
						}
						/* full predicate on the winning group; a tie between groups or an inadmissible winner needs the re-scan */
						if (bg >= 0) {
							const int j = grp[bg].x;
							const float4 X = stage[4 * bg], Y = stage[4 * bg + 1], Z = stage[4 * bg + 2], Lw = stage[4 * bg + 3];
							const float4 c0 = make_float4(X.x, Y.x, Z.x, Lw.x), c1 = make_float4(X.y, Y.y, Z.y, Lw.y),
									c2 = make_float4(X.z, Y.z, Z.z, Lw.z), c3 = make_float4(X.w, Y.w, Z.w, Lw.w);
							const float d0 = nn_dist(qx, qy, qz, c0), d1 = nn_dist(qx, qy, qz, c1), d2 = nn_dist(qx, qy, qz, c2),
									d3 = nn_dist(qx, qy, qz, c3);
							M3D_NN_CONSIDER(d0, c0, j)
							M3D_NN_CONSIDER(d1, c1, j + 1)
							M3D_NN_CONSIDER(d2, c2, j + 2)
							M3D_NN_CONSIDER(d3, c3, j + 3)
							if (best_d != rb) flag = true;
						}
						if (__any_sync(full, flag)) {
							for (int g = 0; g < ngrp; g++) {
								const int4 gi = grp[g];
								const int gx = uxl + (gi.z & 0xffff), gy = uyl + (gi.z >> 16), gz = uzl + gi.w;
								const bool use = flag && mine && gx >= cx0 && gx <= cx1 && gy >= cy0 && gy <= cy1 && gz >= cz0 && gz <= cz1;
								if (use) {
#pragma unroll
									for (int t = 0; t < 4; t++) {
										const float *src = stagef + (g << 4) + t;
										const float4 c0 = make_float4(src[0], src[4], src[8], src[12]);
										const float d0 = nn_dist(qx, qy, qz, c0);
										M3D_NN_CONSIDER(d0, c0, gi.x + t)
									}
								}
							}
						}
						__syncwarp();
					}
				}
				hxl = uxl; hxh = uxh; hyl = uyl; hyh = uyh; hzl = uzl; hzh = uzh;
				/* every cell of this hull has now been evaluated by all lanes (in this round or, for the cells skipped as
				 * part of the previous hull, before): a lane whose box for its CURRENT limit lies inside the hull is done */
				if (unsettled && prune && lim > rho2) {
					int bxl, bxh, byl, byh, bzl, bzh;
					fine_box(lim, bxl, bxh, byl, byh, bzl, bzh);
					if (bxl >= uxl && bxh <= uxh && byl >= uyl && byh <= uyh && bzl >= uzl && bzh <= uzh) unsettled = false;
				}
			}
			if (unsettled && (!prune || lim <= rho2)) unsettled = false;   /* everything at or below the limit was inside this round's box */
			rho2 = f_mul(rho2, 4.0f);
		}
	}
#undef M3D_NN_CONSIDER

	/* the winner: its record in the candidate set carries the point as stored in the scan (local frame) and its original
	 * index — one gather instead of hash[l] -> cloud[index] (lesson_16.cu:640-647, gpu6DSLAM.cpp:367-369) */
	int result = -1;
	float4 rec = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
	if (best_l != kNNNone && best_l >= 0 && best_l < n_first) {
		if (best_j >= 0) {
			rec = __ldg(cs.loc + best_j);
			result = __float_as_int(rec.w);
		} else {
			result = (int)__ldg(s_vals + best_l);
			if (obs_rec) { rec = __ldg(src_xyzl + result); rec.w = __int_as_float(result); }
		}
	}
	if (qi < n_second) {
		if (obs_rec) obs_rec[qi] = rec;
		if (nn_out) nn_out[q_perm ? __ldg(q_perm + qi) : (uint32_t)qi] = result;
	}
	if (eval_counter) {
		unsigned int tot = __reduce_add_sync(full, evals);
		if (lane == 0 && tot) atomicAdd(eval_counter, (unsigned long long)tot);
	}
	if (label_counts) {   /* per-label match counts (gpu6DSLAM.cpp:323-357): warp ballots, one atomic per label per warp */
		bool hit = qi < n_second && result >= 0;
#pragma unroll
		for (int Lb = 0; Lb < 4; Lb++) {
			unsigned m = __ballot_sync(full, hit && label == Lb);
			if (m && lane == Lb) atomicAdd(&label_counts[Lb], (unsigned long long)__popc(m));
		}
	}
}

/* correspondences from the search's per-query records (query order) into the caller's order:
 * out[perm[i]] = index carried by rec[i] (perm == 0: identity) */
__global__ void k_scatter_nn(const uint32_t *__restrict__ perm, int n, const float4 *__restrict__ rec, int *__restrict__ out)
{
	pdl_enter();
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		out[perm ? __ldg(perm + i) : (uint32_t)i] = __float_as_int(__ldg(reinterpret_cast<const float *>(rec + i) + 3));
}

/* gather a stored scan into query order: out[i] = in[perm[i]] */
__global__ void k_gather_perm(const uint32_t *__restrict__ perm, int n, const float4 *__restrict__ in_xyzl, const float4 *__restrict__ in_nrm,
		float4 *__restrict__ out_xyzl, float4 *__restrict__ out_nrm)
{
	pdl_enter();
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t v = __ldg(perm + i);
		out_xyzl[i] = __ldg(in_xyzl + v);
		out_nrm[i] = __ldg(in_nrm + v);
	}
}

/* ---- normal equations: fused fp64 reduction, never materialising A / P / AtP ------------------------------
 * (replaces kernel_fill_A_l_cuda, kernel_cudaCompute_AtP and both DGEMMs; lesson_16.cu:245-439, AXB:407-428)
 *
 * With A_k = -[I | J_k] and J_k linear in the local point p0, all of AtPA / AtPl follow from
 * pose-independent weighted raw moments:
 *   S = sum w,  M1 = sum w p0,  M2 = sum w p0 p0^T (6),  L1 = sum w l,  L2 = sum w p0 l^T (9)   -> 22 sums.
 * Every thread accumulates them in registers over a grid-stride loop, warps reduce with shuffles, the block
 * writes one partial row, and the LAST block to finish (atomic ticket) adds the rows in fixed order, forms the
 * 6x6 system for the current Euler angles, applies the observation gate, runs the Cholesky solve and updates the
 * pose — so the iteration never returns to the host. */

struct Moments {
	double v[kMomentCount];
	__device__ __forceinline__ void clear()
	{
#pragma unroll
		for (int i = 0; i < kMomentCount; i++) v[i] = 0.0;
	}
	__device__ __forceinline__ void add(double w, double x, double y, double z, double lx, double ly, double lz)
	{
		double wx = w * x, wy = w * y, wz = w * z;
		v[0] += w;
		v[1] += wx; v[2] += wy; v[3] += wz;
		v[4] = fma(wx, x, v[4]); v[5] = fma(wx, y, v[5]); v[6] = fma(wx, z, v[6]);
		v[7] = fma(wy, y, v[7]); v[8] = fma(wy, z, v[8]); v[9] = fma(wz, z, v[9]);
		v[10] = fma(w, lx, v[10]); v[11] = fma(w, ly, v[11]); v[12] = fma(w, lz, v[12]);
		v[13] = fma(wx, lx, v[13]); v[14] = fma(wx, ly, v[14]); v[15] = fma(wx, lz, v[15]);
		v[16] = fma(wy, lx, v[16]); v[17] = fma(wy, ly, v[17]); v[18] = fma(wy, lz, v[18]);
		v[19] = fma(wz, lx, v[19]); v[20] = fma(wz, ly, v[20]); v[21] = fma(wz, lz, v[21]);
		v[22] += 1.0;
	}
};

/* 6x6 system (28-double packing) from moments and Euler angles.  R = Rx(om) Ry(fi) Rz(ka) (lesson_16.cu:278-293);
 * C[r][c] is the coefficient 3-vector of J[r][c] = dR p0 / d(om,fi,ka) (lesson_16.cu:310-353). */
__device__ __host__ inline void moments_to_neq(const double *mo, double om, double fi, double ka, double *neq)
{
	double so = sin(om), co = cos(om), sf = sin(fi), cf = cos(fi), sk = sin(ka), ck = cos(ka);
	double R11 = cf * ck, R12 = -cf * sk;
	double R21 = co * sk + so * sf * ck, R22 = co * ck - so * sf * sk, R23 = -so * cf;
	double R31 = so * sk - co * sf * ck, R32 = so * ck + co * sf * sk, R33 = co * cf;
	double C[3][3][3] = {
		{{0, 0, 0}, {-sf * ck, sf * sk, cf}, {R12, -R11, 0}},
		{{-R31, -R32, -R33}, {so * cf * ck, -so * cf * sk, so * sf}, {R22, -R21, 0}},
		{{R21, R22, R23}, {-co * cf * ck, co * cf * sk, -co * sf}, {R32, -R31, 0}}};
	double S = mo[0];
	const double *M1 = mo + 1;
	double M2[3][3] = {{mo[4], mo[5], mo[6]}, {mo[5], mo[7], mo[8]}, {mo[6], mo[8], mo[9]}};
	const double *L1 = mo + 10;
	const double *L2 = mo + 13;   /* L2[i*3+r] = sum w p0_i l_r */
	double N[6][6], b[6];
	for (int i = 0; i < 6; i++) { b[i] = 0; for (int j = 0; j < 6; j++) N[i][j] = 0; }
	for (int r = 0; r < 3; r++) {
		N[r][r] = S;
		b[r] = -L1[r];
		for (int c = 0; c < 3; c++) {
			double s = 0;
			for (int i = 0; i < 3; i++) s += C[r][c][i] * M1[i];
			N[r][3 + c] = s;
		}
	}
	for (int c = 0; c < 3; c++) {
		for (int c2 = c; c2 < 3; c2++) {
			double s = 0;
			for (int r = 0; r < 3; r++)
				for (int i = 0; i < 3; i++) {
					double t = 0;
					for (int j = 0; j < 3; j++) t += M2[i][j] * C[r][c2][j];
					s += C[r][c][i] * t;
				}
			N[3 + c][3 + c2] = s;
		}
		double s = 0;
		for (int r = 0; r < 3; r++)
			for (int i = 0; i < 3; i++) s += C[r][c][i] * L2[i * 3 + r];
		b[3 + c] = -s;
	}
	int k = 0;
	for (int i = 0; i < 6; i++)
		for (int j = i; j < 6; j++) neq[k++] = N[i][j];
	for (int i = 0; i < 6; i++) neq[k++] = b[i];
	neq[k] = mo[22];
}

/* Lower Cholesky + two triangular solves (linearSolverCHOL, AXB:484-539) of the dof-subsystem of a packed system.
 * dof 6: all unknowns; dof 4: {tx,ty,tz,ka} (fill_A_l_4DOFcuda keeps columns 0,1,2,5, lesson_16.cu:493-495).
 * Fully unrolled so the whole factorisation lives in registers; one sqrt and one reciprocal per column.
 * Returns 0 or M3DREG_E_NOT_SPD. */
template <int DOF>
__device__ __forceinline__ int solve_packed_n(const double *neq, double *x)
{
	/* packed index of (i, j), i <= j, in the 6x6 upper triangle */
	auto pk = [](int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); };
	auto sel = [](int i) { return DOF == 6 ? i : (i < 3 ? i : 5); };
	double L[DOF][DOF], inv[DOF], y[DOF];
#pragma unroll
	for (int j = 0; j < DOF; j++) {
		double d = neq[pk(sel(j), sel(j))];
#pragma unroll
		for (int c = 0; c < j; c++) d -= L[j][c] * L[j][c];
		if (!(d > 0.0)) return M3DREG_E_NOT_SPD;
		d = sqrt(d);
		L[j][j] = d;
		inv[j] = 1.0 / d;
#pragma unroll
		for (int i = j + 1; i < DOF; i++) {
			double s = neq[pk(sel(j), sel(i))];
#pragma unroll
			for (int c = 0; c < j; c++) s -= L[i][c] * L[j][c];
			L[i][j] = s * inv[j];
		}
	}
#pragma unroll
	for (int i = 0; i < DOF; i++) {
		double s = neq[21 + sel(i)];
#pragma unroll
		for (int c = 0; c < i; c++) s -= L[i][c] * y[c];
		y[i] = s * inv[i];
	}
#pragma unroll
	for (int i = DOF - 1; i >= 0; i--) {
		double s = y[i];
#pragma unroll
		for (int c = i + 1; c < DOF; c++) s -= L[c][i] * x[c];
		x[i] = s * inv[i];
	}
	return 0;
}

__device__ inline int solve_packed(const double *neq, int dof, double *x)
{
	if (dof == 6) return solve_packed_n<6>(neq, x);
	double x4[4] = {0, 0, 0, 0};
	int st = solve_packed_n<4>(neq, x4);
	x[0] = x4[0]; x[1] = x4[1]; x[2] = x4[2]; x[3] = x4[3];
	return st;
}

/* Matrix4ToEuler / EulerToMatrix (cudaWrapper.cpp:470-514), row-major 4x4.  Double transcendental functions
 * rounded to float, identical text in oracle/m3d_oracle.c so both sides round the same way; upstream's Eigen float
 * quaternion path cannot be reproduced to the last ulp (Eigen absent) — tolerance parity, see DESIGN.md. */
__device__ __host__ inline void matrix4_to_euler(const float *m, float *omfika, float *xyz)
{
	const double kPi = 3.14159265358979323846;
	double trX, trY;
	if (m[0] > 0.0) omfika[1] = (float)asin((double)m[2]);
	else omfika[1] = (float)(kPi - asin((double)m[2]));
	double C = cos((double)omfika[1]);
	if (fabs(C) > 0.005) {
		trX = m[10] / C; trY = -m[6] / C;
		omfika[0] = (float)atan2(trY, trX);
		trX = m[0] / C; trY = -m[1] / C;
		omfika[2] = (float)atan2(trY, trX);
	} else {
		omfika[0] = 0.0f;
		trX = m[5]; trY = m[4];
		omfika[2] = (float)atan2(trY, trX);
	}
	xyz[0] = m[3]; xyz[1] = m[7]; xyz[2] = m[11];
}

__device__ __host__ inline void euler_to_matrix(const float *omfika, const float *xyz, float *m)
{
	float hx = 0.5f * omfika[0], hy = 0.5f * omfika[1], hz = 0.5f * omfika[2];
	float ax = (float)sin((double)hx), aw = (float)cos((double)hx);
	float by = (float)sin((double)hy), bw = (float)cos((double)hy);
	float cz = (float)sin((double)hz), cw = (float)cos((double)hz);
#ifdef __CUDA_ARCH__
#define M3D_MUL(a, b) __fmul_rn(a, b)
#define M3D_ADD(a, b) __fadd_rn(a, b)
#define M3D_SUB(a, b) __fsub_rn(a, b)
#else
#define M3D_MUL(a, b) ((a) * (b))
#define M3D_ADD(a, b) ((a) + (b))
#define M3D_SUB(a, b) ((a) - (b))
#endif
	float w1 = M3D_MUL(aw, bw), x1 = M3D_MUL(ax, bw), y1 = M3D_MUL(aw, by), z1 = M3D_MUL(ax, by);
	float w = M3D_SUB(M3D_MUL(w1, cw), M3D_MUL(z1, cz));
	float x = M3D_ADD(M3D_MUL(x1, cw), M3D_MUL(y1, cz));
	float y = M3D_SUB(M3D_MUL(y1, cw), M3D_MUL(x1, cz));
	float z = M3D_ADD(M3D_MUL(w1, cz), M3D_MUL(z1, cw));
	float tx = M3D_MUL(2.0f, x), ty = M3D_MUL(2.0f, y), tz = M3D_MUL(2.0f, z);
	float twx = M3D_MUL(tx, w), twy = M3D_MUL(ty, w), twz = M3D_MUL(tz, w);
	float txx = M3D_MUL(tx, x), txy = M3D_MUL(ty, x), txz = M3D_MUL(tz, x);
	float tyy = M3D_MUL(ty, y), tyz = M3D_MUL(tz, y), tzz = M3D_MUL(tz, z);
	m[0] = M3D_SUB(1.0f, M3D_ADD(tyy, tzz)); m[1] = M3D_SUB(txy, twz); m[2] = M3D_ADD(txz, twy); m[3] = xyz[0];
	m[4] = M3D_ADD(txy, twz); m[5] = M3D_SUB(1.0f, M3D_ADD(txx, tzz)); m[6] = M3D_SUB(tyz, twx); m[7] = xyz[1];
	m[8] = M3D_SUB(txz, twy); m[9] = M3D_ADD(tyz, twx); m[10] = M3D_SUB(1.0f, M3D_ADD(txx, tyy)); m[11] = xyz[2];
	m[12] = 0.0f; m[13] = 0.0f; m[14] = 0.0f; m[15] = 1.0f;
#undef M3D_MUL
#undef M3D_ADD
#undef M3D_SUB
}

__device__ __forceinline__ void pose_prepare_warp(PoseState *ps, int lane, const float *m_regs);

/* Start of an iteration of registerLastArrivedScan (gpu6DSLAM.cpp:276-291): Euler round trip of the stored pose. */
__device__ inline void pose_prepare(PoseState *ps)
{
	float of[3], t[3];
	matrix4_to_euler(ps->m, of, t);
	euler_to_matrix(of, t, ps->pose1);
	ps->pose6[0] = t[0]; ps->pose6[1] = t[1]; ps->pose6[2] = t[2];
	ps->pose6[3] = of[0]; ps->pose6[4] = of[1]; ps->pose6[5] = of[2];
}

__global__ void k_pose_prepare(PoseState *ps)
{
	pdl_enter();
	if (blockIdx.x == 0 && threadIdx.x < 32) pose_prepare_warp(ps, threadIdx.x, nullptr);
}

/* Start of a fused loop: the pose state is cleared and the caller's pose arrives as a KERNEL ARGUMENT — a host-to-device
 * copy of 300 bytes would queue on the copy engine behind whatever upload is running (the host-buffer iteration keeps the
 * second cloud's 42 MB in flight at that moment) — then the first iteration's Euler round trip. */
struct Pose16 { float m[16]; };
__global__ void k_pose_init(PoseState *ps, const Pose16 pose, unsigned long long *label_counts, int *flags, unsigned int *ticket)
{
	pdl_enter();
	if (blockIdx.x != 0 || threadIdx.x >= 32) return;
	const int lane = threadIdx.x;
	if (lane < 4) label_counts[lane] = 0ull;
	if (lane < FLAG_COUNT) flags[lane] = 0;
	if (lane == 0) *ticket = 0u;
	unsigned int *w = reinterpret_cast<unsigned int *>(ps);
	for (int k = lane; k < (int)(sizeof(PoseState) / sizeof(unsigned int)); k += 32) w[k] = 0u;
	__syncwarp();
	if (lane < 16) ps->m[lane] = pose.m[lane];
	__syncwarp();
	pose_prepare_warp(ps, lane, nullptr);
}

/* ---- warp-cooperative versions of the serial tail ------------------------------------------------------------
 * The last block's tail (6x6 system, Cholesky, pose update, two Euler conversions) is a chain of ~20 double
 * precision transcendental calls when one thread runs it; here the calls that do not depend on each other are
 * spread over lanes executing the SAME code (no divergence), which cuts the chain to 6 call latencies.  Results are
 * bit-identical to the serial helpers above (same operations on the same operands). */
__device__ __forceinline__ double shfl_f64(double v, int src)
{
	return __shfl_sync(0xffffffffu, v, src);
}

/* every lane ends up with the 3x4 part of the matrix in r[12] (row-major [R|t]); lane 0 also stores the 4x4 to m (if any) */
__device__ __forceinline__ void euler_to_matrix_warp(const float *omfika, const float *xyz, float *m, int lane, float *r = nullptr)
{
	/* lanes 0..2: sin/cos of the three half angles */
	float h = 0.5f * omfika[lane < 3 ? lane : 0];
	float sn = (float)sin((double)h), cs = (float)cos((double)h);
	float ax = __shfl_sync(0xffffffffu, sn, 0), aw = __shfl_sync(0xffffffffu, cs, 0);
	float by = __shfl_sync(0xffffffffu, sn, 1), bw = __shfl_sync(0xffffffffu, cs, 1);
	float cz = __shfl_sync(0xffffffffu, sn, 2), cw = __shfl_sync(0xffffffffu, cs, 2);
	float w1 = __fmul_rn(aw, bw), x1 = __fmul_rn(ax, bw), y1 = __fmul_rn(aw, by), z1 = __fmul_rn(ax, by);
	float w = __fsub_rn(__fmul_rn(w1, cw), __fmul_rn(z1, cz));
	float x = __fadd_rn(__fmul_rn(x1, cw), __fmul_rn(y1, cz));
	float y = __fsub_rn(__fmul_rn(y1, cw), __fmul_rn(x1, cz));
	float z = __fadd_rn(__fmul_rn(w1, cz), __fmul_rn(z1, cw));
	float tx = __fmul_rn(2.0f, x), ty = __fmul_rn(2.0f, y), tz = __fmul_rn(2.0f, z);
	float twx = __fmul_rn(tx, w), twy = __fmul_rn(ty, w), twz = __fmul_rn(tz, w);
	float txx = __fmul_rn(tx, x), txy = __fmul_rn(ty, x), txz = __fmul_rn(tz, x);
	float tyy = __fmul_rn(ty, y), tyz = __fmul_rn(tz, y), tzz = __fmul_rn(tz, z);
	float q[12];
	q[0] = __fsub_rn(1.0f, __fadd_rn(tyy, tzz)); q[1] = __fsub_rn(txy, twz); q[2] = __fadd_rn(txz, twy); q[3] = xyz[0];
	q[4] = __fadd_rn(txy, twz); q[5] = __fsub_rn(1.0f, __fadd_rn(txx, tzz)); q[6] = __fsub_rn(tyz, twx); q[7] = xyz[1];
	q[8] = __fsub_rn(txz, twy); q[9] = __fadd_rn(tyz, twx); q[10] = __fsub_rn(1.0f, __fadd_rn(txx, tyy)); q[11] = xyz[2];
	if (r) {
#pragma unroll
		for (int k = 0; k < 12; k++) r[k] = q[k];
	}
	if (m && lane == 0) {
#pragma unroll
		for (int k = 0; k < 12; k++) m[k] = q[k];
		m[12] = 0.0f; m[13] = 0.0f; m[14] = 0.0f; m[15] = 1.0f;
	}
	__syncwarp();
}

/* m: 16 floats readable by every lane (shared or global, already visible); results returned in every lane. */
__device__ __forceinline__ void matrix4_to_euler_warp(const float *m, float *omfika, float *xyz, int lane)
{
	const double kPi = 3.14159265358979323846;
	float fi;
	if (m[0] > 0.0) fi = (float)asin((double)m[2]);
	else fi = (float)(kPi - asin((double)m[2]));
	double C = cos((double)fi);
	float om, ka;
	if (fabs(C) > 0.005) {
		/* lane 0: om = atan2(-m6/C, m10/C); lane 1: ka = atan2(-m1/C, m0/C) */
		double trX = (lane == 0 ? m[10] : m[0]) / C, trY = -(lane == 0 ? m[6] : m[1]) / C;
		float a = (float)atan2(trY, trX);
		om = __shfl_sync(0xffffffffu, a, 0);
		ka = __shfl_sync(0xffffffffu, a, 1);
	} else {
		om = 0.0f;
		ka = (float)atan2((double)m[4], (double)m[5]);
	}
	omfika[0] = om; omfika[1] = fi; omfika[2] = ka;
	xyz[0] = m[3]; xyz[1] = m[7]; xyz[2] = m[11];
}

/* Start of an iteration (gpu6DSLAM.cpp:276-291) by one warp: Euler round trip of the stored pose into pose1 / pose6.
 * m_regs (12 floats, every lane): the 3x4 part of ps->m when the caller still holds it in registers (0: read ps->m). */
__device__ __forceinline__ void pose_prepare_warp(PoseState *ps, int lane, const float *m_regs = nullptr)
{
	float of[3], t[3];
	if (m_regs) matrix4_to_euler_warp(m_regs, of, t, lane);      /* two inlined copies: a register array must not meet a pointer select */
	else matrix4_to_euler_warp(ps->m, of, t, lane);
	euler_to_matrix_warp(of, t, ps->pose1, lane);
	if (lane == 0) {
		ps->pose6[0] = t[0]; ps->pose6[1] = t[1]; ps->pose6[2] = t[2];
		ps->pose6[3] = of[0]; ps->pose6[4] = of[1]; ps->pose6[5] = of[2];
	}
	__syncwarp();
}

/* 6x6 system from the 24 moments, one output per lane (27 outputs), same formula for every lane:
 * with A = -[I | J], J[r][c] = C[r][c] . p0, every column a of A is, in row r, -(E[r][a] . (1, p0)) for a 4-vector
 * E[r][a] ( (delta_ra,0,0,0) for a < 3, (0, C[r][a-3]) otherwise ), hence
 *     N[a][b] = sum_r E[r][a]^T Mext E[r][b],   rhs[a] = - sum_r E[r][a]^T Lext[:, r],
 * Mext = [[S, M1^T], [M1, M2]] (4x4), Lext = [[L1^T], [L2]] (4x3). */
__device__ inline void moments_to_neq_warp(const double *mo, double om, double fi, double ka, double *neq, int lane)
{
	/* lanes 0..2: sincos of om, fi, ka */
	double ang = lane == 0 ? om : (lane == 1 ? fi : ka), sn, cs;
	sincos(ang, &sn, &cs);
	double so = shfl_f64(sn, 0), co = shfl_f64(cs, 0), sf = shfl_f64(sn, 1), cf = shfl_f64(cs, 1), sk = shfl_f64(sn, 2), ck = shfl_f64(cs, 2);
	double R11 = cf * ck, R12 = -cf * sk;
	double R21 = co * sk + so * sf * ck, R22 = co * ck - so * sf * sk, R23 = -so * cf;
	double R31 = so * sk - co * sf * ck, R32 = so * ck + co * sf * sk, R33 = co * cf;
	const double C[3][3][3] = {
		{{0, 0, 0}, {-sf * ck, sf * sk, cf}, {R12, -R11, 0}},
		{{-R31, -R32, -R33}, {so * cf * ck, -so * cf * sk, so * sf}, {R22, -R21, 0}},
		{{R21, R22, R23}, {-co * cf * ck, co * cf * sk, -co * sf}, {R32, -R31, 0}}};
	/* output index -> (a, b): k < 21 walks the upper triangle row by row, 21..26 are the right-hand side */
	int a = 0, b = 0;
	bool rhs = lane >= 21;
	if (!rhs) { int k = lane, row = 0; while (k >= 6 - row) { k -= 6 - row; row++; } a = row; b = row + k; }
	else a = lane - 21;
	const double Mext[4][4] = {{mo[0], mo[1], mo[2], mo[3]}, {mo[1], mo[4], mo[5], mo[6]}, {mo[2], mo[5], mo[7], mo[8]}, {mo[3], mo[6], mo[8], mo[9]}};
	double acc = 0.0;
#pragma unroll
	for (int r = 0; r < 3; r++) {
		double ea[4], eb[4];
#pragma unroll
		for (int i = 0; i < 4; i++) {
			double ca = 0.0, cb = 0.0;
#pragma unroll
			for (int c = 0; c < 3; c++) {   /* select C[r][a-3][i-1] / C[r][b-3][i-1] without dynamic indexing */
				if (i > 0 && a == 3 + c) ca = C[r][c][i - 1];
				if (i > 0 && b == 3 + c) cb = C[r][c][i - 1];
			}
			ea[i] = (a < 3) ? ((i == 0 && a == r) ? 1.0 : 0.0) : ca;
			eb[i] = (b < 3) ? ((i == 0 && b == r) ? 1.0 : 0.0) : cb;
		}
		if (!rhs) {
#pragma unroll
			for (int i = 0; i < 4; i++) {
				double t = Mext[i][0] * eb[0] + Mext[i][1] * eb[1] + Mext[i][2] * eb[2] + Mext[i][3] * eb[3];
				acc += ea[i] * t;
			}
		} else {
			/* Lext[:, r] = (L1[r], L2[0*3+r], L2[1*3+r], L2[2*3+r]) */
			acc -= ea[0] * mo[10 + r] + ea[1] * mo[13 + r] + ea[2] * mo[16 + r] + ea[3] * mo[19 + r];
		}
	}
	if (lane < 27) neq[lane] = acc;
	if (lane == 27) neq[27] = mo[22];
	__syncwarp();
}

/* Per-observation sources for the moment reduction: load(i, raw) issues the (independent, coalesced) loads of
 * observation i and says whether there is one; finish(raw, ...) is the arithmetic.  A thread keeps several in flight. */
struct ObsFromRec {  /* fused path: the search's per-query records + the queries (gpu6DSLAM.cpp:323-398 done on the device) */
	const float4 *rec;          /* {x0, y0, z0 of the matched first-cloud point in ITS LOCAL frame, bits of its index or -1} */
	const float4 *q_xyzl;       /* queries (second cloud, global)           */
	const float *m;             /* pose the first cloud was transformed with this iteration (device, row-major 4x4) */
	const unsigned long long *label_counts;   /* 4 counters per segment */
	const int *seg_of_chunk;    /* batched sweep: segment of query i / kSegChunk (0: one segment) */
	int n_segs;
	float weight[4];
	float r[12];                /* m's 3x4 part, loaded once per thread by prepare() */
	struct Raw { float4 p2, p0; int seg; };
	__device__ __forceinline__ void prepare()
	{
#pragma unroll
		for (int k = 0; k < 12; k++) r[k] = __ldg(m + k);
	}
	__device__ __forceinline__ void prefetch(int i) const      /* the next trip's two streams into L2 while this trip computes */
	{
		asm volatile("prefetch.global.L2 [%0];" :: "l"(rec + i));
		asm volatile("prefetch.global.L2 [%0];" :: "l"(q_xyzl + i));
	}
	__device__ __forceinline__ bool load(int i, Raw &raw) const
	{
		raw.p0 = __ldg(rec + i); raw.p2 = __ldg(q_xyzl + i);
		raw.seg = seg_of_chunk ? __ldg(seg_of_chunk + i / kSegChunk) : 0;
		return __float_as_int(raw.p0.w) >= 0;
	}
	__device__ __forceinline__ void finish(const Raw &raw, const float *wl, double &w, double &x, double &y, double &z,
			double &lx, double &ly, double &lz) const
	{
		int label = __float_as_int(raw.p2.w);
		w = (label >= 0 && label < 4) ? (double)wl[4 * raw.seg + label] : 0.0;
		x = raw.p0.x; y = raw.p0.y; z = raw.p0.z;
		/* the matched point in the global frame, recomputed with the transform's exact operation sequence (same bits as
		 * the transformed cloud the search ran on) instead of a second gather */
		float p1x = __fadd_rn(r[3], __fmaf_rn(r[2], raw.p0.z, __fmaf_rn(r[0], raw.p0.x, __fmul_rn(r[1], raw.p0.y))));
		float p1y = __fadd_rn(r[7], __fmaf_rn(r[6], raw.p0.z, __fmaf_rn(r[4], raw.p0.x, __fmul_rn(r[5], raw.p0.y))));
		float p1z = __fadd_rn(r[11], __fmaf_rn(r[10], raw.p0.z, __fmaf_rn(r[8], raw.p0.x, __fmul_rn(r[9], raw.p0.y))));
		lx = (double)__fsub_rn(p1x, raw.p2.x); ly = (double)__fsub_rn(p1y, raw.p2.y); lz = (double)__fsub_rn(p1z, raw.p2.z);
	}
};

struct ObsFromList { /* stage-level path: the reference's obs_nn_t array */
	const m3dreg_obs_nn *obs;
	struct Raw { float v[7]; };
	__device__ __forceinline__ void prepare() {}
	__device__ __forceinline__ void prefetch(int) const {}
	__device__ __forceinline__ bool load(int i, Raw &r) const
	{
		const float *o = reinterpret_cast<const float *>(obs + i);
#pragma unroll
		for (int k = 0; k < 7; k++) r.v[k] = __ldg(o + k);
		return true;
	}
	__device__ __forceinline__ void finish(const Raw &r, const float *, double &w, double &x, double &y, double &z,
			double &lx, double &ly, double &lz) const
	{
		lx = r.v[0]; ly = r.v[1]; lz = r.v[2]; x = r.v[3]; y = r.v[4]; z = r.v[5]; w = r.v[6];
	}
};

/* What the last block does once all partial rows are in. */
struct FinalizeArgs {
	PoseState *ps;            /* pose state to read Euler angles from / update (may be 0: only write neq_out)      */
	double *neq_out;          /* 28 doubles, overwritten (mode 0) or accumulated into (mode 1)                      */
	int accumulate;           /* 1: neq_out += (sweep)                                                               */
	int solve;                /* 1: gate + Cholesky + pose update + next-iteration pose_prepare                      */
	int dof;
	int obs_threshold;
	const double *pose6_in;   /* Euler angles when ps == 0                                                           */
	uint32_t *bounds_reset;   /* reset for the next iteration (may be 0)                                             */
	unsigned long long *label_counts_reset;
	int label_count_sets;     /* 4-counter sets to reset (0 counts as 1) */
};

/* What warp 0 of the last block does with the finished 28-double system `neq` (shared memory): publish / accumulate
 * it, and (fused loop) gate on the observation count, Cholesky, pose update, Euler round trip for the next iteration,
 * resets. */
__device__ inline void neq_tail_warp(const double *neq, const FinalizeArgs &fin, int lane, const double *p6_in = nullptr)
{
	if (fin.neq_out && lane < kNeqCount) fin.neq_out[lane] = fin.accumulate ? fin.neq_out[lane] + neq[lane] : neq[lane];
	if (fin.solve && fin.ps) {
		PoseState *ps = fin.ps;
		if (lane < kNeqCount) ps->neq[lane] = neq[lane];
		long long n_obs = (long long)(neq[27] + 0.5);
		int status = M3DREG_E_TOO_FEW_OBS;
		double x[6] = {0, 0, 0, 0, 0, 0};
		if (n_obs > (long long)fin.obs_threshold) status = solve_packed(neq, fin.dof, x);     /* gpu6DSLAM.cpp:402; uniform over lanes */
		double p6[6];
#pragma unroll
		for (int k = 0; k < 6; k++) p6[k] = p6_in ? p6_in[k] : ps->pose6[k];
		__syncwarp();
		float mr[12];
		bool have_m = false;
		if (status == 0) {
			/* registerLS tail (cudaWrapper.cpp:574-579 / 641-646) + EulerToMatrix (gpu6DSLAM.cpp:408-413) */
			p6[0] += x[0]; p6[1] += x[1]; p6[2] += x[2];
			if (fin.dof == 6) { p6[3] += x[3]; p6[4] += x[4]; p6[5] += x[5]; }
			else p6[5] += x[3];
			float of[3] = {(float)p6[3], (float)p6[4], (float)p6[5]};
			float t[3] = {(float)p6[0], (float)p6[1], (float)p6[2]};
			euler_to_matrix_warp(of, t, ps->m, lane, mr);
			have_m = true;
		}
		if (lane == 0) {
			for (int k = 0; k < 6; k++) ps->x[k] = x[k];
			ps->n_obs = n_obs;
			ps->status = status;
			ps->iterations += 1;
		}
		__syncwarp();
		pose_prepare_warp(ps, lane, have_m ? mr : nullptr);   /* next iteration's Euler round trip (the new pose is still in registers) */
	}
	if (lane == 0) {
		if (fin.bounds_reset) {
			fin.bounds_reset[0] = fin.bounds_reset[1] = fin.bounds_reset[2] = 0xFFFFFFFFu;
			fin.bounds_reset[3] = fin.bounds_reset[4] = fin.bounds_reset[5] = 0u;
		}
	}
	if (fin.label_counts_reset) {
		const int nc = 4 * (fin.label_count_sets > 0 ? fin.label_count_sets : 1);
		for (int k = lane; k < nc; k += 32) fin.label_counts_reset[k] = 0ull;
	}
}

constexpr int kNeqThreads = 256;
constexpr int kNeqInFlight = 4;

template <class Src>
__global__ void __launch_bounds__(kNeqThreads) k_normal_equations(const Src src_in, int n, double *__restrict__ partials,
		unsigned int *__restrict__ ticket, FinalizeArgs fin)
{
	pdl_enter();
	Src src = src_in;
	src.prepare();
	__shared__ double sm[kNeqThreads / 32][kMomentCount];
	__shared__ float wl[4 * kMaxSegs];
	__shared__ bool is_last;
	if constexpr (std::is_same<Src, ObsFromRec>::value) {
		static_assert(kNeqThreads / 32 * 3 == kMomentCount, "final reduction: three columns per warp");
		static_assert(4 * kMaxSegs <= kNeqThreads, "one thread per (segment, label) weight");
		if ((int)threadIdx.x < 4 * src.n_segs) {
			unsigned long long c = src.label_counts[threadIdx.x];
			/* P = weight / count, float / int -> float (gpu6DSLAM.cpp:377-393) */
			wl[threadIdx.x] = c ? __fdiv_rn(src.weight[threadIdx.x & 3], (float)(int)c) : 0.0f;
		}
	} else {
		if (threadIdx.x < 4) wl[threadIdx.x] = 0.0f;
	}
	__syncthreads();
	Moments mo;
	mo.clear();
	{
		/* kNeqInFlight observations per thread in flight: all loads (coalesced streams), then the arithmetic */
		const int stride = gridDim.x * blockDim.x;
		for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += kNeqInFlight * stride) {
			bool tok[kNeqInFlight];
			typename Src::Raw raw[kNeqInFlight];
#pragma unroll
			for (int k = 0; k < kNeqInFlight; k++) { int i = i0 + k * stride; tok[k] = i < n && src.load(i, raw[k]); }
#pragma unroll
			for (int k = 0; k < kNeqInFlight; k++) { int i = i0 + (kNeqInFlight + k) * stride; if (i < n) src.prefetch(i); }
#pragma unroll
			for (int k = 0; k < kNeqInFlight; k++) {
				if (tok[k]) {
					double w, x, y, z, lx, ly, lz;
					src.finish(raw[k], wl, w, x, y, z, lx, ly, lz);
					mo.add(w, x, y, z, lx, ly, lz);
				}
			}
		}
	}
	int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < kMomentCount; k++) {
		double s = warp_sum(mo.v[k]);
		if (lane == 0) sm[wid][k] = s;
	}
	__syncthreads();
	if (threadIdx.x < kMomentCount) {
		double s = 0;
#pragma unroll
		for (int k = 0; k < kNeqThreads / 32; k++) s += sm[k][threadIdx.x];
		partials[(size_t)blockIdx.x * kPartialCols + threadIdx.x] = s;
		__threadfence();      /* only the writers fence (release side of the ticket) */
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned int t = atomicAdd(ticket, 1u);
		is_last = (t == gridDim.x - 1);
		__threadfence();      /* acquire side: one fence by the thread that took the ticket; the rows are read past L1 */
	}
	__syncthreads();
	if (!is_last) return;
#ifdef M3D_NEQ_TIMING
	unsigned long long tq0, tq1, tq2, tq3;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq0));
#endif
	/* the Euler angles the system is formed for: requested now, needed after the reduction */
	const double *p6g = fin.ps ? fin.ps->pose6 : fin.pose6_in;
	double p6r[6];
#pragma unroll
	for (int k = 0; k < 6; k++) p6r[k] = wid == 0 ? __ldcg(p6g + k) : 0.0;
	/* deterministic final reduction, fixed order: warp w owns columns 3w .. 3w+2, lane l adds rows l, l + 32, ... (ten
	 * rows x three columns of loads in flight per lane: one L2 latency per 320 rows), then a butterfly over the lanes */
	__shared__ double tot[kMomentCount];
	{
		double s0 = 0.0, s1 = 0.0, s2 = 0.0;
		const int c0 = 3 * wid;
		for (unsigned int base = 0; base < gridDim.x; base += 320) {
			double v[10][3];
#pragma unroll
			for (int k = 0; k < 10; k++) {
				const unsigned int b = base + 32 * k + lane;
				const bool in = b < gridDim.x;
				const double *row = partials + (size_t)(in ? b : 0) * kPartialCols + c0;
				v[k][0] = in ? __ldcg(row) : 0.0; v[k][1] = in ? __ldcg(row + 1) : 0.0; v[k][2] = in ? __ldcg(row + 2) : 0.0;
			}
#pragma unroll
			for (int k = 0; k < 10; k++) { s0 += v[k][0]; s1 += v[k][1]; s2 += v[k][2]; }
		}
		s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
		if (lane == 0) { tot[c0] = s0; tot[c0 + 1] = s1; tot[c0 + 2] = s2; }
	}
	__syncthreads();
	if (wid == 0) {
		__shared__ double neq[kNeqCount];
		if (lane == 0) *ticket = 0;
#ifdef M3D_NEQ_TIMING
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq1));
#endif
		moments_to_neq_warp(tot, p6r[3], p6r[4], p6r[5], neq, lane);
#ifdef M3D_NEQ_TIMING
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq2));
#endif
		neq_tail_warp(neq, fin, lane, p6r);
#ifdef M3D_NEQ_TIMING
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq3));
		if (lane == 0) printf("neq tail ns: reduce %llu moments_to_neq %llu tail %llu\n", tq1 - tq0, tq2 - tq1, tq3 - tq2);
#endif
	}
}

/* Standalone device Cholesky (m3dreg_solve_chol): column-major dof x dof in, x out. */
__global__ void k_solve_dense(const double *A, const double *b, int dof, double *x, int *status)
{
	pdl_enter();
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	double neq[kNeqCount];
	const int sel6[6] = {0, 1, 2, 3, 4, 5}, sel4[4] = {0, 1, 2, 5};
	const int *sel = dof == 6 ? sel6 : sel4;
	for (int k = 0; k < kNeqCount; k++) neq[k] = 0;
	/* embed into the 6-DOF packing; unused rows get identity so the packed solver's selection is well defined */
	double full[6][6];
	for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) full[i][j] = (i == j) ? 1.0 : 0.0;
	for (int i = 0; i < dof; i++) for (int j = 0; j < dof; j++) full[sel[i]][sel[j]] = A[i + j * dof];
	int k = 0;
	for (int i = 0; i < 6; i++) for (int j = i; j < 6; j++) neq[k++] = full[i][j];
	for (int i = 0; i < dof; i++) neq[21 + sel[i]] = b[i];
	double xs[6] = {0, 0, 0, 0, 0, 0};
	*status = solve_packed(neq, dof, xs);
	for (int i = 0; i < dof; i++) x[i] = xs[i];
}

/* Sweep solve (registerAll tail, gpu6DSLAM.cpp:572-593) for scans [begin,end): one thread per scan. */
__global__ void k_sweep_solve(const double *__restrict__ neq, int begin, int end, float *__restrict__ poses,
		int dof, int obs_threshold, int *__restrict__ status_out)
{
	pdl_enter();
	int s = begin + blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= end) return;
	float *m = poses + 16 * (size_t)s;
	float of[3], t[3];
	matrix4_to_euler(m, of, t);
	double p6[6] = {t[0], t[1], t[2], of[0], of[1], of[2]};
	const double *q = neq + (size_t)s * kNeqCount;
	long long n_obs = (long long)(q[27] + 0.5);
	int status = M3DREG_E_TOO_FEW_OBS;
	if (n_obs > (long long)obs_threshold) {
		double x[6] = {0, 0, 0, 0, 0, 0};
		double loc[kNeqCount];
		for (int k = 0; k < kNeqCount; k++) loc[k] = q[k];
		status = solve_packed(loc, dof, x);
		if (status == 0) {
			p6[0] += x[0]; p6[1] += x[1]; p6[2] += x[2];
			if (dof == 6) { p6[3] += x[3]; p6[4] += x[4]; p6[5] += x[5]; }
			else p6[5] += x[3];
		}
	}
	float of2[3] = {(float)p6[3], (float)p6[4], (float)p6[5]};
	float t2[3] = {(float)p6[0], (float)p6[1], (float)p6[2]};
	euler_to_matrix(of2, t2, m);     /* replaced by its round trip even when the gate fails (gpu6DSLAM.cpp:586-593) */
	if (status_out) status_out[s] = status;
}

__global__ void k_zero_f64(double *p, int n)
{
	pdl_enter();
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0.0;
}

/* ================================ NDT (point-to-distribution) ====================================================
 * Not in the reference; definition of record: oracle/m3d_oracle.c orc_ndt_normal_equations.
 * Per bucket of the gridded (moving) cloud: mean / covariance of its points' GLOBAL coordinates, accumulated
 * relative to the cell centre in fp64, and the mean of their LOCAL coordinates; W = (Sigma + eps I)^-1.
 * All queries that fall into a bucket share its (mu, W, Jacobian), so the query pass only needs a COUNT and a
 * coordinate SUM per bucket; the normal equations are then a reduction over buckets:
 *     N += cnt * A^T W A,   rhs += A^T W (cnt*mu_g - sum q),   A = -[I | J(mu_l)].
 * Layout: iacc[b*12 ..] = {sum p'(3), sum p'p'^T(6), sum p_local(3)} as fixed-point int64 -> finalised into
 *         acc[b*12 ..] = {mu_g(3), mu_l(3), W(6: xx,xy,xz,yy,yz,zz)}; W.xx == 0 marks an unusable bucket.
 *         qacc[b*4..] = {cnt, sum (q - cell centre)(3) in 2^-40 m} as int64.
 * Deterministic: every sum that is accumulated with atomics is an INTEGER. */
constexpr int kNdtMinPoints = 5;
constexpr double kNdtRegRel = 0.05;

__global__ void k_ndt_zero(long long *__restrict__ iacc, long long *__restrict__ qacc, const m3dreg_grid_params *__restrict__ gp, int zero_acc)
{
	pdl_enter();
	long long nb = gp->number_of_buckets;
	long long total = nb * (zero_acc ? 16 : 4);
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		if (i < nb * 4) qacc[i] = 0;
		else iacc[i - nb * 4] = 0;
	}
}

__device__ __forceinline__ void cell_centre(uint32_t key, const m3dreg_grid_params *gp, double &cx, double &cy, double &cz)
{
	int nby = gp->number_of_buckets_Y, nbz = gp->number_of_buckets_Z;
	int ix = (int)(key / (uint32_t)(nby * nbz)), iy = (int)((key / (uint32_t)nbz) % (uint32_t)nby), iz = (int)(key % (uint32_t)nbz);
	cx = (double)gp->bounding_box_min_X + ((double)ix + 0.5) * (double)gp->resolution_X;
	cy = (double)gp->bounding_box_min_Y + ((double)iy + 0.5) * (double)gp->resolution_Y;
	cz = (double)gp->bounding_box_min_Z + ((double)iz + 0.5) * (double)gp->resolution_Z;
}

/* Segmented warp reduction over runs of equal keys (lanes with key 0xFFFFFFFF are idle): after the call the FIRST
 * lane of every run holds the run total. */
template <int NV>
__device__ __forceinline__ void warp_segmented_sum(uint32_t key, double (&v)[NV], bool &is_head)
{
	const unsigned full = 0xffffffffu;
	int lane = threadIdx.x & 31;
	uint32_t prev = __shfl_up_sync(full, key, 1);
	is_head = (lane == 0) || (prev != key);
	/* a run is identified by the lane of its head, not by its key: the same key may re-occur later in the warp
	 * (unsorted queries), and comparing keys would then add a later run into an earlier one */
	unsigned heads = __ballot_sync(full, is_head);
	int seg = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		int s2 = __shfl_down_sync(full, seg, o);
		bool take = (lane + o < 32) && (s2 == seg);
#pragma unroll
		for (int i = 0; i < NV; i++) {
			double t = __shfl_down_sync(full, v[i], o);
			if (take) v[i] += t;
		}
	}
}

/* Fixed-point scales of the NDT sums (chosen on the host so that 2^22 points of one bucket cannot overflow 63 bits):
 * s1 for coordinates relative to the cell centre, s2 for their products, sl for local coordinates. */
struct NdtScales { double s1, s2, sl; };

/* Per-bucket sums of the gridded cloud, one thread per SORTED position: coordinates relative to the cell centre, their
 * products and the local coordinates as 64-bit FIXED-POINT integers; runs of equal keys inside a warp are added with
 * shuffles, one integer atomic per run piece and sum.  Integer addition is associative: the sums — hence everything
 * downstream — do not depend on scheduling (round 1 used fp64 atomicAdd: not reproducible run to run), and the work is
 * spread evenly whatever the bucket sizes (a warp per bucket took 700 us on the 1 M-point scan: one bucket near the
 * sensor holds 20 000 points).  iacc[b*12..] = {sum p'(3), sum p'p'^T(6), sum p_local(3)}. */
__global__ void __launch_bounds__(256) k_ndt_accumulate_points(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int n,
		const float4 *__restrict__ g_xyzl, const float4 *__restrict__ l_xyzl, const m3dreg_grid_params *__restrict__ gp,
		long long *__restrict__ iacc, NdtScales sc)
{
	pdl_enter();
	if (gp->number_of_buckets <= 0) return;
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	int nround = (n + 31) & ~31;
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nround; p += gridDim.x * blockDim.x) {
		uint32_t key = 0xFFFFFFFFu;
		long long v[12];
#pragma unroll
		for (int i = 0; i < 12; i++) v[i] = 0;
		if (p < n) {
			key = __ldg(keys + p);
			uint32_t idx = __ldg(vals + p);
			float4 g = __ldg(g_xyzl + idx), l = __ldg(l_xyzl + idx);
			double cx, cy, cz;
			cell_centre(key, gp, cx, cy, cz);
			double x = (double)g.x - cx, y = (double)g.y - cy, z = (double)g.z - cz;
			v[0] = __double2ll_rn(x * sc.s1); v[1] = __double2ll_rn(y * sc.s1); v[2] = __double2ll_rn(z * sc.s1);
			v[3] = __double2ll_rn(x * x * sc.s2); v[4] = __double2ll_rn(x * y * sc.s2); v[5] = __double2ll_rn(x * z * sc.s2);
			v[6] = __double2ll_rn(y * y * sc.s2); v[7] = __double2ll_rn(y * z * sc.s2); v[8] = __double2ll_rn(z * z * sc.s2);
			v[9] = __double2ll_rn((double)l.x * sc.sl); v[10] = __double2ll_rn((double)l.y * sc.sl); v[11] = __double2ll_rn((double)l.z * sc.sl);
		}
		uint32_t prev = __shfl_up_sync(full, key, 1);
		bool is_head = (lane == 0) || (prev != key);
		unsigned heads = __ballot_sync(full, is_head);
		int seg = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			int s2 = __shfl_down_sync(full, seg, o);
			bool take = (lane + o < 32) && (s2 == seg);
#pragma unroll
			for (int i = 0; i < 12; i++) {
				long long t = __shfl_down_sync(full, v[i], o);
				if (take) v[i] += t;
			}
		}
		if (is_head && key != 0xFFFFFFFFu) {
			unsigned long long *a = reinterpret_cast<unsigned long long *>(iacc + (size_t)key * 12);
#pragma unroll
			for (int i = 0; i < 12; i++) atomicAdd(a + i, (unsigned long long)v[i]);
		}
	}
}

__device__ __host__ inline bool sym3_inverse(const double *S, double *W)
{
	double a = S[0], b = S[1], c = S[2], d = S[3], e = S[4], f = S[5];
	double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
	double det = a * c00 + b * c01 + c * c02;
	if (!(det > 0.0)) return false;
	double id = 1.0 / det;
	W[0] = c00 * id; W[1] = c01 * id; W[2] = c02 * id;
	W[3] = (a * f - c * c) * id; W[4] = (b * c - a * e) * id; W[5] = (a * d - b * b) * id;
	return true;
}

__global__ void k_ndt_finalize_buckets(const m3dreg_bucket *__restrict__ buckets, const m3dreg_grid_params *__restrict__ gp,
		const long long *__restrict__ iacc, double *__restrict__ acc, NdtScales sc)
{
	pdl_enter();
	long long nb = gp->number_of_buckets;
	double res = (double)gp->resolution_X;
	double eps = (kNdtRegRel * res) * (kNdtRegRel * res);
	for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (long long)gridDim.x * blockDim.x) {
		int n = buckets[b].number_of_points;
		double *a = acc + (size_t)b * 12;
		a[6] = 0.0;                                  /* unusable until proven otherwise */
		if (n < kNdtMinPoints) continue;
		const long long *ia = iacc + (size_t)b * 12;
		double s[3] = {(double)ia[0] / sc.s1, (double)ia[1] / sc.s1, (double)ia[2] / sc.s1};
		double ss[6] = {(double)ia[3] / sc.s2, (double)ia[4] / sc.s2, (double)ia[5] / sc.s2, (double)ia[6] / sc.s2, (double)ia[7] / sc.s2, (double)ia[8] / sc.s2};
		double sl[3] = {(double)ia[9] / sc.sl, (double)ia[10] / sc.sl, (double)ia[11] / sc.sl};
		double inv = 1.0 / n, d = 1.0 / (n - 1);
		double m[3] = {s[0] * inv, s[1] * inv, s[2] * inv};
		double S[6], W[6];
		S[0] = (ss[0] - n * m[0] * m[0]) * d + eps; S[1] = (ss[1] - n * m[0] * m[1]) * d; S[2] = (ss[2] - n * m[0] * m[2]) * d;
		S[3] = (ss[3] - n * m[1] * m[1]) * d + eps; S[4] = (ss[4] - n * m[1] * m[2]) * d; S[5] = (ss[5] - n * m[2] * m[2]) * d + eps;
		if (!sym3_inverse(S, W)) continue;
		double cx, cy, cz;
		cell_centre((uint32_t)b, gp, cx, cy, cz);
		a[0] = m[0] + cx; a[1] = m[1] + cy; a[2] = m[2] + cz;
		a[3] = sl[0] * inv; a[4] = sl[1] * inv; a[5] = sl[2] * inv;
#pragma unroll
		for (int i = 0; i < 6; i++) a[6 + i] = W[i];
	}
}

/* Queries per bucket: a count and the coordinate sum RELATIVE TO THE CELL CENTRE in 2^-40 m fixed point, added with
 * 64-bit INTEGER atomics — integer addition is associative, so the sums do not depend on the order the warps arrive in
 * (fp64 atomics did).  |q - centre| < 1.5 res per axis and at most 2^20 queries of a bucket fit 63 bits for res <= 4 m;
 * the quantum (9e-13 m) is far below the float coordinates' own resolution.  qacc[b*4..] = {count, sx, sy, sz} as int64. */
constexpr double kNdtFixScale = 1099511627776.0;      /* 2^40 */

__global__ void __launch_bounds__(256) k_ndt_accumulate_queries(const float4 *__restrict__ q_xyzl, int n2,
		const m3dreg_grid_params *__restrict__ gp, const double *__restrict__ acc, long long *__restrict__ qacc)
{
	pdl_enter();
	long long nb = gp->number_of_buckets;
	if (nb <= 0) return;
	float mnx = gp->bounding_box_min_X, mny = gp->bounding_box_min_Y, mnz = gp->bounding_box_min_Z;
	float rx = gp->resolution_X, ry = gp->resolution_Y, rz = gp->resolution_Z;
	int nby = gp->number_of_buckets_Y, nbz = gp->number_of_buckets_Z;
	int nround = (n2 + 31) & ~31;
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
		uint32_t key = 0xFFFFFFFFu;
		long long v[4] = {0, 0, 0, 0};
		if (i < n2) {
			float4 p = __ldg(q_xyzl + i);
			bool inside = !(p.x < mnx || p.x > gp->bounding_box_max_X) && !(p.y < mny || p.y > gp->bounding_box_max_Y) &&
					!(p.z < mnz || p.z > gp->bounding_box_max_Z);
			if (inside) {
				int h = cell_of(p.x, mnx, rx) * nby * nbz + cell_of(p.y, mny, ry) * nbz + cell_of(p.z, mnz, rz);
				if (h >= 0 && (long long)h < nb && __ldg(acc + (size_t)h * 12 + 6) > 0.0) {
					key = (uint32_t)h;
					double cx, cy, cz;
					cell_centre(key, gp, cx, cy, cz);
					v[0] = 1;
					v[1] = __double2ll_rn(((double)p.x - cx) * kNdtFixScale);
					v[2] = __double2ll_rn(((double)p.y - cy) * kNdtFixScale);
					v[3] = __double2ll_rn(((double)p.z - cz) * kNdtFixScale);
				}
			}
		}
		/* runs of equal keys inside the warp are added up first (integer: any order gives the same sum) */
		uint32_t prev = __shfl_up_sync(full, key, 1);
		bool is_head = (lane == 0) || (prev != key);
		unsigned heads = __ballot_sync(full, is_head);
		int seg = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			int s2 = __shfl_down_sync(full, seg, o);
			bool take = (lane + o < 32) && (s2 == seg);
#pragma unroll
			for (int k = 0; k < 4; k++) {
				long long t = __shfl_down_sync(full, v[k], o);
				if (take) v[k] += t;
			}
		}
		if (is_head && key != 0xFFFFFFFFu) {
			unsigned long long *a = reinterpret_cast<unsigned long long *>(qacc + (size_t)key * 4);
#pragma unroll
			for (int k = 0; k < 4; k++) atomicAdd(a + k, (unsigned long long)v[k]);
		}
	}
}

__global__ void __launch_bounds__(kNeqThreads) k_ndt_normal_equations(const double *__restrict__ acc, const long long *__restrict__ qacc,
		const m3dreg_grid_params *__restrict__ gp, double *__restrict__ partials, unsigned int *__restrict__ ticket, FinalizeArgs fin)
{
	pdl_enter();
	__shared__ double sm[kNeqThreads / 32][kPartialCols];
	__shared__ bool is_last;
	long long nb = gp->number_of_buckets;
	const double *p6 = fin.ps ? fin.ps->pose6 : fin.pose6_in;
	double om = p6[3], fi = p6[4], ka = p6[5];
	double so = sin(om), co = cos(om), sf = sin(fi), cf = cos(fi), sk = sin(ka), ck = cos(ka);
	double R11 = cf * ck, R12 = -cf * sk;
	double R21 = co * sk + so * sf * ck, R22 = co * ck - so * sf * sk, R23 = -so * cf;
	double R31 = so * sk - co * sf * ck, R32 = so * ck + co * sf * sk, R33 = co * cf;
	const double C[3][3][3] = {
		{{0, 0, 0}, {-sf * ck, sf * sk, cf}, {R12, -R11, 0}},
		{{-R31, -R32, -R33}, {so * cf * ck, -so * cf * sk, so * sf}, {R22, -R21, 0}},
		{{R21, R22, R23}, {-co * cf * ck, co * cf * sk, -co * sf}, {R32, -R31, 0}}};
	double sum[kPartialCols];
#pragma unroll
	for (int k = 0; k < kPartialCols; k++) sum[k] = 0.0;
	for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (long long)gridDim.x * blockDim.x) {
		const double *a = acc + (size_t)b * 12;
		const long long *qa = qacc + (size_t)b * 4;
		const double cnt = (double)qa[0];
		if (!(a[6] > 0.0) || !(cnt > 0.0)) continue;
		double ccx, ccy, ccz;
		cell_centre((uint32_t)b, gp, ccx, ccy, ccz);
		/* sum of the bucket's queries = count x centre + fixed-point sum of the offsets */
		const double sqx = cnt * ccx + (double)qa[1] / kNdtFixScale, sqy = cnt * ccy + (double)qa[2] / kNdtFixScale,
				sqz = cnt * ccz + (double)qa[3] / kNdtFixScale;
		double A[3][6];
#pragma unroll
		for (int r = 0; r < 3; r++)
#pragma unroll
			for (int c = 0; c < 3; c++) {
				A[r][c] = (r == c) ? -1.0 : 0.0;
				A[r][3 + c] = -(C[r][c][0] * a[3] + C[r][c][1] * a[4] + C[r][c][2] * a[5]);
			}
		double W[3][3] = {{a[6], a[7], a[8]}, {a[7], a[9], a[10]}, {a[8], a[10], a[11]}};
		double sl[3] = {cnt * a[0] - sqx, cnt * a[1] - sqy, cnt * a[2] - sqz};
		int k = 0;
#pragma unroll
		for (int i = 0; i < 6; i++) {
			double wa[3];
#pragma unroll
			for (int r = 0; r < 3; r++) wa[r] = A[0][i] * W[0][r] + A[1][i] * W[1][r] + A[2][i] * W[2][r];
#pragma unroll
			for (int j = i; j < 6; j++) sum[k++] += cnt * (wa[0] * A[0][j] + wa[1] * A[1][j] + wa[2] * A[2][j]);
			sum[21 + i] += wa[0] * sl[0] + wa[1] * sl[1] + wa[2] * sl[2];
		}
		sum[27] += cnt;
	}
	int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < kPartialCols; k++) {
		double s = warp_sum(sum[k]);
		if (lane == 0) sm[wid][k] = s;
	}
	__syncthreads();
	if (threadIdx.x < kPartialCols) {
		double s = 0;
#pragma unroll
		for (int k = 0; k < kNeqThreads / 32; k++) s += sm[k][threadIdx.x];
		partials[(size_t)blockIdx.x * kPartialCols + threadIdx.x] = s;
	}
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned int t = atomicAdd(ticket, 1u);
		is_last = (t == gridDim.x - 1);
	}
	__syncthreads();
	if (!is_last) return;
	__threadfence();
	__shared__ double tot[kPartialCols];
	for (int col = wid; col < kPartialCols; col += kNeqThreads / 32) {
		double s = 0;
		for (unsigned int b = lane; b < gridDim.x; b += 32) s += __ldcg(partials + (size_t)b * kPartialCols + col);
		s = warp_sum(s);
		if (lane == 0) tot[col] = s;
	}
	__syncthreads();
	if (wid == 0) {
		if (lane == 0) *ticket = 0;
		neq_tail_warp(tot, fin, lane);
	}
}

} /* namespace m3d */
