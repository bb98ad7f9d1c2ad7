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

namespace m3d {

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

/* ---- AoS (reference 40-byte point) <-> SoA ----------------------------------------------------------- */
__global__ void k_unpack_points(const m3dreg_point *__restrict__ in, int n, float4 *__restrict__ xyzl, float4 *__restrict__ nrm)
{
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
	if (threadIdx.x < 3) bounds[threadIdx.x] = 0xFFFFFFFFu;
	else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;
}

/* Rigid transform of an AoS cloud, bit-compatible with the reference's device kernel (lesson_16.cu:1341-1367). */
__global__ void k_transform_aos(const m3dreg_point *__restrict__ in, m3dreg_point *__restrict__ out, int n,
		float r00, float r01, float r02, float t0, float r10, float r11, float r12, float t1,
		float r20, float r21, float r22, float t2)
{
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
	float r00 = __ldg(m + 0), r01 = __ldg(m + 1), r02 = __ldg(m + 2), t0 = __ldg(m + 3);
	float r10 = __ldg(m + 4), r11 = __ldg(m + 5), r12 = __ldg(m + 6), t1 = __ldg(m + 7);
	float r20 = __ldg(m + 8), r21 = __ldg(m + 9), r22 = __ldg(m + 10), t2 = __ldg(m + 11);
	float mnx = INFINITY, mny = INFINITY, mnz = INFINITY, mxx = -INFINITY, mxy = -INFINITY, mxz = -INFINITY;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		float4 p = __ldg(in_xyzl + i);
		float x = __fadd_rn(t0, __fmaf_rn(r02, p.z, __fmaf_rn(r00, p.x, __fmul_rn(r01, p.y))));
		float y = __fadd_rn(t1, __fmaf_rn(r12, p.z, __fmaf_rn(r10, p.x, __fmul_rn(r11, p.y))));
		float z = __fadd_rn(t2, __fmaf_rn(r22, p.z, __fmaf_rn(r20, p.x, __fmul_rn(r21, p.y))));
		out_xyzl[i] = make_float4(x, y, z, p.w);
		if (out_nrm) {      /* the fused loop rotates only the candidates' normals (k_build_candidates), not all of them */
			float4 q = __ldg(in_nrm + i);
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
	if (WITH_BOUNDS) block_bounds_commit(mnx, mny, mnz, mxx, mxy, mxz, bounds);
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
	bool ok = !(nbx <= 0 || nby <= 0 || nbz <= 0 || nb > 2147483647LL || nb > bucket_cap);
	g.number_of_buckets = ok ? nb : 0;
	return ok;
}

__device__ __forceinline__ int cell_of(float v, float mn, float res)
{
	return (int)__fdiv_rn(__fsub_rn(v, mn), res);     /* sub.f32, div.rn.f32, cvt.rzi.s32.f32 */
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
__global__ void __launch_bounds__(kSortThreads) k_grid_head(const float4 *__restrict__ xyzl, int n, const uint32_t *__restrict__ bounds,
		float rx, float ry, float rz, float ext, long long bucket_cap, m3dreg_grid_params *__restrict__ gp_out, int *__restrict__ flags,
		unsigned int *__restrict__ cell_count, m3dreg_bucket *__restrict__ buckets, uint32_t *__restrict__ keys,
		int tiles, int passes, uint32_t *__restrict__ hist)
{
	__shared__ uint32_t sh[kRadixSize];
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
#pragma unroll
	for (int j = 0; j < ITEMS; j++) {
		int i = base + j * kSortThreads + threadIdx.x;
		int bin = -1;
		if (i < n) {
			float4 p = __ldg(xyzl + i);
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
	if (gp && gp->number_of_buckets <= 0) return;
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int nround = (n + 31) & ~31;      /* whole warps stay together for the ballots below */
	const uint32_t k0 = __ldg(keys), k1 = n > 1 ? __ldg(keys + 1) : k0;
	const bool has_quirk = n > 1 && k0 != k1;
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nround; p += gridDim.x * blockDim.x) {
		bool listed = false, valid = p < n;
		uint32_t k = 0xFFFFFFFFu;
		bool quirk = false;
		if (valid) {
			k = __ldg(keys + p);
			if (table_out) {
				m3dreg_hash_element h;
				h.index_of_point = (int)__ldg(vals + p);
				h.index_of_bucket = (int)k;
				table_out[p] = h;
			}
			quirk = has_quirk && k == k1;
			bool run_start = (p == 0) || (__ldg(keys + p - 1) != k);
			bool run_end = (p == n - 1) || (__ldg(keys + p + 1) != k);
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

/* Same list from an externally supplied bucket table (stage-level m3dreg_nn_search). */
__global__ void k_list_cells(const uint32_t *__restrict__ keys, int n, const m3dreg_bucket *__restrict__ buckets,
		uint32_t *__restrict__ cell_list, unsigned int *__restrict__ cell_count)
{
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
 * the compact arrays (a bucket's candidates always fit inside its own range, so no prefix sum is needed) and —
 * because the search result is the lexicographic minimum of (dist, sorted position), which is independent of the
 * order in which candidates are looked at — stores them in a SPATIAL order: counting sort by
 * (label & 3, Morton code of the 4x4x4 sub-cell inside the bucket).  Every 8 consecutive candidates form a block
 * with an axis-aligned bounding box and a label mask; the search only touches blocks whose box can contain
 * something closer than the current best.  Each record keeps its sorted position l for the tie-break and result.
 * Two sets exist when the INNER and OUTER caps differ (different strides). */
struct CandSet {
	float4 *xyzl;     /* {x, y, z, label bits}                               */
	float4 *nrm;      /* {nx, ny, nz, bits of l = position in the sorted table (hashElement index)} */
	float4 *mlo;      /* per block: {min x, min y, min z, label mask bits}   */
	float4 *mhi;      /* per block: {max x, max y, max z, 0}                 */
};

constexpr int kCandBlock = 8;
constexpr int kBuildWarps = 4;
constexpr int kBuildBins = 256;
constexpr int kBuildPerLane = 4;   /* candidates per lane in flight (loads issued together) */

__device__ __forceinline__ int candidate_stride(int npts, int cap)
{
	int iter = 1;
	if (cap < npts) { iter = npts / cap; if (iter <= 0) iter = 1; }
	return iter;
}

__device__ __forceinline__ uint32_t label_bit(int label) { return 1u << (label & 31); }

struct CellFrame { float ox, oy, oz, sx, sy, sz; };   /* bucket origin and 4/resolution (layout only, not parity relevant) */

__device__ __forceinline__ uint32_t cand_bin(const float4 &p, const CellFrame &f)
{
	int ux = min(3, max(0, (int)((p.x - f.ox) * f.sx)));
	int uy = min(3, max(0, (int)((p.y - f.oy) * f.sy)));
	int uz = min(3, max(0, (int)((p.z - f.oz) * f.sz)));
	uint32_t mx = (ux & 1) | ((ux & 2) << 2), my = (uy & 1) | ((uy & 2) << 2), mz = (uz & 1) | ((uz & 2) << 2);
	return ((uint32_t)(__float_as_int(p.w) & 3) << 6) | mx | (my << 1) | (mz << 2);
}

struct NormalRotation { float r[9]; int on; };   /* rotation applied to the candidates' normals (same rounding as the transform kernels) */

__device__ __forceinline__ void build_cell_candidates(const uint32_t *__restrict__ vals, const float4 *__restrict__ src_xyzl,
		const float4 *__restrict__ src_nrm, const NormalRotation &rot, int begin, int npts, int cap, const CellFrame &f, const CandSet &set,
		uint32_t *hist, int lane)
{
	const unsigned full = 0xffffffffu;
	if (npts <= 0) return;
	if (cap <= 0) {    /* nothing of this bucket may be looked at in this role */
		if (lane == 0) set.mhi[begin] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
		return;
	}
	const int iter = candidate_stride(npts, cap);
	const int ncand = (npts + iter - 1) / iter;
	const uint32_t lt = (1u << lane) - 1u;
	/* pass 1: bin histogram */
#pragma unroll
	for (int k = 0; k < kBuildBins / 32; k++) hist[lane + 32 * k] = 0;
	__syncwarp();
	for (int c0 = 0; c0 < ncand; c0 += 32 * kBuildPerLane) {
		float4 p[kBuildPerLane];
#pragma unroll
		for (int k = 0; k < kBuildPerLane; k++) {
			int cc = c0 + k * 32 + lane;
			if (cc < ncand) p[k] = __ldg(src_xyzl + __ldg(vals + begin + cc * iter));
		}
#pragma unroll
		for (int k = 0; k < kBuildPerLane; k++) {
			int cc = c0 + k * 32 + lane;
			if (c0 + k * 32 >= ncand) break;
			bool valid = cc < ncand;
			uint32_t bin = valid ? cand_bin(p[k], f) : (0x100u + lane);
			uint32_t peers = __match_any_sync(full, bin);
			if (valid && lane == __ffs(peers) - 1) hist[bin] += __popc(peers);
			__syncwarp();
		}
	}
	/* exclusive scan of the 256 bins: 8 consecutive bins per lane */
	{
		uint32_t v[8], sum = 0;
#pragma unroll
		for (int k = 0; k < 8; k++) { v[k] = hist[lane * 8 + k]; sum += v[k]; }
		uint32_t incl = sum;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			uint32_t t = __shfl_up_sync(full, incl, o);
			if (lane >= o) incl += t;
		}
		uint32_t run = incl - sum;
		__syncwarp();
#pragma unroll
		for (int k = 0; k < 8; k++) { hist[lane * 8 + k] = run; run += v[k]; }
		__syncwarp();
	}
	/* pass 2: stable placement (ascending sorted position inside a bin) */
	for (int c0 = 0; c0 < ncand; c0 += 32 * kBuildPerLane) {
		float4 p[kBuildPerLane], nr[kBuildPerLane];
		int ls[kBuildPerLane];
#pragma unroll
		for (int k = 0; k < kBuildPerLane; k++) {
			int cc = c0 + k * 32 + lane;
			if (cc < ncand) {
				ls[k] = begin + cc * iter;
				uint32_t v = __ldg(vals + ls[k]);
				p[k] = __ldg(src_xyzl + v);
				nr[k] = __ldg(src_nrm + v);
			}
		}
#pragma unroll
		for (int k = 0; k < kBuildPerLane; k++) {
			int cc = c0 + k * 32 + lane;
			if (c0 + k * 32 >= ncand) break;
			bool valid = cc < ncand;
			uint32_t bin = valid ? cand_bin(p[k], f) : (0x100u + lane);
			uint32_t peers = __match_any_sync(full, bin);
			int leader = __ffs(peers) - 1;
			uint32_t old = 0;
			if (valid && lane == leader) { old = hist[bin]; hist[bin] = old + __popc(peers); }
			old = __shfl_sync(full, old, leader);
			if (valid) {
				int pos = begin + (int)(old + __popc(peers & lt));
				set.xyzl[pos] = p[k];
				float4 nn = nr[k];
				if (rot.on) {
					nn.x = __fmaf_rn(rot.r[2], nr[k].z, __fmaf_rn(rot.r[0], nr[k].x, __fmul_rn(rot.r[1], nr[k].y)));
					nn.y = __fmaf_rn(rot.r[5], nr[k].z, __fmaf_rn(rot.r[3], nr[k].x, __fmul_rn(rot.r[4], nr[k].y)));
					nn.z = __fmaf_rn(rot.r[8], nr[k].z, __fmaf_rn(rot.r[6], nr[k].x, __fmul_rn(rot.r[7], nr[k].y)));
				}
				set.nrm[pos] = make_float4(nn.x, nn.y, nn.z, __int_as_float(ls[k]));
			}
			__syncwarp();
		}
	}
	__syncwarp();
	/* block boxes + label masks (re-read through L2: the records were written by other lanes of this warp) */
	const int nblocks = (ncand + kCandBlock - 1) / kCandBlock;
	for (int j = lane; j < nblocks; j += 32) {
		float lox = INFINITY, loy = INFINITY, loz = INFINITY, hix = -INFINITY, hiy = -INFINITY, hiz = -INFINITY;
		uint32_t mask = 0;
#pragma unroll
		for (int k = 0; k < kCandBlock; k++) {
			int idx = j * kCandBlock + k;
			if (idx < ncand) {
				float4 c = __ldcg(set.xyzl + begin + idx);
				lox = fminf(lox, c.x); loy = fminf(loy, c.y); loz = fminf(loz, c.z);
				hix = fmaxf(hix, c.x); hiy = fmaxf(hiy, c.y); hiz = fmaxf(hiz, c.z);
				mask |= label_bit(__float_as_int(c.w));
			}
		}
		set.mlo[begin + j] = make_float4(lox, loy, loz, __uint_as_float(mask));
		set.mhi[begin + j] = make_float4(hix, hiy, hiz, __int_as_float(j == 0 ? ncand : 0));   /* the first header carries the candidate count */
	}
}

/* One warp per searchable bucket of the compact list k_finalize_grid / k_list_cells left behind. */
__global__ void __launch_bounds__(kBuildWarps * 32, 4) k_build_candidates(const uint32_t *__restrict__ vals,
		const m3dreg_grid_params *__restrict__ gp, const m3dreg_bucket *__restrict__ buckets,
		const uint32_t *__restrict__ cell_list, const unsigned int *__restrict__ cell_count,
		const float4 *__restrict__ src_xyzl, const float4 *__restrict__ src_nrm, const float *__restrict__ nrm_m, int max_inner, int max_outer,
		CandSet ci, CandSet co, int two_sets)
{
	__shared__ uint32_t s_hist[kBuildWarps][kBuildBins];
	NormalRotation rot;
	rot.on = nrm_m != nullptr;
#pragma unroll
	for (int k = 0; k < 9; k++) rot.r[k] = rot.on ? __ldg(nrm_m + (k / 3) * 4 + (k % 3)) : 0.0f;    /* row-major 4x4 */
	if (gp->number_of_buckets <= 0) return;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int nby = gp->number_of_buckets_Y, nbz = gp->number_of_buckets_Z;
	const float mnx = gp->bounding_box_min_X, mny = gp->bounding_box_min_Y, mnz = gp->bounding_box_min_Z;
	const float rx = gp->resolution_X, ry = gp->resolution_Y, rz = gp->resolution_Z;
	const unsigned int ncells = *cell_count;
	const unsigned int nwarps = gridDim.x * kBuildWarps;
	for (unsigned int t = blockIdx.x * kBuildWarps + w; t < ncells; t += nwarps) {
		int c = (int)__ldg(cell_list + t);
		const int *bp = reinterpret_cast<const int *>(buckets + c);
		int c_begin = __ldg(bp), c_n = __ldg(bp + 2);
		int ix = c / (nby * nbz), iy = (c / nbz) % nby, iz = c % nbz;
		CellFrame f;
		f.ox = mnx + (float)ix * rx; f.oy = mny + (float)iy * ry; f.oz = mnz + (float)iz * rz;
		f.sx = 4.0f / rx; f.sy = 4.0f / ry; f.sz = 4.0f / rz;
		build_cell_candidates(vals, src_xyzl, src_nrm, rot, c_begin, c_n, max_inner, f, ci, s_hist[w], lane);
		if (two_sets) build_cell_candidates(vals, src_xyzl, src_nrm, rot, c_begin, c_n, max_outer, f, co, s_hist[w], lane);
	}
}

/* split an externally supplied reference-layout table into key / value streams (stage-level m3dreg_nn_search) */
__global__ void k_split_table(const m3dreg_hash_element *__restrict__ table, int n, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
		m3dreg_hash_element h = table[p];
		keys[p] = (uint32_t)h.index_of_bucket;
		vals[p] = (uint32_t)h.index_of_point;
	}
}

/* ---- semantic nearest neighbour (kernel_semanticNearestNeighborSearch, lesson_16.cu:531-703) ------------- */

/* The reference's angle gate (lesson_16.cu:666-676): acos(dot)*180.0f/M_PI, |.| < 90.0f, with acos the CUDA
 * float acosf.  As a function of the f32 dot product its accepted set is exactly the interval
 *      0x328885AC (1.589327e-08) <= dot <= 1.0f
 * (NaN, |dot| > 1, zero and negative dots are rejected; tiny positive dots still round to >= 90.0f degrees).
 * tests/test_gpu_gate.py proves this equal to the upstream expression for ALL 2^32 float bit patterns on the
 * device, so the two compares below are bit-exact and ~60 instructions (acosf + an f64 divide) cheaper. */
constexpr uint32_t kAngleGateMinBits = 0x328885ACu;
__device__ __forceinline__ bool angle_gate(float dot)
{
	return dot >= __uint_as_float(kAngleGateMinBits) && dot <= 1.0f;
}

struct NNQuery {
	float x, y, z, nx, ny, nz, r2;
	int label;
	float best;
	int best_l;
};

/* Visit order in the reference is ascending sorted position l (cells are visited in ascending linear index and
 * the table is sorted by it) and its update is a strict `<`, so the reference result is the lexicographic minimum
 * of (dist, l) over the admissible candidates.  Keeping that pair lets candidates be looked at in ANY order, lets
 * whole groups be skipped when provably useless, and makes evaluating extra candidates harmless — which is what
 * allows a whole warp to walk one candidate list together.
 *
 * Warp-cooperative search: every lane holds one query; the warp stages up to 4 candidate blocks (32 records) in its
 * 512-byte slice of shared memory with one coalesced LDG.128 per lane.
 *   HOT  loop (straight-line): broadcast LDS.128 + 3 FADD + FMUL + 2 FFMA, then a branch-free running minimum over
 *        the candidates whose label matches, plus a flag that records an exact tie at the minimum.
 *   COLD step (once per staged group): if the group minimum can beat min(best, r^2), fetch that candidate's sorted
 *        position and normal, apply the angle gate and the exact (dist, l) comparison.  Should the gate reject it, or
 *        a tie have been seen (staging order is spatial, not ascending l), the lane re-scans the group sequentially
 *        with the full predicate. */
__device__ __forceinline__ float nn_dist(float qx, float qy, float qz, const float4 &c)
{
	float dx = __fsub_rn(qx, c.x), dy = __fsub_rn(qy, c.y), dz = __fsub_rn(qz, c.z);
	return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ void nn_group_min(const NNQuery &q, const float4 *stage, int base, float &bmin, int &bidx, bool &tie)
{
#pragma unroll
	for (int kk = 0; kk < kCandBlock; kk++) {
		float4 c = stage[base + kk];
		float d = nn_dist(q.x, q.y, q.z, c);
		bool ok = __float_as_int(c.w) == q.label;
		bool better = ok && (d < bmin);
		bool eq = ok && (d == bmin);
		tie = better ? false : (tie || eq);
		bmin = better ? d : bmin;
		bidx = better ? base + kk : bidx;
	}
}

/* Stage the blocks b[0..nsel) (block indices inside the bucket) and evaluate them for every lane. */
__device__ __forceinline__ void nn_eval_blocks(NNQuery &q, bool need, const CandSet &set, int begin, int ncand,
		int b0, int b1, int b2, int b3, int nsel, float4 *stage, int lane, unsigned int &evals)
{
	const int g = lane >> 3;
	const int blk = g == 0 ? b0 : (g == 1 ? b1 : (g == 2 ? b2 : b3));
	const int idx = blk * kCandBlock + (lane & 7);
	const bool valid = g < nsel && idx < ncand;
	__syncwarp();
	stage[lane] = valid ? __ldg(set.xyzl + begin + idx) : make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(0x7fffffff));
	__syncwarp();
	evals += (unsigned int)(nsel * kCandBlock);
	float bmin = INFINITY;
	int bidx = -1;
	bool tie = false;
#pragma unroll 1
	for (int g8 = 0; g8 < nsel; g8++) nn_group_min(q, stage, g8 * kCandBlock, bmin, bidx, tie);
	/* cold step */
	if (need && bidx >= 0 && bmin <= fminf(q.best, q.r2)) {
		bool rescan = tie;
		if (!rescan) {
			int gg = bidx >> 3;
			int cand = begin + (gg == 0 ? b0 : (gg == 1 ? b1 : (gg == 2 ? b2 : b3))) * kCandBlock + (bidx & 7);
			float4 cn = __ldg(set.nrm + cand);
			int l = __float_as_int(cn.w);
			if (bmin < q.best || (bmin == q.best && l < q.best_l)) {
				float dot = __fmaf_rn(q.nz, cn.z, __fmaf_rn(q.nx, cn.x, __fmul_rn(q.ny, cn.y)));
				if (angle_gate(dot)) { q.best = bmin; q.best_l = l; }
				else rescan = true;     /* group minimum inadmissible (rare: opposite faces of thin structures) */
			}
		}
		if (rescan) {
			for (int kk = 0; kk < nsel * kCandBlock; kk++) {
				float4 c = stage[kk];
				float d = nn_dist(q.x, q.y, q.z, c);
				if (__float_as_int(c.w) == q.label && d <= q.r2 && d <= q.best) {
					int gg = kk >> 3;
					int cand = begin + (gg == 0 ? b0 : (gg == 1 ? b1 : (gg == 2 ? b2 : b3))) * kCandBlock + (kk & 7);
					float4 cn2 = __ldg(set.nrm + cand);
					int l2 = __float_as_int(cn2.w);
					if (d < q.best || l2 < q.best_l) {
						float dot2 = __fmaf_rn(q.nz, cn2.z, __fmaf_rn(q.nx, cn2.x, __fmul_rn(q.ny, cn2.y)));
						if (angle_gate(dot2)) { q.best = d; q.best_l = l2; }
					}
				}
			}
		}
	}
}

/* Search one bucket's candidate set for the lanes flagged `need`.
 * Buckets with <= 4 blocks are staged whole.  Larger ones are searched in rounds of growing radius rho: the warp
 * forms the bounding box of the queries that are still UNSETTLED (min(best, r^2) > rho^2 of the previous round),
 * inflates it by rho, and lane j tests block j's box (and label mask) against it; blocks that pass and were not yet
 * staged are evaluated by all lanes.  Exactness: a candidate c with dist(q, c) <= lim_q has |q.x - c.x| <= sqrt(lim_q)
 * up to two float roundings (dist >= fl(dx*dx) by monotonicity of fma/mul), so with R = sqrt_ru(rho^2) * (1 + 2^-20),
 * box bounds rounded outwards and lim_q <= rho^2, c lies inside the inflated box and its block's box overlaps it.
 * A lane whose limit does not exceed rho^2 therefore has nothing left to find in this bucket; the others go on to
 * the next round (rho doubles) until rho^2 covers every remaining limit. */
__device__ __forceinline__ void nn_visit_set(NNQuery &q, bool need, int begin, const CandSet &set,
		float trial_unit2, int prune, float4 *stage, int lane, unsigned int &evals)
{
	const unsigned full = 0xffffffffu;
	if (!__any_sync(full, need)) return;
	/* block headers of the first 32 blocks (lanes beyond the bucket's block count read slack that is never used);
	 * the candidate count rides in the first header */
	float4 mlo = __ldg(set.mlo + begin + lane), mhi = __ldg(set.mhi + begin + lane);
	const int ncand = __float_as_int(__shfl_sync(full, mhi.w, 0));
	if (ncand <= 0) return;
	const int nblocks = (ncand + kCandBlock - 1) / kCandBlock;
	const bool flat = nblocks <= 4 || !prune;
	const float trial2 = trial_unit2 / (float)ncand;
	const uint32_t ox = f2o(q.x), oy = f2o(q.y), oz = f2o(q.z);
	const uint32_t lbit = label_bit(q.label);
#pragma unroll 1
	for (int cb = 0; cb < nblocks; cb += 32) {
		const int nblk = min(32, nblocks - cb);
		if (cb > 0 && !flat && lane < nblk) { mlo = __ldg(set.mlo + begin + cb + lane); mhi = __ldg(set.mhi + begin + cb + lane); }
		unsigned staged = 0;
		bool U = need;
		float rho2 = -1.0f;
#pragma unroll 1
		for (;;) {
			unsigned m;
			float maxlim = 0.0f;
			if (flat) {
				m = nblk == 32 ? 0xffffffffu : ((1u << nblk) - 1u);
			} else {
				/* non-negative floats order like their bit patterns */
				unsigned lv = U ? __float_as_uint(fminf(q.best, q.r2)) + 1u : 0u;
				unsigned mv = __reduce_max_sync(full, lv);
				if (mv == 0u) break;
				maxlim = __uint_as_float(mv - 1u);
				rho2 = rho2 < 0.0f ? fminf(maxlim, trial2) : fminf(maxlim, rho2 * 4.0f);
				float R = __fmul_ru(__fsqrt_ru(rho2), 1.00000095367431640625f);
				float bxl = __fsub_rd(o2f(__reduce_min_sync(full, U ? ox : 0xFFFFFFFFu)), R);
				float byl = __fsub_rd(o2f(__reduce_min_sync(full, U ? oy : 0xFFFFFFFFu)), R);
				float bzl = __fsub_rd(o2f(__reduce_min_sync(full, U ? oz : 0xFFFFFFFFu)), R);
				float bxh = __fadd_ru(o2f(__reduce_max_sync(full, U ? ox : 0u)), R);
				float byh = __fadd_ru(o2f(__reduce_max_sync(full, U ? oy : 0u)), R);
				float bzh = __fadd_ru(o2f(__reduce_max_sync(full, U ? oz : 0u)), R);
				unsigned labels = __reduce_or_sync(full, U ? lbit : 0u);
				bool hit = lane < nblk && (__float_as_uint(mlo.w) & labels) &&
						!(mlo.x > bxh || mhi.x < bxl || mlo.y > byh || mhi.y < byl || mlo.z > bzh || mhi.z < bzl);
				m = __ballot_sync(full, hit) & ~staged;
				staged |= m;
			}
#pragma unroll 1
			while (m) {
				int b0 = __ffs(m) - 1; m &= m - 1;
				int nsel = 1, b1 = 0, b2 = 0, b3 = 0;
				if (m) { b1 = __ffs(m) - 1; m &= m - 1; nsel = 2; }
				if (m) { b2 = __ffs(m) - 1; m &= m - 1; nsel = 3; }
				if (m) { b3 = __ffs(m) - 1; m &= m - 1; nsel = 4; }
				nn_eval_blocks(q, need, set, begin, ncand, cb + b0, cb + b1, cb + b2, cb + b3, nsel, stage, lane, evals);
			}
			if (flat || rho2 >= maxlim) break;
			U = U && (fminf(q.best, q.r2) > rho2);
		}
	}
}

/* Conservative per-axis gap between the query and the slab of cells at offset -1 / +1 (rounded DOWN to float).
 * A point stored in cell c satisfies trunc(fl(fl(v-min)/res)) == c; with two roundings of relative error 2^-24,
 * (v-min) < ix*res*(1+2^-21) for cells <= ix-1 and (v-min) >= (ix+1)*res*(1-2^-21) for cells >= ix+1.
 * A factor 2^-20 is used.  fl(q - v) is the correctly rounded true difference, rounding is monotone and the gap
 * is a float, so |fl(q-v)| >= gap, hence fma(gz,gz,fma(gx,gx,gy*gy)) <= the reference's dist for every
 * candidate of that cell: skipping a cell whose bound exceeds the current best (or r^2) cannot change the result. */
__device__ __forceinline__ void axis_gaps(float q, float mn, float res, int ic, float &g_lo, float &g_hi)
{
	const float up_f = 1.00000095367431640625f, dn_f = 0.99999904632568359375f;     /* 1 +- 2^-20 */
	float up = __fmul_ru(__fmul_ru((float)ic, res), up_f);            /* exclusive upper bound of cells <= ic-1, rounded up   */
	float lo = __fmul_rd(__fmul_rd((float)(ic + 1), res), dn_f);      /* inclusive lower bound of cells >= ic+1, rounded down */
	g_lo = fmaxf(0.0f, __fsub_rd(__fsub_rd(q, mn), up));
	g_hi = fmaxf(0.0f, __fsub_rd(lo, __fsub_ru(q, mn)));
}

struct NNLane {          /* per-lane search state besides the query itself */
	int home, ix, iy, iz;
	float gxl, gxh, gyl, gyh, gzl, gzh;
	unsigned done;       /* bit o = (dx+1)*9 + (dy+1)*3 + (dz+1): neighbour bucket already handled (or pruned) */
};

struct NNGrid {
	long long nb;
	int nbx, nby, nbz;
	const m3dreg_bucket *buckets;
	CandSet ci, co;
	int max_inner, max_outer, two_sets, prune;
	float trial_unit2;
};

/* Bucket (cx,cy,cz) for EVERY lane that has it in its 27-neighbourhood and has not handled it yet — whatever the
 * lane's home bucket is, so a bucket is staged once per warp, not once per group of equal homes. */
__device__ __forceinline__ void nn_visit_cell(NNQuery &q, NNLane &s, const NNGrid &G, int cx, int cy, int cz,
		float4 *stage, int lane, unsigned int &evals)
{
	int dx = cx - s.ix, dy = cy - s.iy, dz = cz - s.iz;
	bool in27 = s.home >= 0 && (unsigned)(dx + 1) <= 2u && (unsigned)(dy + 1) <= 2u && (unsigned)(dz + 1) <= 2u;
	int o = (dx + 1) * 9 + (dy + 1) * 3 + (dz + 1);
	bool part = in27 && !((s.done >> (o & 31)) & 1u);
	if (part) s.done |= 1u << o;
	float gx = dx < 0 ? s.gxl : (dx > 0 ? s.gxh : 0.0f), gy = dy < 0 ? s.gyl : (dy > 0 ? s.gyh : 0.0f), gz = dz < 0 ? s.gzl : (dz > 0 ? s.gzh : 0.0f);
	float lbd = __fmaf_rn(gz, gz, __fmaf_rn(gx, gx, __fmul_rn(gy, gy)));
	bool need = part && (!G.prune || !(lbd > q.best || lbd > q.r2));
	if (!__any_sync(0xffffffffu, need)) return;
	int cell = (cx * G.nby + cy) * G.nbz + cz;
	int begin = __ldg(reinterpret_cast<const int *>(G.buckets + cell));
	if (begin < 0) return;       /* empty bucket, or the quirk bucket the reference cannot see (number_of_points == 0) */
	const bool inner = (o == 13);
#pragma unroll 1
	for (int pass = 0; pass < (G.two_sets ? 2 : 1); pass++) {
		/* equal caps: one shared set for every role; different caps: INNER set for the home role, OUTER for the rest */
		bool mine = G.two_sets ? (need && (pass == 0 ? inner : !inner)) : need;
		nn_visit_set(q, mine, begin, pass == 0 ? G.ci : G.co, G.trial_unit2, G.prune, stage, lane, evals);
	}
}

/* Semantic NN, one query per lane, candidate groups shared by the warp.
 * Queries are expected in a spatially coherent order (the scan store keeps a (label, Morton)-sorted copy of every
 * scan, and a rigid transform preserves coherence), so the 32 queries of a warp cover a small patch of one surface.
 * Phase A visits every distinct home bucket of the warp; phase B the neighbour buckets some lane cannot exclude by
 * the lower bound above with its best-so-far.
 * q_perm (may be null = identity) maps the query's position to its index in the caller's order: nn_out is written
 * in the caller's order (the reference's layout), nn_seq (may be null) in query-array order for the next stage. */
constexpr int kNNThreads = 128;

__global__ void __launch_bounds__(kNNThreads, 6) k_nn_search(const float4 *__restrict__ q_xyzl, const float4 *__restrict__ q_nrm,
		const uint32_t *__restrict__ q_perm, int n_second, CandSet ci, CandSet co,
		const uint32_t *__restrict__ s_vals, int n_first,
		const m3dreg_bucket *__restrict__ buckets, const m3dreg_grid_params *__restrict__ gp,
		float search_radius, int max_inner, int max_outer, int prune,
		int *__restrict__ nn_out, int *__restrict__ nn_seq, unsigned long long *__restrict__ label_counts,
		unsigned long long *__restrict__ eval_counter)
{
	__shared__ float4 s_stage[kNNThreads / 32][32];
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	float4 *stage = s_stage[threadIdx.x >> 5];
	unsigned int evals = 0;
	int qi = blockIdx.x * blockDim.x + threadIdx.x;
	NNGrid G;
	G.nb = gp->number_of_buckets;
	G.nbx = gp->number_of_buckets_X; G.nby = gp->number_of_buckets_Y; G.nbz = gp->number_of_buckets_Z;
	G.buckets = buckets; G.ci = ci; G.co = co;
	G.max_inner = max_inner; G.max_outer = max_outer; G.two_sets = (max_inner != max_outer); G.prune = prune;
	float mnx = gp->bounding_box_min_X, mny = gp->bounding_box_min_Y, mnz = gp->bounding_box_min_Z;
	float rx = gp->resolution_X, ry = gp->resolution_Y, rz = gp->resolution_Z;
	G.trial_unit2 = 1.0f * fmaxf(rx, fmaxf(ry, rz)) * fmaxf(rx, fmaxf(ry, rz));   /* first-round radius^2 = res^2 / candidates: about one candidate spacing */
	NNQuery q;
	q.x = q.y = q.z = q.nx = q.ny = q.nz = 0.0f;
	q.r2 = __fmul_rn(search_radius, search_radius);
	q.label = -1;
	q.best = 100000000.0f;
	q.best_l = 0x7fffffff;
	NNLane s;
	s.home = -1; s.ix = s.iy = s.iz = 0; s.done = 0;
	s.gxl = s.gxh = s.gyl = s.gyh = s.gzl = s.gzh = 0.0f;
	if (qi < n_second && G.nb > 0) {
		float4 p = __ldg(q_xyzl + qi), pn = __ldg(q_nrm + qi);
		q.x = p.x; q.y = p.y; q.z = p.z; q.nx = pn.x; q.ny = pn.y; q.nz = pn.z;
		q.label = __float_as_int(p.w);
		bool inside = !(p.x < mnx || p.x > gp->bounding_box_max_X) && !(p.y < mny || p.y > gp->bounding_box_max_Y) &&
				!(p.z < mnz || p.z > gp->bounding_box_max_Z);
		if (inside) {
			s.ix = cell_of(p.x, mnx, rx); s.iy = cell_of(p.y, mny, ry); s.iz = cell_of(p.z, mnz, rz);
			int h = s.ix * G.nby * G.nbz + s.iy * G.nbz + s.iz;
			if (h >= 0 && (long long)h < G.nb) s.home = h;
		}
	}
	/* per-axis gaps to the slabs at offset -1 / +1 (0 for the own slab) */
	if (prune && s.home >= 0) {
		axis_gaps(q.x, mnx, rx, s.ix, s.gxl, s.gxh);
		axis_gaps(q.y, mny, ry, s.iy, s.gyl, s.gyh);
		axis_gaps(q.z, mnz, rz, s.iz, s.gzl, s.gzh);
	}
	/* phase A: every distinct home bucket of the warp; phase B: the neighbour buckets that can still matter with the
	 * limits phase A left behind.  One visit call site keeps the kernel small enough for the instruction cache. */
	unsigned remaining = __ballot_sync(full, s.home >= 0);
	unsigned todo = 0;
	bool phase_b = false;
#pragma unroll 1
	for (;;) {
		int cx, cy, cz;
		if (remaining) {
			int leader = __ffs(remaining) - 1;
			int h = __shfl_sync(full, s.home, leader);
			cx = __shfl_sync(full, s.ix, leader); cy = __shfl_sync(full, s.iy, leader); cz = __shfl_sync(full, s.iz, leader);
			remaining &= ~__ballot_sync(full, s.home == h);
		} else {
			if (!phase_b) {
				phase_b = true;
				/* a neighbour can only matter if one of the six face gaps is within the limit: usually none is */
				float lim = fminf(q.best, q.r2);
				float gmin = fminf(fminf(fminf(s.gxl, s.gxh), fminf(s.gyl, s.gyh)), fminf(s.gzl, s.gzh));
				bool near_face = s.home >= 0 && (!prune || !(__fmul_rn(gmin, gmin) > lim));
				if (!__any_sync(full, near_face)) break;
				if (near_face) {
					/* superset of the buckets worth a visit (sum of squared gaps with a reassociation margin; the exact
					 * bound is re-checked when the bucket is visited); out-of-grid neighbours get an infinite gap */
					const float limm = prune ? __fmul_ru(lim, 1.00000095367431640625f) : 3.0e38f;     /* finite: out-of-grid stays excluded */
					float sx[3], sy[3], sz[3];
					sx[0] = s.ix > 0 ? __fmul_rd(s.gxl, s.gxl) : INFINITY; sx[1] = 0.0f; sx[2] = s.ix + 1 < G.nbx ? __fmul_rd(s.gxh, s.gxh) : INFINITY;
					sy[0] = s.iy > 0 ? __fmul_rd(s.gyl, s.gyl) : INFINITY; sy[1] = 0.0f; sy[2] = s.iy + 1 < G.nby ? __fmul_rd(s.gyh, s.gyh) : INFINITY;
					sz[0] = s.iz > 0 ? __fmul_rd(s.gzl, s.gzl) : INFINITY; sz[1] = 0.0f; sz[2] = s.iz + 1 < G.nbz ? __fmul_rd(s.gzh, s.gzh) : INFINITY;
					if (!prune) {
#pragma unroll
						for (int a = 0; a < 3; a += 2) { if (sx[a] != INFINITY) sx[a] = 0.0f; if (sy[a] != INFINITY) sy[a] = 0.0f; if (sz[a] != INFINITY) sz[a] = 0.0f; }
					}
#pragma unroll
					for (int i = 0; i < 3; i++)
#pragma unroll
						for (int j = 0; j < 3; j++) {
							float sxy = __fadd_rd(sx[i], sy[j]);
#pragma unroll
							for (int k = 0; k < 3; k++) todo |= (__fadd_rd(sxy, sz[k]) <= limm) ? (1u << (i * 9 + j * 3 + k)) : 0u;
						}
					todo &= ~s.done;
				}
			}
			unsigned pending = __ballot_sync(full, todo != 0u);
			if (!pending) break;
			int leader = __ffs(pending) - 1;
			int o = __ffs(__shfl_sync(full, todo, leader)) - 1;
			cx = __shfl_sync(full, s.ix, leader) + (o / 9 - 1); cy = __shfl_sync(full, s.iy, leader) + ((o / 3) % 3 - 1);
			cz = __shfl_sync(full, s.iz, leader) + (o % 3 - 1);
		}
		nn_visit_cell(q, s, G, cx, cy, cz, stage, lane, evals);
		todo &= ~s.done;
	}
	if (eval_counter && lane == 0 && evals) atomicAdd(eval_counter, (unsigned long long)evals);
	int result = -1;
	if (q.best_l != 0x7fffffff && q.best_l < n_first) result = (int)__ldg(s_vals + q.best_l);
	if (qi < n_second) {
		if (nn_seq) nn_seq[qi] = result;
		nn_out[q_perm ? __ldg(q_perm + qi) : (uint32_t)qi] = result;
	}
	if (label_counts) {   /* per-label match counts (gpu6DSLAM.cpp:323-357): warp ballots, one atomic per label per warp */
		bool hit = qi < n_second && result >= 0;
#pragma unroll
		for (int L = 0; L < 4; L++) {
			unsigned m = __ballot_sync(full, hit && q.label == L);
			if (m && lane == L) atomicAdd(&label_counts[L], (unsigned long long)__popc(m));
		}
	}
}

/* gather a stored scan into query order: out[i] = in[perm[i]] */
__global__ void k_gather_perm(const uint32_t *__restrict__ perm, int n, const float4 *__restrict__ in_xyzl, const float4 *__restrict__ in_nrm,
		float4 *__restrict__ out_xyzl, float4 *__restrict__ out_nrm)
{
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

__device__ inline void pose_prepare_warp(PoseState *ps, int lane);

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
	if (blockIdx.x == 0 && threadIdx.x < 32) pose_prepare_warp(ps, threadIdx.x);
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

__device__ inline void euler_to_matrix_warp(const float *omfika, const float *xyz, float *m, int lane)
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
	if (lane == 0) {
		m[0] = __fsub_rn(1.0f, __fadd_rn(tyy, tzz)); m[1] = __fsub_rn(txy, twz); m[2] = __fadd_rn(txz, twy); m[3] = xyz[0];
		m[4] = __fadd_rn(txy, twz); m[5] = __fsub_rn(1.0f, __fadd_rn(txx, tzz)); m[6] = __fsub_rn(tyz, twx); m[7] = xyz[1];
		m[8] = __fsub_rn(txz, twy); m[9] = __fadd_rn(tyz, twx); m[10] = __fsub_rn(1.0f, __fadd_rn(txx, tyy)); m[11] = xyz[2];
		m[12] = 0.0f; m[13] = 0.0f; m[14] = 0.0f; m[15] = 1.0f;
	}
	__syncwarp();
}

/* m: 16 floats readable by every lane (shared or global, already visible); results returned in every lane. */
__device__ inline void matrix4_to_euler_warp(const float *m, float *omfika, float *xyz, int lane)
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

/* Start of an iteration (gpu6DSLAM.cpp:276-291) by one warp: Euler round trip of ps->m into pose1 / pose6. */
__device__ inline void pose_prepare_warp(PoseState *ps, int lane)
{
	float of[3], t[3];
	matrix4_to_euler_warp(ps->m, of, t, lane);
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

/* Per-observation sources for the moment reduction.  Three steps so that a thread can keep several observations in
 * flight: token(i) (first load), fetch(i, token, raw) (dependent gathers), finish(raw, ...) (arithmetic). */
struct ObsFromNN {   /* fused path: nn[] + clouds (gpu6DSLAM.cpp:323-398 done on the device) */
	const int *nn;
	const float4 *q_xyzl;       /* queries (second cloud, global)           */
	const float4 *g_xyzl;       /* first cloud, global, original order      */
	const float4 *l_xyzl;       /* first cloud, local, original order       */
	const unsigned long long *label_counts;
	float weight[4];
	struct Raw { float4 p2, p1, p0; };
	__device__ __forceinline__ int token(int i) const { return __ldg(nn + i); }
	__device__ __forceinline__ void fetch(int i, int j, Raw &r) const
	{
		r.p2 = __ldg(q_xyzl + i); r.p1 = __ldg(g_xyzl + j); r.p0 = __ldg(l_xyzl + j);
	}
	__device__ __forceinline__ void finish(const Raw &r, const float *wl, double &w, double &x, double &y, double &z,
			double &lx, double &ly, double &lz) const
	{
		int label = __float_as_int(r.p2.w);
		w = (label >= 0 && label < 4) ? (double)wl[label] : 0.0;
		x = r.p0.x; y = r.p0.y; z = r.p0.z;
		lx = (double)__fsub_rn(r.p1.x, r.p2.x); ly = (double)__fsub_rn(r.p1.y, r.p2.y); lz = (double)__fsub_rn(r.p1.z, r.p2.z);
	}
};

struct ObsFromList { /* stage-level path: the reference's obs_nn_t array */
	const m3dreg_obs_nn *obs;
	struct Raw { float v[7]; };
	__device__ __forceinline__ int token(int) const { return 0; }
	__device__ __forceinline__ void fetch(int i, int, Raw &r) const
	{
		const float *o = reinterpret_cast<const float *>(obs + i);
#pragma unroll
		for (int k = 0; k < 7; k++) r.v[k] = __ldg(o + k);
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
};

/* What warp 0 of the last block does with the finished 28-double system `neq` (shared memory): publish / accumulate
 * it, and (fused loop) gate on the observation count, Cholesky, pose update, Euler round trip for the next iteration,
 * resets. */
__device__ inline void neq_tail_warp(const double *neq, const FinalizeArgs &fin, int lane)
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
		for (int k = 0; k < 6; k++) p6[k] = ps->pose6[k];
		__syncwarp();
		if (status == 0) {
			/* registerLS tail (cudaWrapper.cpp:574-579 / 641-646) + EulerToMatrix (gpu6DSLAM.cpp:408-413) */
			p6[0] += x[0]; p6[1] += x[1]; p6[2] += x[2];
			if (fin.dof == 6) { p6[3] += x[3]; p6[4] += x[4]; p6[5] += x[5]; }
			else p6[5] += x[3];
			float of[3] = {(float)p6[3], (float)p6[4], (float)p6[5]};
			float t[3] = {(float)p6[0], (float)p6[1], (float)p6[2]};
			euler_to_matrix_warp(of, t, ps->m, lane);
		}
		if (lane == 0) {
			for (int k = 0; k < 6; k++) ps->x[k] = x[k];
			ps->n_obs = n_obs;
			ps->status = status;
			ps->iterations += 1;
		}
		__syncwarp();
		pose_prepare_warp(ps, lane);   /* next iteration's Euler round trip */
	}
	if (lane == 0) {
		if (fin.bounds_reset) {
			fin.bounds_reset[0] = fin.bounds_reset[1] = fin.bounds_reset[2] = 0xFFFFFFFFu;
			fin.bounds_reset[3] = fin.bounds_reset[4] = fin.bounds_reset[5] = 0u;
		}
		if (fin.label_counts_reset) {
			fin.label_counts_reset[0] = fin.label_counts_reset[1] = fin.label_counts_reset[2] = fin.label_counts_reset[3] = 0ull;
		}
	}
}

constexpr int kNeqThreads = 256;
constexpr int kNeqInFlight = 4;

template <class Src>
__global__ void __launch_bounds__(kNeqThreads) k_normal_equations(Src src, int n, double *__restrict__ partials,
		unsigned int *__restrict__ ticket, FinalizeArgs fin)
{
	__shared__ double sm[kNeqThreads / 32][kMomentCount];
	__shared__ float wl[4];
	__shared__ bool is_last;
	if (threadIdx.x < 4) {
		float w = 0.0f;
		if constexpr (std::is_same<Src, ObsFromNN>::value) {
			unsigned long long c = src.label_counts[threadIdx.x];
			/* P = weight / count, float / int -> float (gpu6DSLAM.cpp:377-393) */
			w = c ? __fdiv_rn(src.weight[threadIdx.x], (float)(int)c) : 0.0f;
		}
		wl[threadIdx.x] = w;
	}
	__syncthreads();
	Moments mo;
	mo.clear();
	{
		/* kNeqInFlight observations per thread in flight: index loads, then the dependent gathers, then the arithmetic */
		const int stride = gridDim.x * blockDim.x;
		for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += kNeqInFlight * stride) {
			int tok[kNeqInFlight];
			typename Src::Raw raw[kNeqInFlight];
#pragma unroll
			for (int k = 0; k < kNeqInFlight; k++) { int i = i0 + k * stride; tok[k] = i < n ? src.token(i) : -1; }
#pragma unroll
			for (int k = 0; k < kNeqInFlight; k++) { int i = i0 + k * stride; if (tok[k] >= 0) src.fetch(i, tok[k], raw[k]); }
#pragma unroll
			for (int k = 0; k < kNeqInFlight; k++) {
				if (tok[k] >= 0) {
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
	/* deterministic final reduction: fixed row order */
	__shared__ double tot[kMomentCount];
	for (int col = wid; col < kMomentCount; col += kNeqThreads / 32) {     /* warp per column, lanes stride the rows */
		double s = 0;
		for (unsigned int b = lane; b < gridDim.x; b += 32) s += __ldcg(partials + (size_t)b * kPartialCols + col);
		s = warp_sum(s);
		if (lane == 0) tot[col] = s;
	}
	__syncthreads();
	if (wid == 0) {
		__shared__ double neq[kNeqCount];
		if (lane == 0) *ticket = 0;
		const double *p6 = fin.ps ? fin.ps->pose6 : fin.pose6_in;
		moments_to_neq_warp(tot, p6[3], p6[4], p6[5], neq, lane);
		neq_tail_warp(neq, fin, lane);
	}
}

/* Standalone device Cholesky (m3dreg_solve_chol): column-major dof x dof in, x out. */
__global__ void k_solve_dense(const double *A, const double *b, int dof, double *x, int *status)
{
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
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0.0;
}

/* ================================ NDT (point-to-distribution) ====================================================
 * Not in the reference; definition of record: oracle/m3d_oracle.c orc_ndt_normal_equations.
 * Per bucket of the gridded (moving) cloud: mean / covariance of its points' GLOBAL coordinates, accumulated
 * relative to the cell centre in fp64, and the mean of their LOCAL coordinates; W = (Sigma + eps I)^-1.
 * All queries that fall into a bucket share its (mu, W, Jacobian), so the query pass only needs a COUNT and a
 * coordinate SUM per bucket; the normal equations are then a reduction over buckets:
 *     N += cnt * A^T W A,   rhs += A^T W (cnt*mu_g - sum q),   A = -[I | J(mu_l)].
 * Layout: acc[b*12 ..] = {sum p'(3), sum p'p'^T(6), sum p_local(3)} -> finalised in place to
 *         {mu_g(3), mu_l(3), W(6: xx,xy,xz,yy,yz,zz)}; W.xx == 0 marks an unusable bucket.  qacc[b*4..] = {cnt, sum q(3)}. */
constexpr int kNdtMinPoints = 5;
constexpr double kNdtRegRel = 0.05;

__global__ void k_ndt_zero(double *__restrict__ acc, double *__restrict__ qacc, const m3dreg_grid_params *__restrict__ gp, int zero_acc)
{
	long long nb = gp->number_of_buckets;
	long long total = nb * (zero_acc ? 16 : 4);
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
		if (i < nb * 4) qacc[i] = 0.0;
		else acc[i - nb * 4] = 0.0;
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

__global__ void __launch_bounds__(256) k_ndt_accumulate_points(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int n,
		const float4 *__restrict__ g_xyzl, const float4 *__restrict__ l_xyzl, const m3dreg_grid_params *__restrict__ gp,
		double *__restrict__ acc)
{
	if (gp->number_of_buckets <= 0) return;
	int nround = (n + 31) & ~31;
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nround; p += gridDim.x * blockDim.x) {
		uint32_t key = 0xFFFFFFFFu;
		double v[12];
#pragma unroll
		for (int i = 0; i < 12; i++) v[i] = 0.0;
		if (p < n) {
			key = __ldg(keys + p);
			uint32_t idx = __ldg(vals + p);
			float4 g = __ldg(g_xyzl + idx), l = __ldg(l_xyzl + idx);
			double cx, cy, cz;
			cell_centre(key, gp, cx, cy, cz);
			double x = (double)g.x - cx, y = (double)g.y - cy, z = (double)g.z - cz;
			v[0] = x; v[1] = y; v[2] = z;
			v[3] = x * x; v[4] = x * y; v[5] = x * z; v[6] = y * y; v[7] = y * z; v[8] = z * z;
			v[9] = l.x; v[10] = l.y; v[11] = l.z;
		}
		bool head;
		warp_segmented_sum<12>(key, v, head);
		if (head && key != 0xFFFFFFFFu) {
			double *a = acc + (size_t)key * 12;
#pragma unroll
			for (int i = 0; i < 12; i++) atomicAdd(a + i, v[i]);
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
		double *__restrict__ acc)
{
	long long nb = gp->number_of_buckets;
	double res = (double)gp->resolution_X;
	double eps = (kNdtRegRel * res) * (kNdtRegRel * res);
	for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (long long)gridDim.x * blockDim.x) {
		int n = buckets[b].number_of_points;
		double *a = acc + (size_t)b * 12;
		if (n < kNdtMinPoints) { a[6] = 0.0; continue; }
		double s[3] = {a[0], a[1], a[2]}, ss[6] = {a[3], a[4], a[5], a[6], a[7], a[8]}, sl[3] = {a[9], a[10], a[11]};
		double inv = 1.0 / n, d = 1.0 / (n - 1);
		double m[3] = {s[0] * inv, s[1] * inv, s[2] * inv};
		double S[6], W[6];
		S[0] = (ss[0] - n * m[0] * m[0]) * d + eps; S[1] = (ss[1] - n * m[0] * m[1]) * d; S[2] = (ss[2] - n * m[0] * m[2]) * d;
		S[3] = (ss[3] - n * m[1] * m[1]) * d + eps; S[4] = (ss[4] - n * m[1] * m[2]) * d; S[5] = (ss[5] - n * m[2] * m[2]) * d + eps;
		if (!sym3_inverse(S, W)) { a[6] = 0.0; continue; }
		double cx, cy, cz;
		cell_centre((uint32_t)b, gp, cx, cy, cz);
		a[0] = m[0] + cx; a[1] = m[1] + cy; a[2] = m[2] + cz;
		a[3] = sl[0] * inv; a[4] = sl[1] * inv; a[5] = sl[2] * inv;
#pragma unroll
		for (int i = 0; i < 6; i++) a[6 + i] = W[i];
	}
}

__global__ void __launch_bounds__(256) k_ndt_accumulate_queries(const float4 *__restrict__ q_xyzl, int n2,
		const m3dreg_grid_params *__restrict__ gp, const double *__restrict__ acc, double *__restrict__ qacc)
{
	long long nb = gp->number_of_buckets;
	if (nb <= 0) return;
	float mnx = gp->bounding_box_min_X, mny = gp->bounding_box_min_Y, mnz = gp->bounding_box_min_Z;
	float rx = gp->resolution_X, ry = gp->resolution_Y, rz = gp->resolution_Z;
	int nby = gp->number_of_buckets_Y, nbz = gp->number_of_buckets_Z;
	int nround = (n2 + 31) & ~31;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
		uint32_t key = 0xFFFFFFFFu;
		double v[4] = {0.0, 0.0, 0.0, 0.0};
		if (i < n2) {
			float4 p = __ldg(q_xyzl + i);
			bool inside = !(p.x < mnx || p.x > gp->bounding_box_max_X) && !(p.y < mny || p.y > gp->bounding_box_max_Y) &&
					!(p.z < mnz || p.z > gp->bounding_box_max_Z);
			if (inside) {
				int h = cell_of(p.x, mnx, rx) * nby * nbz + cell_of(p.y, mny, ry) * nbz + cell_of(p.z, mnz, rz);
				if (h >= 0 && (long long)h < nb && __ldg(acc + (size_t)h * 12 + 6) > 0.0) {
					key = (uint32_t)h;
					v[0] = 1.0; v[1] = p.x; v[2] = p.y; v[3] = p.z;
				}
			}
		}
		bool head;
		warp_segmented_sum<4>(key, v, head);
		if (head && key != 0xFFFFFFFFu) {
			double *a = qacc + (size_t)key * 4;
#pragma unroll
			for (int k = 0; k < 4; k++) atomicAdd(a + k, v[k]);
		}
	}
}

__global__ void __launch_bounds__(kNeqThreads) k_ndt_normal_equations(const double *__restrict__ acc, const double *__restrict__ qacc,
		const m3dreg_grid_params *__restrict__ gp, double *__restrict__ partials, unsigned int *__restrict__ ticket, FinalizeArgs fin)
{
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
		const double *qa = qacc + (size_t)b * 4;
		double cnt = qa[0];
		if (!(a[6] > 0.0) || !(cnt > 0.0)) continue;
		double A[3][6];
#pragma unroll
		for (int r = 0; r < 3; r++)
#pragma unroll
			for (int c = 0; c < 3; c++) {
				A[r][c] = (r == c) ? -1.0 : 0.0;
				A[r][3 + c] = -(C[r][c][0] * a[3] + C[r][c][1] * a[4] + C[r][c][2] * a[5]);
			}
		double W[3][3] = {{a[6], a[7], a[8]}, {a[7], a[9], a[10]}, {a[8], a[10], a[11]}};
		double sl[3] = {cnt * a[0] - qa[1], cnt * a[1] - qa[2], cnt * a[2] - qa[3]};
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
