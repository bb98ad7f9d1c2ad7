/* grid_build.cuh — the regular-grid decomposition of one iteration as ONE persistent kernel.
 *
 * Replaces, per iteration of the fused loop, the reference's cudaCalculateGridParams + cudaCalculateGrid
 * (src/lesson_16.cu:23-243: 3x thrust::minmax_element, 6 D2H copies, host float math, kernel_initializeIndByKey,
 * kernel_getIndexOfBucketForPoints, thrust::sort, kernel_initializeBuckets / updateBuckets /
 * countNumberOfPointsForBuckets / copyKeys, each followed by cudaDeviceSynchronize) and the CPU transform of the first
 * cloud (src/gpu6DSLAM.cpp:635-663) — round 1 spent eight launches (66 us on the 1 M-point pair) on it.
 *
 * One block per SM, all co-resident (cooperative launch), phases separated by a software grid barrier (~1 us) instead
 * of kernel boundaries.  Block c owns the contiguous point range [c*T, (c+1)*T):
 *   P1  transform of the scan by the pose (registers only, nothing stored) + bounding box               | barrier
 *   P2  grid parameters (every block, identical float ops), bucket key of every point, first-digit
 *       histogram of the block's range, dense per-bucket point counts (integer atomics)                    | barrier
 *   P3  stable LSD scatter by digit 0; next digit's histogram accumulated where the items land;
 *       bucket-count scan, level 1 (each block: its slice of the dense table)                              | barrier
 *   P4  bucket-count scan, level 2 -> dense {begin,end,n} table with the reference's first-element quirk
 *       (src/lesson_16.cu:148-158), list of searchable buckets; LSD scatter by digit 1 (if any)           | barrier
 *   (P4' LSD scatter by digit 2 for > 2^16 buckets                                                         | barrier)
 *   (the sorted table always ends in buffer 1: the first pass reads from buffer 0 or 1 by the parity of the pass count)
 *   P5  candidate sets (one warp per searchable bucket), transforming the candidates' points and normals on
 *       the fly.
 * The dense table comes from the COUNTS (begin = exclusive scan), not from run detection on the sorted keys, so it is
 * ready one phase before the sort ends; the sort only has to produce the permutation.  Digits are
 * ceil(bits / passes) <= 8 bits wide with bits = bit width of the actual bucket count (read on the device): 2 passes up
 * to 65 536 buckets, 3 beyond.  All cross-block data written inside the launch is read with ld.global.cg (L1 is not
 * coherent across SMs); the read-only inputs (stored scan, pose) go through the non-coherent path.
 * Results are bit-identical to the multi-kernel path (and hence to the reference): same float ops for the keys, a
 * stable sort, integer counting. */
#pragma once
#include "m3dreg_kernels.cuh"

namespace m3d {

constexpr int kGbThreads = 1024;
constexpr int kGbWarps = kGbThreads / 32;
constexpr int kGbItems = 7;                               /* keys per thread and sub-tile */
constexpr int kGbSub = kGbThreads * kGbItems;             /* sub-tile: 7168 keys */
constexpr int kGbRadix = 256;

struct GridBlockRec {       /* level-1 result of the bucket-count scan for one block's slice of the dense table */
	int total;              /* points in the slice                         */
	int occ;                /* occupied buckets in the slice               */
	int k0, c0;             /* smallest occupied bucket of the slice (INT_MAX: none) and its count */
	int k1;                 /* second smallest occupied bucket (INT_MAX: none) */
	int pad[3];
};

constexpr int kGbCacheMaxBytes = 160 * 1024;             /* dynamic shared memory for the block's transformed points */

struct GridBuildArgs {
	const float4 *lx, *ln;          /* the scan being gridded, original order (local frame unless pose == 0) */
	int n;
	int cache;                      /* the block's T transformed points fit in dynamic shared memory (3 float arrays of T) */
	unsigned long long *dbg;        /* optional: globaltimer of block 0 at the phase boundaries (diagnostics), 0 = off */
	const float *pose;              /* DEVICE row-major 4x4 applied to lx / ln; 0 = identity (cloud already global) */
	float res, ext;
	long long bucket_cap;           /* entries allocated for buckets / bcount */
	int max_inner, max_outer;
	int build_cands;                /* 0: grid only (NDT) */
	float4 *g_xyzl;                 /* optional: the transformed cloud in original order (NDT, exports); 0 = not kept */
	uint32_t *keys[2], *vals[2];
	uint32_t *hist;                 /* 3 matrices [256][Gp], Gp = gridDim rounded up to 4 (digit-major: a block reads rows with 128-bit loads) */
	int *bcount;                    /* dense per-bucket counts: all zero on entry, all zero on exit */
	int *bbegin;                    /* scratch, bucket_cap ints: slice-relative begin */
	m3dreg_bucket *buckets;
	uint32_t *cell_list;
	unsigned int *cell_count;
	GridBlockRec *brec;             /* gridDim records */
	m3dreg_grid_params *gp;
	int *flags;
	uint32_t *bounds;               /* 6 ordered-uint bounds: reset state on entry, reset again on exit */
	unsigned int *bar;              /* [0] arrival count, [1] generation */
	CandSet ci, co;
};

/* diagnostics: %globaltimer of one thread into dbg[slot] (dbg == 0: nothing) */
__device__ __forceinline__ void gb_stamp(unsigned long long *dbg, int slot, bool who)
{
	if (dbg && who) {
		unsigned long long t;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
		dbg[slot] = t;
	}
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
{
	unsigned int v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

/* Sense-reversing grid barrier over all blocks of a co-resident grid.  bar[0] = arrivals, bar[1] = generation; the last
 * arriver resets the count and bumps the generation, so the pair is back in its initial state after every barrier and
 * no launch depends on what an earlier one left behind.  `gen` is the generation this block saw at kernel entry (read
 * before its first arrival, hence before that barrier can complete). */
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int &gen)
{
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(bar, 1u) == gridDim.x - 1) {
			atomicExch(bar, 0u);
			__threadfence();
			atomicAdd(bar + 1, 1u);
		} else {
			while (ld_acquire_u32(bar + 1) == gen) { }
		}
		__threadfence();
	}
	gen++;
	__syncthreads();
}

__device__ __forceinline__ int gb_bits_for(long long nb)
{
	int bits = 1;
	while (bits < 31 && (1LL << bits) < nb) bits++;
	return bits;
}

/* exclusive scan of one value per thread over the block (kGbThreads threads); returns the exclusive prefix, *total the
 * block sum.  s_scan: kGbWarps + 1 entries.  64-bit so that two counters can ride in one scan (count | flag << 40). */
__device__ __forceinline__ long long gb_block_excl_scan(long long v, long long *s_scan, long long *total)
{
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	long long incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		long long t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += t;
	}
	__syncthreads();
	if (lane == 31) s_scan[w] = incl;
	__syncthreads();
	if (w == 0) {
		long long x = s_scan[lane], xi = x;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			long long t = __shfl_up_sync(0xffffffffu, xi, o);
			if (lane >= o) xi += t;
		}
		s_scan[lane] = xi - x;
		if (lane == 31) s_scan[32] = xi;
	}
	__syncthreads();
	*total = s_scan[32];
	return s_scan[w] + incl - v;
}

/* One stable LSD pass over the block's range: keys_in/vals_in -> keys_out/vals_out by digit (key >> shift) & mask.
 * hist = this pass's [256][Gp] matrix of per-block digit counts (complete); hist_next (0 for the last pass) gets
 * the next digit's counts per DESTINATION block.  vals_in == 0: implicit original indices.  keys_smem != 0: the block's
 * keys are still in shared memory from the key phase (first pass: a block sorts the range it made the keys of). */
__device__ __noinline__ void gb_scatter_pass(const uint32_t *keys_in, const uint32_t *keys_smem, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
		int n, int T, int shift, int dbits, const uint32_t *hist, uint32_t *hist_next, uint32_t (*s_w)[kGbRadix], uint32_t *s_gbase,
		long long *s_scan, unsigned long long *dbg = nullptr)
{
	const bool dbg_who = blockIdx.x == 0 && threadIdx.x == 0;
	gb_stamp(dbg, 16, dbg_who);
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const uint32_t mask = (1u << dbits) - 1u;
	const int G = (int)gridDim.x, c = (int)blockIdx.x;
	/* a) where this block's digits start: sum of the digit's counts in earlier blocks + all smaller digits everywhere.
	 * Only the 2^dbits digits in use are read; (1024 >> dbits) threads share a digit's row, each adding every tpd-th
	 * 128-bit piece (all of a thread's loads in flight at once), then a butterfly over those threads */
	{
		const int lt = dbits < 5 ? 5 : dbits;                      /* log2 of digits handled (>= 32 so that a digit's threads sit in one warp) */
		const int tpd = kGbThreads >> lt;                          /* threads per digit: 4 (8 bits) .. 32 (5 bits) */
		const int d = threadIdx.x / tpd, qd = threadIdx.x % tpd;
		const int Gp = (G + 3) & ~3;
		uint32_t pre = 0, tot = 0;
		if (d < (1 << dbits)) {
			const uint4 *row = reinterpret_cast<const uint4 *>(hist + (size_t)d * Gp);
#pragma unroll 5
			for (int j = qd; j < (Gp >> 2); j += tpd) {
				const uint4 v = __ldcg(row + j);
				const int r0 = j << 2;
				tot += v.x + v.y + v.z + v.w;
				pre += (r0 < c ? v.x : 0u) + (r0 + 1 < c ? v.y : 0u) + (r0 + 2 < c ? v.z : 0u) + (r0 + 3 < c ? v.w : 0u);
			}
		}
		for (int o = 1; o < tpd; o <<= 1) {
			tot += __shfl_xor_sync(0xffffffffu, tot, o);
			pre += __shfl_xor_sync(0xffffffffu, pre, o);
		}
		long long total;
		const long long ex = gb_block_excl_scan(qd == 0 ? (long long)tot : 0LL, s_scan, &total);      /* thread order = digit order */
		if (qd == 0 && d < kGbRadix) s_gbase[d] = (uint32_t)((int)ex + (int)pre);
		__syncthreads();
	}
	gb_stamp(dbg, 17, dbg_who);
	const int begin = c * T, end = min(n, begin + T);
	const uint32_t lt = (1u << lane) - 1u;
	for (int sub = begin; sub < end; sub += kGbSub) {
#pragma unroll
		for (int k = 0; k < kGbRadix / 32; k++) s_w[w][k * 32 + lane] = 0;      /* each warp clears its own row */
		__syncwarp();
		const int wbase = sub + w * (32 * kGbItems);
		uint32_t key[kGbItems], val[kGbItems], rank[kGbItems];
#pragma unroll
		for (int j = 0; j < kGbItems; j++) {
			const int i = wbase + j * 32 + lane;
			const bool valid = i < end;
			key[j] = valid ? (keys_smem ? keys_smem[i - begin] : __ldcg(keys_in + i)) : 0xFFFFFFFFu;
			val[j] = valid ? (vals_in ? __ldcg(vals_in + i) : (uint32_t)i) : 0u;
		}
#pragma unroll
		for (int j = 0; j < kGbItems; j++) {
			const int i = wbase + j * 32 + lane;
			const bool valid = i < end;
			const uint32_t d = (key[j] >> shift) & mask;
			const uint32_t mk = valid ? d : (0x100u + lane);               /* idle lanes: private match groups */
			const uint32_t peers = __match_any_sync(0xffffffffu, mk);
			const int leader = __ffs(peers) - 1;
			uint32_t old = 0;
			if (lane == leader && valid) {
				old = s_w[w][d];
				s_w[w][d] = old + __popc(peers);
			}
			old = __shfl_sync(0xffffffffu, old, leader);
			rank[j] = old + __popc(peers & lt);
			__syncwarp();
		}
		__syncthreads();
		gb_stamp(dbg, 18, dbg_who);
		uint32_t sub_tot = 0;
		if (threadIdx.x < 256) {     /* exclusive scan over the warps for digit = threadIdx.x */
			uint32_t run = 0;
#pragma unroll 8
			for (int k = 0; k < kGbWarps; k++) {
				const uint32_t cnt = s_w[k][threadIdx.x];
				s_w[k][threadIdx.x] = run;
				run += cnt;
			}
			sub_tot = run;
		}
		__syncthreads();
		gb_stamp(dbg, 19, dbg_who);
#pragma unroll
		for (int j = 0; j < kGbItems; j++) {
			const int i = wbase + j * 32 + lane;
			if (i < end) {
				const uint32_t d = (key[j] >> shift) & mask;
				const uint32_t pos = s_gbase[d] + s_w[w][d] + rank[j];
				keys_out[pos] = key[j];
				vals_out[pos] = val[j];
				rank[j] = pos;
			}
		}
		gb_stamp(dbg, 20, dbg_who);
		if (hist_next) {
			/* next pass: one atomic per group of equal (destination block, next digit) inside the warp */
#pragma unroll
			for (int j = 0; j < kGbItems; j++) {
				const int i = wbase + j * 32 + lane;
				const bool valid = i < end;
				const uint32_t slot = valid ? ((key[j] >> (shift + dbits)) & mask) * (uint32_t)((G + 3) & ~3) + rank[j] / (uint32_t)T : (0xFFFFFF00u + lane);
				const uint32_t peers = __match_any_sync(0xffffffffu, slot);
				if (valid && lane == __ffs(peers) - 1) atomicAdd(hist_next + slot, (uint32_t)__popc(peers));
			}
		}
		__syncthreads();
		gb_stamp(dbg, 21, dbg_who);
		if (threadIdx.x < 256) s_gbase[threadIdx.x] += sub_tot;
		__syncthreads();
	}
}

#define M3D_GB_STAMP(slot) do { if (a.dbg && c == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.dbg[slot] = t_; } } while (0)

__global__ void __launch_bounds__(kGbThreads, 1) k_grid_build(const GridBuildArgs a)
{
	pdl_enter();
	extern __shared__ float s_pts[];                               /* cache != 0: x[T], y[T], z[T] of the block's transformed points; x is later the keys */
	__shared__ uint32_t s_raw[kGbWarps * (kBuildTabMax + 7)];      /* scatter: [32][256] warp digit counts; candidates: [32][264] bin tables */
	__shared__ uint32_t s_gbase[kGbRadix];
	__shared__ long long s_scan[kGbWarps + 1];
	__shared__ int s_rec[8];
	uint32_t (*s_w)[kGbRadix] = reinterpret_cast<uint32_t (*)[kGbRadix]>(s_raw);
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int G = (int)gridDim.x, c = (int)blockIdx.x;
	const int Gp = (G + 3) & ~3;
	const int n = a.n;
	unsigned int gen = 0;
	if (threadIdx.x == 0) gen = ld_acquire_u32(a.bar + 1);
	M3D_GB_STAMP(0);
	/* block ranges: T a multiple of 32 so that warps never straddle two blocks' ranges */
	const int T = (((n + G - 1) / G) + 31) & ~31;
	const int begin = min(n, c * T), end = min(n, begin + T);
	float *s_x = s_pts, *s_y = s_pts + T, *s_z = s_pts + 2 * T;
	uint32_t *s_keys = reinterpret_cast<uint32_t *>(s_pts);

	PointXform xf;
	xf.on = a.pose != nullptr;
#pragma unroll
	for (int k = 0; k < 12; k++) xf.r[k] = xf.on ? __ldg(a.pose + k) : 0.0f;

	/* ---- P1: bounding box of the transformed cloud (replaces 3x thrust::minmax_element, lesson_16.cu:34-45);
	 *      four independent loads per thread in flight; the transformed points stay in shared memory for P2 ---- */
	{
		float mnx = INFINITY, mny = INFINITY, mnz = INFINITY, mxx = -INFINITY, mxy = -INFINITY, mxz = -INFINITY;
		for (int i0 = begin + (int)threadIdx.x; i0 < end; i0 += 4 * kGbThreads) {
			float4 p[4];
#pragma unroll
			for (int k = 0; k < 4; k++) { const int i = i0 + k * kGbThreads; p[k] = __ldg(a.lx + (i < end ? i : i0)); }
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const int i = i0 + k * kGbThreads;
				if (i < end) {
					const float4 q = xform_point(xf, p[k]);
					mnx = fminf(mnx, q.x); mny = fminf(mny, q.y); mnz = fminf(mnz, q.z);
					mxx = fmaxf(mxx, q.x); mxy = fmaxf(mxy, q.y); mxz = fmaxf(mxz, q.z);
					if (a.cache) { s_x[i - begin] = q.x; s_y[i - begin] = q.y; s_z[i - begin] = q.z; }
				}
			}
		}
		block_bounds_commit(mnx, mny, mnz, mxx, mxy, mxz, a.bounds);
	}
	M3D_GB_STAMP(1);
	grid_barrier(a.bar, gen);
	M3D_GB_STAMP(2);

	/* ---- P2: grid parameters, keys, first-digit histogram, dense bucket counts ---- */
	m3dreg_grid_params g;
	bool ok;
	{
		uint32_t b[6];
#pragma unroll
		for (int k = 0; k < 6; k++) b[k] = __ldcg(a.bounds + k);
		ok = grid_params_from_bounds_dev(b, a.res, a.res, a.res, a.ext, a.bucket_cap, g);
	}
	if (c == 0 && threadIdx.x == 0) {
		*a.gp = g;
		if (!ok) atomicExch(&a.flags[FLAG_ERROR], M3DREG_E_TOO_MANY_BUCKETS);
	}
	if (!ok) return;      /* every block takes the same decision: nobody is left waiting (the host re-arms the bounds) */
	const long long nb = g.number_of_buckets;
	const int bits = gb_bits_for(nb);
	const int passes = (bits + 7) / 8;
	const int dbits = (bits + passes - 1) / passes;
	const uint32_t dmask = (1u << dbits) - 1u;
	const int start = (passes & 1) ? 0 : 1;      /* the sorted table always ends up in buffer 1 */
	uint32_t *hist0 = a.hist, *hist1 = a.hist + (size_t)Gp * kGbRadix, *hist2 = a.hist + 2 * (size_t)Gp * kGbRadix;
	{
		if (threadIdx.x < kGbRadix) {
			hist1[(size_t)threadIdx.x * Gp + c] = 0;
			hist2[(size_t)threadIdx.x * Gp + c] = 0;
		}
		/* per-WARP digit histograms (s_w rows): a block's points fall into a handful of buckets, and 32 warps adding to
		 * the same few shared-memory words serialise (measured: this phase took 10 us with one block histogram) */
#pragma unroll
		for (int k = 0; k < kGbRadix / 32; k++) s_w[w][k * 32 + lane] = 0;
		if (c == 0) {                     /* padding columns G..Gp-1 of the three matrices (the buffer is shared with other sorts) */
			for (int t = (int)threadIdx.x; t < 3 * kGbRadix * (Gp - G); t += kGbThreads)
				a.hist[(size_t)(t / (Gp - G)) * Gp + G + t % (Gp - G)] = 0;
		}
		__syncwarp();
		const int nby = g.number_of_buckets_Y, nbz = g.number_of_buckets_Z;
		const int span = ((end - begin) + kGbThreads - 1) / kGbThreads * kGbThreads;      /* whole warps stay together for the ballots */
		for (int o = (int)threadIdx.x; o < span; o += kGbThreads) {
			const int i = begin + o;
			int key = -1;
			if (i < end) {
				float4 p;
				if (a.cache) p = make_float4(s_x[o], s_y[o], s_z[o], 0.0f);
				else p = xform_point(xf, __ldg(a.lx + i));
				const int ix = cell_of(p.x, g.bounding_box_min_X, a.res), iy = cell_of(p.y, g.bounding_box_min_Y, a.res),
						iz = cell_of(p.z, g.bounding_box_min_Z, a.res);
				key = ix * nby * nbz + iy * nbz + iz;
				if (a.cache) s_keys[o] = (uint32_t)key;      /* over x[o], which only this thread reads */
				else a.keys[start][i] = (uint32_t)key;
				if (a.g_xyzl) { p.w = a.cache ? __ldg(&a.lx[i].w) : p.w; a.g_xyzl[i] = p; }
			}
			/* runs of equal keys inside the warp: one shared-memory and one global atomic per run */
			const int prev = __shfl_up_sync(0xffffffffu, key, 1);
			const bool head = (lane == 0) || (prev != key);
			const unsigned heads = __ballot_sync(0xffffffffu, head);
			if (head && key >= 0) {
				const unsigned later = heads & ~((2u << lane) - 1u);
				const int len = (later ? __ffs(later) - 1 : 32) - lane;      /* idle lanes (key -1) only trail the last valid run */
				atomicAdd(&s_w[w][(uint32_t)key & dmask], (uint32_t)len);     /* contention only among the run heads of this warp */
				atomicAdd(a.bcount + key, len);
			}
		}
		__syncthreads();
		if (threadIdx.x < kGbRadix) {
			uint32_t s = 0;
#pragma unroll 8
			for (int k = 0; k < kGbWarps; k++) s += s_w[k][threadIdx.x];
			hist0[(size_t)threadIdx.x * Gp + c] = s;
		}
		__syncthreads();      /* s_w is reused by the first sort pass */
	}
	M3D_GB_STAMP(3);
	grid_barrier(a.bar, gen);
	M3D_GB_STAMP(4);

	/* ---- P3: LSD pass 0 + bucket-count scan level 1 ---- */
	const int S = (int)((nb + G - 1) / G);                       /* dense-table slice per block */
	const long long sb = (long long)c * S;
	const int sn = (int)max(0LL, min((long long)S, nb - sb));      /* buckets in this block's slice */
	{
		int run = 0, occ = 0;
		if (threadIdx.x == 0) { s_rec[0] = 0x7fffffff; s_rec[1] = 0; s_rec[2] = 0x7fffffff; }
		__syncthreads();
		for (int o = 0; o < sn; o += kGbThreads) {
			const int bi = o + (int)threadIdx.x;
			const int cnt = bi < sn ? __ldcg(a.bcount + sb + bi) : 0;
			long long total;
			const long long ex = gb_block_excl_scan((long long)cnt | (cnt > 0 ? (1LL << 40) : 0LL), s_scan, &total);
			if (bi < sn) a.bbegin[sb + bi] = run + (int)(ex & ((1LL << 40) - 1));
			/* the two smallest occupied buckets of the slice = occupied ranks 0 and 1 */
			const int orank = occ + (int)(ex >> 40);
			if (cnt > 0 && orank == 0) { s_rec[0] = (int)(sb + bi); s_rec[1] = cnt; }
			if (cnt > 0 && orank == 1) s_rec[2] = (int)(sb + bi);
			run += (int)(total & ((1LL << 40) - 1));
			occ += (int)(total >> 40);
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			GridBlockRec r;
			r.total = run; r.occ = occ; r.k0 = s_rec[0]; r.c0 = s_rec[1]; r.k1 = s_rec[2]; r.pad[0] = r.pad[1] = r.pad[2] = 0;
			a.brec[c] = r;
		}
	}
	M3D_GB_STAMP(15);
	gb_scatter_pass(a.keys[start], a.cache ? s_keys : nullptr, nullptr, a.keys[start ^ 1], a.vals[start ^ 1], n, T, 0, dbits, hist0, passes > 1 ? hist1 : nullptr, s_w, s_gbase, s_scan, a.dbg);
	M3D_GB_STAMP(5);
	grid_barrier(a.bar, gen);
	M3D_GB_STAMP(6);

	/* ---- P4: bucket-count scan level 2 -> dense table + searchable-bucket list; LSD pass 1 ---- */
	{
		/* every block reduces the G slice records (warp 0, ceil(G/32) records per lane) */
		__shared__ int s_l2[8];
		if (w == 0) {
			const int per = (G + 31) / 32;
			int tot_before = 0, occ_before = 0, occ_all = 0;
			int k0 = 0x7fffffff, c0 = 0, k1 = 0x7fffffff;
			int my_tot = 0, my_occ = 0, my_tb = 0, my_ob = 0;
			/* lane-local pass over its records (block order = lane-major), then a scan over lanes */
			int lk0 = 0x7fffffff, lc0 = 0, lk1 = 0x7fffffff;
			for (int k = 0; k < per; k++) {
				const int r = lane * per + k;
				if (r < G) {
					const int4 v = __ldcg(reinterpret_cast<const int4 *>(a.brec + r));       /* total, occ, k0, c0 */
					const int rk1 = __ldcg(&a.brec[r].k1);
					if (r < c) { my_tb += v.x; my_ob += v.y; }
					my_tot += v.x; my_occ += v.y;
					/* merge (lk0, lk1) with (v.z, rk1): records are in increasing key order */
					if (lk0 == 0x7fffffff) { lk0 = v.z; lc0 = v.w; lk1 = rk1; }
					else if (lk1 == 0x7fffffff) lk1 = v.z;
				}
			}
			tot_before = __reduce_add_sync(0xffffffffu, my_tb);
			occ_before = __reduce_add_sync(0xffffffffu, my_ob);
			occ_all = __reduce_add_sync(0xffffffffu, my_occ);
			(void)my_tot;
			/* first two occupied buckets overall: the first lane holding one, then its second or the next lane's first */
			const unsigned has = __ballot_sync(0xffffffffu, lk0 != 0x7fffffff);
			if (has) {
				const int f = __ffs(has) - 1;
				k0 = __shfl_sync(0xffffffffu, lk0, f);
				c0 = __shfl_sync(0xffffffffu, lc0, f);
				k1 = __shfl_sync(0xffffffffu, lk1, f);
				const unsigned rest = has & ~((2u << f) - 1u);
				if (k1 == 0x7fffffff && rest) k1 = __shfl_sync(0xffffffffu, lk0, __ffs(rest) - 1);
			}
			if (lane == 0) {
				s_l2[0] = tot_before; s_l2[1] = occ_before; s_l2[2] = occ_all; s_l2[3] = k0; s_l2[4] = c0; s_l2[5] = k1;
			}
		}
		__syncthreads();
		const int tot_before = s_l2[0], occ_all = s_l2[2], k0 = s_l2[3], c0 = s_l2[4], k1 = s_l2[5];
		int occ_before = s_l2[1];
		/* kernel_updateBuckets' first-element quirk (lesson_16.cu:148-158): when sorted element 0 is alone in its bucket,
		 * the run starting at position 1 (bucket k1) never gets index_begin and keeps number_of_points 0; its index_end
		 * is a write race upstream (1 vs run end) — we store the run end */
		const bool has_quirk = n > 1 && c0 == 1 && k1 != 0x7fffffff;
		(void)k0;
		if (has_quirk && (long long)k1 < sb) occ_before -= 1;
		if (c == 0 && threadIdx.x == 0) *a.cell_count = (unsigned int)(occ_all - (has_quirk ? 1 : 0));
		int listed_run = 0;
		for (int o = 0; o < sn; o += kGbThreads) {
			const int bi = o + (int)threadIdx.x;
			int cnt = 0, bg = 0;
			if (bi < sn) {
				cnt = __ldcg(a.bcount + sb + bi);
				bg = tot_before + __ldcg(a.bbegin + sb + bi);
				a.bcount[sb + bi] = 0;                                 /* clean for the next launch */
			}
			const bool quirk = has_quirk && bi < sn && (sb + bi) == (long long)k1;
			const bool listed = cnt > 0 && !quirk;
			if (bi < sn) {
				m3dreg_bucket rec;
				if (cnt > 0) { rec.index_begin = quirk ? -1 : bg; rec.index_end = bg + cnt; rec.number_of_points = quirk ? 0 : cnt; }
				else { rec.index_begin = -1; rec.index_end = -1; rec.number_of_points = 0; }
				int *bp = reinterpret_cast<int *>(a.buckets + sb + bi);
				bp[0] = rec.index_begin; bp[1] = rec.index_end; bp[2] = rec.number_of_points;
			}
			long long total;
			const long long ex = gb_block_excl_scan(listed ? 1LL : 0LL, s_scan, &total);
			if (listed) a.cell_list[occ_before + listed_run + (int)ex] = (uint32_t)(sb + bi);
			listed_run += (int)total;
		}
	}
	if (passes > 1)
		gb_scatter_pass(a.keys[start ^ 1], nullptr, a.vals[start ^ 1], a.keys[start], a.vals[start], n, T, dbits, dbits, hist1, passes > 2 ? hist2 : nullptr, s_w, s_gbase, s_scan);
	M3D_GB_STAMP(7);
	grid_barrier(a.bar, gen);
	M3D_GB_STAMP(8);
	if (passes > 2) {
		gb_scatter_pass(a.keys[start], nullptr, a.vals[start], a.keys[start ^ 1], a.vals[start ^ 1], n, T, 2 * dbits, dbits, hist2, nullptr, s_w, s_gbase, s_scan);
		grid_barrier(a.bar, gen);
	}
	M3D_GB_STAMP(9);
	if (c == 0 && threadIdx.x == 0) {
		a.bounds[0] = a.bounds[1] = a.bounds[2] = 0xFFFFFFFFu;      /* everybody read them before the second barrier */
		a.bounds[3] = a.bounds[4] = a.bounds[5] = 0u;
	}

	/* ---- P5: candidate sets, one warp per searchable bucket ---- */
	if (!a.build_cands) return;
	{
		uint32_t (*s_hist)[kBuildTabMax + 7] = reinterpret_cast<uint32_t (*)[kBuildTabMax + 7]>(s_raw);
		NormalRotation rot;
		rot.on = xf.on;
#pragma unroll
		for (int k = 0; k < 9; k++) rot.r[k] = xf.r[(k / 3) * 4 + (k % 3)];
		const int nby = g.number_of_buckets_Y, nbz = g.number_of_buckets_Z;
		CellGeom cg;
		cg.mnx = g.bounding_box_min_X; cg.mny = g.bounding_box_min_Y; cg.mnz = g.bounding_box_min_Z;
		cg.rx = a.res; cg.ry = a.res; cg.rz = a.res;
		const int tables = nn_tables_usable(a.max_inner, a.max_outer) ? 1 : 0;
		const bool two_sets = a.max_inner != a.max_outer;
		const unsigned int ncells = __ldcg(a.cell_count);
		const unsigned int nwarps = (unsigned int)G * kGbWarps;
		const uint32_t *vals = a.vals[1];
		/* warp-major over blocks: consecutive list entries go to different SMs */
		for (unsigned int t = (unsigned int)w * G + c; t < ncells; t += nwarps) {
			if (a.dbg && c == 0 && w == 12 && lane == 0 && t == (unsigned int)w * G + c) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.dbg[22] = t_; }
			const int cell = (int)__ldcg(a.cell_list + t);
			const int *bp = reinterpret_cast<const int *>(a.buckets + cell);
			const int c_begin = __ldcg(bp), c_n = __ldcg(bp + 2);
			cg.cx = cell / (nby * nbz); cg.cy = (cell / nbz) % nby; cg.cz = cell % nbz;
			/* diagnostics: the phases of one mid-list bucket (block 0, warp 12) */
			unsigned long long *bdbg = (a.dbg && c == 0 && w == 12 && t == (unsigned int)w * G + c) ? a.dbg : nullptr;
			if (bdbg && lane == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); bdbg[23] = t_; }
			build_cell_candidates<true>(vals, a.lx, a.ln, nullptr, rot, xf, c_begin, c_n, a.max_inner, tables, cg, a.ci, s_hist[w], lane, bdbg);
			if (two_sets) build_cell_candidates<true>(vals, a.lx, a.ln, nullptr, rot, xf, c_begin, c_n, a.max_outer, tables, cg, a.co, s_hist[w], lane);
		}
	}
	__syncthreads();
	M3D_GB_STAMP(10);
}

} /* namespace m3d */
