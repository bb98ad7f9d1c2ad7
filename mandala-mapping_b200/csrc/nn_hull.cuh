/* nn_hull.cuh — semantic nearest neighbour, warp-shared hull search, second generation (k_nn_search_hull).
 *
 * Same answer as kernel_semanticNearestNeighborSearch (src/lesson_16.cu:531-703) and as k_nn_search / nn_query()
 * (nn_core.cuh): the lexicographic minimum of (dist, l) over the admissible candidates of the query's 27-neighbourhood.
 * Same algorithm as round 1's k_nn_search_grid (m3dreg_kernels.cuh — kept for A/B runs, M3DREG_NN_V7=1):
 *   1. rounds of growing radius rho; every unsettled lane computes the conservative box of fine columns (res / 4) that
 *      holds all candidates with dist <= min(limit, rho^2); the warp takes the HULL of the boxes (6 REDUX);
 *   2. the lanes look up the hull's fine cells side by side (bucket record + bin table, once per cell and warp) and
 *      compact the non-empty bins of the warp's label into a segment list;
 *   3. the segments' candidates are staged in shared memory by a flat coalesced copy, groups of four transposed;
 *   4. all lanes evaluate all staged groups (broadcast LDS.128 + packed f32x2 arithmetic), branch-free running minimum
 *      that remembers the winning GROUP and flags exact ties between groups;
 *   5. one cold step per lane applies the full predicate to the winning group; a tie or an inadmissible winner re-scans;
 *   6. a lane is settled when its limit is covered by rho^2 or its box for the current limit lies inside the hull.
 * What changed against round 1's kernel, and what measuring it on B200 showed (C2, 1 048 576 queries; profiles/):
 *   - PERSISTENT warps: the grid is one wave (resident blocks x SM count), every warp fetches chunks of 32 queries from an
 *     atomic counter until none is left; the launch-constant prologue runs once per warp, everything that depends only on
 *     the launch (4 / res, r^2, first radius) comes in as kernel arguments;
 *   - the per-lane margins, neighbourhood clamps and the query normal are recomputed / re-read where they are used instead
 *     of living in registers; the previous round's hull (whose cells are skipped, exactness argument in
 *     m3dreg_kernels.cuh) is kept packed in four registers; the box helper is out of line (two call sites);
 *   - the cold step issues its (up to four) normal loads and the query normal together, one L2 latency instead of four
 *     dependent ones; the winner's record comes from the candidate set's `loc` stream (point in the local frame +
 *     original index): one gather instead of table[l] -> cloud[index];
 *   - 7 KB instead of 11.8 KB of shared memory per two-warp block (64 cells, 128 staged candidates): what the blocks do
 *     not take stays L1 cache, and the L1 hit rate is what moved with occupancy (56 % at 24 warps per SM, 33 % at 32);
 *   - the evaluation counter is a template parameter (profiling only).
 * MEASURED: 12 / 14 / 16 resident blocks per SM (80 / 72 / 64 registers) give 115.8 / 114.2 / 117.1 us against 118.6 us
 * for round 1's kernel — issue slots stay 57 % busy whatever the number of resident warps (more warps = longer
 * scoreboard stalls: same instructions per second).  57.6 M warp instructions per launch at that rate is the kernel's
 * time; the remaining lever is the instruction count itself, see DESIGN.md section 5.
 * Exactness: a lane only ever takes the minimum over candidates of buckets in its own 27-neighbourhood (tested per
 * group when the hull leaves some lane's neighbourhood); looking at MORE of those than the reference does not change a
 * minimum, and every candidate with dist <= limit lies inside the lane's box (nn_core.cuh: col_floor / col_ceil). */
#pragma once
#include "m3dreg_kernels.cuh"

namespace m3d {

#ifndef M3D_NNH_MINBLOCKS
#define M3D_NNH_MINBLOCKS 12
#endif
constexpr int kNNHThreads = 64;
constexpr int kNNHWarps = kNNHThreads / 32;
#ifndef M3D_NNH_CELLS
#define M3D_NNH_CELLS 64
#endif
#ifndef M3D_NNH_STAGE
#define M3D_NNH_STAGE 128
#endif
constexpr int kNNHCells = M3D_NNH_CELLS;  /* hull cells looked up per chunk (segment list capacity) */
constexpr int kNNHStage = M3D_NNH_STAGE;  /* candidates staged per batch (multiple of 4); shared memory per warp =
                                           * 16 B x (cells + stage + stage / 4): what the blocks do not take stays L1 cache */

struct NNHullArgs {
	const float4 *q_xyzl, *q_nrm;
	const uint32_t *q_perm;
	int n_second;
	CandSet cs;
	const uint32_t *s_vals;
	int n_first;
	const m3dreg_bucket *buckets;
	const m3dreg_grid_params *gp;
	float search_radius;
	float r2;                   /* fl(radius * radius), lesson_16.cu:553 */
	float iwx, iwy, iwz;        /* fl(4 / res): fine columns per metre */
	float rho2_first;           /* squared radius of the first round */
	int cap, prune;
	NNTuning tune;
	int *nn_out;
	float4 *obs_rec;
	const float4 *src_xyzl;
	unsigned long long *label_counts;
	unsigned long long *eval_counter;
	const int *seg_of_chunk;
	unsigned int *work;         /* [0] next chunk of 32 queries, [1] warps that ran out of work: both zero between launches */
	/* profiling build with M3DREG_NN_DIAG=1 only (tools/nn_tail.py): [0] ~(earliest start), [1] latest chunk end, [2] sum of warp exits,
	 * [3] warps, [4] longest chunk (ns), [5] sum of chunk times, [6] chunks, [7] chunks > 20 us, [8] chunks > 50 us, [9] time in chunks
	 * that searched per lane, [10] such chunks */
	unsigned long long *diag;
};

/* the lane's conservative box of fine columns for dist <= tau (nn_query()'s box), clamped to its 27-neighbourhood */
struct NNBox { int xl, xh, yl, yh, zl, zh; };

/* out of line: two call sites of ~110 instructions each, and the hot path of a chunk should stay near the 32 KB
 * instruction cache (B300_MICROARCH.md: L1.5 I-cache) */
__device__ __noinline__ NNBox nnh_box(float tau, int prune, float qx, float qy, float qz, float mnx, float mny, float mnz,
		float iwx, float iwy, float iwz, int ix, int iy, int iz, int nbx, int nby, int nbz)
{
	/* R >= sqrt(tau) * (1 + 2^-20) is all the proof needs: tau * rsqrt(tau) is within 2^-21 of sqrt(tau) (MUFU.RSQ: 2 ulp),
	 * the factor 1 + 2^-13 covers that with three orders of magnitude to spare and costs a box 0.01 % wider */
	const float R = !prune ? INFINITY : (tau > 1.0e-30f ? f_fma(f_mul(tau, rsqrtf(tau)), 1.0001220703125f, 1.0e-18f) : 1.1e-15f);
	const float mgx = f_fma(f_mul(fabsf(qx) + fabsf(mnx), iwx), 3.814697265625e-06f, 9.765625e-04f);
	const float mgy = f_fma(f_mul(fabsf(qy) + fabsf(mny), iwy), 3.814697265625e-06f, 9.765625e-04f);
	const float mgz = f_fma(f_mul(fabsf(qz) + fabsf(mnz), iwz), 3.814697265625e-06f, 9.765625e-04f);
	NNBox b;
	b.xl = col_floor(f_sub(qx, R), mnx, iwx, mgx); b.xh = col_ceil(f_add(qx, R), mnx, iwx, mgx);
	b.yl = col_floor(f_sub(qy, R), mny, iwy, mgy); b.yh = col_ceil(f_add(qy, R), mny, iwy, mgy);
	b.zl = col_floor(f_sub(qz, R), mnz, iwz, mgz); b.zh = col_ceil(f_add(qz, R), mnz, iwz, mgz);
	/* the 27-neighbourhood with edge clamping (lesson_16.cu:588-608), as a box of fine columns */
	b.xl = max(b.xl, (ix > 0 ? ix - 1 : ix) << 2); b.xh = min(b.xh, ((ix != nbx - 1 ? ix + 1 : ix) << 2) + 3);
	b.yl = max(b.yl, (iy > 0 ? iy - 1 : iy) << 2); b.yh = min(b.yh, ((iy != nby - 1 ? iy + 1 : iy) << 2) + 3);
	b.zl = max(b.zl, (iz > 0 ? iz - 1 : iz) << 2); b.zh = min(b.zh, ((iz != nbz - 1 ? iz + 1 : iz) << 2) + 3);
	return b;
}

/* does fine cell (gx, gy, gz) belong to a bucket of the 27-neighbourhood of home cell (ix, iy, iz)? */
__device__ __forceinline__ bool nnh_in_neighbourhood(int gx, int gy, int gz, int ix, int iy, int iz, int nbx, int nby, int nbz)
{
	return gx >= ((ix > 0 ? ix - 1 : ix) << 2) && gx <= ((ix != nbx - 1 ? ix + 1 : ix) << 2) + 3 &&
			gy >= ((iy > 0 ? iy - 1 : iy) << 2) && gy <= ((iy != nby - 1 ? iy + 1 : iy) << 2) + 3 &&
			gz >= ((iz > 0 ? iz - 1 : iz) << 2) && gz <= ((iz != nbz - 1 ? iz + 1 : iz) << 2) + 3;
}

__device__ __noinline__ int2 nnh_query_fallback(const NNHullArgs *a, float qx, float qy, float qz, int label, int qi)
{
	const m3dreg_grid_params *gp = a->gp;
	NNParams P;
	P.mnx = gp->bounding_box_min_X; P.mny = gp->bounding_box_min_Y; P.mnz = gp->bounding_box_min_Z;
	P.mxx = gp->bounding_box_max_X; P.mxy = gp->bounding_box_max_Y; P.mxz = gp->bounding_box_max_Z;
	P.rx = gp->resolution_X; P.ry = gp->resolution_Y; P.rz = gp->resolution_Z;
	P.nbx = gp->number_of_buckets_X; P.nby = gp->number_of_buckets_Y; P.nbz = gp->number_of_buckets_Z;
	P.nb = gp->number_of_buckets;
	P.buckets = a->buckets;
	P.ci = a->cs; P.co = a->cs;
	P.cap_in = a->cap; P.cap_out = a->cap;
	nn_params_finish(P, a->search_radius, a->prune ? 1 : 0);
	unsigned int ev = 0;
	const float4 pn = __ldg(a->q_nrm + qi);
	const int l = nn_query(P, make_float4(qx, qy, qz, __int_as_float(label)), pn, ev);
	return make_int2(l, (int)ev);
}

/* full predicate of the reference on one candidate (lesson_16.cu:658-686) with the (dist, l) order made explicit;
 * N = its {normal, label} record */
#define M3D_NNH_CONSIDER(D, LBITS, N, J)                                                                               \
	if ((D) <= lim) {                                                                                                  \
		const int l_ = __float_as_int(LBITS);                                                                           \
		if (((D) < best_d || l_ < best_l) && __float_as_int((N).w) == label) {                                         \
			const float dot_ = f_fma(pn.z, (N).z, f_fma(pn.x, (N).x, f_mul(pn.y, (N).y)));                              \
			if (angle_gate(dot_)) { best_d = (D); best_l = l_; best_j = (J); lim = (D); }                              \
		}                                                                                                               \
	}

template <bool COUNT>
__global__ void __launch_bounds__(kNNHThreads, M3D_NNH_MINBLOCKS) k_nn_search_hull(const __grid_constant__ NNHullArgs a)
{
	pdl_enter();
	__shared__ int4 s_segs[kNNHWarps][kNNHCells];               /* {first candidate, count, hull cell x | y << 16, hull cell z} */
	__shared__ float4 s_cand[kNNHWarps][kNNHStage];             /* staged candidates, per group of four: {x0..x3}, {y0..y3}, {z0..z3}, {l0..l3} */
	__shared__ int4 s_grp[kNNHWarps][kNNHStage / 4];            /* per group: {index of its first candidate, valid, hull cell x | y << 16, z} */
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const unsigned lt_mask = (1u << lane) - 1u;
	int4 *segs = s_segs[threadIdx.x >> 5];
	float4 *stage = s_cand[threadIdx.x >> 5];
	float *stagef = reinterpret_cast<float *>(stage);
	int4 *grp = s_grp[threadIdx.x >> 5];
	const m3dreg_grid_params *__restrict__ gp = a.gp;
	const float mnx = gp->bounding_box_min_X, mny = gp->bounding_box_min_Y, mnz = gp->bounding_box_min_Z;
	const int nbx = gp->number_of_buckets_X, nby = gp->number_of_buckets_Y, nbz = gp->number_of_buckets_Z;
	const long long nb = gp->number_of_buckets;
	const int tables = (nn_tables_usable(a.cap, a.cap) && nn_columns_usable(nbx, nby, nbz)) ? 1 : 0;
	const float4 *__restrict__ cx = a.cs.xyzl;
	const float4 *__restrict__ cn = a.cs.nrm;
	const int n_chunks = (a.n_second + 31) >> 5;

	for (;;) {      /* persistent warp: one chunk of 32 queries per trip */
	int chunk = 0;
	if (lane == 0) chunk = (int)atomicAdd(a.work, 1u);
	chunk = __shfl_sync(full, chunk, 0);
	if (chunk >= n_chunks) break;
	const int qi = (chunk << 5) + lane;
	unsigned long long t_chunk = 0;
	bool chunk_fell_back = false;
	if (COUNT && a.diag) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_chunk)); if (lane == 0) atomicMax(a.diag, ~t_chunk); }

	unsigned int evals = 0;
	int best_l = kNNNone, best_j = -1, label = -1;                     /* best_j: the winner's slot in the candidate set (-1: found by nn_query()) */
	float best_d = 100000000.0f;                                        /* lesson_16.cu:597 */
	float lim = fminf(a.r2, 99999992.0f);
	float qx = 0.0f, qy = 0.0f, qz = 0.0f;
	int ix = 0, iy = 0, iz = 0;
	bool active = false;
	if (qi < a.n_second && nb > 0 && a.cap > 0) {
		const float4 p = __ldg(a.q_xyzl + qi);
		qx = p.x; qy = p.y; qz = p.z;
		label = __float_as_int(p.w);
		/* lesson_16.cu:562-583 */
		if (!(qx < mnx || qx > gp->bounding_box_max_X || qy < mny || qy > gp->bounding_box_max_Y || qz < mnz || qz > gp->bounding_box_max_Z)) {
			ix = cell_of(qx, mnx, gp->resolution_X); iy = cell_of(qy, mny, gp->resolution_Y); iz = cell_of(qz, mnz, gp->resolution_Z);
			const int home = ix * nby * nbz + iy * nbz + iz;
			active = home >= 0 && (long long)home < nb && lim >= 0.0f;
		}
	}
	if (!nn_columns_usable(nbx, nby, nbz)) {                            /* warp-uniform */
		if (active) { const int2 fr = nnh_query_fallback(&a, qx, qy, qz, label, qi); best_l = fr.x; evals += (unsigned int)fr.y; }
		active = false;
		if (COUNT) chunk_fell_back = true;
	}

	unsigned todo = __ballot_sync(full, active);
	while (todo) {                                                      /* one pass per label present in the warp (almost always one) */
		const int L = __shfl_sync(full, label, __ffs(todo) - 1);
		const bool mine = active && label == L;
		bool unsettled = mine;
		todo &= ~__ballot_sync(full, mine);
		float rho2 = a.rho2_first;
		/* previous round's hull, packed: low corner + extents (x 8 bits, y and z 12 bits; a larger hull is simply not
		 * remembered — skipping its cells is an optimisation, never needed for the answer) */
		int hxl = 0, hyl = 0, hzl = 0;
		unsigned hext = 0;                                              /* (dx) | (dy << 8) | (dz << 20), 0 = none */
		for (int round = 0; round < 80; round++) {
			if (!__any_sync(full, unsettled)) break;
			NNBox b;
			b.xl = b.yl = b.zl = 0x7fffffff; b.xh = b.yh = b.zh = -0x7fffffff;
			if (unsettled) {
				b = nnh_box(a.prune ? fminf(lim, rho2) : lim, a.prune, qx, qy, qz, mnx, mny, mnz, a.iwx, a.iwy, a.iwz, ix, iy, iz, nbx, nby, nbz);
				if (b.xl > b.xh || b.yl > b.yh || b.zl > b.zh) { b.xl = b.yl = b.zl = 0x7fffffff; b.xh = b.yh = b.zh = -0x7fffffff; }
			}
			const int uxl = __reduce_min_sync(full, b.xl), uxh = __reduce_max_sync(full, b.xh);
			const int uyl = __reduce_min_sync(full, b.yl), uyh = __reduce_max_sync(full, b.yh);
			const int uzl = __reduce_min_sync(full, b.zl), uzh = __reduce_max_sync(full, b.zh);
			if (uxl <= uxh) {
				const int dx = uxh - uxl + 1, dy = uyh - uyl + 1, dz = uzh - uzl + 1;
				{
					const long long ncell = (long long)dx * dy * dz;
					const int own = b.xl <= b.xh ? (b.xh - b.xl + 1) * (b.yh - b.yl + 1) * (b.zh - b.zl + 1) : 0;
					const int own_max = __reduce_max_sync(full, own);
					if (dx > kNNHCells || dy > 32767 || dz > 32767 || (ncell > a.tune.hull_min && ncell > (long long)a.tune.hull_ratio * own_max)) {
						/* scattered warp: per-lane search from scratch */
						if (COUNT) {
							const unsigned fb = __ballot_sync(full, unsettled);
							if (lane == 0) atomicAdd(a.eval_counter + 1, (unsigned long long)__popc(fb));
						}
						if (unsettled) { const int2 fr = nnh_query_fallback(&a, qx, qy, qz, label, qi); best_l = fr.x; best_j = -1; evals += (unsigned int)fr.y; }
						unsettled = false;
						if (COUNT) chunk_fell_back = true;
						break;
					}
				}
				/* may every lane look at every bucket the hull touches? (always, unless the radius exceeds the bucket size) */
				const bool nb_all = __all_sync(full, !mine || (nnh_in_neighbourhood(uxl, uyl, uzl, ix, iy, iz, nbx, nby, nbz) &&
						nnh_in_neighbourhood(uxh, uyh, uzh, ix, iy, iz, nbx, nby, nbz)));
				const int nrows = dy * dz, rpc = __float2int_rz(__fdividef((float)kNNHCells + 0.5f, (float)dx));   /* = kNNHCells / dx for 1 <= dx <= kNNHCells */
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
							/* cells of the previous round's hull were evaluated by every lane then (unsigned compares: below the
							 * low corner wraps to a huge value) */
							const bool in_old = (unsigned)(gx - hxl) < (hext & 0xffu) && (unsigned)(gy - hyl) < ((hext >> 8) & 0xfffu) &&
									(unsigned)(gz - hzl) < (hext >> 20);
							const int cell = ((gx >> 2) * nby + (gy >> 2)) * nbz + (gz >> 2);
							const int *rec = reinterpret_cast<const int *>(a.buckets + cell);
							int npts = 0, begin = -1;
							if (!in_old) { npts = __ldg(rec + 2); begin = __ldg(rec); }
							if (npts > 0 && begin >= 0) {           /* lesson_16.cu:615-616 (also the quirk bucket) */
								const int level = tables ? nn_level(npts) : -1;
								const int sh = level < 0 ? 2 : 2 - level;
								const int am = (1 << sh) - 1;
								/* a bin of a coarser bucket covers 2^sh fine cells per axis: listed from the first one the hull holds (a
								 * skipped representative means the bin touched the previous hull, where it was listed whole) */
								const bool rep = (ax == 0 || !(gx & am)) && (ay == 0 || !(gy & am)) && (az == 0 || !(gz & am));
								if (rep) {
									if (level < 0) {                /* no table: the whole bucket, in walk order */
										const int iter = candidate_stride(npts, a.cap);
										start = begin; cnt = (npts + iter - 1) / iter;
									} else {
										const int bin = nn_bin(L, (gx & 3) >> sh, (gy & 3) >> sh, (gz & 3) >> sh, level);
										const unsigned short *tab = a.cs.tab + 2 * (size_t)begin + bin;
										const int s = __ldg(tab), e = __ldg(tab + 1);
										start = begin + s; cnt = e - s;
									}
									rel_xy = ax | (ay << 16); rel_z = az;
								}
							}
						}
						const unsigned m = __ballot_sync(full, cnt > 0);
						if (cnt > 0) segs[nseg + __popc(m & lt_mask)] = make_int4(start, cnt, rel_xy, rel_z);
						nseg += __popc(m);
					}
					__syncwarp();
					/* 3./4. batches of at most kNNHStage staged candidates */
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
						const unsigned fit = __ballot_sync(full, k0 + lane < nseg && incl <= kNNHStage);
						const int ntake = __popc(fit);                  /* segments are taken in order: fit is a prefix mask */
						int ncand;
						if (ntake == 0) {                               /* the first segment alone exceeds a batch: take a part of it */
							const int4 s0 = segs[k0];
							__syncwarp();
							if (lane == 0) segs[k0] = make_int4(s0.x + kNNHStage, s0.y - kNNHStage, s0.z, s0.w);
#pragma unroll
							for (int g = lane; g < kNNHStage / 4; g += 32) grp[g] = make_int4(s0.x + 4 * g, 4, s0.z, s0.w);
							ncand = kNNHStage;
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
						{
							const unsigned long long qx2 = f2_pack(qx, qx), qy2 = f2_pack(qy, qy), qz2 = f2_pack(qz, qz);
							if (nb_all) {
								if (COUNT && mine) evals += (unsigned int)ncand;
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
									const bool use = nnh_in_neighbourhood(uxl + (gi.z & 0xffff), uyl + (gi.z >> 16), uzl + gi.w, ix, iy, iz, nbx, nby, nbz);
									if (COUNT && mine && use) evals += 4u;
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
						}
						/* full predicate on the winning group: its four {normal, label} records and the query's normal go out
						 * together; a tie between groups or an inadmissible winner needs the re-scan */
						float4 pn = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
						const bool need_pn = __any_sync(full, bg >= 0 || flag);      /* the re-scan below needs it too */
						if (need_pn && mine) pn = __ldg(a.q_nrm + qi);
						if (bg >= 0) {
							const int4 gi = grp[bg];
							const int j = gi.x;
							const float4 n0 = __ldg(cn + j), n1 = __ldg(cn + j + (gi.y > 1 ? 1 : 0)), n2 = __ldg(cn + j + (gi.y > 2 ? 2 : 0)),
									n3 = __ldg(cn + j + (gi.y > 3 ? 3 : 0));
							const float4 X = stage[4 * bg], Y = stage[4 * bg + 1], Z = stage[4 * bg + 2], Lw = stage[4 * bg + 3];
							const float d0 = nn_dist(qx, qy, qz, make_float4(X.x, Y.x, Z.x, 0.0f)), d1 = nn_dist(qx, qy, qz, make_float4(X.y, Y.y, Z.y, 0.0f)),
									d2 = nn_dist(qx, qy, qz, make_float4(X.z, Y.z, Z.z, 0.0f)), d3 = nn_dist(qx, qy, qz, make_float4(X.w, Y.w, Z.w, 0.0f));
							/* padding slots hold +inf coordinates: their distance is +inf or NaN, never <= lim */
							M3D_NNH_CONSIDER(d0, Lw.x, n0, j)
							M3D_NNH_CONSIDER(d1, Lw.y, n1, j + 1)
							M3D_NNH_CONSIDER(d2, Lw.z, n2, j + 2)
							M3D_NNH_CONSIDER(d3, Lw.w, n3, j + 3)
							if (best_d != rb) flag = true;
						}
						if (__any_sync(full, flag)) {
							for (int g = 0; g < ngrp; g++) {
								const int4 gi = grp[g];
								const bool use = flag && mine && (nb_all || nnh_in_neighbourhood(uxl + (gi.z & 0xffff), uyl + (gi.z >> 16), uzl + gi.w, ix, iy, iz, nbx, nby, nbz));
								if (use) {
#pragma unroll
									for (int t = 0; t < 4; t++) {
										const float *src = stagef + (g << 4) + t;
										const float d0 = nn_dist(qx, qy, qz, make_float4(src[0], src[4], src[8], 0.0f));
										if (d0 <= lim) {
											const float4 n0 = __ldg(cn + gi.x + t);
											M3D_NNH_CONSIDER(d0, src[12], n0, gi.x + t)
										}
									}
								}
							}
						}
						__syncwarp();
					}
				}
				hxl = uxl; hyl = uyl; hzl = uzl;
				hext = (dx <= 255 && dy <= 4095 && dz <= 4095) ? ((unsigned)dx | ((unsigned)dy << 8) | ((unsigned)dz << 20)) : 0u;
				/* every cell of this hull has now been evaluated by all lanes (in this round or, for the cells skipped as part
				 * of the previous hull, before): a lane whose box for its CURRENT limit lies inside the hull is done */
				if (unsettled && a.prune && lim > rho2) {
					const NNBox e = nnh_box(lim, a.prune, qx, qy, qz, mnx, mny, mnz, a.iwx, a.iwy, a.iwz, ix, iy, iz, nbx, nby, nbz);
					if (e.xl >= uxl && e.xh <= uxh && e.yl >= uyl && e.yh <= uyh && e.zl >= uzl && e.zh <= uzh) unsettled = false;
				}
			}
			if (unsettled && (!a.prune || lim <= rho2)) unsettled = false;   /* everything at or below the limit was inside this round's box */
			rho2 = f_mul(rho2, 4.0f);
		}
	}

	/* the winner: its record in the candidate set carries the point as stored in the scan (local frame) and its original
	 * index — one gather instead of hash[l] -> cloud[index] (lesson_16.cu:640-647, gpu6DSLAM.cpp:367-369) */
	int result = -1;
	float4 rec = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
	if (best_l != kNNNone && best_l >= 0 && best_l < a.n_first) {
		if (best_j >= 0) {
			rec = __ldg(a.cs.loc + best_j);
			result = __float_as_int(rec.w);
		} else {
			result = (int)__ldg(a.s_vals + best_l);
			if (a.obs_rec) { rec = __ldg(a.src_xyzl + result); rec.w = __int_as_float(result); }
		}
	}
	if (qi < a.n_second) {
		if (a.obs_rec) a.obs_rec[qi] = rec;
		if (a.nn_out) a.nn_out[a.q_perm ? __ldg(a.q_perm + qi) : (uint32_t)qi] = result;
	}
	if (COUNT) {
		const unsigned int tot = __reduce_add_sync(full, evals);
		if (lane == 0 && tot) atomicAdd(a.eval_counter, (unsigned long long)tot);
	}
	if (a.label_counts) {   /* per-label match counts (gpu6DSLAM.cpp:323-357): warp ballots, one atomic per label per warp */
		unsigned long long *lc = a.label_counts;
		if (a.seg_of_chunk) lc += 4 * __ldg(a.seg_of_chunk + (chunk << 5) / kSegChunk);      /* 32 queries never straddle segments */
		const bool hit = qi < a.n_second && result >= 0;
#pragma unroll
		for (int Lb = 0; Lb < 4; Lb++) {
			const unsigned m = __ballot_sync(full, hit && label == Lb);
			if (m && lane == Lb) atomicAdd(&lc[Lb], (unsigned long long)__popc(m));
		}
	}
	if (COUNT && a.diag && lane == 0) {
		unsigned long long t1;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		const unsigned long long dt = t1 - t_chunk;
		atomicMax(a.diag + 4, dt); atomicAdd(a.diag + 5, dt); atomicAdd(a.diag + 6, 1ull);
		if (dt > 20000ull) atomicAdd(a.diag + 7, 1ull);
		if (dt > 50000ull) atomicAdd(a.diag + 8, 1ull);
		if (chunk_fell_back) { atomicAdd(a.diag + 9, dt); atomicAdd(a.diag + 10, 1ull); }
		atomicMax(a.diag + 1, t1);
	}
	}      /* persistent loop */
	if (COUNT && a.diag && lane == 0) {
		unsigned long long t1;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
		atomicAdd(a.diag + 2, t1 - ~a.diag[0]); atomicAdd(a.diag + 3, 1ull);
	}
	/* the last warp to run out of work re-arms the counters for the next launch */
	if (lane == 0) {
		const unsigned int total_warps = gridDim.x * kNNHWarps;
		if (atomicAdd(a.work + 1, 1u) == total_warps - 1) { a.work[0] = 0u; a.work[1] = 0u; __threadfence(); }
	}
}
#undef M3D_NNH_CONSIDER

} /* namespace m3d */
