/* nn_core.cuh — the semantic nearest-neighbour search of ONE query over sub-binned candidate sets.
 *
 * Written as __host__ __device__ code on purpose: the device kernel (k_nn_search, m3dreg_kernels.cuh) calls
 * nn_query() once per thread, and tests/csrc/nn_emul.cpp compiles the very same function with g++ to check the
 * search logic against the CPU oracle without a GPU.  The product never runs the host instantiation.
 *
 * Reference semantics (kernel_semanticNearestNeighborSearch, src/lesson_16.cu:531-703): for the query's home
 * bucket and its (edge-clamped) 26 neighbours, walk sorted positions l = begin, begin+s, ... with
 * s = n / cap (cap = INNER for the home bucket, OUTER otherwise; lesson_16.cu:618-640) and keep a candidate iff
 *      label equal  &&  |acos(n_p . n_q) * 180/pi| < 90  &&  dist <= r^2  &&  dist < best      (lesson_16.cu:658-686)
 * Visit order is ascending l and the update is a strict '<', so the answer is the LEXICOGRAPHIC MINIMUM of (dist, l)
 * over the admissible candidates.  Carrying that pair makes the answer independent of the order in which candidates
 * are looked at, which is what allows the layout below.
 *
 * Candidate layout (built every iteration by k_build_candidates): only the <= 2*cap-1 candidates per bucket the
 * reference would ever look at are kept, in the bucket's own [begin, begin+ncand) range.  A bucket with enough points
 * is cut into S x S x S sub-cells (S = 4, 2 or 1) and its candidates are stored grouped by
 *      bin = (label & 3) * S^3 + (uz * S + uy) * S + ux
 * with a table of exclusive bin offsets (u16) living in the bucket's own range of a side array.  Sub-cell columns
 * are defined by EXACT float comparisons against thresholds b_k = fma(float(c*S + k), res/S, min) that the build and
 * the search compute with the same expression, so "candidate in column u" is an exact statement about its coordinate.
 *
 * Search: rounds of growing radius.  In a round with limit tau (= min(current limit, rho^2)) every candidate with
 * dist <= tau lies inside the axis-aligned box q +- R, R >= sqrt(tau) (rounded outwards); because cell and column
 * functions are monotone in the coordinate, the bins that can hold such a candidate are exactly a box of bins, and the
 * bins of the previous round's box are skipped.  The search stops when the limit no longer exceeds rho^2: then every
 * candidate at or below the limit has been looked at.  Nothing here is approximate: results are bit-identical to the
 * reference on every parity case (tests/test_gpu_stages.py, tests/test_nn_emul.py). */
#pragma once
#include <stdint.h>
#include <math.h>
#include <vector_types.h>
#include "../../include/m3dreg.h"

#if defined(__CUDACC__)
#define M3D_HD __host__ __device__ __forceinline__
#else
#define M3D_HD inline
#endif

namespace m3d {

/* ---- float operations with the reference's roundings (device: explicit intrinsics, host: IEEE ops compiled with
 *      -ffp-contract=off; directed roundings are widened by one ulp on the host, which only makes pruning weaker) ---- */
#ifdef __CUDA_ARCH__
M3D_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
M3D_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
M3D_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
M3D_HD float f_div(float a, float b) { return __fdiv_rn(a, b); }
M3D_HD float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
M3D_HD float f_add_up(float a, float b) { return __fadd_ru(a, b); }
M3D_HD float f_sub_up(float a, float b) { return __fsub_ru(a, b); }
M3D_HD float f_sub_dn(float a, float b) { return __fsub_rd(a, b); }
M3D_HD float f_mul_up(float a, float b) { return __fmul_ru(a, b); }
M3D_HD float f_mul_dn(float a, float b) { return __fmul_rd(a, b); }
M3D_HD float f_sqrt_up(float a) { return __fsqrt_ru(a); }
M3D_HD int f_bits(float a) { return __float_as_int(a); }
M3D_HD float f_from_bits(int a) { return __int_as_float(a); }
#define M3D_LDG(p) __ldg(p)
#else
M3D_HD float f_add(float a, float b) { volatile float r = a + b; return r; }
M3D_HD float f_sub(float a, float b) { volatile float r = a - b; return r; }
M3D_HD float f_mul(float a, float b) { volatile float r = a * b; return r; }
M3D_HD float f_div(float a, float b) { volatile float r = a / b; return r; }
M3D_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
M3D_HD float f_add_up(float a, float b) { return nextafterf(f_add(a, b), INFINITY); }
M3D_HD float f_sub_up(float a, float b) { return nextafterf(f_sub(a, b), INFINITY); }
M3D_HD float f_sub_dn(float a, float b) { return nextafterf(f_sub(a, b), -INFINITY); }
M3D_HD float f_mul_up(float a, float b) { return nextafterf(f_mul(a, b), INFINITY); }
M3D_HD float f_mul_dn(float a, float b) { return nextafterf(f_mul(a, b), -INFINITY); }
M3D_HD float f_sqrt_up(float a) { return nextafterf(sqrtf(a), INFINITY); }
M3D_HD int f_bits(float a) { union { float f; int i; } u; u.f = a; return u.i; }
M3D_HD float f_from_bits(int a) { union { float f; int i; } u; u.i = a; return u.f; }
#define M3D_LDG(p) (*(p))
#endif

/* cell index along one axis: sub.f32, div.rn.f32, cvt.rzi.s32.f32 (lesson_16.cu:124-126, 578-580) */
M3D_HD int cell_of(float v, float mn, float res)
{
	return (int)f_div(f_sub(v, mn), res);
}

/* stride of the candidate walk inside a bucket (lesson_16.cu:628-635) */
M3D_HD int candidate_stride(int npts, int cap)
{
	int iter = 1;
	if (cap < npts) { iter = npts / cap; if (iter <= 0) iter = 1; }
	return iter;
}

/* The reference's angle gate (lesson_16.cu:666-676): acos(dot)*180.0f/M_PI, |.| < 90.0f, with acos the CUDA float
 * acosf.  As a function of the f32 dot product its accepted set is exactly the interval
 *      0x328885AC (1.589327e-08) <= dot <= 1.0f
 * (NaN, |dot| > 1, zero and negative dots are rejected; tiny positive dots still round to >= 90.0f degrees).
 * tests/test_gpu_gate.py proves this equal to the upstream expression for ALL 2^32 float bit patterns on the device,
 * so the two compares below are bit-exact and ~60 instructions (acosf + an f64 divide) cheaper. */
constexpr uint32_t kAngleGateMinBits = 0x328885ACu;
M3D_HD bool angle_gate(float dot)
{
	return dot >= f_from_bits((int)kAngleGateMinBits) && dot <= 1.0f;
}

/* dist = fma(dz,dz, fma(dx,dx, dy*dy)) (lesson_16.cu:658-660, PTX-verified association) */
M3D_HD float nn_dist(float qx, float qy, float qz, const float4 &c)
{
	float dx = f_sub(qx, c.x), dy = f_sub(qy, c.y), dz = f_sub(qz, c.z);
	return f_fma(dz, dz, f_fma(dx, dx, f_mul(dy, dy)));
}

/* Conservative per-axis gap between the query and the slab of cells at offset -1 / +1 (rounded DOWN to float).
 * A point stored in cell c satisfies trunc(fl(fl(v-min)/res)) == c; with two roundings of relative error 2^-24,
 * (v-min) < ix*res*(1+2^-21) for cells <= ix-1 and (v-min) >= (ix+1)*res*(1-2^-21) for cells >= ix+1.
 * A factor 2^-20 is used.  fl(q - v) is the correctly rounded true difference, rounding is monotone and the gap
 * is a float, so |fl(q-v)| >= gap for every point of those cells. */
M3D_HD void axis_gaps(float q, float mn, float res, int ic, float &g_lo, float &g_hi)
{
	const float up_f = 1.00000095367431640625f, dn_f = 0.99999904632568359375f;     /* 1 +- 2^-20 */
	float up = f_mul_up(f_mul_up((float)ic, res), up_f);            /* exclusive upper bound of cells <= ic-1, rounded up   */
	float lo = f_mul_dn(f_mul_dn((float)(ic + 1), res), dn_f);      /* inclusive lower bound of cells >= ic+1, rounded down */
	g_lo = fmaxf(0.0f, f_sub_dn(f_sub_dn(q, mn), up));
	g_hi = fmaxf(0.0f, f_sub_dn(lo, f_sub_up(q, mn)));
}

/* ---- candidate layout ------------------------------------------------------------------------------------------ */
struct CandSet {
	float4 *xyzl;            /* {x, y, z, bits of l = position in the sorted table (hashElement index)} */
	float4 *nrm;             /* {nx, ny, nz, label bits}                                              */
	unsigned short *tab;     /* bin offset tables: the bucket with index_begin b owns tab[2b, 2b + 2n) */
};

/* log2 of the sub-cells per axis for a bucket with npts points, -1 = no table (candidates stored in walk order).
 * The table has 4*S^3 + 1 u16 entries and must fit the 2*npts entries the bucket owns. */
M3D_HD int nn_level(int npts)
{
	return npts >= 129 ? 2 : (npts >= 17 ? 1 : (npts >= 3 ? 0 : -1));
}

/* u16 offsets can address a bucket's candidates only while 2*cap - 1 <= 65535 */
M3D_HD bool nn_tables_usable(int cap_inner, int cap_outer)
{
	int c = cap_inner > cap_outer ? cap_inner : cap_outer;
	return c <= 32768;
}

M3D_HD float nn_subcell_width(float res, int level)
{
	return level == 2 ? f_mul(res, 0.25f) : (level == 1 ? f_mul(res, 0.5f) : res);
}

/* column of coordinate v inside the bucket whose cell index along this axis is c: the number of thresholds
 * b_k = fma(float(c*S + k), w, min), k = 1..S-1, that v reaches.  Monotone in v. */
M3D_HD int nn_col(float v, float mn, float w, int c, int level)
{
	const int S = 1 << level, cb = c << level;
	int u = 0;
	for (int k = 1; k < S; k++) u += (v >= f_fma((float)(cb + k), w, mn)) ? 1 : 0;
	return u;
}

M3D_HD int nn_bin(int label, int ux, int uy, int uz, int level)
{
	return ((label & 3) << (3 * level)) + (((uz << level) + uy) << level) + ux;
}

/* ---- one query --------------------------------------------------------------------------------------------------- */
struct NNParams {
	float mnx, mny, mnz, mxx, mxy, mxz, rx, ry, rz;
	int nbx, nby, nbz;
	long long nb;
	const m3dreg_bucket *buckets;
	CandSet ci, co;            /* INNER (home bucket) and OUTER (neighbours) sets; the same set when the caps are equal */
	int cap_in, cap_out;
	int tables;                /* nn_tables_usable(cap_in, cap_out) */
	int prune;                 /* 0: one round over the whole neighbourhood (equivalence test only) */
	float r2;                  /* fl(radius * radius) */
	float rho2_first;          /* squared radius of the first non-trivial round */
};

constexpr int kNNNone = 0x7fffffff;

struct NNBest {
	float best_d;
	int best_l;
	float lim;                 /* a candidate can only matter if dist <= lim */
};

/* Returns the sorted position l of the reference's answer, or kNNNone.  p = {x,y,z,label bits}, pn = normal. */
M3D_HD int nn_query(const NNParams &P, const float4 &p, const float4 &pn, unsigned int &evals)
{
	const float qx = p.x, qy = p.y, qz = p.z;
	const int label = f_bits(p.w);
	/* lesson_16.cu:562-576 */
	if (qx < P.mnx || qx > P.mxx || qy < P.mny || qy > P.mxy || qz < P.mnz || qz > P.mxz) return kNNNone;
	const int ix = cell_of(qx, P.mnx, P.rx), iy = cell_of(qy, P.mny, P.ry), iz = cell_of(qz, P.mnz, P.rz);
	const int home = ix * P.nby * P.nbz + iy * P.nbz + iz;
	if (!(home >= 0 && (long long)home < P.nb)) return kNNNone;         /* lesson_16.cu:583 */

	NNBest b;
	b.best_d = 100000000.0f;                                            /* lesson_16.cu:597 */
	b.best_l = kNNNone;
	b.lim = fminf(P.r2, 99999992.0f);                                   /* dist <= r2 && dist < 1e8 */
	if (!(b.lim >= 0.0f)) return kNNNone;

	float gxl, gxh, gyl, gyh, gzl, gzh;
	axis_gaps(qx, P.mnx, P.rx, ix, gxl, gxh);
	axis_gaps(qy, P.mny, P.ry, iy, gyl, gyh);
	axis_gaps(qz, P.mnz, P.rz, iz, gzl, gzh);
	const bool has_xl = ix > 0, has_xh = ix != P.nbx - 1;               /* edge clamping, lesson_16.cu:588-595 */
	const bool has_yl = iy > 0, has_yh = iy != P.nby - 1;
	const bool has_zl = iz > 0, has_zh = iz != P.nbz - 1;

	float R_old = -1.0f;
	float rho2 = 0.0f;
	for (int round = 0; round < 80; round++) {
		/* every candidate with dist <= tau has |fl(q - c)| <= sqrt(tau) up to two roundings per axis (dist >= fl(d*d)),
		 * hence lies in [q - R, q + R] with R rounded outwards and a 2^-20 margin */
		const float tau = P.prune ? fminf(b.lim, rho2) : b.lim;
		const float R = P.prune ? f_add_up(f_mul_up(f_sqrt_up(tau), 1.00000095367431640625f), 1.0e-18f) : INFINITY;
		const float lox = f_sub_dn(qx, R), hix = f_add_up(qx, R);
		const float loy = f_sub_dn(qy, R), hiy = f_add_up(qy, R);
		const float loz = f_sub_dn(qz, R), hiz = f_add_up(qz, R);
		const bool have_old = R_old >= 0.0f;
		const float olox = f_sub_dn(qx, R_old), ohix = f_add_up(qx, R_old);
		const float oloy = f_sub_dn(qy, R_old), ohiy = f_add_up(qy, R_old);
		const float oloz = f_sub_dn(qz, R_old), ohiz = f_add_up(qz, R_old);
		const int dxlo = (has_xl && gxl <= R) ? -1 : 0, dxhi = (has_xh && gxh <= R) ? 1 : 0;
		const int dylo = (has_yl && gyl <= R) ? -1 : 0, dyhi = (has_yh && gyh <= R) ? 1 : 0;
		const int dzlo = (has_zl && gzl <= R) ? -1 : 0, dzhi = (has_zh && gzh <= R) ? 1 : 0;
		const int odxlo = (has_xl && gxl <= R_old) ? -1 : 0, odxhi = (has_xh && gxh <= R_old) ? 1 : 0;
		const int odylo = (has_yl && gyl <= R_old) ? -1 : 0, odyhi = (has_yh && gyh <= R_old) ? 1 : 0;
		const int odzlo = (has_zl && gzl <= R_old) ? -1 : 0, odzhi = (has_zh && gzh <= R_old) ? 1 : 0;

		for (int dx = dxlo; dx <= dxhi; dx++)
		for (int dy = dylo; dy <= dyhi; dy++)
		for (int dz = dzlo; dz <= dzhi; dz++) {
			const bool inner = (dx | dy | dz) == 0;
			const int cap = inner ? P.cap_in : P.cap_out;                /* lesson_16.cu:618-626 */
			if (cap <= 0) continue;
			const int cell = home + (dx * P.nby + dy) * P.nbz + dz;
			const int *rec = reinterpret_cast<const int *>(P.buckets + cell);
			const int npts = M3D_LDG(rec + 2);
			if (npts <= 0) continue;                                      /* lesson_16.cu:615-616 (also the quirk bucket) */
			const int begin = M3D_LDG(rec);
			if (begin < 0) continue;
			const bool b_old = have_old && dx >= odxlo && dx <= odxhi && dy >= odylo && dy <= odyhi && dz >= odzlo && dz <= odzhi;
			const float4 *cx = inner ? P.ci.xyzl : P.co.xyzl;
			const float4 *cn = inner ? P.ci.nrm : P.co.nrm;
			const int level = P.tables ? nn_level(npts) : -1;
			int xlo = 0, xhi = 0, ylo = 0, yhi = 0, zlo = 0, zhi = 0;
			int oxlo = 0, oxhi = -1, oylo = 0, oyhi = -1, ozlo = 0, ozhi = -1;
			int flat_n = 0, bin_l = 0;
			const unsigned short *tab = nullptr;
			if (level < 0) {
				if (b_old) continue;
				const int iter = candidate_stride(npts, cap);
				flat_n = (npts + iter - 1) / iter;
			} else {
				const float wx = nn_subcell_width(P.rx, level), wy = nn_subcell_width(P.ry, level), wz = nn_subcell_width(P.rz, level);
				xlo = nn_col(lox, P.mnx, wx, ix + dx, level); xhi = nn_col(hix, P.mnx, wx, ix + dx, level);
				ylo = nn_col(loy, P.mny, wy, iy + dy, level); yhi = nn_col(hiy, P.mny, wy, iy + dy, level);
				zlo = nn_col(loz, P.mnz, wz, iz + dz, level); zhi = nn_col(hiz, P.mnz, wz, iz + dz, level);
				if (b_old) {
					oxlo = nn_col(olox, P.mnx, wx, ix + dx, level); oxhi = nn_col(ohix, P.mnx, wx, ix + dx, level);
					oylo = nn_col(oloy, P.mny, wy, iy + dy, level); oyhi = nn_col(ohiy, P.mny, wy, iy + dy, level);
					ozlo = nn_col(oloz, P.mnz, wz, iz + dz, level); ozhi = nn_col(ohiz, P.mnz, wz, iz + dz, level);
				}
				tab = (inner ? P.ci.tab : P.co.tab) + 2 * (size_t)begin;
				bin_l = (label & 3) << (3 * level);
			}
			for (int uz = zlo; uz <= zhi; uz++)
			for (int uy = ylo; uy <= yhi; uy++) {
				/* up to two runs of candidates per row of bins: the row's part of the new box minus the old box */
				int s0 = 0, e0 = flat_n, s1 = 0, e1 = 0;
				if (level >= 0) {
					const int row = bin_l + (((uz << level) + uy) << level);
					const bool row_old = b_old && uz >= ozlo && uz <= ozhi && uy >= oylo && uy <= oyhi;
					if (!row_old) {
						s0 = M3D_LDG(tab + row + xlo); e0 = M3D_LDG(tab + row + xhi + 1);
					} else {
						s0 = e0 = 0;
						if (xlo < oxlo) { s0 = M3D_LDG(tab + row + xlo); e0 = M3D_LDG(tab + row + oxlo); }
						if (oxhi < xhi) { s1 = M3D_LDG(tab + row + oxhi + 1); e1 = M3D_LDG(tab + row + xhi + 1); }
					}
				}
				for (int seg = 0; seg < 2; seg++) {
					const int s = seg ? s1 : s0, e = seg ? e1 : e0;
					evals += (unsigned int)(e > s ? e - s : 0);
					for (int i = s; i < e; i += 4) {
						/* four candidates in flight; indices past the run repeat its last candidate (harmless) */
						const int j0 = begin + i, j1 = begin + (i + 1 < e ? i + 1 : e - 1), j2 = begin + (i + 2 < e ? i + 2 : e - 1),
								j3 = begin + (i + 3 < e ? i + 3 : e - 1);
						const float4 c0 = M3D_LDG(cx + j0), c1 = M3D_LDG(cx + j1), c2 = M3D_LDG(cx + j2), c3 = M3D_LDG(cx + j3);
						const float d0 = nn_dist(qx, qy, qz, c0), d1 = nn_dist(qx, qy, qz, c1), d2 = nn_dist(qx, qy, qz, c2),
								d3 = nn_dist(qx, qy, qz, c3);
#define M3D_NN_CONSIDER(D, C, J)                                                                                       \
						if ((D) <= b.lim) {                                                                                    \
							/* D <= lim <= best_d: a strictly smaller distance, or the same distance at a smaller position */   \
							const int l_ = f_bits((C).w);                                                                       \
							if ((D) < b.best_d || l_ < b.best_l) {                                                             \
								const float4 n_ = M3D_LDG(cn + (J));                                                            \
								if (f_bits(n_.w) == label) {                                                                    \
									const float dot_ = f_fma(pn.z, n_.z, f_fma(pn.x, n_.x, f_mul(pn.y, n_.y)));                 \
									if (angle_gate(dot_)) { b.best_d = (D); b.best_l = l_; b.lim = (D); }                      \
								}                                                                                               \
							}                                                                                                   \
						}
						M3D_NN_CONSIDER(d0, c0, j0)
						M3D_NN_CONSIDER(d1, c1, j1)
						M3D_NN_CONSIDER(d2, c2, j2)
						M3D_NN_CONSIDER(d3, c3, j3)
#undef M3D_NN_CONSIDER
					}
				}
			}
		}
		if (!P.prune || b.lim <= rho2) break;      /* everything at or below the limit was inside this round's box */
		R_old = R;
		rho2 = rho2 > 0.0f ? f_mul(rho2, 4.0f) : P.rho2_first;
	}
	return b.best_l;
}

} /* namespace m3d */
