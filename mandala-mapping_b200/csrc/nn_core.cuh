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

/* ---- candidate layout ------------------------------------------------------------------------------------------ */
struct CandSet {
	float4 *xyzl;            /* {x, y, z, bits of l = position in the sorted table (hashElement index)} */
	float4 *nrm;             /* {nx, ny, nz, label bits}                                              */
	unsigned short *tab;     /* bin offset tables: the bucket with index_begin b owns tab[2b, 2b + 2n) */
	float4 *loc;             /* {x0, y0, z0 as stored in the scan (LOCAL frame in the fused loops), bits of the original index};
	                          * only the device kernels touch it (nn_query() does not) */
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

/* the search addresses fine columns (4 per bucket and axis) with ints */
M3D_HD bool nn_columns_usable(int nbx, int nby, int nbz)
{
	return nbx <= (1 << 27) && nby <= (1 << 27) && nbz <= (1 << 27);
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
	float iwx, iwy, iwz;       /* 4 / res: fine (level-2) columns per metre, only used for conservative column boxes */
	int nbx, nby, nbz;
	long long nb;
	const m3dreg_bucket *buckets;
	CandSet ci, co;            /* INNER (home bucket) and OUTER (neighbours) sets; the same set when the caps are equal */
	int cap_in, cap_out;
	int tables;                /* nn_tables_usable(cap_in, cap_out, nbx, nby, nbz) */
	int prune;                 /* 0: one round over the whole neighbourhood (equivalence test only) */
	float r2;                  /* fl(radius * radius) */
	float rho2_first;          /* squared radius of the first non-trivial round */
};

constexpr int kNNNone = 0x7fffffff;

/* derived fields of NNParams from the grid fields, the caps and the search radius (same code on device and host) */
M3D_HD void nn_params_finish(NNParams &P, float search_radius, int prune)
{
	P.tables = (nn_tables_usable(P.cap_in, P.cap_out) && nn_columns_usable(P.nbx, P.nby, P.nbz)) ? 1 : 0;
	P.prune = prune;
	P.r2 = f_mul(search_radius, search_radius);                         /* lesson_16.cu:553 */
	P.iwx = f_div(4.0f, P.rx); P.iwy = f_div(4.0f, P.ry); P.iwz = f_div(4.0f, P.rz);
	const float rmin = f_mul(fminf(P.rx, fminf(P.ry, P.rz)), 0.0625f);
	P.rho2_first = fmaxf(f_mul(rmin, rmin), 1.0e-30f);
}

/* Conservative fine-column bounds of a coordinate interval.  The fine column of a stored candidate with coordinate v is
 * U(v) = 4 * cell_of(v) + (thresholds reached inside the cell); its cell boundary and thresholds lie within a few ulps
 * (relative 2^-22 of the magnitudes involved) of the ideal lattice mn + U * res/4.  t = fl(fl(v - mn) * iw) carries a
 * comparable error, so with the margin m = |t| * 2^-18 + mg, mg = (|q| + |mn|) * iw * 2^-18 + 2^-10 (32x the worst
 * case), every candidate with v >= lo has U(v) >= col_floor(lo) and every candidate with v <= hi has
 * U(v) <= col_ceil(hi).  Only a SUPERSET of the bins is needed for exactness, never the exact set. */
M3D_HD int col_floor(float lo, float mn, float iw, float mg)
{
	float t = f_mul(f_sub(lo, mn), iw);
	t = t - (fabsf(t) * 3.814697265625e-06f + mg);
	return (int)fminf(fmaxf(t, -1.0f), 1.0e9f);        /* negative values truncate towards 0 / clamp to -1: callers clamp to >= 0 */
}
M3D_HD int col_ceil(float hi, float mn, float iw, float mg)
{
	float t = f_mul(f_sub(hi, mn), iw);
	t = t + (fabsf(t) * 3.814697265625e-06f + mg);
	return (int)fminf(fmaxf(t, -1.0f), 1.0e9f);
}

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

	float best_d = 100000000.0f;                                        /* lesson_16.cu:597 */
	int best_l = kNNNone;
	float lim = fminf(P.r2, 99999992.0f);                               /* a candidate can only matter if dist <= lim (<= r2, < 1e8) */
	if (!(lim >= 0.0f)) return kNNNone;

	/* the 27-neighbourhood with edge clamping (lesson_16.cu:588-608), as a box of fine columns */
	const int cx0 = (ix > 0 ? ix - 1 : ix) << 2, cx1 = ((ix != P.nbx - 1 ? ix + 1 : ix) << 2) + 3;
	const int cy0 = (iy > 0 ? iy - 1 : iy) << 2, cy1 = ((iy != P.nby - 1 ? iy + 1 : iy) << 2) + 3;
	const int cz0 = (iz > 0 ? iz - 1 : iz) << 2, cz1 = ((iz != P.nbz - 1 ? iz + 1 : iz) << 2) + 3;
	const float mgx = f_fma(f_mul(fabsf(qx) + fabsf(P.mnx), P.iwx), 3.814697265625e-06f, 9.765625e-04f);
	const float mgy = f_fma(f_mul(fabsf(qy) + fabsf(P.mny), P.iwy), 3.814697265625e-06f, 9.765625e-04f);
	const float mgz = f_fma(f_mul(fabsf(qz) + fabsf(P.mnz), P.iwz), 3.814697265625e-06f, 9.765625e-04f);

	int oxl = 1, oxh = 0, oyl = 1, oyh = 0, ozl = 1, ozh = 0;          /* fine-column box of the previous round (empty) */
	float rho2 = P.rho2_first;
	for (int round = 0; round < 80; round++) {
		/* every candidate with dist <= tau has |fl(q - c)| <= sqrt(tau) up to two roundings per axis (dist >= fl(d*d)),
		 * hence lies in [q - R, q + R] with R rounded outwards and a 2^-20 margin */
		const float tau = P.prune ? fminf(lim, rho2) : lim;
		const float R = P.prune ? f_add_up(f_mul_up(f_sqrt_up(tau), 1.00000095367431640625f), 1.0e-18f) : INFINITY;
		int xl = col_floor(f_sub(qx, R), P.mnx, P.iwx, mgx), xh = col_ceil(f_add(qx, R), P.mnx, P.iwx, mgx);
		int yl = col_floor(f_sub(qy, R), P.mny, P.iwy, mgy), yh = col_ceil(f_add(qy, R), P.mny, P.iwy, mgy);
		int zl = col_floor(f_sub(qz, R), P.mnz, P.iwz, mgz), zh = col_ceil(f_add(qz, R), P.mnz, P.iwz, mgz);
		xl = xl > cx0 ? xl : cx0; xh = xh < cx1 ? xh : cx1;
		yl = yl > cy0 ? yl : cy0; yh = yh < cy1 ? yh : cy1;
		zl = zl > cz0 ? zl : cz0; zh = zh < cz1 ? zh : cz1;
		const bool have_old = oxl <= oxh;

		for (int bx = xl >> 2; bx <= (xh >> 2); bx++)
		for (int by = yl >> 2; by <= (yh >> 2); by++)
		for (int bz = zl >> 2; bz <= (zh >> 2); bz++) {
			const bool inner = bx == ix && by == iy && bz == iz;
			const int cap = inner ? P.cap_in : P.cap_out;                /* lesson_16.cu:618-626 */
			if (cap <= 0) continue;
			const int cell = (bx * P.nby + by) * P.nbz + bz;
			const int *rec = reinterpret_cast<const int *>(P.buckets + cell);
			const int npts = M3D_LDG(rec + 2);
			if (npts <= 0) continue;                                      /* lesson_16.cu:615-616 (also the quirk bucket) */
			const int begin = M3D_LDG(rec);
			if (begin < 0) continue;
			const float4 *cx = inner ? P.ci.xyzl : P.co.xyzl;
			const float4 *cn = inner ? P.ci.nrm : P.co.nrm;
			const int level = P.tables ? nn_level(npts) : -1;
			const int fx = bx << 2, fy = by << 2, fz = bz << 2;         /* first fine column of this bucket */
			/* this bucket's part of the previous round's box, in fine columns */
			const int pxl = oxl > fx ? oxl : fx, pxh = oxh < fx + 3 ? oxh : fx + 3;
			const int pyl = oyl > fy ? oyl : fy, pyh = oyh < fy + 3 ? oyh : fy + 3;
			const int pzl = ozl > fz ? ozl : fz, pzh = ozh < fz + 3 ? ozh : fz + 3;
			const bool b_old = have_old && pxl <= pxh && pyl <= pyh && pzl <= pzh;   /* visited before (in part) */
			int xlo = 0, xhi = 0, ylo = 0, yhi = 0, zlo = 0, zhi = 0;
			int oxlo = 0, oxhi = -1, oylo = 0, oyhi = -1, ozlo = 0, ozhi = -1;
			int flat_n = 0, bin_l = 0;
			const unsigned short *tab = nullptr;
			if (level < 0) {
				if (b_old) continue;                                      /* a bucket without a table is looked at as a whole */
				const int iter = candidate_stride(npts, cap);
				flat_n = (npts + iter - 1) / iter;
			} else {
				/* columns of this bucket at its own level: a level-L column is 2^(2-L) fine columns (identical thresholds) */
				const int sh = 2 - level;
				xlo = ((xl > fx ? xl : fx) - fx) >> sh; xhi = ((xh < fx + 3 ? xh : fx + 3) - fx) >> sh;
				ylo = ((yl > fy ? yl : fy) - fy) >> sh; yhi = ((yh < fy + 3 ? yh : fy + 3) - fy) >> sh;
				zlo = ((zl > fz ? zl : fz) - fz) >> sh; zhi = ((zh < fz + 3 ? zh : fz + 3) - fz) >> sh;
				if (b_old) {     /* columns visited by the previous round (all their candidates were looked at) */
					oxlo = (pxl - fx) >> sh; oxhi = (pxh - fx) >> sh;
					oylo = (pyl - fy) >> sh; oyhi = (pyh - fy) >> sh;
					ozlo = (pzl - fz) >> sh; ozhi = (pzh - fz) >> sh;
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
						if ((D) <= lim) {                                                                                      \
							/* D <= lim <= best_d: a strictly smaller distance, or the same distance at a smaller position */   \
							const int l_ = f_bits((C).w);                                                                       \
							if ((D) < best_d || l_ < best_l) {                                                                 \
								const float4 n_ = M3D_LDG(cn + (J));                                                            \
								if (f_bits(n_.w) == label) {                                                                    \
									const float dot_ = f_fma(pn.z, n_.z, f_fma(pn.x, n_.x, f_mul(pn.y, n_.y)));                 \
									if (angle_gate(dot_)) { best_d = (D); best_l = l_; lim = (D); }                            \
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
		if (!P.prune || lim <= rho2) break;      /* everything at or below the limit was inside this round's box */
		oxl = xl; oxh = xh; oyl = yl; oyh = yh; ozl = zl; ozh = zh;
		rho2 = f_mul(rho2, 4.0f);
	}
	return best_l;
}

} /* namespace m3d */
