/* nn_emul.cpp — TEST-ONLY host instantiation of the device search logic (mandala-mapping_b200/csrc/nn_core.cuh).
 *
 * nn_query() is __host__ __device__; this file compiles it with g++ (-ffp-contract=off), builds the candidate layout
 * serially exactly as k_build_candidates specifies it (walk order, bins, u16 offset tables inside the bucket's own
 * range) and runs one query after the other, so the search logic, the layout contract and the pruning proofs can be
 * checked against the CPU oracle on a machine without a GPU (tests/test_nn_emul.py).  The product never links this. */
#include <vector>
#include <algorithm>
#include <cstring>
#include <cstdio>
#include <vector_types.h>
#include <vector_functions.h>

#include "../../mandala-mapping_b200/csrc/nn_core.cuh"

using namespace m3d;

namespace {

struct HostSet {
	std::vector<float4> xyzl, nrm;
	std::vector<unsigned short> tab;
	CandSet view() { CandSet s; s.xyzl = xyzl.data(); s.nrm = nrm.data(); s.tab = tab.data(); return s; }
};

void build_set(HostSet &hs, const m3dreg_point *first, int n1, const m3dreg_hash_element *table, const m3dreg_bucket *buckets,
		const m3dreg_grid_params *gp, int cap, int tables)
{
	hs.xyzl.assign((size_t)n1 + 8, make_float4(NAN, NAN, NAN, 0.0f));     /* untouched slots must never be read */
	hs.nrm.assign((size_t)n1 + 8, make_float4(NAN, NAN, NAN, 0.0f));
	hs.tab.assign(2 * (size_t)n1 + 16, 0xFFFF);
	const int nby = gp->number_of_buckets_Y, nbz = gp->number_of_buckets_Z;
	for (long long c = 0; c < gp->number_of_buckets; c++) {
		const int begin = buckets[c].index_begin, npts = buckets[c].number_of_points;
		if (npts <= 0 || begin < 0 || cap <= 0) continue;
		const int iter = candidate_stride(npts, cap);
		const int ncand = (npts + iter - 1) / iter;
		const int level = tables ? nn_level(npts) : -1;
		const int cx = (int)(c / (nby * nbz)), cy = (int)((c / nbz) % nby), cz = (int)(c % nbz);
		auto rec = [&](int k, float4 &px, float4 &pn) {
			const m3dreg_point &p = first[table[begin + k * iter].index_of_point];
			px = make_float4(p.x, p.y, p.z, f_from_bits(begin + k * iter));
			pn = make_float4(p.normal_x, p.normal_y, p.normal_z, f_from_bits(p.label));
			return p.label;
		};
		if (level < 0) {
			for (int k = 0; k < ncand; k++) rec(k, hs.xyzl[(size_t)begin + k], hs.nrm[(size_t)begin + k]);
			continue;
		}
		const int nbins = 4 << (3 * level);
		const float wx = nn_subcell_width(gp->resolution_X, level), wy = nn_subcell_width(gp->resolution_Y, level),
				wz = nn_subcell_width(gp->resolution_Z, level);
		std::vector<int> bin((size_t)ncand), count((size_t)nbins + 1, 0);
		for (int k = 0; k < ncand; k++) {
			float4 px, pn;
			int label = rec(k, px, pn);
			bin[(size_t)k] = nn_bin(label, nn_col(px.x, gp->bounding_box_min_X, wx, cx, level), nn_col(px.y, gp->bounding_box_min_Y, wy, cy, level),
					nn_col(px.z, gp->bounding_box_min_Z, wz, cz, level), level);
			count[(size_t)bin[(size_t)k]]++;
		}
		std::vector<int> off((size_t)nbins + 1, 0);
		for (int b = 0; b < nbins; b++) off[(size_t)b + 1] = off[(size_t)b] + count[(size_t)b];
		for (int b = 0; b <= nbins; b++) hs.tab[2 * (size_t)begin + b] = (unsigned short)off[(size_t)b];
		std::vector<int> run(off);
		for (int k = 0; k < ncand; k++) {
			int pos = begin + run[(size_t)bin[(size_t)k]]++;
			rec(k, hs.xyzl[(size_t)pos], hs.nrm[(size_t)pos]);
		}
	}
}

} /* namespace */

extern "C" int emul_nn_search(const m3dreg_point *first, int n1, const m3dreg_point *second, int n2,
		const m3dreg_hash_element *table, const m3dreg_bucket *buckets, const m3dreg_grid_params *gp,
		float radius, int max_inner, int max_outer, int prune, int *nn_out, long long *evals_out)
{
	const int tables = nn_tables_usable(max_inner, max_outer) ? 1 : 0;
	HostSet si, so;
	build_set(si, first, n1, table, buckets, gp, max_inner, tables);
	const bool two = max_inner != max_outer;
	if (two) build_set(so, first, n1, table, buckets, gp, max_outer, tables);
	NNParams P;
	P.mnx = gp->bounding_box_min_X; P.mny = gp->bounding_box_min_Y; P.mnz = gp->bounding_box_min_Z;
	P.mxx = gp->bounding_box_max_X; P.mxy = gp->bounding_box_max_Y; P.mxz = gp->bounding_box_max_Z;
	P.rx = gp->resolution_X; P.ry = gp->resolution_Y; P.rz = gp->resolution_Z;
	P.nbx = gp->number_of_buckets_X; P.nby = gp->number_of_buckets_Y; P.nbz = gp->number_of_buckets_Z;
	P.nb = gp->number_of_buckets;
	P.buckets = buckets;
	P.ci = si.view();
	P.co = two ? so.view() : si.view();
	P.cap_in = max_inner; P.cap_out = max_outer;
	nn_params_finish(P, radius, prune);
	long long total = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+:total)
	for (int q = 0; q < n2; q++) {
		const m3dreg_point &s = second[q];
		float4 p = make_float4(s.x, s.y, s.z, f_from_bits(s.label));
		float4 pn = make_float4(s.normal_x, s.normal_y, s.normal_z, 0.0f);
		unsigned int evals = 0;
		int l = P.nb > 0 ? nn_query(P, p, pn, evals) : kNNNone;
		nn_out[q] = (l != kNNNone && l >= 0 && l < n1) ? table[l].index_of_point : -1;
		total += evals;
	}
	if (evals_out) *evals_out = total;
	return 0;
}


/* ---- warp-level emulation of k_nn_search_grid -------------------------------------------------------------------
 * The warp-shared search (m3dreg_kernels.cuh) restated for 32 "lanes" held in arrays: hull of the lanes' boxes, one
 * look-up per hull cell (representative-cell rule for coarser bins, cells of the previous hull skipped), batches of
 * staged groups of four, branch-free group minimum + tie flag, cold full-predicate step, re-scan, hull-based settle
 * test, scattered-warp fallback to nn_query().  It does not share source with the kernel (the kernel is written with
 * warp intrinsics); it exists so that the ALGORITHM's exactness argument is checked against the oracle on a machine
 * without a GPU, on the same cases the GPU parity tests use. */
namespace {

constexpr int kWCells = 128, kWStage = 192, kHullMin = 128, kHullRatio = 8;
/* 1: cells of the previous round's hull are skipped (k_nn_search_grid, round 1); 0: every round looks at its whole hull
 * (k_nn_search_hull) */
int g_skip_old_hull = 0;

struct WLane {
	float qx, qy, qz, pnx, pny, pnz, best_d, lim, mgx, mgy, mgz;
	int label, best_l, cx0, cx1, cy0, cy1, cz0, cz1;
	bool active, mine, unsettled;
	float4 p, pn;
};
struct WSeg { int start, cnt, ax, ay, az; };
struct WGroup { float4 c[4]; int gstart, valid, ax, ay, az; };

inline void w_fine_box(const NNParams &P, const WLane &q, float tau, int &xl, int &xh, int &yl, int &yh, int &zl, int &zh)
{
	const float R = !P.prune ? INFINITY : (tau > 1.0e-30f ? f_fma(sqrtf(tau), 1.0001220703125f, 1.0e-18f) : 1.1e-15f);
	xl = col_floor(f_sub(q.qx, R), P.mnx, P.iwx, q.mgx); xh = col_ceil(f_add(q.qx, R), P.mnx, P.iwx, q.mgx);
	yl = col_floor(f_sub(q.qy, R), P.mny, P.iwy, q.mgy); yh = col_ceil(f_add(q.qy, R), P.mny, P.iwy, q.mgy);
	zl = col_floor(f_sub(q.qz, R), P.mnz, P.iwz, q.mgz); zh = col_ceil(f_add(q.qz, R), P.mnz, P.iwz, q.mgz);
	xl = xl > q.cx0 ? xl : q.cx0; xh = xh < q.cx1 ? xh : q.cx1;
	yl = yl > q.cy0 ? yl : q.cy0; yh = yh < q.cy1 ? yh : q.cy1;
	zl = zl > q.cz0 ? zl : q.cz0; zh = zh < q.cz1 ? zh : q.cz1;
}

inline void w_consider(const CandSet &cs, WLane &q, float d, const float4 &c, int j)
{
	if (d <= q.lim) {
		const int l = f_bits(c.w);
		if (d < q.best_d || l < q.best_l) {
			const float4 n = cs.nrm[j];
			if (f_bits(n.w) == q.label) {
				const float dot = f_fma(q.pnz, n.z, f_fma(q.pnx, n.x, f_mul(q.pny, n.y)));
				if (angle_gate(dot)) { q.best_d = d; q.best_l = l; q.lim = d; }
			}
		}
	}
}

} /* namespace */

extern "C" void emul_set_skip_old_hull(int on) { g_skip_old_hull = on ? 1 : 0; }

extern "C" int emul_nn_search_warp(const m3dreg_point *first, int n1, const m3dreg_point *second, int n2,
		const m3dreg_hash_element *table, const m3dreg_bucket *buckets, const m3dreg_grid_params *gp,
		float radius, int cap, int prune, int *nn_out, long long *fallbacks_out, long long *rescans_out)
{
	const int tables = nn_tables_usable(cap, cap) ? 1 : 0;
	HostSet hs;
	build_set(hs, first, n1, table, buckets, gp, cap, tables);
	NNParams P;
	P.mnx = gp->bounding_box_min_X; P.mny = gp->bounding_box_min_Y; P.mnz = gp->bounding_box_min_Z;
	P.mxx = gp->bounding_box_max_X; P.mxy = gp->bounding_box_max_Y; P.mxz = gp->bounding_box_max_Z;
	P.rx = gp->resolution_X; P.ry = gp->resolution_Y; P.rz = gp->resolution_Z;
	P.nbx = gp->number_of_buckets_X; P.nby = gp->number_of_buckets_Y; P.nbz = gp->number_of_buckets_Z;
	P.nb = gp->number_of_buckets;
	P.buckets = buckets;
	P.ci = hs.view(); P.co = P.ci;
	P.cap_in = cap; P.cap_out = cap;
	nn_params_finish(P, radius, prune);
	const CandSet cs = P.ci;
	long long fallbacks = 0, rescans = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+:fallbacks, rescans)
	for (int base = 0; base < n2; base += 32) {
		WLane L[32];
		for (int l = 0; l < 32; l++) {
			WLane &q = L[l];
			const int qi = base + l;
			q.best_l = kNNNone; q.label = -1; q.best_d = 100000000.0f; q.lim = fminf(P.r2, 99999992.0f);
			q.qx = q.qy = q.qz = q.pnx = q.pny = q.pnz = 0.0f;
			q.active = q.mine = q.unsettled = false;
			int ix = 0, iy = 0, iz = 0;
			if (qi < n2 && P.nb > 0 && cap > 0) {
				const m3dreg_point &s = second[qi];
				q.p = make_float4(s.x, s.y, s.z, f_from_bits(s.label));
				q.pn = make_float4(s.normal_x, s.normal_y, s.normal_z, 0.0f);
				q.qx = s.x; q.qy = s.y; q.qz = s.z; q.pnx = s.normal_x; q.pny = s.normal_y; q.pnz = s.normal_z;
				q.label = s.label;
				if (!(q.qx < P.mnx || q.qx > P.mxx || q.qy < P.mny || q.qy > P.mxy || q.qz < P.mnz || q.qz > P.mxz)) {
					ix = cell_of(q.qx, P.mnx, P.rx); iy = cell_of(q.qy, P.mny, P.ry); iz = cell_of(q.qz, P.mnz, P.rz);
					const int home = ix * P.nby * P.nbz + iy * P.nbz + iz;
					q.active = home >= 0 && (long long)home < P.nb && q.lim >= 0.0f;
				}
			}
			q.cx0 = (ix > 0 ? ix - 1 : ix) << 2; q.cx1 = ((ix != P.nbx - 1 ? ix + 1 : ix) << 2) + 3;
			q.cy0 = (iy > 0 ? iy - 1 : iy) << 2; q.cy1 = ((iy != P.nby - 1 ? iy + 1 : iy) << 2) + 3;
			q.cz0 = (iz > 0 ? iz - 1 : iz) << 2; q.cz1 = ((iz != P.nbz - 1 ? iz + 1 : iz) << 2) + 3;
			q.mgx = f_fma(f_mul(fabsf(q.qx) + fabsf(P.mnx), P.iwx), 3.814697265625e-06f, 9.765625e-04f);
			q.mgy = f_fma(f_mul(fabsf(q.qy) + fabsf(P.mny), P.iwy), 3.814697265625e-06f, 9.765625e-04f);
			q.mgz = f_fma(f_mul(fabsf(q.qz) + fabsf(P.mnz), P.iwz), 3.814697265625e-06f, 9.765625e-04f);
		}
		auto fallback = [&](WLane &q) { unsigned int ev = 0; q.best_l = nn_query(P, q.p, q.pn, ev); fallbacks++; };
		if (!nn_columns_usable(P.nbx, P.nby, P.nbz)) {
			for (int l = 0; l < 32; l++) if (L[l].active) { fallback(L[l]); L[l].active = false; }
		}
		unsigned todo = 0;
		for (int l = 0; l < 32; l++) if (L[l].active) todo |= 1u << l;
		while (todo) {
			const int Lb = L[__builtin_ctz(todo)].label;
			for (int l = 0; l < 32; l++) {
				L[l].mine = L[l].active && L[l].label == Lb;
				L[l].unsettled = L[l].mine;
				if (L[l].mine) todo &= ~(1u << l);
			}
			int hxl = 1, hxh = 0, hyl = 1, hyh = 0, hzl = 1, hzh = 0;
			float rho2 = P.rho2_first;
			for (int round = 0; round < 80; round++) {
				bool any = false;
				for (int l = 0; l < 32; l++) any = any || L[l].unsettled;
				if (!any) break;
				int uxl = 0x7fffffff, uxh = -0x7fffffff, uyl = 0x7fffffff, uyh = -0x7fffffff, uzl = 0x7fffffff, uzh = -0x7fffffff, own_max = 0;
				for (int l = 0; l < 32; l++) {
					if (!L[l].unsettled) continue;
					int xl, xh, yl, yh, zl, zh;
					w_fine_box(P, L[l], P.prune ? fminf(L[l].lim, rho2) : L[l].lim, xl, xh, yl, yh, zl, zh);
					if (xl > xh || yl > yh || zl > zh) continue;
					uxl = std::min(uxl, xl); uxh = std::max(uxh, xh); uyl = std::min(uyl, yl); uyh = std::max(uyh, yh);
					uzl = std::min(uzl, zl); uzh = std::max(uzh, zh);
					own_max = std::max(own_max, (xh - xl + 1) * (yh - yl + 1) * (zh - zl + 1));
				}
				if (uxl <= uxh) {
					const int dx = uxh - uxl + 1, dy = uyh - uyl + 1, dz = uzh - uzl + 1;
					const long long ncell = (long long)dx * dy * dz;
					if (dx > kWCells || dy > 32767 || dz > 32767 || (ncell > kHullMin && ncell > (long long)kHullRatio * own_max)) {
						for (int l = 0; l < 32; l++) if (L[l].unsettled) { fallback(L[l]); L[l].unsettled = false; }
						break;
					}
					bool nb_all = true;
					for (int l = 0; l < 32; l++)
						nb_all = nb_all && (!L[l].mine || (uxl >= L[l].cx0 && uxh <= L[l].cx1 && uyl >= L[l].cy0 && uyh <= L[l].cy1 && uzl >= L[l].cz0 && uzh <= L[l].cz1));
					const int nrows = dy * dz, rpc = kWCells / dx;
					for (int row0 = 0; row0 < nrows; row0 += rpc) {
						const int nr = std::min(rpc, nrows - row0);
						std::vector<WSeg> segs;
						for (int c = 0; c < nr * dx; c++) {
							const int rr = c / dx, ax = c - rr * dx, r = row0 + rr, az = r / dy, ay = r - az * dy;
							const int gx = uxl + ax, gy = uyl + ay, gz = uzl + az;
							if (g_skip_old_hull && gx >= hxl && gx <= hxh && gy >= hyl && gy <= hyh && gz >= hzl && gz <= hzh) continue;
							const int cell = ((gx >> 2) * P.nby + (gy >> 2)) * P.nbz + (gz >> 2);
							const int npts = buckets[cell].number_of_points, begin = buckets[cell].index_begin;
							if (!(npts > 0 && begin >= 0)) continue;
							const int level = P.tables ? nn_level(npts) : -1;
							const int sh = level < 0 ? 2 : 2 - level, am = (1 << sh) - 1;
							if (!((ax == 0 || !(gx & am)) && (ay == 0 || !(gy & am)) && (az == 0 || !(gz & am)))) continue;
							WSeg sg; sg.ax = ax; sg.ay = ay; sg.az = az;
							if (level < 0) {
								const int iter = candidate_stride(npts, cap);
								sg.start = begin; sg.cnt = (npts + iter - 1) / iter;
							} else {
								const int bin = nn_bin(Lb, (gx & 3) >> sh, (gy & 3) >> sh, (gz & 3) >> sh, level);
								const unsigned short *tab = cs.tab + 2 * (size_t)begin + bin;
								sg.start = begin + tab[0]; sg.cnt = (int)tab[1] - (int)tab[0];
							}
							if (sg.cnt > 0) segs.push_back(sg);
						}
						size_t k0 = 0;
						while (k0 < segs.size()) {
							std::vector<WGroup> groups;
							auto push_groups = [&](const WSeg &sg, int start, int cnt) {
								for (int g = 0; g < (cnt + 3) / 4; g++) {
									WGroup G; G.gstart = start + 4 * g; G.valid = std::min(4, cnt - 4 * g); G.ax = sg.ax; G.ay = sg.ay; G.az = sg.az;
									for (int t = 0; t < 4; t++)
										G.c[t] = t < G.valid ? cs.xyzl[G.gstart + t] : make_float4(INFINITY, INFINITY, INFINITY, f_from_bits(kNNNone));
									groups.push_back(G);
								}
							};
							int used = 0;
							size_t k = k0;
							while (k < segs.size() && used + ((segs[k].cnt + 3) & ~3) <= kWStage) { push_groups(segs[k], segs[k].start, segs[k].cnt); used += (segs[k].cnt + 3) & ~3; k++; }
							if (k == k0) {       /* the first segment alone exceeds a batch: take a part of it */
								push_groups(segs[k0], segs[k0].start, kWStage);
								segs[k0].start += kWStage; segs[k0].cnt -= kWStage;
							} else k0 = k;
							bool any_flag = false;
							bool flag[32];
							for (int l = 0; l < 32; l++) {
								WLane &q = L[l];
								float rb = q.mine ? q.lim : -INFINITY;
								int bg = -1;
								flag[l] = false;
								for (size_t g = 0; g < groups.size(); g++) {
									const WGroup &G = groups[g];
									const int gx = uxl + G.ax, gy = uyl + G.ay, gz = uzl + G.az;
									const bool use = nb_all || (gx >= q.cx0 && gx <= q.cx1 && gy >= q.cy0 && gy <= q.cy1 && gz >= q.cz0 && gz <= q.cz1);
									float m4 = fminf(fminf(nn_dist(q.qx, q.qy, q.qz, G.c[0]), nn_dist(q.qx, q.qy, q.qz, G.c[1])),
											fminf(nn_dist(q.qx, q.qy, q.qz, G.c[2]), nn_dist(q.qx, q.qy, q.qz, G.c[3])));
									if (!use) m4 = INFINITY;
									const bool lt = m4 < rb;
									flag[l] = flag[l] || (m4 == rb);
									if (lt) { rb = m4; bg = (int)g; }
								}
								if (bg >= 0) {
									const WGroup &G = groups[(size_t)bg];
									for (int t = 0; t < 4; t++) w_consider(cs, q, nn_dist(q.qx, q.qy, q.qz, G.c[t]), G.c[t], G.gstart + t);
									if (q.best_d != rb) flag[l] = true;
								}
								any_flag = any_flag || flag[l];
							}
							if (any_flag) {
								for (int l = 0; l < 32; l++) {
									WLane &q = L[l];
									if (!(flag[l] && q.mine)) continue;
									rescans++;
									for (const WGroup &G : groups) {
										const int gx = uxl + G.ax, gy = uyl + G.ay, gz = uzl + G.az;
										if (!(gx >= q.cx0 && gx <= q.cx1 && gy >= q.cy0 && gy <= q.cy1 && gz >= q.cz0 && gz <= q.cz1)) continue;
										for (int t = 0; t < 4; t++) w_consider(cs, q, nn_dist(q.qx, q.qy, q.qz, G.c[t]), G.c[t], G.gstart + t);
									}
								}
							}
						}
					}
					hxl = uxl; hxh = uxh; hyl = uyl; hyh = uyh; hzl = uzl; hzh = uzh;
					for (int l = 0; l < 32; l++) {
						WLane &q = L[l];
						if (q.unsettled && P.prune && q.lim > rho2) {
							int bxl, bxh, byl, byh, bzl, bzh;
							w_fine_box(P, q, q.lim, bxl, bxh, byl, byh, bzl, bzh);
							if (bxl >= uxl && bxh <= uxh && byl >= uyl && byh <= uyh && bzl >= uzl && bzh <= uzh) q.unsettled = false;
						}
					}
				}
				for (int l = 0; l < 32; l++)
					if (L[l].unsettled && (!P.prune || L[l].lim <= rho2)) L[l].unsettled = false;
				rho2 = f_mul(rho2, 4.0f);
			}
		}
		for (int l = 0; l < 32 && base + l < n2; l++) {
			const int bl = L[l].best_l;
			nn_out[base + l] = (bl != kNNNone && bl >= 0 && bl < n1) ? table[bl].index_of_point : -1;
		}
	}
	if (fallbacks_out) *fallbacks_out = fallbacks;
	if (rescans_out) *rescans_out = rescans;
	return 0;
}
