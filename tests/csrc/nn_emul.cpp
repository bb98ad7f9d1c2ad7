/* nn_emul.cpp — TEST-ONLY host instantiation of the device search logic (mandala-mapping_b200/csrc/nn_core.cuh).
 *
 * nn_query() is __host__ __device__; this file compiles it with g++ (-ffp-contract=off), builds the candidate layout
 * serially exactly as k_build_candidates specifies it (walk order, bins, u16 offset tables inside the bucket's own
 * range) and runs one query after the other, so the search logic, the layout contract and the pruning proofs can be
 * checked against the CPU oracle on a machine without a GPU (tests/test_nn_emul.py).  The product never links this. */
#include <vector>
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
