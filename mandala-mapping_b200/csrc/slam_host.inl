/* slam_host.inl — host side of the multi-scan sweep (gpu6DSLAM::registerAll, src/gpu6DSLAM.cpp:424-597), part of m3dreg.cu.
 *
 * The reference runs registerAll on one GPU, pair after pair.  Here the (i, j) pairs of a Jacobi sweep are the unit of
 * work of a multi-GPU job: every rank (one m3dreg context per GPU, all scans resident on every GPU) plans the SAME pair
 * list and the SAME partition from the poses alone, accumulates the normal-equation blocks of its own pairs, the
 * n_scans x 28 doubles are summed over the ranks with ONE ncclAllReduce on the context's stream (NVLink / NVSwitch),
 * and every rank solves all scans redundantly (6x6 Cholesky each) — no gather, no second collective.
 *
 * NCCL is resolved at run time (dlsym on the process first — a host application or PyTorch that already carries NCCL
 * keeps ONE copy of it — then dlopen of libnccl.so.2): libm3dreg.so has no link-time dependency on it and single-GPU
 * users never load it. */
#include <dlfcn.h>

namespace {

/* the few NCCL declarations this file needs (nccl.h: ncclResult_t ncclSuccess = 0, ncclDataType_t ncclFloat64 = 8,
 * ncclRedOp_t ncclSum = 0, ncclUniqueId = 128 opaque bytes) */
typedef struct ncclComm *m3d_nccl_comm_t;
struct m3d_nccl_unique_id { char internal[128]; };
typedef int (*pfn_ncclGetUniqueId)(m3d_nccl_unique_id *);
typedef int (*pfn_ncclCommInitRank)(m3d_nccl_comm_t *, int, m3d_nccl_unique_id, int);
typedef int (*pfn_ncclCommDestroy)(m3d_nccl_comm_t);
typedef int (*pfn_ncclAllReduce)(const void *, void *, size_t, int, int, m3d_nccl_comm_t, cudaStream_t);
typedef const char *(*pfn_ncclGetErrorString)(int);

struct NcclApi {
	bool tried = false, ok = false;
	pfn_ncclGetUniqueId get_unique_id = nullptr;
	pfn_ncclCommInitRank comm_init_rank = nullptr;
	pfn_ncclCommDestroy comm_destroy = nullptr;
	pfn_ncclAllReduce all_reduce = nullptr;
};

NcclApi &nccl_api()
{
	static NcclApi api;
	if (api.tried) return api;
	api.tried = true;
	void *h = RTLD_DEFAULT;
	if (!dlsym(h, "ncclAllReduce")) {
		h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!h) return api;
	}
	api.get_unique_id = (pfn_ncclGetUniqueId)dlsym(h, "ncclGetUniqueId");
	api.comm_init_rank = (pfn_ncclCommInitRank)dlsym(h, "ncclCommInitRank");
	api.comm_destroy = (pfn_ncclCommDestroy)dlsym(h, "ncclCommDestroy");
	api.all_reduce = (pfn_ncclAllReduce)dlsym(h, "ncclAllReduce");
	api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_reduce;
	return api;
}

/* Pairs (i, j), i in [first_optimised, n), j != i, whose (Euler round-tripped = stored) translations are closer than the
 * threshold: the reference's gate, float arithmetic (gpu6DSLAM.cpp:432-469).  Ordered by i, then j. */
void slam_gate_pairs(const float *poses, int n_scans, float threshold, int first_optimised, std::vector<int> &pi, std::vector<int> &pj)
{
	pi.clear(); pj.clear();
	for (int i = first_optimised < 0 ? 0 : first_optimised; i < n_scans; i++) {
		const float *a = poses + 16 * (size_t)i;
		for (int j = 0; j < n_scans; j++) {
			if (i == j) continue;
			const float *b = poses + 16 * (size_t)j;
			volatile float dx = a[3] - b[3], dy = a[7] - b[7], dz = a[11] - b[11];
			volatile float xx = dx * dx, yy = dy * dy, zz = dz * dz;
			volatile float s = xx + yy;
			s = s + zz;
			const float dist = sqrtf(s);
			if (dist < threshold) { pi.push_back(i); pj.push_back(j); }
		}
	}
}

/* Owner rank of every pair: whole groups of equal i first (the grid of scan i is then built once), heaviest group to the
 * least loaded rank; groups heavier than 1.25x the ideal share are split into consecutive runs.  Deterministic: every
 * rank computes the same partition from the same poses.  Cost model: points of i + points of j per pair, plus the points
 * of i once per group (its grid build) — or, when the previous sweep of this context measured it, the device time per
 * pair of every scan's group (cost_per_pair, ms; identical on every rank: it is all-reduced). */
void slam_partition(const std::vector<int> &pi, const std::vector<int> &pj, const int *sizes, int world, std::vector<int> &owner,
		const double *cost_per_pair = nullptr)
{
	const size_t np = pi.size();
	owner.assign(np, 0);
	if (world <= 1 || np == 0) return;
	struct Group { double cost; size_t begin, end; };
	std::vector<Group> groups;
	if (cost_per_pair) {      /* measured costs: whole groups, heaviest first, to the least loaded rank */
		double known = 0.0;
		size_t n_known = 0;
		for (size_t k = 0; k < np; k++) if (cost_per_pair[pi[k]] > 0.0) { known += cost_per_pair[pi[k]]; n_known++; }
		const double fallback = n_known ? known / (double)n_known : 1.0;
		double total_ms = 0.0;
		for (size_t p = 0; p < np;) {
			size_t e = p;
			while (e < np && pi[e] == pi[p]) e++;
			const double cpp = cost_per_pair[pi[p]] > 0.0 ? cost_per_pair[pi[p]] : fallback;
			groups.push_back({cpp * (double)(e - p), p, e});
			total_ms += groups.back().cost;
			p = e;
		}
		const double ideal_ms = total_ms / (double)world;
		std::vector<Group> split;
		for (const Group &g : groups) {      /* a group heavier than 1.25x the ideal share is cut into consecutive runs, as below */
			const size_t len = g.end - g.begin;
			if (g.cost > 1.25 * ideal_ms && len > 1) {
				size_t parts = (size_t)ceil(g.cost / ideal_ms);
				if (parts > len) parts = len;
				size_t at = g.begin;
				for (size_t q = 0; q < parts; q++) {
					const size_t l = len / parts + (q < len % parts ? 1 : 0);
					split.push_back({g.cost * (double)l / (double)len, at, at + l});
					at += l;
				}
			} else split.push_back(g);
		}
		std::stable_sort(split.begin(), split.end(), [](const Group &a, const Group &b) { return a.cost > b.cost; });
		std::vector<double> load((size_t)world, 0.0);
		for (const Group &g : split) {
			int r = 0;
			for (int k = 1; k < world; k++) if (load[(size_t)k] < load[(size_t)r]) r = k;
			load[(size_t)r] += g.cost;
			for (size_t k = g.begin; k < g.end; k++) owner[k] = r;
		}
		return;
	}
	double total = 0.0;
	for (size_t p = 0; p < np;) {      /* pairs arrive grouped by i (slam_gate_pairs order) */
		size_t e = p;
		while (e < np && pi[e] == pi[p]) e++;
		for (size_t k = p; k < e; k++) total += (double)sizes[pi[k]] + (double)sizes[pj[k]];
		total += (double)sizes[pi[p]];
		p = e;
	}
	const double ideal = total / (double)world;
	for (size_t p = 0; p < np;) {
		size_t e = p;
		while (e < np && pi[e] == pi[p]) e++;
		double c = (double)sizes[pi[p]];
		for (size_t k = p; k < e; k++) c += (double)sizes[pi[k]] + (double)sizes[pj[k]];
		const size_t len = e - p;
		if (c > 1.25 * ideal && len > 1) {
			size_t parts = (size_t)ceil(c / ideal);
			if (parts > len) parts = len;
			/* consecutive runs, the first (len % parts) one longer (numpy.array_split) */
			size_t at = p;
			for (size_t q = 0; q < parts; q++) {
				const size_t l = len / parts + (q < len % parts ? 1 : 0);
				double cc = (double)sizes[pi[p]];
				for (size_t k = at; k < at + l; k++) cc += (double)sizes[pi[k]] + (double)sizes[pj[k]];
				groups.push_back({cc, at, at + l});
				at += l;
			}
		} else groups.push_back({c, p, e});
		p = e;
	}
	std::stable_sort(groups.begin(), groups.end(), [](const Group &a, const Group &b) { return a.cost > b.cost; });
	std::vector<double> load((size_t)world, 0.0);
	for (const Group &g : groups) {
		int r = 0;
		for (int k = 1; k < world; k++) if (load[(size_t)k] < load[(size_t)r]) r = k;
		load[(size_t)r] += g.cost;
		for (size_t k = g.begin; k < g.end; k++) owner[k] = r;
	}
}

} /* namespace */

extern "C" {

int m3dreg_slam_plan(const float *poses, int n_scans, const int *sizes, float distance_threshold, int first_optimised, int world,
		int *pair_i, int *pair_j, int *owner, int cap)
{
	if (!poses || n_scans <= 0 || !sizes || world <= 0) return M3DREG_E_INVALID_ARG;
	std::vector<int> pi, pj, ow;
	slam_gate_pairs(poses, n_scans, distance_threshold, first_optimised, pi, pj);
	slam_partition(pi, pj, sizes, world, ow);
	const int np = (int)pi.size();
	if (pair_i && pair_j && owner) {
		if (cap < np) return M3DREG_E_SIZE_MISMATCH;
		for (int k = 0; k < np; k++) { pair_i[k] = pi[(size_t)k]; pair_j[k] = pj[(size_t)k]; owner[k] = ow[(size_t)k]; }
	}
	return np;
}

int m3dreg_slam_plan_measured(const float *poses, int n_scans, const int *sizes, float distance_threshold, int first_optimised, int world,
		const double *cost_per_pair, int *pair_i, int *pair_j, int *owner, int cap)
{
	if (!poses || n_scans <= 0 || !sizes || world <= 0 || !cost_per_pair) return M3DREG_E_INVALID_ARG;
	std::vector<int> pi, pj, ow;
	slam_gate_pairs(poses, n_scans, distance_threshold, first_optimised, pi, pj);
	slam_partition(pi, pj, sizes, world, ow, cost_per_pair);
	const int np = (int)pi.size();
	if (pair_i && pair_j && owner) {
		if (cap < np) return M3DREG_E_SIZE_MISMATCH;
		for (int k = 0; k < np; k++) { pair_i[k] = pi[(size_t)k]; pair_j[k] = pj[(size_t)k]; owner[k] = ow[(size_t)k]; }
	}
	return np;
}

int m3dreg_nccl_get_unique_id(void *id128)
{
	if (!id128) return M3DREG_E_INVALID_ARG;
	NcclApi &api = nccl_api();
	if (!api.ok) return M3DREG_E_NO_NCCL;
	m3d_nccl_unique_id id;
	int rc = api.get_unique_id(&id);
	if (rc != 0) return M3DREG_E_NCCL;
	memcpy(id128, &id, sizeof(id));
	return 0;
}

int m3dreg_nccl_init(m3dreg_ctx *c, const void *id128, int rank, int world)
{
	if (!c || !id128 || rank < 0 || world <= 0 || rank >= world) return M3DREG_E_INVALID_ARG;
	NcclApi &api = nccl_api();
	if (!api.ok) return M3DREG_E_NO_NCCL;
	CK(cudaSetDevice(c->dev));
	if (c->nccl_comm && c->nccl_owned) { api.comm_destroy((m3d_nccl_comm_t)c->nccl_comm); c->nccl_comm = nullptr; }
	m3d_nccl_unique_id id;
	memcpy(&id, id128, sizeof(id));
	m3d_nccl_comm_t comm = nullptr;
	int rc = api.comm_init_rank(&comm, world, id, rank);
	if (rc != 0) return M3DREG_E_NCCL;
	c->nccl_comm = comm; c->nccl_owned = true; c->nccl_rank = rank; c->nccl_world = world;
	return 0;
}

int m3dreg_nccl_attach(m3dreg_ctx *c, void *nccl_comm, int rank, int world)
{
	if (!c || rank < 0 || world <= 0 || rank >= world || (world > 1 && !nccl_comm)) return M3DREG_E_INVALID_ARG;
	if (world > 1 && !nccl_api().ok) return M3DREG_E_NO_NCCL;
	if (c->nccl_comm && c->nccl_owned && nccl_api().ok) nccl_api().comm_destroy((m3d_nccl_comm_t)c->nccl_comm);      /* NCCL is only ever looked up by multi-GPU users */
	c->nccl_comm = nccl_comm; c->nccl_owned = false; c->nccl_rank = rank; c->nccl_world = world;
	return 0;
}

int m3dreg_slam_sweep(m3dreg_ctx *c, int n_scans, float *poses, const m3dreg_slam_params *sp, int *status_out, m3dreg_sweep_stats *stats)
{
	if (!c || n_scans <= 0 || !poses || !sp || !valid_params(&sp->reg)) return M3DREG_E_INVALID_ARG;
	if ((size_t)n_scans > c->scans.size()) return M3DREG_E_BAD_SLOT;
	CK(cudaSetDevice(c->dev));
	const int world = c->nccl_world > 0 ? c->nccl_world : 1, rank = c->nccl_world > 0 ? c->nccl_rank : 0;
	int first_opt = sp->first_optimised;
	if (first_opt < 0) first_opt = 0;
	if (first_opt > n_scans) first_opt = n_scans;
	std::vector<int> sizes((size_t)n_scans);
	for (int s = 0; s < n_scans; s++) {
		sizes[(size_t)s] = c->scans[(size_t)s].n;
		if (sizes[(size_t)s] <= 0) return M3DREG_E_BAD_SLOT;
	}
	std::vector<int> pi, pj, owner, mi, mj;
	slam_gate_pairs(poses, n_scans, sp->distance_threshold, first_opt, pi, pj);
	m3dreg_ctx::SlamCosts &costs = c->slam_costs_for(sp->reg);      /* this kind of sweep's last measurement */
	const bool have_costs = world > 1 && costs.per_pair.size() == (size_t)n_scans && !getenv("M3DREG_SLAM_STATIC_PARTITION");
	slam_partition(pi, pj, sizes.data(), world, owner, have_costs ? costs.per_pair.data() : nullptr);
	if (have_costs) c->slam_cost_per_pair_used = costs.per_pair;
	long long my_points = 0, all_points = 0;
	for (size_t k = 0; k < pi.size(); k++) {
		const long long pts = (long long)sizes[(size_t)pi[k]] + sizes[(size_t)pj[k]];
		all_points += pts;
		if (owner[k] == rank) { mi.push_back(pi[k]); mj.push_back(pj[k]); my_points += pts; }
	}
	int e;
	if ((e = c->d_neq.ensure((size_t)n_scans * kNeqCount))) return e;
	CK(cudaEventRecord(c->ev0, c->stream));
	if ((e = m3dreg_sweep_zero(c, c->d_neq.p, n_scans))) return e;
	if (world > 1) {      /* buffers for any share of this pair list (see m3dreg_ctx::sweep_reserve_*) */
		size_t max_batch = 0, max_first = 0;
		long long max_cap = 1;
		for (size_t p = 0; p < pi.size();) {
			size_t e2 = p, total = 0;
			while (e2 < pi.size() && pi[e2] == pi[p]) {
				total += ((size_t)sizes[(size_t)pj[e2]] + kSegChunk - 1) / kSegChunk * kSegChunk;
				e2++;
			}
			const size_t lim = ((size_t)8 << 20) + ((size_t)1 << 22);      /* a batch closes at 8 Mi queries (sweep_accumulate_impl) */
			if (total > lim) total = lim;
			if (total > max_batch) max_batch = total;
			const Scan &A = c->scans[(size_t)pi[p]];
			if ((size_t)A.n > max_first) max_first = (size_t)A.n;
			const long long cap = bucket_capacity_for(A.diag, &sp->reg);
			if (cap > max_cap) max_cap = cap;
			p = e2;
		}
		c->sweep_reserve_segs = pi.size(); c->sweep_reserve_second = max_batch; c->sweep_reserve_first = max_first; c->sweep_reserve_cap = max_cap;
	}
	double *d_group_ms = nullptr;
	if (world > 1) {      /* device time per group of pairs, for the next sweep's partition */
		if ((e = c->d_group_ms.ensure((size_t)n_scans))) return e;
		CK(cudaMemsetAsync(c->d_group_ms.p, 0, (size_t)n_scans * sizeof(double), c->stream));
		d_group_ms = c->d_group_ms.p;
	}
	if (!mi.empty() && (e = sweep_accumulate_impl(c, (int)mi.size(), mi.data(), mj.data(), poses, n_scans, &sp->reg, c->d_neq.p, d_group_ms))) return e;
	CK(cudaEventRecord(c->ev1, c->stream));
	c->sweep_reserve_segs = 0; c->sweep_reserve_second = 0; c->sweep_reserve_first = 0; c->sweep_reserve_cap = 0;
	if (world > 1) {
		NcclApi &api = nccl_api();
		if (!api.ok || !c->nccl_comm) return M3DREG_E_NO_NCCL;
		/* the ONE exchange step of a sweep: n_scans x 28 doubles, in place, on the registration stream */
		int rc = api.all_reduce(c->d_neq.p, c->d_neq.p, (size_t)n_scans * kNeqCount, 8 /* ncclFloat64 */, 0 /* ncclSum */,
				(m3d_nccl_comm_t)c->nccl_comm, c->stream);
		if (rc != 0) return M3DREG_E_NCCL;
	}
	if (!c->ev2) CK(cudaEventCreate(&c->ev2));
	CK(cudaEventRecord(c->ev2, c->stream));
	if (world > 1) {      /* second, tiny exchange (n_scans doubles): every rank learns what every group cost */
		int rc = nccl_api().all_reduce(d_group_ms, d_group_ms, (size_t)n_scans, 8, 0, (m3d_nccl_comm_t)c->nccl_comm, c->stream);
		if (rc != 0) return M3DREG_E_NCCL;
		c->slam_group_ms.assign((size_t)n_scans, 0.0);
		CK(cudaMemcpyAsync(c->slam_group_ms.data(), d_group_ms, (size_t)n_scans * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	}
	if ((e = m3dreg_sweep_solve(c, c->d_neq.p, n_scans, first_opt, n_scans, poses, &sp->reg, status_out))) return e;      /* synchronises */
	if (world > 1) {
		std::vector<int> cnt((size_t)n_scans, 0);
		for (size_t k = 0; k < pi.size(); k++) cnt[(size_t)pi[k]]++;
		costs.per_pair.assign((size_t)n_scans, 0.0);
		for (int s = 0; s < n_scans; s++)
			if (cnt[(size_t)s] > 0 && c->slam_group_ms[(size_t)s] > 0.0) costs.per_pair[(size_t)s] = c->slam_group_ms[(size_t)s] / (double)cnt[(size_t)s];
	}
	if (getenv("M3DREG_SLAM_DEBUG")) {      /* how well did the plan predict this rank's share? */
		float ms_acc = 0.0f;
		cudaEventElapsedTime(&ms_acc, c->ev0, c->ev1);
		double predicted = 0.0, measured = 0.0;
		if (have_costs) for (size_t k = 0; k < pi.size(); k++) if (owner[k] == rank) predicted += c->slam_cost_per_pair_used.empty() ? 0.0 : c->slam_cost_per_pair_used[(size_t)pi[k]];
		if (world > 1) { std::vector<int> seen((size_t)n_scans, 0); for (size_t k = 0; k < mi.size(); k++) if (!seen[(size_t)mi[k]]) { seen[(size_t)mi[k]] = 1; measured += c->slam_group_ms[(size_t)mi[k]]; } }
		fprintf(stderr, "[m3dreg sweep] rank %d/%d pairs %zu of %zu plan %s predicted %.3f ms accumulate %.3f ms (sum of own groups' all-reduced times %.3f)\n",
				rank, world, mi.size(), pi.size(), have_costs ? "measured" : "static", predicted, ms_acc, measured);
	}
	if (stats) {
		float ms_acc = 0.0f, ms_red = 0.0f;
		cudaEventElapsedTime(&ms_acc, c->ev0, c->ev1);
		cudaEventElapsedTime(&ms_red, c->ev1, c->ev2);
		stats->n_pairs = (int64_t)pi.size();
		stats->n_pairs_mine = (int64_t)mi.size();
		stats->points_all = all_points;
		stats->points_mine = my_points;
		stats->accumulate_ms = ms_acc;
		stats->allreduce_ms = ms_red;
		stats->rank = rank;
		stats->world = world;
	}
	return 0;
}

int m3dreg_slam_copy_neq(m3dreg_ctx *c, double *neq_out, int n_scans)
{
	if (!c || !neq_out || n_scans <= 0 || (size_t)n_scans * kNeqCount > c->d_neq.cap) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	CK(cudaMemcpyAsync(neq_out, c->d_neq.p, (size_t)n_scans * kNeqCount * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

} /* extern "C" */
