/* m3dreg.cu — context, persistent arena and the C ABI (include/m3dreg.h) over the sm_100a kernels.
 *
 * Host-side structure mirrors the reference's L1 wrapper (src/cudaWrapper.cpp: CCudaWrapper) but
 *   - device buffers live in a per-context arena that only grows (the reference cudaMalloc/cudaFree's 13+9+18
 *     times per ICP iteration, cudaWrapper.cpp:360-420,523-572, CCUDAAXBSolverWrapper.cpp:407-539),
 *   - everything is issued on one stream with no cudaDeviceSynchronize between kernels (45 in lesson_16.cu),
 *   - the iteration state (pose, bounds, grid parameters, normal equations) stays on the device.
 * There is no CPU fallback: without an sm_100 device m3dreg_create fails.
 */
#include <type_traits>
#include <vector>
#include <utility>
#include <new>
#include <cstring>
#include <cstdio>
#include <algorithm>
#include <chrono>

#include "m3dreg_kernels.cuh"
#include "grid_build.cuh"
#include "nn_hull.cuh"
#include "preproc.cuh"

using namespace m3d;

#define CK(expr)                                                      \
	do {                                                              \
		cudaError_t e__ = (expr);                                     \
		if (e__ != cudaSuccess) return (int)e__;                      \
	} while (0)

namespace {

struct LocalBox { float mn[3], mx[3], diag; };     /* a scan's bounding box in its local frame */

struct Scan {
	float4 *xyzl = nullptr;     /* original order: the order the grid's tie-break (ascending index) refers to   */
	float4 *nrm = nullptr;
	float4 *sx = nullptr;       /* cell-sorted copy, used when the scan plays the QUERY role (warp coherence)     */
	float4 *sn = nullptr;
	uint32_t *perm = nullptr;   /* perm[sorted position] = original index                                         */
	int n = 0;
	size_t cap = 0;
	float diag = 0.0f;          /* diagonal of the local bounding box: bounds the extent of the scan under ANY rigid pose */
	float bb_min[3] = {0, 0, 0}, bb_max[3] = {0, 0, 0};   /* local bounding box (host copy, read back once at upload) */
	LocalBox box() const
	{
		LocalBox b;
		for (int k = 0; k < 3; k++) { b.mn[k] = bb_min[k]; b.mx[k] = bb_max[k]; }
		b.diag = diag;
		return b;
	}
	void release()
	{
		if (xyzl) cudaFree(xyzl);
		if (nrm) cudaFree(nrm);
		if (sx) cudaFree(sx);
		if (sn) cudaFree(sn);
		if (perm) cudaFree(perm);
		xyzl = nrm = sx = sn = nullptr; perm = nullptr; n = 0; cap = 0;
	}
};

template <class T>
struct DevBuf {
	T *p = nullptr;
	size_t cap = 0;
	int ensure(size_t n)
	{
		if (n <= cap) return 0;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = n + n / 8 + 64;
		cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
		if (e != cudaSuccess) return (int)e;
		cap = want;
		return 0;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostSmall {          /* pinned host mirror of the small device block */
	PoseState ps;
	uint32_t bounds[8];
	m3dreg_grid_params gp;
	int flags[FLAG_COUNT];
	unsigned long long label_counts[4];
	double scratch[64];
	float mats[32];
};

} /* namespace */

struct m3dreg_ctx {
	int dev = 0;
	int sm_count = 148;
	cudaStream_t own_stream = nullptr;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	/* host-buffer iteration (m3dreg_icp_iteration_host): the two clouds go up on a copy stream of their own; the second one
	 * (the queries) is only waited for after the grid of the first has been built, so box pass, sort and candidate sets run
	 * under its transfer.  q_deferred_* = the unpack of the queries that icp_iteration_device issues at that point. */
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t ev_copy[3] = {nullptr, nullptr, nullptr};
	cudaEvent_t q_deferred_ev = nullptr;
	const m3dreg_point *q_deferred_aos = nullptr;
	int q_deferred_n = 0;
	int64_t launches = 0;
	int prune = 1;               /* exact bucket pruning in the NN search (0 only for the equivalence test) */

	std::vector<Scan> scans;

	/* arena */
	DevBuf<float4> g_xyzl, g_nrm, ci_xyzl, ci_nrm, co_xyzl, co_nrm, q_xyzl, q_nrm, l_xyzl, l_nrm;
	DevBuf<unsigned short> ci_tab, co_tab;           /* per-bucket bin offset tables of the candidate sets (nn_core.cuh) */
	DevBuf<float4> ci_loc, co_loc;                   /* candidates' local coordinates + original index (moment epilogue of the search) */
	DevBuf<int> bcount, bbegin;                      /* grid megakernel: dense per-bucket counts (zero between launches), slice-relative begins */
	DevBuf<GridBlockRec> brec;
	int coop_pdl = -1;                               /* cooperative + programmatic launch accepted together? (-1: not tried yet) */
	bool gb_attr_set = false;                        /* k_grid_build's dynamic shared memory limit raised */
	unsigned long long *gb_dbg = nullptr;            /* 16 phase time stamps of k_grid_build (only written while profiling) */
	DevBuf<uint32_t> keys[2], vals[2], hist, digit_tot, cell_list;
	DevBuf<m3dreg_bucket> buckets;
	DevBuf<int> nn;
	DevBuf<float4> obs_rec;                          /* per query, query order: matched point (local frame) + its index (k_nn_search*) */
	int grid_mega = 0;                               /* env M3DREG_GRID_MEGA=1: the single-launch grid build (grid_build.cuh) instead of the
	                                                  * PDL-chained kernels — measured slower on B200 at every size, kept for A/B runs */
	int nn_v7 = 0;                                   /* env M3DREG_NN_V7=1: round 1's k_nn_search_grid instead of k_nn_search_hull (A/B runs) */
	cudaError_t launch_err = cudaSuccess;            /* first failed kernel launch since the last report */
	DevBuf<m3dreg_point> aos_a, aos_b, pp_aos;       /* pp_*: pre-registration steps (preproc_host.inl) */
	DevBuf<unsigned char> pp_markers;
	DevBuf<int> pp_tiles;
	DevBuf<m3dreg_obs_nn> obs;
	DevBuf<double> partials, ndt_acc;
	DevBuf<long long> ndt_qacc;                      /* per bucket {count, fixed-point coordinate sums} of the NDT query pass */
	DevBuf<long long> ndt_iacc;                      /* per bucket 12 fixed-point sums of the gridded cloud (NDT statistics) */
	float act_local_mag = 1.0f;                      /* largest |coordinate| of the gridded scan in its local frame (NDT fixed-point scale) */
	DevBuf<m3dreg_hash_element> table;
	DevBuf<float> d_poses1;      /* sweep: round-tripped poses, 16 floats per scan */
	DevBuf<double> d_pose6;      /* sweep: tx,ty,tz,om,fi,ka per scan             */
	DevBuf<int> d_sweep_status;  /* sweep: per-scan solve status                  */
	DevBuf<double> d_neq;        /* m3dreg_slam_sweep: n_scans x 28 normal-equation blocks (the all-reduce buffer) */
	/* multi-GPU sweeps: measured device time of every scan's group of pairs (ms, %globaltimer stamps at the group boundaries),
	 * all-reduced after the sweep; the NEXT sweep's partition balances these instead of point counts (the cost of a pair
	 * depends on how much the two scans overlap, which point counts do not show) */
	DevBuf<double> d_group_ms;
	unsigned long long *stamp_last = nullptr;
	std::vector<double> slam_group_ms, slam_cost_per_pair_used;
	/* measured cost per pair of every scan's group, kept per sweep KIND (mode, bucket size, search radius): an NDT sweep costs
	 * a fifth of an ICP sweep and spreads differently over the scans, and drivers alternate them (C5) — each kind plans with
	 * its own last measurement.  Four kinds are remembered (least recently used replaced). */
	struct SlamCosts { int mode = -1; float bucket = 0.0f, radius = 0.0f; unsigned long long age = 0; std::vector<double> per_pair; };
	SlamCosts slam_costs[4];
	unsigned long long slam_costs_clock = 0;
	SlamCosts &slam_costs_for(const m3dreg_reg_params &r)
	{
		SlamCosts *lru = &slam_costs[0];
		for (auto &k : slam_costs) {
			if (k.mode == r.mode && k.bucket == r.bucket_size && k.radius == r.search_radius) { k.age = ++slam_costs_clock; return k; }
			if (k.age < lru->age) lru = &k;
		}
		lru->mode = r.mode; lru->bucket = r.bucket_size; lru->radius = r.search_radius; lru->per_pair.clear(); lru->age = ++slam_costs_clock;
		return *lru;
	}
	void slam_costs_clear() { for (auto &k : slam_costs) { k.mode = -1; k.per_pair.clear(); k.age = 0; } }
	/* m3dreg_slam_sweep sizes the sweep's buffers for the WHOLE pair list, not for this rank's share: the share changes from
	 * sweep to sweep under the measured-cost partition, and a buffer that grows synchronises (and re-pins host memory) */
	size_t sweep_reserve_segs = 0, sweep_reserve_second = 0, sweep_reserve_first = 0;
	long long sweep_reserve_cap = 0;
	void *nccl_comm = nullptr;   /* ncclComm_t of this rank (m3dreg_nccl_init / m3dreg_nccl_attach), 0 = single GPU */
	bool nccl_owned = false;
	int nccl_rank = 0, nccl_world = 0;
	cudaEvent_t ev2 = nullptr;

	/* batched sweep step: segment table, chunk -> segment map, per-segment label counters, pinned staging */
	DevBuf<SweepSeg> d_segs;
	DevBuf<int> d_seg_of_chunk;
	DevBuf<unsigned long long> d_seg_counts;
	void *h_sweep = nullptr;     /* pinned: a sweep's round-tripped poses + every batch's segment table */
	size_t h_sweep_bytes = 0;

	/* small device block */
	PoseState *ps = nullptr;
	uint32_t *bounds = nullptr;
	m3dreg_grid_params *gp = nullptr;
	int *flags = nullptr;
	unsigned long long *label_counts = nullptr;
	unsigned int *ticket = nullptr;
	unsigned int *cell_count = nullptr;           /* number of searchable buckets in the compact list */
	unsigned int *grid_bar = nullptr;             /* grid barrier of k_grid_build: arrival count, generation */
	unsigned int *nn_work = nullptr;              /* k_nn_search_hull's work counters (next chunk, idle warps): zero between launches */
	int nn_blocks_per_sm[2] = {0, 0};             /* resident blocks per SM of k_nn_search_hull<false|true> (occupancy API, once) */
	unsigned long long *eval_counter = nullptr;   /* candidates staged by the NN search (warp-level), diagnostic */
	int use_pdl = 1;             /* programmatic dependent launch for every kernel (env M3DREG_NO_PDL=1 disables) */
	int nn_diag = 0;             /* env M3DREG_NN_DIAG=1: per-chunk time stamps of the profiling search kernel (tools/nn_tail.py) */
	int nn_per_thread = 0;       /* test switch (env M3DREG_NN_PER_THREAD=1): k_nn_search instead of k_nn_search_grid */
	NNTuning nn_tune = {16, 128, 8};   /* heuristics of k_nn_search_grid (env M3DREG_NN_RHO_DIV / _HULL_MIN / _HULL_RATIO override) */
	/* The pairs of a registerAll sweep are scans up to the pair gate (10 m) apart: a good part of the queries has no partner
	 * inside the search radius and pays every doubling of the round radius up to it.  A larger first radius and larger
	 * shared hulls suit that regime — measured on B200: 100 HDL-32E scans x 65 536 points 91.4 -> 74 ms per sweep, 12
	 * rotating-SICK scans x 1 M points 22.7 -> 19.6 ms — while the same-viewpoint pair loops are faster with the values above
	 * (1 M-point pair 116 vs 145 us).  Any values give the exact answer.  Env M3DREG_NN_SWEEP_RHO_DIV / _HULL_MIN / _HULL_RATIO. */
	NNTuning nn_tune_sweep = {8, 512, 32};
	double *scratch = nullptr;   /* 64 doubles */
	float *mats = nullptr;       /* 32 floats  */
	HostSmall *h = nullptr;      /* pinned */

	/* active fused loop (icp_begin .. icp_end) */
	bool active = false;
	const float4 *act_lx = nullptr, *act_ln = nullptr;
	const uint32_t *act_perm = nullptr;
	int act_n1 = 0, act_n2 = 0, act_sort_bits = 0;
	m3dreg_reg_params act_prm;

	/* per-stage profiling */
	bool profiling = false;
	cudaEvent_t pev[M3DREG_STAGE_COUNT + 1] = {nullptr, nullptr, nullptr, nullptr, nullptr};
	float stage_ms[M3DREG_STAGE_COUNT] = {0, 0, 0, 0};
	int stage_iters = 0;

	/* what the last fused iteration left behind (export hooks) */
	int last_n_first = 0, last_n_second = 0, last_sorted = 0;
	bool last_valid = false, last_nn_valid = true;
	double *neq_out_ext = nullptr;                /* m3dreg_icp_set_neq_out: extra destination of the fused loop's block */
	/* CUDA-graph replay of the fused iteration (m3dreg_icp_step): the ten PDL-chained launches of one iteration are captured
	 * once per loop (every kernel argument is a device pointer or a constant of the loop: the pose lives on the device) and
	 * replayed per step — one driver call instead of ten.  Up to two instantiations, keyed by the extra normal-equation
	 * destination (a multi-GPU caller alternates two buffers).  Env M3DREG_NO_GRAPH=1: plain launches. */
	struct IterGraph { cudaGraphExec_t exec = nullptr; double *neq_out = nullptr; int launches = 0; unsigned long long age = 0; };
	IterGraph it_graph[2];
	unsigned long long it_graph_clock = 0;
	bool it_graph_failed = false;
	int use_graph = 1;
	bool nn_pending = false;                      /* obs_rec (query order) is newer than nn (caller order) */
	const uint32_t *nn_pending_perm = nullptr;
};

namespace {

inline int grid_for(const m3dreg_ctx *c, long long n, int threads, int per_sm = 8)
{
	long long b = (n + threads - 1) / threads;
	long long cap = (long long)c->sm_count * per_sm;
	if (b > cap) b = cap;
	if (b < 1) b = 1;
	return (int)b;
}

/* Box pass: two 512-thread blocks per SM, every thread streams its points four loads at a time — each block ends with six
 * atomics on the same six words, and 1 184 small blocks measured 14 us where 148 large ones (the single-launch variant's
 * first phase) took 4.6 us for the same 16 MB. */
inline int box_pass_blocks(const m3dreg_ctx *c, long long n)
{
	long long b = (n + 511) / 512;
	long long cap = (long long)c->sm_count * 2;
	return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

/* Every kernel goes out with programmatic stream serialisation (see pdl_enter() in m3dreg_kernels.cuh): its blocks
 * are scheduled as the previous kernel's blocks exit and wait in griddepcontrol.wait for its completion.
 * M3DREG_NO_PDL=1 in the environment falls back to plain stream order. */
template <class... KArgs, class... Args>
inline void launch_kernel(m3dreg_ctx *c, void (*kernel)(KArgs...), int grid, int block, Args &&... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)grid, 1, 1);
	cfg.blockDim = dim3((unsigned)block, 1, 1);
	cfg.dynamicSmemBytes = 0;
	cfg.stream = c->stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = c->use_pdl ? 1 : 0;
	cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
	if (e != cudaSuccess && c->launch_err == cudaSuccess) c->launch_err = e;
	c->launches++;
}
#define LAUNCH(ctx, kernel, grid, block, ...) launch_kernel((ctx), kernel, (grid), (block), __VA_ARGS__)

/* status of the launches issued since the last call (launch errors are sticky in the context until reported) */
inline int launch_status(m3dreg_ctx *c)
{
	cudaError_t e = c->launch_err;
	c->launch_err = cudaSuccess;
	if (e == cudaSuccess) e = cudaGetLastError();
	return (int)e;
}

/* k_grid_build: one block per SM, all co-resident (cooperative launch: the kernel synchronises across the grid), chained
 * to the previous kernel with programmatic stream serialisation like every other launch when the driver takes both
 * attributes together (probed once per context). */
inline void launch_grid_build(m3dreg_ctx *c, const GridBuildArgs &a)
{
	if (!c->gb_attr_set) {
		cudaFuncSetAttribute(k_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize, kGbCacheMaxBytes);
		c->gb_attr_set = true;
	}
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)c->sm_count, 1, 1);
	cfg.blockDim = dim3((unsigned)kGbThreads, 1, 1);
	const int T = (((a.n + c->sm_count - 1) / c->sm_count) + 31) & ~31;      /* same expression as the kernel's */
	cfg.dynamicSmemBytes = a.cache ? (size_t)T * 12 : 0;
	cfg.stream = c->stream;
	cudaLaunchAttribute attr[2];
	attr[0].id = cudaLaunchAttributeCooperative;
	attr[0].val.cooperative = 1;
	attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[1].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cudaError_t e = cudaErrorUnknown;
	if (c->use_pdl && c->coop_pdl != 0) {
		cfg.numAttrs = 2;
		e = cudaLaunchKernelEx(&cfg, k_grid_build, a);
		if (c->coop_pdl < 0) c->coop_pdl = (e == cudaSuccess) ? 1 : 0;
		if (e != cudaSuccess) cudaGetLastError();
	}
	if (e != cudaSuccess) {
		cfg.numAttrs = 1;
		e = cudaLaunchKernelEx(&cfg, k_grid_build, a);
	}
	if (e != cudaSuccess && c->launch_err == cudaSuccess) c->launch_err = e;
	c->launches++;
}

int bits_for(long long nb);
long long sort_limit(const m3dreg_ctx *c, int bits);

int bits_for(long long nb)
{
	int bits = 1;
	while (bits < 31 && (1LL << bits) < nb) bits++;
	return bits;
}

/* Host replica of the tail of cudaCalculateGridParams (lesson_16.cu:64-91).  volatile stores keep every
 * operation a separately rounded IEEE float op whatever the host compiler's contraction setting. */
int grid_params_from_bounds(const float mn[3], const float mx[3], float rx, float ry, float rz, float ext, m3dreg_grid_params *out)
{
	volatile float mxx = mx[0], mxy = mx[1], mxz = mx[2], mnx = mn[0], mny = mn[1], mnz = mn[2];
	mxx = mxx + ext; mnx = mnx - ext;
	mxy = mxy + ext; mny = mny - ext;
	mxz = mxz + ext; mnz = mnz - ext;
	volatile float dx = mxx - mnx, dy = mxy - mny, dz = mxz - mnz;
	volatile float qx = dx / rx, qy = dy / ry, qz = dz / rz;
	volatile float fx = qx + 1.0f, fy = qy + 1.0f, fz = qz + 1.0f;
	int nbx = (int)fx, nby = (int)fy, nbz = (int)fz;
	memset(out, 0, sizeof(*out));
	out->bounding_box_min_X = mnx; out->bounding_box_min_Y = mny; out->bounding_box_min_Z = mnz;
	out->bounding_box_max_X = mxx; out->bounding_box_max_Y = mxy; out->bounding_box_max_Z = mxz;
	out->number_of_buckets_X = nbx; out->number_of_buckets_Y = nby; out->number_of_buckets_Z = nbz;
	out->resolution_X = rx; out->resolution_Y = ry; out->resolution_Z = rz;
	long long nb = (long long)nbx * nby * nbz;
	out->number_of_buckets = nb;
	if (nbx <= 0 || nby <= 0 || nbz <= 0 || nb > 2147483647LL || !nn_columns_usable(nbx, nby, nbz)) return M3DREG_E_TOO_MANY_BUCKETS;
	return 0;
}

float o2f_host(uint32_t o)
{
	uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
	float f;
	memcpy(&f, &u, 4);
	return f;
}

int ensure_first(m3dreg_ctx *c, size_t n)
{
	int e;
	if ((e = c->g_xyzl.ensure(n))) return e;
	if ((e = c->g_nrm.ensure(n))) return e;
	if ((e = c->digit_tot.ensure(kRadixSize))) return e;
	for (int k = 0; k < 2; k++) {
		if ((e = c->keys[k].ensure(n))) return e;
		if ((e = c->vals[k].ensure(n))) return e;
	}
	size_t tiles = (n + 1023) / 1024 + 1;
	size_t hist_need = 4 * tiles * kRadixSize;                       /* one digit-by-tile matrix per radix pass */
	if (hist_need < (size_t)3 * c->sm_count * kGbRadix) hist_need = (size_t)3 * c->sm_count * kGbRadix;   /* k_grid_build: [block][digit] x 3 passes */
	if ((e = c->hist.ensure(hist_need))) return e;
	if ((e = c->brec.ensure((size_t)c->sm_count))) return e;
	if ((e = c->cell_list.ensure(n))) return e;                      /* searchable buckets: at most one per point */
	return 0;
}

/* Dense bucket table + the megakernel's count arrays for `cap` buckets.  bcount must be all zero between launches:
 * a (re)allocation is cleared here, k_grid_build leaves it clean, rearm_grid() clears it after a rejected grid. */
int ensure_buckets(m3dreg_ctx *c, size_t cap, bool ndt)
{
	int e;
	if ((e = c->buckets.ensure(cap))) return e;
	if (c->bcount.cap < c->buckets.cap) {
		if ((e = c->bcount.ensure(c->buckets.cap))) return e;
		cudaError_t ce = cudaMemsetAsync(c->bcount.p, 0, c->bcount.cap * sizeof(int), c->stream);
		if (ce != cudaSuccess) return (int)ce;
	}
	if ((e = c->bbegin.ensure(c->buckets.cap))) return e;
	if (ndt) {
		if ((e = c->ndt_acc.ensure(c->buckets.cap * 12))) return e;
		if ((e = c->ndt_qacc.ensure(c->buckets.cap * 4))) return e;
		if ((e = c->ndt_iacc.ensure(c->buckets.cap * 12))) return e;
	}
	return 0;
}

/* Bucket capacity that holds the grid of a scan under ANY rigid pose: every axis extent of the transformed cloud is at
 * most the diagonal of its local bounding box, so nb_axis <= (diag + 2 ext) / res + 1 (+2: float slack).  No read-back,
 * no synchronisation (round 1 sized the table from the initial box: one D2H + sync per fused loop and per scan of a
 * sweep, and a pose that drifted out of the margin aborted the loop). */
long long bucket_capacity_for(float diag, const m3dreg_reg_params *prm)
{
	double per_axis = floor(((double)diag * 1.0001 + 2.0 * (double)prm->bbox_extension) / (double)prm->bucket_size) + 3.0;
	if (!(per_axis >= 1.0)) per_axis = 1.0;
	double cap = per_axis * per_axis * per_axis;
	if (cap > 2147483647.0) cap = 2147483647.0;
	return (long long)cap;
}

/* Radix bits for the grid of a scan at (about) the given pose, WITHOUT looking at the device: the transformed cloud lies
 * inside the transformed local bounding box, whose axis-aligned extent bounds every nb_axis; `margin` extra cells per
 * axis absorb the pose changes of a registration loop.  The sort runs ceil(bits / 8) passes and therefore orders any
 * grid of up to 2^(8 passes) buckets correctly — that (and the allocated table) is what the key kernel checks the actual
 * bucket count against on the device (sort_limit()). */
int planned_sort_bits(const float bb_min[3], const float bb_max[3], const float *pose16, const m3dreg_reg_params *prm, int margin)
{
	double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
	for (int k = 0; k < 8; k++) {
		const double x = (k & 1) ? bb_max[0] : bb_min[0], y = (k & 2) ? bb_max[1] : bb_min[1], z = (k & 4) ? bb_max[2] : bb_min[2];
		for (int r = 0; r < 3; r++) {
			const double v = pose16 ? (double)pose16[4 * r] * x + (double)pose16[4 * r + 1] * y + (double)pose16[4 * r + 2] * z + (double)pose16[4 * r + 3]
					: (r == 0 ? x : (r == 1 ? y : z));
			if (v < lo[r]) lo[r] = v;
			if (v > hi[r]) hi[r] = v;
		}
	}
	double cap = 1.0;
	for (int r = 0; r < 3; r++) {
		double nb = floor(((hi[r] - lo[r]) * 1.0001 + 2.0 * (double)prm->bbox_extension) / (double)prm->bucket_size) + 2.0 + (double)margin;
		if (!(nb >= 1.0)) nb = 1.0;
		cap *= nb;
	}
	if (cap > 2147483647.0) cap = 2147483647.0;
	return bits_for((long long)cap);
}

/* largest bucket count a sort planned for `bits` radix bits orders correctly, capped by the allocated table */
long long sort_limit(const m3dreg_ctx *c, int bits)
{
	int passes = (bits + kRadixBits - 1) / kRadixBits;
	if (passes < 1) passes = 1;
	long long lim = passes * kRadixBits >= 31 ? 2147483647LL : (1LL << (passes * kRadixBits));
	return lim < (long long)c->buckets.cap ? lim : (long long)c->buckets.cap;
}

int ensure_second(m3dreg_ctx *c, size_t n)
{
	int e;
	if ((e = c->q_xyzl.ensure(n))) return e;
	if ((e = c->q_nrm.ensure(n))) return e;
	if ((e = c->nn.ensure(n))) return e;
	if ((e = c->obs_rec.ensure(n))) return e;
	return 0;
}

int ensure_partials(m3dreg_ctx *c)
{
	return c->partials.ensure((size_t)c->sm_count * 8 * kPartialCols);
}

#ifndef M3D_SORT_ITEMS
#define M3D_SORT_ITEMS 8
#endif
constexpr int kSortItemsBig = M3D_SORT_ITEMS;   /* keys per thread of a sort tile for large inputs */

struct SortPlan { int items, tiles, passes; };

SortPlan plan_sort(int n, int bits)
{
	SortPlan sp;
	sp.passes = (bits + kRadixBits - 1) / kRadixBits;
	if (sp.passes < 1) sp.passes = 1;
	sp.items = n >= (1 << 19) ? kSortItemsBig : 4;
	sp.tiles = (n + kSortThreads * sp.items - 1) / (kSortThreads * sp.items);
	return sp;
}

/* stable LSD radix sort of keys[0] (+ vals[0], or the implicit indices 0..n-1 when vals_implicit) by the low `bits`
 * bits; returns the index (0/1) of the buffers holding the result.  gp (device, may be null) lets every pass no-op
 * when the grid was rejected on the device.  hist0_ready: the first pass's histogram matrix was already produced
 * (and the later ones zeroed) by k_grid_head.  Pass p+1's histogram is accumulated by pass p's scatter. */
int sort_by_bucket(m3dreg_ctx *c, int n, int bits, const m3dreg_grid_params *gp, bool hist0_ready = false, bool vals_implicit = false)
{
	SortPlan sp = plan_sort(n, bits);
	const size_t mat = (size_t)kRadixSize * sp.tiles;
	const bool big = sp.items == kSortItemsBig;
	int cur = 0;
	if (!hist0_ready) {
		if (sp.passes > 1) cudaMemsetAsync(c->hist.p + mat, 0, (size_t)(sp.passes - 1) * mat * sizeof(uint32_t), c->stream);
		if (big) LAUNCH(c, k_radix_hist<kSortItemsBig>, sp.tiles, kSortThreads, c->keys[0].p, n, 0, sp.tiles, c->hist.p, gp);
		else LAUNCH(c, k_radix_hist<4>, sp.tiles, kSortThreads, c->keys[0].p, n, 0, sp.tiles, c->hist.p, gp);
	}
	for (int p = 0; p < sp.passes; p++) {
		int shift = p * kRadixBits;
		uint32_t *h = c->hist.p + (size_t)p * mat;
		uint32_t *hn = p + 1 < sp.passes ? h + mat : nullptr;
		const uint32_t *vin = (p == 0 && vals_implicit) ? nullptr : c->vals[cur].p;
		LAUNCH(c, k_radix_scan, kRadixSize, 256, h, sp.tiles, c->digit_tot.p, gp);
		if (big) LAUNCH(c, k_radix_scatter<kSortItemsBig>, sp.tiles, kSortThreads, c->keys[cur].p, vin, c->keys[cur ^ 1].p, c->vals[cur ^ 1].p, n, shift, sp.tiles, h, c->digit_tot.p, gp, hn);
		else LAUNCH(c, k_radix_scatter<4>, sp.tiles, kSortThreads, c->keys[cur].p, vin, c->keys[cur ^ 1].p, c->vals[cur ^ 1].p, n, shift, sp.tiles, h, c->digit_tot.p, gp, hn);
		cur ^= 1;
	}
	return cur;
}

void host_roundtrip_pose(const float *m, float *pose1, double *pose6)
{
	float of[3], t[3];
	matrix4_to_euler(m, of, t);
	euler_to_matrix(of, t, pose1);
	if (pose6) {
		pose6[0] = t[0]; pose6[1] = t[1]; pose6[2] = t[2];
		pose6[3] = of[0]; pose6[4] = of[1]; pose6[5] = of[2];
	}
}

/* Candidate sets: the OUTER set is only materialised when the caps differ. */
int ensure_candidates(m3dreg_ctx *c, size_t n1, int max_inner, int max_outer)
{
	int e;
	if ((e = c->ci_xyzl.ensure(n1))) return e;
	if ((e = c->ci_nrm.ensure(n1))) return e;
	if ((e = c->ci_loc.ensure(n1))) return e;
	if ((e = c->ci_tab.ensure(2 * n1 + 8))) return e;
	if (max_inner != max_outer) {
		if ((e = c->co_xyzl.ensure(n1))) return e;
		if ((e = c->co_nrm.ensure(n1))) return e;
		if ((e = c->co_loc.ensure(n1))) return e;
		if ((e = c->co_tab.ensure(2 * n1 + 8))) return e;
	}
	return 0;
}

CandSet cand_set(m3dreg_ctx *c, bool outer)
{
	CandSet s;
	if (outer) { s.xyzl = c->co_xyzl.p; s.nrm = c->co_nrm.p; s.tab = c->co_tab.p; s.loc = c->co_loc.p; }
	else { s.xyzl = c->ci_xyzl.p; s.nrm = c->ci_nrm.p; s.tab = c->ci_tab.p; s.loc = c->ci_loc.p; }
	return s;
}

/* gp must already be on the device (c->gp); src = the gridded cloud in original order (global frame); loc_src = the same
 * cloud in the frame the moment reduction wants (local; 0 = src); cell_list / c->cell_count hold the searchable buckets. */
void build_candidates(m3dreg_ctx *c, const uint32_t *vals, const m3dreg_bucket *buckets, const uint32_t *cell_list,
		const float4 *src_xyzl, const float4 *src_nrm, const float4 *loc_src, const float *nrm_m, bool xform_points, int max_inner, int max_outer)
{
	bool two = max_inner != max_outer;
	LAUNCH(c, k_build_candidates, c->sm_count * 7, kBuildWarps * 32, vals, c->gp, buckets, cell_list, c->cell_count, src_xyzl, src_nrm, loc_src, nrm_m,
			xform_points ? 1 : 0, max_inner, max_outer, cand_set(c, false), cand_set(c, two), two ? 1 : 0);
}

/* src_xyzl: the first cloud as the moment reduction wants it (local frame in the fused loops), original order;
 * res: the grid's resolution per axis (what c->gp holds on the device) */
void launch_nn(m3dreg_ctx *c, const uint32_t *q_perm, int n2, const uint32_t *vals, int n1, const m3dreg_bucket *buckets, const float res[3],
		float radius, int max_inner, int max_outer, int prune, int *nn_out, float4 *obs_rec, const float4 *src_xyzl,
		unsigned long long *label_counts, const int *seg_of_chunk = nullptr)
{
	bool two = max_inner != max_outer;
	if (nn_out == c->nn.p) c->nn_pending = false;      /* the caller-order buffer is being written directly */
	if (!two && !c->nn_per_thread && !c->nn_v7) {      /* one candidate set (the reference's default caps): warp-shared hull search */
		NNHullArgs a = {};
		a.q_xyzl = c->q_xyzl.p; a.q_nrm = c->q_nrm.p; a.q_perm = q_perm; a.n_second = n2;
		a.cs = cand_set(c, false); a.s_vals = vals; a.n_first = n1; a.buckets = buckets; a.gp = c->gp;
		a.search_radius = radius;
		/* launch constants, same IEEE operations as nn_params_finish() / the round-1 kernel computed per thread */
		{
			volatile float r2 = radius * radius, iwx = 4.0f / res[0], iwy = 4.0f / res[1], iwz = 4.0f / res[2];
			volatile float rmin = fminf(res[0], fminf(res[1], res[2])) / (float)(c->nn_tune.rho_div > 0 ? c->nn_tune.rho_div : 16);
			volatile float rho2 = rmin * rmin;
			a.r2 = r2; a.iwx = iwx; a.iwy = iwy; a.iwz = iwz; a.rho2_first = fmaxf(rho2, 1.0e-30f);
		}
		a.cap = max_outer; a.prune = prune; a.tune = c->nn_tune;
		a.nn_out = nn_out; a.obs_rec = obs_rec; a.src_xyzl = src_xyzl; a.label_counts = label_counts;
		a.eval_counter = c->profiling ? c->eval_counter : nullptr; a.seg_of_chunk = seg_of_chunk;
		a.work = c->nn_work;
		if (c->profiling && c->nn_diag) { cudaMemsetAsync(c->gb_dbg, 0, 16 * sizeof(unsigned long long), c->stream); a.diag = c->gb_dbg; }
		/* persistent warps: one wave of resident blocks, chunks of 32 queries handed out by an atomic counter */
		const int pi = c->profiling ? 1 : 0;
		if (c->nn_blocks_per_sm[pi] <= 0) {
			int nb = 0;
			cudaError_t oe = pi ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_nn_search_hull<true>, kNNHThreads, 0)
					: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_nn_search_hull<false>, kNNHThreads, 0);
			c->nn_blocks_per_sm[pi] = (oe == cudaSuccess && nb > 0) ? nb : 8;
		}
		const int chunks = (n2 + 31) / 32;
		int blocks = c->sm_count * c->nn_blocks_per_sm[pi];
		if (blocks > (chunks + kNNHWarps - 1) / kNNHWarps) blocks = (chunks + kNNHWarps - 1) / kNNHWarps;
		if (blocks < 1) blocks = 1;
		if (c->profiling) LAUNCH(c, k_nn_search_hull<true>, blocks, kNNHThreads, a);
		else LAUNCH(c, k_nn_search_hull<false>, blocks, kNNHThreads, a);
		return;
	}
	if (!two && !c->nn_per_thread) {
		LAUNCH(c, k_nn_search_grid, (n2 + kNNGThreads - 1) / kNNGThreads, kNNGThreads, c->q_xyzl.p, c->q_nrm.p, q_perm, n2,
				cand_set(c, false), vals, n1, buckets, c->gp, radius, max_outer, prune, c->nn_tune, nn_out, obs_rec, src_xyzl, label_counts,
				c->profiling ? c->eval_counter : nullptr, seg_of_chunk);
		return;
	}
	LAUNCH(c, k_nn_search, (n2 + kNNThreads - 1) / kNNThreads, kNNThreads, c->q_xyzl.p, c->q_nrm.p, q_perm, n2,
			cand_set(c, false), cand_set(c, two), vals, n1, buckets, c->gp, radius, max_inner, max_outer, prune, nn_out, obs_rec, src_xyzl,
			label_counts, c->profiling ? c->eval_counter : nullptr, seg_of_chunk);
}

/* Multi-kernel grid build (PDL-chained launches): params (device) from the reduced box, keys, stable LSD sort, dense
 * table, candidate sets.  (src_xyzl, src_nrm) = the cloud in original order: with pose != 0 the stored scan in its local
 * frame, transformed on the fly by the key pass and the candidate gather exactly as the box pass did (the transformed
 * cloud is never stored); with pose == 0 a cloud that is already global.  bounds must already hold the reduced box.
 * Measured against the single-launch variant (build_grid_mega) on B200: 59 + 17 us vs 82 us on the 1 M-point pair,
 * 49 + 8 vs 106 us on the 65 k-point pair — this is the default. */
void build_grid_legacy(m3dreg_ctx *c, const float4 *src_xyzl, const float4 *src_nrm, int n1, const float *pose, const m3dreg_reg_params *prm, int sort_bits)
{
	SortPlan sp = plan_sort(n1, sort_bits);
	if (sp.items == kSortItemsBig)
		LAUNCH(c, k_grid_head<kSortItemsBig>, sp.tiles, kSortThreads, src_xyzl, pose, n1, c->bounds, prm->bucket_size, prm->bucket_size, prm->bucket_size,
				prm->bbox_extension, sort_limit(c, sort_bits), c->gp, c->flags, c->cell_count, c->buckets.p, c->keys[0].p, sp.tiles, sp.passes, c->hist.p);
	else
		LAUNCH(c, k_grid_head<4>, sp.tiles, kSortThreads, src_xyzl, pose, n1, c->bounds, prm->bucket_size, prm->bucket_size, prm->bucket_size,
				prm->bbox_extension, sort_limit(c, sort_bits), c->gp, c->flags, c->cell_count, c->buckets.p, c->keys[0].p, sp.tiles, sp.passes, c->hist.p);
	int cur = sort_by_bucket(c, n1, sort_bits, c->gp, true, true);
	LAUNCH(c, k_finalize_grid, grid_for(c, n1, 256), 256, c->keys[cur].p, c->vals[cur].p, n1, c->gp, c->buckets.p,
			(m3dreg_hash_element *)nullptr, c->cell_list.p, c->cell_count);
	if (prm->mode != M3DREG_MODE_NDT)
		build_candidates(c, c->vals[cur].p, c->buckets.p, c->cell_list.p, src_xyzl, src_nrm, (const float4 *)nullptr, pose, pose != nullptr, prm->max_inner, prm->max_outer);
	c->last_sorted = cur;
}

/* The grid of one iteration as ONE launch (grid_build.cuh): transform of (lx, ln) by the device pose + bounding box,
 * grid parameters, keys, stable sort, dense bucket table, candidate sets.  pose == 0: the cloud is already global.
 * The transformed cloud is only kept (g_xyzl) when keep_global is set (NDT statistics). */
void build_grid_mega(m3dreg_ctx *c, const float4 *lx, const float4 *ln, int n1, const float *pose, const m3dreg_reg_params *prm, bool keep_global)
{
	GridBuildArgs a = {};
	a.lx = lx; a.ln = ln; a.n = n1; a.pose = pose;
	{
		const int T = (((n1 + c->sm_count - 1) / c->sm_count) + 31) & ~31;
		a.cache = ((size_t)T * 12 <= (size_t)kGbCacheMaxBytes) ? 1 : 0;
	}
	a.dbg = c->profiling ? c->gb_dbg : nullptr;
	a.res = prm->bucket_size; a.ext = prm->bbox_extension;
	a.bucket_cap = (long long)c->buckets.cap;
	a.max_inner = prm->max_inner; a.max_outer = prm->max_outer;
	a.build_cands = prm->mode != M3DREG_MODE_NDT ? 1 : 0;
	a.g_xyzl = keep_global ? c->g_xyzl.p : nullptr;
	a.keys[0] = c->keys[0].p; a.keys[1] = c->keys[1].p; a.vals[0] = c->vals[0].p; a.vals[1] = c->vals[1].p;
	a.hist = c->hist.p; a.bcount = c->bcount.p; a.bbegin = c->bbegin.p; a.buckets = c->buckets.p;
	a.cell_list = c->cell_list.p; a.cell_count = c->cell_count; a.brec = c->brec.p; a.gp = c->gp; a.flags = c->flags;
	a.bounds = c->bounds; a.bar = c->grid_bar;
	a.ci = cand_set(c, false); a.co = cand_set(c, prm->max_inner != prm->max_outer);
	launch_grid_build(c, a);
	c->last_sorted = 1;
}

/* Reads the reduced bounds back (one sync), sizes the dense bucket table with a margin so the box may drift
 * while the pose converges, and returns the radix bit count for that capacity. */
int plan_buckets(m3dreg_ctx *c, const m3dreg_reg_params *prm, int *sort_bits)
{
	CK(cudaMemcpyAsync(c->h->bounds, c->bounds, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	float mn[3], mx[3];
	for (int k = 0; k < 3; k++) { mn[k] = o2f_host(c->h->bounds[k]); mx[k] = o2f_host(c->h->bounds[3 + k]); }
	m3dreg_grid_params gp;
	int st = grid_params_from_bounds(mn, mx, prm->bucket_size, prm->bucket_size, prm->bucket_size, prm->bbox_extension, &gp);
	if (st) return st;
	long long cap = (long long)(gp.number_of_buckets_X + 4) * (gp.number_of_buckets_Y + 4) * (gp.number_of_buckets_Z + 4);
	if (cap > 2147483647LL) cap = 2147483647LL;
	if (gp.number_of_buckets > cap) return M3DREG_E_TOO_MANY_BUCKETS;
	int e = ensure_buckets(c, (size_t)cap, prm->mode == M3DREG_MODE_NDT);
	if (e) return e;
	*sort_bits = bits_for((long long)c->buckets.cap);
	return 0;
}

int check_flags(m3dreg_ctx *c)
{
	CK(cudaMemcpyAsync(c->h->flags, c->flags, sizeof(int) * FLAG_COUNT, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	int f = c->h->flags[FLAG_ERROR];
	if (f) {
		CK(cudaMemsetAsync(c->flags, 0, sizeof(int) * FLAG_COUNT, c->stream));
	}
	return f;
}

bool valid_params(const m3dreg_reg_params *p)
{
	if (!p) return false;
	if (!(p->bucket_size > 0.0f) || !(p->search_radius >= 0.0f)) return false;
	if (p->dof != 6 && p->dof != 4) return false;
	if (p->mode != M3DREG_MODE_ICP && p->mode != M3DREG_MODE_NDT) return false;
	return true;
}

/* NDT: per-bucket statistics of the gridded cloud (once per grid) and the query pass + bucket reduction (per pair). */
/* fixed-point scales: |p - centre| <= res / 2 (+ rounding), |local| <= local_mag; up to 2^22 points per bucket */
NdtScales ndt_scales(float res, float local_mag)
{
	double p = 1.0;
	while (p * 4.0 < (double)res) p *= 2.0;
	double q = 1.0;
	while (q < (double)local_mag) q *= 2.0;
	NdtScales sc;
	sc.s1 = 274877906944.0 / p;            /* 2^38 / p        */
	sc.s2 = 274877906944.0 / (p * p);
	sc.sl = 549755813888.0 / q;            /* 2^39 / q        */
	return sc;
}

void ndt_bucket_stats(m3dreg_ctx *c, const float4 *lx, int n1, const m3dreg_reg_params *prm, float local_mag)
{
	int cur = c->last_sorted;
	const NdtScales sc = ndt_scales(prm->bucket_size, local_mag);
	LAUNCH(c, k_ndt_zero, grid_for(c, (long long)c->buckets.cap * 16, 256), 256, c->ndt_iacc.p, c->ndt_qacc.p, c->gp, 1);
	LAUNCH(c, k_ndt_accumulate_points, grid_for(c, n1, 256), 256, c->keys[cur].p, c->vals[cur].p, n1, c->g_xyzl.p, lx, c->gp, c->ndt_iacc.p, sc);
	LAUNCH(c, k_ndt_finalize_buckets, grid_for(c, (long long)c->buckets.cap, 256), 256, c->buckets.p, c->gp, c->ndt_iacc.p, c->ndt_acc.p, sc);
}

void ndt_queries_and_reduce(m3dreg_ctx *c, int n2, const FinalizeArgs &fin, bool zero_qacc)
{
	if (zero_qacc) LAUNCH(c, k_ndt_zero, grid_for(c, (long long)c->buckets.cap * 4, 256), 256, c->ndt_iacc.p, c->ndt_qacc.p, c->gp, 0);
	LAUNCH(c, k_ndt_accumulate_queries, grid_for(c, n2, 256), 256, c->q_xyzl.p, n2, c->gp, c->ndt_acc.p, c->ndt_qacc.p);
	LAUNCH(c, k_ndt_normal_equations, grid_for(c, (long long)c->buckets.cap, kNeqThreads, 2), kNeqThreads, c->ndt_acc.p, c->ndt_qacc.p, c->gp,
			c->partials.p, c->ticket, fin);
}

/* nn[] in the caller's order from the last fused iteration's query-order records */
void materialize_nn(m3dreg_ctx *c)
{
	if (!c->nn_pending) return;
	LAUNCH(c, k_scatter_nn, grid_for(c, c->last_n_second, 256), 256, c->nn_pending_perm, c->last_n_second, c->obs_rec.p, c->nn.p);
	c->nn_pending = false;
}

void stage_events_collect(m3dreg_ctx *c)
{
	cudaEventRecord(c->pev[4], c->stream);
	cudaEventSynchronize(c->pev[4]);
	for (int k = 0; k < M3DREG_STAGE_COUNT; k++) {
		float ms = 0.0f;
		cudaEventElapsedTime(&ms, c->pev[k], c->pev[k + 1]);
		c->stage_ms[k] += ms;
	}
	c->stage_iters++;
}

/* One registerLastArrivedScan iteration, fully on the device, nothing read back: box pass, key pass, sort passes, bucket
 * table, candidate sets, search, moment reduction + solve (PDL-chained launches).  first local cloud = (lx, ln), queries
 * already in q_*. */
void drop_iteration_graphs(m3dreg_ctx *c)
{
	for (auto &g : c->it_graph) {
		if (g.exec) cudaGraphExecDestroy(g.exec);
		g = m3dreg_ctx::IterGraph();
	}
	c->it_graph_failed = false;
}

void icp_iteration_device(m3dreg_ctx *c, const float4 *lx, const float4 *ln, int n1, int n2,
		const m3dreg_reg_params *prm, int sort_bits)
{
	const bool prof = c->profiling;
	const bool ndt = prm->mode == M3DREG_MODE_NDT;
	if (prof) cudaEventRecord(c->pev[0], c->stream);
	if (!c->grid_mega) {
		/* box pass: transform in registers, nothing stored (NDT keeps the transformed cloud for its bucket statistics) */
		LAUNCH(c, k_transform_soa<true>, box_pass_blocks(c, n1), 512, lx, ln, n1, c->ps->pose1, ndt ? c->g_xyzl.p : (float4 *)nullptr, (float4 *)nullptr, c->bounds);
		if (prof) cudaEventRecord(c->pev[1], c->stream);
		build_grid_legacy(c, lx, ln, n1, c->ps->pose1, prm, sort_bits);
	} else {
		if (prof) cudaEventRecord(c->pev[1], c->stream);      /* the transform is part of the grid launch */
		build_grid_mega(c, lx, ln, n1, c->ps->pose1, prm, ndt);
	}
	if (c->q_deferred_ev) {      /* host-buffer call: the queries' upload was left running under the grid build */
		cudaStreamWaitEvent(c->stream, c->q_deferred_ev, 0);
		LAUNCH(c, k_unpack_points, (c->q_deferred_n + 255) / 256, 256, c->q_deferred_aos, c->q_deferred_n, c->q_xyzl.p, c->q_nrm.p);
		c->q_deferred_ev = nullptr;
	}
	FinalizeArgs fin = {};
	fin.ps = c->ps; fin.neq_out = c->neq_out_ext; fin.accumulate = 0; fin.solve = 1; fin.dof = prm->dof;
	fin.obs_threshold = prm->obs_threshold; fin.pose6_in = nullptr; fin.bounds_reset = c->bounds;
	if (ndt) {
		ndt_bucket_stats(c, lx, n1, prm, c->act_local_mag);
		if (prof) { cudaEventRecord(c->pev[2], c->stream); cudaEventRecord(c->pev[3], c->stream); }
		fin.label_counts_reset = nullptr;
		ndt_queries_and_reduce(c, n2, fin, false);
		if (prof) stage_events_collect(c);
		c->last_n_first = n1; c->last_n_second = n2; c->last_valid = true; c->last_nn_valid = false;
		return;
	}
	if (prof) cudaEventRecord(c->pev[2], c->stream);
	/* the caller-order copy of the correspondences is only materialised when somebody asks for it (materialize_nn) */
	const float res3[3] = {prm->bucket_size, prm->bucket_size, prm->bucket_size};
	launch_nn(c, c->act_perm, n2, c->vals[c->last_sorted].p, n1, c->buckets.p, res3, prm->search_radius, prm->max_inner, prm->max_outer, 1,
			nullptr, c->obs_rec.p, lx, c->label_counts);
	c->nn_pending = true; c->nn_pending_perm = c->act_perm;
	ObsFromRec src = {};
	src.n_segs = 1;
	src.rec = c->obs_rec.p; src.q_xyzl = c->q_xyzl.p; src.m = c->ps->pose1; src.label_counts = c->label_counts;
	for (int k = 0; k < 4; k++) src.weight[k] = prm->weight[k];
	fin.label_counts_reset = c->label_counts;
	if (prof) cudaEventRecord(c->pev[3], c->stream);
	LAUNCH(c, k_normal_equations<ObsFromRec>, grid_for(c, n2, kNeqThreads, 2), kNeqThreads, src, n2, c->partials.p, c->ticket, fin);
	c->last_nn_valid = true;
	if (prof) stage_events_collect(c);
	c->last_n_first = n1; c->last_n_second = n2; c->last_valid = true;
}

void fill_stats(m3dreg_ctx *c, m3dreg_icp_stats *stats, float ms)
{
	if (!stats) return;
	const PoseState &ps = c->h->ps;
	stats->iterations_run = ps.iterations;
	stats->last_status = ps.status;
	stats->n_obs_last = ps.n_obs;
	stats->n_buckets_last = c->h->gp.number_of_buckets;
	for (int k = 0; k < 6; k++) stats->x_last[k] = ps.x[k];
	stats->device_ms = ms;
}

} /* namespace */

extern "C" {

int m3dreg_version(void) { return M3DREG_VERSION; }

const char *m3dreg_status_string(int status)
{
	switch (status) {
	case M3DREG_OK: return "ok";
	case M3DREG_E_INVALID_ARG: return "invalid argument";
	case M3DREG_E_TOO_MANY_BUCKETS: return "bucket count exceeds int32 / planned capacity";
	case M3DREG_E_NOT_SPD: return "normal equations not positive definite";
	case M3DREG_E_TOO_FEW_OBS: return "too few observations";
	case M3DREG_E_BAD_SLOT: return "bad scan slot";
	case M3DREG_E_NO_DEVICE: return "no sm_100 CUDA device";
	case M3DREG_E_SIZE_MISMATCH: return "size mismatch";
	case M3DREG_E_NO_NCCL: return "NCCL not available (libnccl.so.2 not found) or no communicator attached";
	case M3DREG_E_NCCL: return "NCCL call failed";
	case M3DREG_E_IO: return "file could not be read, written or parsed";
	default: break;
	}
	if (status > 0) return cudaGetErrorString((cudaError_t)status);
	return "unknown status";
}

int m3dreg_create(m3dreg_ctx **out, int cuda_device)
{
	if (!out) return M3DREG_E_INVALID_ARG;
	*out = nullptr;
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); return M3DREG_E_NO_DEVICE; }
	if (cuda_device < 0 || cuda_device >= count) return M3DREG_E_NO_DEVICE;
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, cuda_device));
	if (prop.major != 10) return M3DREG_E_NO_DEVICE;     /* the fatbin only carries sm_100a SASS */
	CK(cudaSetDevice(cuda_device));
	m3dreg_ctx *c = new (std::nothrow) m3dreg_ctx();
	if (!c) return (int)cudaErrorMemoryAllocation;
	c->dev = cuda_device;
	c->sm_count = prop.multiProcessorCount;
	{ const char *e = getenv("M3DREG_NO_PDL"); c->use_pdl = (e && e[0] == '1') ? 0 : 1; }
	{ const char *e = getenv("M3DREG_NN_PER_THREAD"); c->nn_per_thread = (e && e[0] == '1') ? 1 : 0; }
	{ const char *e = getenv("M3DREG_GRID_MEGA"); c->grid_mega = (e && e[0] == '1') ? 1 : 0; }
	{ const char *e = getenv("M3DREG_NN_V7"); c->nn_v7 = (e && e[0] == '1') ? 1 : 0; }
	{ const char *e = getenv("M3DREG_NN_RHO_DIV"); if (e && atoi(e) > 0) c->nn_tune.rho_div = atoi(e); }
	{ const char *e = getenv("M3DREG_NN_HULL_MIN"); if (e && atoi(e) > 0) c->nn_tune.hull_min = atoi(e); }
	{ const char *e = getenv("M3DREG_NN_HULL_RATIO"); if (e && atoi(e) > 0) c->nn_tune.hull_ratio = atoi(e); }
	{ const char *e = getenv("M3DREG_NO_GRAPH"); if (e && atoi(e) != 0) c->use_graph = 0; }
	{ const char *e = getenv("M3DREG_NN_DIAG"); if (e) c->nn_diag = atoi(e) != 0 ? 1 : 0; }
	{ const char *e = getenv("M3DREG_NN_SWEEP_RHO_DIV"); if (e && atoi(e) > 0) c->nn_tune_sweep.rho_div = atoi(e); }
	{ const char *e = getenv("M3DREG_NN_SWEEP_HULL_MIN"); if (e && atoi(e) > 0) c->nn_tune_sweep.hull_min = atoi(e); }
	{ const char *e = getenv("M3DREG_NN_SWEEP_HULL_RATIO"); if (e && atoi(e) > 0) c->nn_tune_sweep.hull_ratio = atoi(e); }
	cudaError_t e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
	if (e != cudaSuccess) { delete c; return (int)e; }
	c->stream = c->own_stream;
	cudaEventCreate(&c->ev0);
	cudaEventCreate(&c->ev1);
	size_t small = sizeof(PoseState) + 8 * sizeof(uint32_t) + sizeof(m3dreg_grid_params) + FLAG_COUNT * sizeof(int) +
			4 * sizeof(unsigned long long) + 64 + 64 * sizeof(double) + 32 * sizeof(float) + 1536;
	char *blk = nullptr;
	e = cudaMalloc((void **)&blk, small);
	if (e != cudaSuccess) { m3dreg_destroy(c); return (int)e; }
	cudaMemset(blk, 0, small);
	size_t off = 0;
	auto take = [&](size_t bytes) { char *p = blk + off; off += (bytes + 15) & ~(size_t)15; return p; };
	c->ps = (PoseState *)take(sizeof(PoseState));
	c->gp = (m3dreg_grid_params *)take(sizeof(m3dreg_grid_params));
	c->scratch = (double *)take(64 * sizeof(double));
	c->label_counts = (unsigned long long *)take(4 * sizeof(unsigned long long));
	c->bounds = (uint32_t *)take(8 * sizeof(uint32_t));
	c->flags = (int *)take(FLAG_COUNT * sizeof(int));
	c->ticket = (unsigned int *)take(16);
	c->cell_count = (unsigned int *)take(16);
	c->grid_bar = (unsigned int *)take(16);
	c->gb_dbg = (unsigned long long *)take(32 * sizeof(unsigned long long));
	c->nn_work = (unsigned int *)take(16);
	c->stamp_last = (unsigned long long *)take(16);
	c->eval_counter = (unsigned long long *)take(16);
	c->mats = (float *)take(32 * sizeof(float));
	e = cudaMallocHost((void **)&c->h, sizeof(HostSmall));
	if (e != cudaSuccess) { m3dreg_destroy(c); return (int)e; }
	memset(c->h, 0, sizeof(HostSmall));
	int rc = ensure_partials(c);
	if (rc) { m3dreg_destroy(c); return rc; }
	*out = c;
	return M3DREG_OK;
}

void m3dreg_destroy(m3dreg_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->dev);
	if (c->own_stream) cudaStreamSynchronize(c->own_stream);
	drop_iteration_graphs(c);
	m3dreg_nccl_attach(c, nullptr, 0, 1);      /* destroys a communicator this context created */
	c->d_neq.release();
	if (c->ev2) cudaEventDestroy(c->ev2);
	for (auto &s : c->scans) s.release();
	c->g_xyzl.release(); c->g_nrm.release(); c->ci_xyzl.release(); c->ci_nrm.release(); c->co_xyzl.release(); c->co_nrm.release(); c->digit_tot.release();
	c->ci_tab.release(); c->co_tab.release(); c->ci_loc.release(); c->co_loc.release();
	c->bcount.release(); c->bbegin.release(); c->brec.release();
	c->q_xyzl.release(); c->q_nrm.release(); c->l_xyzl.release(); c->l_nrm.release();
	for (int k = 0; k < 2; k++) { c->keys[k].release(); c->vals[k].release(); }
	c->hist.release(); c->buckets.release(); c->nn.release(); c->obs_rec.release(); c->cell_list.release(); c->aos_a.release(); c->aos_b.release();
	c->obs.release(); c->partials.release(); c->ndt_acc.release(); c->ndt_qacc.release(); c->ndt_iacc.release(); c->table.release(); c->d_poses1.release(); c->d_pose6.release(); c->d_sweep_status.release();
	c->d_segs.release(); c->d_seg_of_chunk.release(); c->d_seg_counts.release();
	c->pp_aos.release(); c->pp_markers.release(); c->pp_tiles.release(); c->d_group_ms.release();
	if (c->h_sweep) cudaFreeHost(c->h_sweep);
	if (c->ps) cudaFree(c->ps);   /* base of the small block */
	if (c->h) cudaFreeHost(c->h);
	for (int k = 0; k <= M3DREG_STAGE_COUNT; k++) if (c->pev[k]) cudaEventDestroy(c->pev[k]);
	for (auto &ev : c->ev_copy) if (ev) cudaEventDestroy(ev);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->ev0) cudaEventDestroy(c->ev0);
	if (c->ev1) cudaEventDestroy(c->ev1);
	if (c->own_stream) cudaStreamDestroy(c->own_stream);
	delete c;
}

int m3dreg_warm_up(m3dreg_ctx *c)
{
	if (!c) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	LAUNCH(c, k_reset_bounds, 1, 32, c->bounds);
	CK(cudaStreamSynchronize(c->stream));
	return (int)cudaGetLastError();
}

int m3dreg_set_stream(m3dreg_ctx *c, void *s)
{
	if (!c) return M3DREG_E_INVALID_ARG;
	CK(cudaStreamSynchronize(c->stream));
	c->stream = s ? (cudaStream_t)s : c->own_stream;
	return 0;      /* captured iteration graphs stay valid: a graph is launched into whatever stream is current */
}

void *m3dreg_get_stream(m3dreg_ctx *c) { return c ? (void *)c->stream : nullptr; }

int m3dreg_synchronize(m3dreg_ctx *c)
{
	if (!c) return M3DREG_E_INVALID_ARG;
	CK(cudaStreamSynchronize(c->stream));
	return (int)cudaGetLastError();
}

int64_t m3dreg_launch_count(const m3dreg_ctx *c) { return c ? c->launches : 0; }

int m3dreg_get_nn_evaluations(m3dreg_ctx *c, uint64_t *count_out, int reset)
{
	if (!c || !count_out) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	CK(cudaMemcpyAsync(c->h->label_counts, c->eval_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
	if (reset) CK(cudaMemsetAsync(c->eval_counter, 0, sizeof(unsigned long long), c->stream));
	CK(cudaStreamSynchronize(c->stream));
	*count_out = (uint64_t)c->h->label_counts[0];
	return 0;
}

int m3dreg_get_nn_fallbacks(m3dreg_ctx *c, uint64_t *count_out, int reset)
{
	if (!c || !count_out) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	CK(cudaMemcpyAsync(c->h->label_counts, c->eval_counter + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
	if (reset) CK(cudaMemsetAsync(c->eval_counter + 1, 0, sizeof(unsigned long long), c->stream));
	CK(cudaStreamSynchronize(c->stream));
	*count_out = (uint64_t)c->h->label_counts[0];
	return 0;
}

int m3dreg_get_grid_phase_ns(m3dreg_ctx *c, uint64_t *stamps_out)
{
	if (!c || !stamps_out) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	CK(cudaStreamSynchronize(c->stream));
	CK(cudaMemcpy(stamps_out, c->gb_dbg, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	return 0;
}

int m3dreg_set_pruning(m3dreg_ctx *c, int enabled)
{
	if (!c) return M3DREG_E_INVALID_ARG;
	c->prune = enabled ? 1 : 0;
	return 0;
}

/* ---- stage-level entry points ------------------------------------------------------------------------ */

int m3dreg_calculate_grid_params(m3dreg_ctx *c, const m3dreg_point *d_cloud, int n, float rx, float ry, float rz, float ext,
		m3dreg_grid_params *out)
{
	if (!c || !d_cloud || n <= 0 || !out || !(rx > 0) || !(ry > 0) || !(rz > 0)) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	LAUNCH(c, k_reset_bounds, 1, 32, c->bounds);
	LAUNCH(c, k_bounds_aos, grid_for(c, n, 256), 256, d_cloud, n, c->bounds);
	CK(cudaMemcpyAsync(c->h->bounds, c->bounds, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	float mn[3], mx[3];
	for (int k = 0; k < 3; k++) { mn[k] = o2f_host(c->h->bounds[k]); mx[k] = o2f_host(c->h->bounds[3 + k]); }
	return grid_params_from_bounds(mn, mx, rx, ry, rz, ext, out);
}

int m3dreg_calculate_grid(m3dreg_ctx *c, const m3dreg_point *d_cloud, int n, const m3dreg_grid_params *params,
		m3dreg_bucket *d_buckets, m3dreg_hash_element *d_table)
{
	if (!c || !d_cloud || n <= 0 || !params || !d_buckets || !d_table) return M3DREG_E_INVALID_ARG;
	if (params->number_of_buckets <= 0 || params->number_of_buckets > 2147483647LL) return M3DREG_E_TOO_MANY_BUCKETS;
	CK(cudaSetDevice(c->dev));
	int e = ensure_first(c, (size_t)n);
	if (e) return e;
	c->h->gp = *params;
	CK(cudaMemcpyAsync(c->gp, &c->h->gp, sizeof(m3dreg_grid_params), cudaMemcpyHostToDevice, c->stream));
	LAUNCH(c, k_keys_aos, grid_for(c, n, 256), 256, d_cloud, n, c->gp, c->keys[0].p, c->vals[0].p);
	int cur = sort_by_bucket(c, n, bits_for(params->number_of_buckets), nullptr);
	LAUNCH(c, k_init_buckets, grid_for(c, params->number_of_buckets * 3, 256), 256, d_buckets, (const m3dreg_grid_params *)nullptr,
			(long long)params->number_of_buckets);
	LAUNCH(c, k_finalize_grid, grid_for(c, n, 256), 256, c->keys[cur].p, c->vals[cur].p, n, (const m3dreg_grid_params *)nullptr, d_buckets, d_table,
			(uint32_t *)nullptr, (unsigned int *)nullptr);
	c->last_valid = false;
	CK(cudaStreamSynchronize(c->stream));   /* c->h->gp is reused by the next call */
	return (int)cudaGetLastError();
}

int m3dreg_nn_search(m3dreg_ctx *c, const m3dreg_point *d_first, int n1, const m3dreg_point *d_second, int n2,
		const m3dreg_hash_element *d_table, const m3dreg_bucket *d_buckets, const m3dreg_grid_params *params,
		float search_radius, int max_inner, int max_outer, int *d_nn)
{
	if (!c || !d_first || !d_second || n1 <= 0 || n2 <= 0 || !d_table || !d_buckets || !params || !d_nn) return M3DREG_E_INVALID_ARG;
	if (!nn_columns_usable(params->number_of_buckets_X, params->number_of_buckets_Y, params->number_of_buckets_Z)) return M3DREG_E_TOO_MANY_BUCKETS;
	CK(cudaSetDevice(c->dev));
	int e = ensure_first(c, (size_t)n1);
	if (e) return e;
	if ((e = ensure_second(c, (size_t)n2))) return e;
	c->h->gp = *params;
	CK(cudaMemcpyAsync(c->gp, &c->h->gp, sizeof(m3dreg_grid_params), cudaMemcpyHostToDevice, c->stream));
	LAUNCH(c, k_unpack_points, (n1 + 255) / 256, 256, d_first, n1, c->g_xyzl.p, c->g_nrm.p);
	LAUNCH(c, k_unpack_points, (n2 + 255) / 256, 256, d_second, n2, c->q_xyzl.p, c->q_nrm.p);
	if ((e = ensure_candidates(c, (size_t)n1, max_inner, max_outer))) return e;
	LAUNCH(c, k_split_table, grid_for(c, n1, 256), 256, d_table, n1, c->keys[0].p, c->vals[0].p);
	CK(cudaMemsetAsync(c->cell_count, 0, sizeof(unsigned int), c->stream));
	LAUNCH(c, k_list_cells, grid_for(c, n1, 256), 256, c->keys[0].p, n1, d_buckets, c->cell_list.p, c->cell_count);
	build_candidates(c, c->vals[0].p, d_buckets, c->cell_list.p, c->g_xyzl.p, c->g_nrm.p, (const float4 *)nullptr, (const float *)nullptr, false, max_inner, max_outer);
	const float res3[3] = {params->resolution_X, params->resolution_Y, params->resolution_Z};
	launch_nn(c, nullptr, n2, c->vals[0].p, n1, d_buckets, res3, search_radius, max_inner, max_outer, c->prune, d_nn, (float4 *)nullptr, c->g_xyzl.p, nullptr);
	c->last_valid = false;
	CK(cudaStreamSynchronize(c->stream));
	return (int)cudaGetLastError();
}

int m3dreg_transform(m3dreg_ctx *c, const m3dreg_point *d_in, m3dreg_point *d_out, int n, const float *m)
{
	if (!c || !d_in || !d_out || n <= 0 || !m) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	LAUNCH(c, k_transform_aos, (n + 255) / 256, 256, d_in, d_out, n, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11]);
	/* drop-in for cudaTransformPointCloud, which ends in cudaDeviceSynchronize (lesson_16.cu:1369-1384): the caller may
	 * touch d_out from any stream when this returns */
	CK(cudaStreamSynchronize(c->stream));
	return launch_status(c);
}

static int normal_equations_device(m3dreg_ctx *c, const m3dreg_obs_nn *d_obs, int n_obs, const double *pose6)
{
	/* result: packed 28 doubles in c->scratch[8..36); pose6 staged in c->scratch[0..6) */
	memcpy(c->h->scratch, pose6, 6 * sizeof(double));
	CK(cudaMemcpyAsync(c->scratch, c->h->scratch, 6 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	ObsFromList src;
	src.obs = d_obs;
	FinalizeArgs fin = {};
	fin.ps = nullptr; fin.neq_out = c->scratch + 8; fin.accumulate = 0; fin.solve = 0; fin.dof = 6; fin.obs_threshold = 0;
	fin.pose6_in = c->scratch; fin.bounds_reset = nullptr; fin.label_counts_reset = nullptr;
	LAUNCH(c, k_normal_equations<ObsFromList>, grid_for(c, n_obs, kNeqThreads, 4), kNeqThreads, src, n_obs, c->partials.p, c->ticket, fin);
	return 0;
}

static void unpack_neq(const double *neq, int dof, double *AtPA, double *AtPl)
{
	const int sel6[6] = {0, 1, 2, 3, 4, 5}, sel4[4] = {0, 1, 2, 5};
	const int *sel = dof == 6 ? sel6 : sel4;
	double full[6][6];
	int k = 0;
	for (int i = 0; i < 6; i++)
		for (int j = i; j < 6; j++) { full[i][j] = neq[k]; full[j][i] = neq[k]; k++; }
	for (int i = 0; i < dof; i++) {
		if (AtPl) AtPl[i] = neq[21 + sel[i]];
		if (AtPA) for (int j = 0; j < dof; j++) AtPA[i + j * dof] = full[sel[i]][sel[j]];
	}
}

int m3dreg_normal_equations(m3dreg_ctx *c, const m3dreg_obs_nn *d_obs, int n_obs, const double *pose6, int dof,
		double *AtPA_out, double *AtPl_out)
{
	if (!c || !d_obs || n_obs <= 0 || !pose6 || (dof != 6 && dof != 4)) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	int e = normal_equations_device(c, d_obs, n_obs, pose6);
	if (e) return e;
	CK(cudaMemcpyAsync(c->h->scratch + 8, c->scratch + 8, kNeqCount * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	unpack_neq(c->h->scratch + 8, dof, AtPA_out, AtPl_out);
	return (int)cudaGetLastError();
}

int m3dreg_solve_chol(m3dreg_ctx *c, const double *AtPA, const double *AtPl, int dof, double *x_out)
{
	if (!c || !AtPA || !AtPl || !x_out || (dof != 6 && dof != 4)) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	memcpy(c->h->scratch, AtPA, sizeof(double) * dof * dof);
	memcpy(c->h->scratch + 36, AtPl, sizeof(double) * dof);
	CK(cudaMemcpyAsync(c->scratch, c->h->scratch, 42 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	LAUNCH(c, k_solve_dense, 1, 32, c->scratch, c->scratch + 36, dof, c->scratch + 48, c->flags + 1);
	CK(cudaMemcpyAsync(c->h->scratch + 48, c->scratch + 48, 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaMemcpyAsync(c->h->flags, c->flags, FLAG_COUNT * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	memcpy(x_out, c->h->scratch + 48, sizeof(double) * dof);
	return c->h->flags[1];
}

int m3dreg_solve_observations(m3dreg_ctx *c, const m3dreg_obs_nn *d_obs, int n_obs, const double *pose6, int dof, double *x_out)
{
	if (!c || !d_obs || n_obs <= 0 || !pose6 || !x_out || (dof != 6 && dof != 4)) return M3DREG_E_INVALID_ARG;
	double N[36], b[6];
	int e = m3dreg_normal_equations(c, d_obs, n_obs, pose6, dof, N, b);
	if (e) return e;
	return m3dreg_solve_chol(c, N, b, dof, x_out);
}

/* ---- CCudaWrapper-level entry points on host buffers --------------------------------------------------- */

int m3dreg_semantic_nn_host(m3dreg_ctx *c, const m3dreg_point *first, int n1, const m3dreg_point *second, int n2,
		float search_radius, float bucket_size, float ext, int max_inner, int max_outer, int *nn_out)
{
	if (!c || !first || !second || n1 <= 0 || n2 <= 0 || !nn_out || !(bucket_size > 0)) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	int e;
	if ((e = c->aos_a.ensure((size_t)n1))) return e;
	if ((e = c->aos_b.ensure((size_t)n2))) return e;
	if ((e = ensure_first(c, (size_t)n1))) return e;
	if ((e = ensure_second(c, (size_t)n2))) return e;
	CK(cudaMemcpyAsync(c->aos_a.p, first, (size_t)n1 * sizeof(m3dreg_point), cudaMemcpyHostToDevice, c->stream));
	CK(cudaMemcpyAsync(c->aos_b.p, second, (size_t)n2 * sizeof(m3dreg_point), cudaMemcpyHostToDevice, c->stream));
	LAUNCH(c, k_reset_bounds, 1, 32, c->bounds);
	LAUNCH(c, k_bounds_aos, grid_for(c, n1, 256), 256, c->aos_a.p, n1, c->bounds);
	LAUNCH(c, k_unpack_points, (n1 + 255) / 256, 256, c->aos_a.p, n1, c->g_xyzl.p, c->g_nrm.p);
	LAUNCH(c, k_unpack_points, (n2 + 255) / 256, 256, c->aos_b.p, n2, c->q_xyzl.p, c->q_nrm.p);
	m3dreg_reg_params prm;
	memset(&prm, 0, sizeof(prm));
	prm.bucket_size = bucket_size; prm.bbox_extension = ext; prm.search_radius = search_radius;
	prm.max_inner = max_inner; prm.max_outer = max_outer;
	if ((e = ensure_candidates(c, (size_t)n1, max_inner, max_outer))) return e;
	int sort_bits = 0;
	if ((e = plan_buckets(c, &prm, &sort_bits))) return e;      /* sizes the bucket table from the box (one read-back) */
	if (!c->grid_mega) build_grid_legacy(c, c->g_xyzl.p, c->g_nrm.p, n1, (const float *)nullptr, &prm, sort_bits);
	else {
		LAUNCH(c, k_reset_bounds, 1, 32, c->bounds);              /* k_grid_build reduces the box itself */
		build_grid_mega(c, c->g_xyzl.p, c->g_nrm.p, n1, (const float *)nullptr, &prm, false);
	}
	const float res3[3] = {bucket_size, bucket_size, bucket_size};
	launch_nn(c, nullptr, n2, c->vals[c->last_sorted].p, n1, c->buckets.p, res3, search_radius, max_inner, max_outer, c->prune,
			c->nn.p, (float4 *)nullptr, c->g_xyzl.p, nullptr);
	CK(cudaMemcpyAsync(nn_out, c->nn.p, (size_t)n2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	c->last_n_first = n1; c->last_n_second = n2; c->last_valid = true;
	int f = check_flags(c);
	if (f) return f;
	return (int)cudaGetLastError();
}

int m3dreg_register_ls_host(m3dreg_ctx *c, const m3dreg_obs_nn *obs, int n_obs, double *pose6, int dof, double *x_out)
{
	if (!c || !obs || n_obs <= 0 || !pose6 || (dof != 6 && dof != 4)) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	int e = c->obs.ensure((size_t)n_obs);
	if (e) return e;
	CK(cudaMemcpyAsync(c->obs.p, obs, (size_t)n_obs * sizeof(m3dreg_obs_nn), cudaMemcpyHostToDevice, c->stream));
	double x[6] = {0, 0, 0, 0, 0, 0};
	e = m3dreg_solve_observations(c, c->obs.p, n_obs, pose6, dof, x);
	if (e) return e;
	/* cudaWrapper.cpp:574-579 / 641-646 */
	pose6[0] += x[0]; pose6[1] += x[1]; pose6[2] += x[2];
	if (dof == 6) { pose6[3] += x[3]; pose6[4] += x[4]; pose6[5] += x[5]; }
	else pose6[5] += x[3];
	if (x_out) memcpy(x_out, x, sizeof(double) * dof);
	return 0;
}

void m3dreg_matrix4_to_euler(const float *m, float *omfika, float *xyz) { matrix4_to_euler(m, omfika, xyz); }
void m3dreg_euler_to_matrix(const float *omfika, const float *xyz, float *m) { euler_to_matrix(omfika, xyz, m); }

/* ---- scan store ------------------------------------------------------------------------------------------ */

/* (label, Morton)-sorted copy of a scan for its QUERY role: stable radix sort by label | Morton code of a fine local
 * grid (cell = extent/512, at least 6.25 cm).  Any order is correct — the NN result does not depend on query order —
 * this one makes the 32 queries of a warp one small patch of one semantic surface. */
static int presort_scan(m3dreg_ctx *c, Scan &s, const m3dreg_point *d_aos)
{
	int n = s.n, e;
	if ((e = ensure_first(c, (size_t)n))) return e;
	LAUNCH(c, k_reset_bounds, 1, 32, c->bounds);
	LAUNCH(c, k_bounds_aos, grid_for(c, n, 256), 256, d_aos, n, c->bounds);
	CK(cudaMemcpyAsync(c->h->bounds, c->bounds, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	float mn[3], mx[3], ext = 0.0f;
	double d2 = 0.0;
	for (int k = 0; k < 3; k++) {
		mn[k] = o2f_host(c->h->bounds[k]); mx[k] = o2f_host(c->h->bounds[3 + k]);
		if (mx[k] - mn[k] > ext) ext = mx[k] - mn[k];
		d2 += ((double)mx[k] - (double)mn[k]) * ((double)mx[k] - (double)mn[k]);
		s.bb_min[k] = mn[k]; s.bb_max[k] = mx[k];
	}
	s.diag = (float)sqrt(d2);
	float res = ext / 511.0f;
	if (!(res > 0.0625f)) res = 0.0625f;
	LAUNCH(c, k_keys_presort, grid_for(c, n, 256), 256, d_aos, n, mn[0], mn[1], mn[2], 1.0f / res, c->keys[0].p, c->vals[0].p);
	int cur = sort_by_bucket(c, n, 29, nullptr);
	CK(cudaMemcpyAsync(s.perm, c->vals[cur].p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
	LAUNCH(c, k_gather_perm, grid_for(c, n, 256), 256, s.perm, n, s.xyzl, s.nrm, s.sx, s.sn);
	CK(cudaStreamSynchronize(c->stream));
	return (int)cudaGetLastError();
}

int m3dreg_scan_upload(m3dreg_ctx *c, int slot, const m3dreg_point *src, int n, int src_on_device)
{
	if (!c || slot < 0 || !src || n <= 0) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	if ((size_t)slot >= c->scans.size()) c->scans.resize((size_t)slot + 1);
	Scan &s = c->scans[(size_t)slot];
	if ((size_t)n > s.cap) {
		s.release();
		CK(cudaMalloc((void **)&s.xyzl, (size_t)n * sizeof(float4)));
		CK(cudaMalloc((void **)&s.nrm, (size_t)n * sizeof(float4)));
		CK(cudaMalloc((void **)&s.sx, (size_t)n * sizeof(float4)));
		CK(cudaMalloc((void **)&s.sn, (size_t)n * sizeof(float4)));
		CK(cudaMalloc((void **)&s.perm, (size_t)n * sizeof(uint32_t)));
		s.cap = (size_t)n;
	}
	const m3dreg_point *d_src = src;
	if (!src_on_device) {
		int e = c->aos_a.ensure((size_t)n);
		if (e) return e;
		CK(cudaMemcpyAsync(c->aos_a.p, src, (size_t)n * sizeof(m3dreg_point), cudaMemcpyHostToDevice, c->stream));
		d_src = c->aos_a.p;
	}
	LAUNCH(c, k_unpack_points, (n + 255) / 256, 256, d_src, n, s.xyzl, s.nrm);
	s.n = n;
	c->slam_costs_clear();
	c->active = false;
	c->last_valid = false;
	return presort_scan(c, s, d_src);
}

int m3dreg_scan_size(const m3dreg_ctx *c, int slot)
{
	if (!c || slot < 0 || (size_t)slot >= c->scans.size()) return M3DREG_E_BAD_SLOT;
	return c->scans[(size_t)slot].n;
}

int m3dreg_scan_clear(m3dreg_ctx *c)
{
	if (!c) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	CK(cudaStreamSynchronize(c->stream));
	for (auto &s : c->scans) s.release();
	c->scans.clear();
	c->slam_costs_clear();      /* measured pair costs belong to the scans that are gone */
	return 0;
}

/* ---- fused loops ------------------------------------------------------------------------------------------- */

/* bounding box of an AoS cloud on the device (one read-back; host-buffer entry points only) */
static int local_box_aos(m3dreg_ctx *c, const m3dreg_point *d_aos, int n, LocalBox *box)
{
	LAUNCH(c, k_reset_bounds, 1, 32, c->bounds);
	LAUNCH(c, k_bounds_aos, grid_for(c, n, 256), 256, d_aos, n, c->bounds);
	CK(cudaMemcpyAsync(c->h->bounds, c->bounds, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	double d2 = 0.0;
	for (int k = 0; k < 3; k++) {
		box->mn[k] = o2f_host(c->h->bounds[k]); box->mx[k] = o2f_host(c->h->bounds[3 + k]);
		const double d = (double)box->mx[k] - (double)box->mn[k];
		d2 += d * d;
	}
	box->diag = (float)sqrt(d2);
	return 0;
}

/* box = the first cloud's LOCAL bounding box (Scan::box()) */
static int icp_begin_internal(m3dreg_ctx *c, const float4 *lx, const float4 *ln, int n1, int n2, const float *pose_first,
		const m3dreg_reg_params *prm, const LocalBox &box)
{
	int e;
	if ((e = ensure_candidates(c, (size_t)n1, prm->max_inner, prm->max_outer))) return e;
	static_assert(sizeof(PoseState) % sizeof(unsigned int) == 0, "k_pose_init clears the state word by word");
	drop_iteration_graphs(c);      /* a new loop: other clouds, sizes, parameters */
	Pose16 p16;
	memcpy(p16.m, pose_first, 16 * sizeof(float));
	/* state, label counters, flags and ticket cleared, pose in (as a kernel argument: no copy-engine traffic), first Euler round trip */
	LAUNCH(c, k_pose_init, 1, 32, c->ps, p16, c->label_counts, c->flags, c->ticket);
	/* Nothing is read back to plan the loop (round 1 transformed the cloud once and synchronised to size the table):
	 * the bucket table holds the grid of this scan under ANY pose (its extent never exceeds the local box's diagonal), and
	 * the number of sort passes follows from the local box rotated by the initial pose plus a margin of four cells per
	 * axis for the pose changes of the loop. */
	if ((e = ensure_buckets(c, (size_t)bucket_capacity_for(box.diag, prm), prm->mode == M3DREG_MODE_NDT))) return e;
	const int sort_bits = planned_sort_bits(box.mn, box.mx, pose_first, prm, 4);
	c->act_local_mag = 1.0f;
	for (int k = 0; k < 3; k++) c->act_local_mag = fmaxf(c->act_local_mag, fmaxf(fabsf(box.mn[k]), fabsf(box.mx[k])));
	LAUNCH(c, k_reset_bounds, 1, 32, c->bounds);
	c->act_lx = lx; c->act_ln = ln; c->act_n1 = n1; c->act_n2 = n2; c->act_sort_bits = sort_bits; c->act_prm = *prm;
	c->active = true;
	return 0;
}

/* `iterations` fused iterations of the active loop: graph replay (see m3dreg_ctx::it_graph) or plain launches */
static int run_iterations(m3dreg_ctx *c, int iterations)
{
	int it = 0;
	/* capturing and instantiating the graph costs about as much host time as three iterations of plain launches, and a
	 * deferred query upload (host-buffer call, one iteration) must not be baked into a graph */
	if (c->use_graph && !c->profiling && !c->it_graph_failed && iterations >= 4 && !c->q_deferred_ev) {
		m3dreg_ctx::IterGraph *g = nullptr;
		for (auto &cand : c->it_graph) if (cand.exec && cand.neq_out == c->neq_out_ext) g = &cand;
		if (!g) {      /* capture one iteration (nothing runs while capturing) */
			m3dreg_ctx::IterGraph *slot = &c->it_graph[0];
			for (auto &cand : c->it_graph) if (!cand.exec || cand.age < slot->age) { slot = &cand; if (!cand.exec) break; }
			if (slot->exec) { cudaGraphExecDestroy(slot->exec); *slot = m3dreg_ctx::IterGraph(); }
			const int64_t launches_before = c->launches;
			cudaGraph_t graph = nullptr;
			if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
				icp_iteration_device(c, c->act_lx, c->act_ln, c->act_n1, c->act_n2, &c->act_prm, c->act_sort_bits);
				const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
				cudaGraphExec_t exec = nullptr;
				if (ce == cudaSuccess && graph && c->launch_err == cudaSuccess && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
					slot->exec = exec; slot->neq_out = c->neq_out_ext; slot->launches = (int)(c->launches - launches_before);
					g = slot;
				}
				if (graph) cudaGraphDestroy(graph);
			}
			c->launches = launches_before;
			if (!g) { cudaGetLastError(); c->launch_err = cudaSuccess; c->it_graph_failed = true; }      /* this loop runs on plain launches */
		}
		if (g) {
			g->age = ++c->it_graph_clock;
			for (; it < iterations; it++) {
				const cudaError_t ce = cudaGraphLaunch(g->exec, c->stream);
				if (ce != cudaSuccess) return (int)ce;
				c->launches += g->launches;
			}
		}
	}
	for (; it < iterations; it++)
		icp_iteration_device(c, c->act_lx, c->act_ln, c->act_n1, c->act_n2, &c->act_prm, c->act_sort_bits);
	return (int)cudaGetLastError();
}

static int icp_end_internal(m3dreg_ctx *c, float *pose_first_out, m3dreg_icp_stats *stats, float ms)
{
	CK(cudaMemcpyAsync(&c->h->ps, c->ps, sizeof(PoseState), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaMemcpyAsync(&c->h->gp, c->gp, sizeof(m3dreg_grid_params), cudaMemcpyDeviceToHost, c->stream));
	int f = check_flags(c);
	fill_stats(c, stats, ms);
	if (f) return f;
	if (pose_first_out) memcpy(pose_first_out, c->h->ps.m, 16 * sizeof(float));
	return (int)cudaGetLastError();
}

/* env M3DREG_HOST_TRACE=1: host time stamps of the host-buffer iteration's steps on stderr (tools/e2e_probe.py) */
static bool host_trace_on()
{
	static const bool on = getenv("M3DREG_HOST_TRACE") && atoi(getenv("M3DREG_HOST_TRACE")) != 0;
	return on;
}
static double host_now_us()
{
	return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#define HOST_TRACE(label) do { if (host_trace_on()) fprintf(stderr, "[m3dreg host trace] %-34s %12.1f us\n", label, host_now_us()); } while (0)

static int icp_loop(m3dreg_ctx *c, const float4 *lx, const float4 *ln, int n1, int n2, float *pose_first,
		const m3dreg_reg_params *prm, int iterations, m3dreg_icp_stats *stats, const LocalBox &box)
{
	int e = icp_begin_internal(c, lx, ln, n1, n2, pose_first, prm, box);
	if (e) return e;
	CK(cudaEventRecord(c->ev0, c->stream));
	HOST_TRACE("loop: begin enqueued");
	if ((e = run_iterations(c, iterations))) return e;
	CK(cudaEventRecord(c->ev1, c->stream));
	HOST_TRACE("loop: iterations enqueued");
	CK(cudaEventSynchronize(c->ev1));
	HOST_TRACE("loop: iterations done");
	float ms = 0.0f;
	cudaEventElapsedTime(&ms, c->ev0, c->ev1);
	e = icp_end_internal(c, pose_first, stats, ms);
	HOST_TRACE("loop: state read back");
	return e;
}

static int stage_queries(m3dreg_ctx *c, int second_slot, const float *pose_second)
{
	const Scan &B = c->scans[(size_t)second_slot];
	/* queries: second scan transformed once by the Euler round trip of its pose (gpu6DSLAM.cpp:295-307) */
	host_roundtrip_pose(pose_second, c->h->mats, nullptr);
	CK(cudaMemcpyAsync(c->mats, c->h->mats, 16 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
	LAUNCH(c, k_transform_soa<false>, grid_for(c, B.n, 256), 256, B.sx, B.sn, B.n, c->mats, c->q_xyzl.p, c->q_nrm.p, (uint32_t *)nullptr);
	c->act_perm = B.perm;
	return 0;
}

int m3dreg_icp_begin(m3dreg_ctx *c, int first_slot, int second_slot, const float *pose_first, const float *pose_second,
		const m3dreg_reg_params *prm)
{
	if (!c || !pose_first || !pose_second || !valid_params(prm)) return M3DREG_E_INVALID_ARG;
	if (first_slot < 0 || second_slot < 0 || (size_t)first_slot >= c->scans.size() || (size_t)second_slot >= c->scans.size())
		return M3DREG_E_BAD_SLOT;
	const Scan &A = c->scans[(size_t)first_slot], &B = c->scans[(size_t)second_slot];
	if (A.n <= 0 || B.n <= 0) return M3DREG_E_BAD_SLOT;
	CK(cudaSetDevice(c->dev));
	int e;
	if ((e = ensure_first(c, (size_t)A.n))) return e;
	if ((e = ensure_second(c, (size_t)B.n))) return e;
	if ((e = stage_queries(c, second_slot, pose_second))) return e;
	return icp_begin_internal(c, A.xyzl, A.nrm, A.n, B.n, pose_first, prm, A.box());
}

int m3dreg_icp_step(m3dreg_ctx *c, int iterations)
{
	if (!c || iterations < 0) return M3DREG_E_INVALID_ARG;
	if (!c->active) return M3DREG_E_BAD_SLOT;
	CK(cudaSetDevice(c->dev));
	return run_iterations(c, iterations);
}

int m3dreg_icp_end(m3dreg_ctx *c, float *pose_first_out, m3dreg_icp_stats *stats)
{
	if (!c) return M3DREG_E_INVALID_ARG;
	if (!c->active) return M3DREG_E_BAD_SLOT;
	CK(cudaSetDevice(c->dev));
	c->active = false;
	return icp_end_internal(c, pose_first_out, stats, 0.0f);
}

int m3dreg_icp_copy_neq(m3dreg_ctx *c, double *d_dst)
{
	if (!c || !d_dst) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	CK(cudaMemcpyAsync(d_dst, c->ps->neq, kNeqCount * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
	return 0;
}

int m3dreg_icp_set_neq_out(m3dreg_ctx *c, double *d_dst)
{
	if (!c) return M3DREG_E_INVALID_ARG;
	c->neq_out_ext = d_dst;
	return 0;
}

int m3dreg_set_profiling(m3dreg_ctx *c, int enabled)
{
	if (!c) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	if (enabled && !c->pev[0])
		for (int k = 0; k <= M3DREG_STAGE_COUNT; k++) CK(cudaEventCreate(&c->pev[k]));
	c->profiling = enabled != 0;
	for (int k = 0; k < M3DREG_STAGE_COUNT; k++) c->stage_ms[k] = 0.0f;
	c->stage_iters = 0;
	return 0;
}

int m3dreg_get_stage_ms(m3dreg_ctx *c, float *ms_out, int *iterations_out)
{
	if (!c || !ms_out) return M3DREG_E_INVALID_ARG;
	for (int k = 0; k < M3DREG_STAGE_COUNT; k++) { ms_out[k] = c->stage_ms[k]; c->stage_ms[k] = 0.0f; }
	if (iterations_out) *iterations_out = c->stage_iters;
	c->stage_iters = 0;
	return 0;
}

int m3dreg_icp_pair(m3dreg_ctx *c, int first_slot, int second_slot, float *pose_first, const float *pose_second,
		const m3dreg_reg_params *prm, int iterations, m3dreg_icp_stats *stats)
{
	if (!c || !pose_first || !pose_second || !valid_params(prm) || iterations < 0) return M3DREG_E_INVALID_ARG;
	if (first_slot < 0 || second_slot < 0 || (size_t)first_slot >= c->scans.size() || (size_t)second_slot >= c->scans.size())
		return M3DREG_E_BAD_SLOT;
	const Scan &A = c->scans[(size_t)first_slot], &B = c->scans[(size_t)second_slot];
	if (A.n <= 0 || B.n <= 0) return M3DREG_E_BAD_SLOT;
	CK(cudaSetDevice(c->dev));
	int e;
	if ((e = ensure_first(c, (size_t)A.n))) return e;
	if ((e = ensure_second(c, (size_t)B.n))) return e;
	if ((e = stage_queries(c, second_slot, pose_second))) return e;
	e = icp_loop(c, A.xyzl, A.nrm, A.n, B.n, pose_first, prm, iterations, stats, A.box());
	c->active = false;
	return e;
}

int m3dreg_icp_iteration_host(m3dreg_ctx *c, const m3dreg_point *first_local, int n1, const m3dreg_point *second_global, int n2,
		float *pose_first, const m3dreg_reg_params *prm, int *nn_out, m3dreg_icp_stats *stats)
{
	if (!c || !first_local || !second_global || n1 <= 0 || n2 <= 0 || !pose_first || !valid_params(prm)) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	int e;
	if ((e = c->aos_a.ensure((size_t)n1))) return e;
	if ((e = c->aos_b.ensure((size_t)n2))) return e;
	if ((e = c->l_xyzl.ensure((size_t)n1))) return e;
	if ((e = c->l_nrm.ensure((size_t)n1))) return e;
	if ((e = ensure_first(c, (size_t)n1))) return e;
	if ((e = ensure_second(c, (size_t)n2))) return e;
	if (!c->copy_stream) {
		CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
		for (auto &ev : c->ev_copy) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
	}
	/* uploads on the copy stream, ordered after whatever the registration stream still does with the staging buffers
	 * (env M3DREG_HOST_OVERLAP=0: on the registration stream itself, i.e. nothing runs under the transfers — A/B runs) */
	HOST_TRACE("host iteration: entry");
	static const bool overlap = !(getenv("M3DREG_HOST_OVERLAP") && atoi(getenv("M3DREG_HOST_OVERLAP")) == 0);
	cudaStream_t cs = overlap ? c->copy_stream : c->stream;
	CK(cudaEventRecord(c->ev_copy[0], c->stream));
	CK(cudaStreamWaitEvent(cs, c->ev_copy[0], 0));
	CK(cudaMemcpyAsync(c->aos_a.p, first_local, (size_t)n1 * sizeof(m3dreg_point), cudaMemcpyHostToDevice, cs));
	CK(cudaEventRecord(c->ev_copy[1], cs));
	CK(cudaMemcpyAsync(c->aos_b.p, second_global, (size_t)n2 * sizeof(m3dreg_point), cudaMemcpyHostToDevice, cs));
	CK(cudaEventRecord(c->ev_copy[2], cs));
	CK(cudaStreamWaitEvent(c->stream, c->ev_copy[1], 0));
	LAUNCH(c, k_unpack_points, (n1 + 255) / 256, 256, c->aos_a.p, n1, c->l_xyzl.p, c->l_nrm.p);
	c->act_perm = nullptr;
	LocalBox box;
	HOST_TRACE("host iteration: copies enqueued");
	e = local_box_aos(c, c->aos_a.p, n1, &box);
	HOST_TRACE("host iteration: local box known");
	if (!e) {
		/* the queries are unpacked by the iteration itself, after the grid of the first cloud is built (icp_iteration_device) */
		c->q_deferred_ev = c->ev_copy[2]; c->q_deferred_aos = c->aos_b.p; c->q_deferred_n = n2;
		e = icp_loop(c, c->l_xyzl.p, c->l_nrm.p, n1, n2, pose_first, prm, 1, stats, box);
	}
	c->active = false;
	if (c->q_deferred_ev || e) {      /* not consumed (error path): the caller's buffers must not be read after the return */
		c->q_deferred_ev = nullptr;
		cudaStreamSynchronize(c->copy_stream);
	}
	if (e) return e;
	if (nn_out) {
		if (prm->mode == M3DREG_MODE_NDT) CK(cudaMemsetAsync(c->nn.p, 0xFF, (size_t)n2 * sizeof(int), c->stream));
		else materialize_nn(c);
		CK(cudaMemcpyAsync(nn_out, c->nn.p, (size_t)n2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	HOST_TRACE("host iteration: nn read back");
	return (int)cudaGetLastError();
}

int m3dreg_export_last_grid(m3dreg_ctx *c, m3dreg_grid_params *params_out, m3dreg_hash_element *table_out, int table_cap,
		m3dreg_bucket *buckets_out, int64_t buckets_cap)
{
	if (!c) return M3DREG_E_INVALID_ARG;
	if (!c->last_valid) return M3DREG_E_BAD_SLOT;
	CK(cudaSetDevice(c->dev));
	CK(cudaMemcpyAsync(&c->h->gp, c->gp, sizeof(m3dreg_grid_params), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	if (params_out) *params_out = c->h->gp;
	int n = c->last_n_first;
	if (table_out) {
		if (table_cap < n) return M3DREG_E_SIZE_MISMATCH;
		std::vector<uint32_t> k((size_t)n), v((size_t)n);
		CK(cudaMemcpyAsync(k.data(), c->keys[c->last_sorted].p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
		CK(cudaMemcpyAsync(v.data(), c->vals[c->last_sorted].p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		for (int i = 0; i < n; i++) { table_out[i].index_of_point = (int)v[(size_t)i]; table_out[i].index_of_bucket = (int)k[(size_t)i]; }
	}
	if (buckets_out) {
		if (buckets_cap < c->h->gp.number_of_buckets) return M3DREG_E_SIZE_MISMATCH;
		CK(cudaMemcpyAsync(buckets_out, c->buckets.p, (size_t)c->h->gp.number_of_buckets * sizeof(m3dreg_bucket), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
	}
	return 0;
}

int m3dreg_export_last_nn(m3dreg_ctx *c, int *nn_out, int nn_cap)
{
	if (!c || !nn_out) return M3DREG_E_INVALID_ARG;
	if (!c->last_valid || !c->last_nn_valid) return M3DREG_E_BAD_SLOT;   /* NDT has no correspondences */
	if (nn_cap < c->last_n_second) return M3DREG_E_SIZE_MISMATCH;
	CK(cudaSetDevice(c->dev));
	materialize_nn(c);
	CK(cudaMemcpyAsync(nn_out, c->nn.p, (size_t)c->last_n_second * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return 0;
}

/* ---- multi-scan sweep (registerAll) ------------------------------------------------------------------------ */

int m3dreg_sweep_zero(m3dreg_ctx *c, double *d_neq, int n_scans)
{
	if (!c || !d_neq || n_scans <= 0) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	LAUNCH(c, k_zero_f64, grid_for(c, (long long)n_scans * kNeqCount, 256), 256, d_neq, n_scans * kNeqCount);
	return (int)cudaGetLastError();
}

/* pinned staging of a sweep: grows only (growth synchronises; steady-state sweeps never do) */
static int ensure_sweep_staging(m3dreg_ctx *c, size_t bytes)
{
	if (c->h_sweep_bytes >= bytes) return 0;
	CK(cudaStreamSynchronize(c->stream));
	if (c->h_sweep) cudaFreeHost(c->h_sweep);
	c->h_sweep = nullptr; c->h_sweep_bytes = 0;
	CK(cudaMallocHost(&c->h_sweep, bytes + bytes / 2));
	c->h_sweep_bytes = bytes + bytes / 2;
	return 0;
}

__global__ void k_group_stamp(double *__restrict__ group_ms, int finished_group, unsigned long long *__restrict__ last)
{
	pdl_enter();
	if (threadIdx.x != 0) return;
	unsigned long long now;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
	if (finished_group >= 0) group_ms[finished_group] += (double)(now - *last) * 1.0e-6;
	*last = now;
}

static int sweep_accumulate_impl(m3dreg_ctx *c, int n_pairs, const int *pair_i, const int *pair_j, const float *poses, int n_scans,
		const m3dreg_reg_params *prm, double *d_neq, double *d_group_ms);

int m3dreg_sweep_accumulate(m3dreg_ctx *c, int n_pairs, const int *pair_i, const int *pair_j, const float *poses, int n_scans,
		const m3dreg_reg_params *prm, double *d_neq)
{
	return sweep_accumulate_impl(c, n_pairs, pair_i, pair_j, poses, n_scans, prm, d_neq, nullptr);
}

/* d_group_ms (n_scans doubles on the device, may be 0): += the device time of every scan's group of pairs */
static int sweep_accumulate_impl(m3dreg_ctx *c, int n_pairs, const int *pair_i, const int *pair_j, const float *poses, int n_scans,
		const m3dreg_reg_params *prm, double *d_neq, double *d_group_ms)
{
	if (!c || n_pairs < 0 || (n_pairs && (!pair_i || !pair_j)) || !poses || n_scans <= 0 || !valid_params(prm) || !d_neq)
		return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	int e;
	for (int p = 0; p < n_pairs; p++) {
		int i = pair_i[p], j = pair_j[p];
		if (i < 0 || j < 0 || i >= n_scans || j >= n_scans || (size_t)i >= c->scans.size() || (size_t)j >= c->scans.size() ||
				c->scans[(size_t)i].n <= 0 || c->scans[(size_t)j].n <= 0 || i == j)
			return M3DREG_E_BAD_SLOT;
	}
	const bool ndt = prm->mode == M3DREG_MODE_NDT;
	/* ---- plan (host only): pairs grouped by i so the grid of scan i is built once (the reference rebuilds it per j,
	 * gpu6DSLAM.cpp:478); ICP: the neighbours j of one i are searched in batches — one transform, ONE search and ONE moment
	 * reduction over the concatenated queries of up to kMaxSegs neighbours instead of three small launches per pair.
	 * Everything the device needs (round-tripped poses, every batch's segment table) is staged in ONE pinned block and
	 * uploaded up front, every buffer is sized up front: the loop below only launches kernels — no read-back, no
	 * synchronisation, no allocation (round 1 synchronised per scan and per batch). */
	std::vector<int> order((size_t)n_pairs);
	for (int p = 0; p < n_pairs; p++) order[(size_t)p] = p;
	/* Inside a group the FARTHEST neighbours come first.  Queries of a distant neighbour mostly have no partner inside the
	 * search radius and pay every round of the search (chunks of 90-280 us where the mean is 36 us), and the search hands
	 * its chunks out in query order: with the neighbours in index order the distant ones sat at the end of the launch and
	 * 38 % of it was tail (mean warp out of work at 262 of 420 us, tools/nn_tail_sweep.py).  Longest first needs no
	 * measurement here — the distance between the poses is the proxy.  Env M3DREG_SWEEP_INDEX_ORDER=1: index order (A/B). */
	std::vector<float> far((size_t)n_pairs, 0.0f);
	if (!getenv("M3DREG_SWEEP_INDEX_ORDER"))
		for (int p = 0; p < n_pairs; p++) {
			const float *a = poses + 16 * (size_t)pair_i[p], *b = poses + 16 * (size_t)pair_j[p];
			const float dx = a[3] - b[3], dy = a[7] - b[7], dz = a[11] - b[11];
			far[(size_t)p] = dx * dx + dy * dy + dz * dz;
		}
	std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
		if (pair_i[a] != pair_i[b]) return pair_i[a] < pair_i[b];
		return far[(size_t)a] > far[(size_t)b];
	});
	const size_t kBatchQueries = (size_t)8 << 20;       /* queries per batch (incl. padding): 8 Mi x 36 B of query state */
	struct Batch { int i, seg0, nseg, j_single; size_t total; };
	std::vector<Batch> batches;
	std::vector<SweepSeg> all_segs;
	size_t max_first = 0, max_second = 0;
	long long max_cap = 1;
	for (int q = 0; q < n_pairs;) {
		const int i = pair_i[order[(size_t)q]];
		const Scan &A = c->scans[(size_t)i];
		if ((size_t)A.n > max_first) max_first = (size_t)A.n;
		const long long cap = bucket_capacity_for(A.diag, prm);
		if (cap > max_cap) max_cap = cap;
		Batch bt;
		bt.i = i; bt.seg0 = (int)all_segs.size(); bt.nseg = 0; bt.total = 0; bt.j_single = pair_j[order[(size_t)q]];
		if (ndt) {      /* NDT: one pair per step (the per-bucket query sums are per pair) */
			bt.total = (size_t)c->scans[(size_t)bt.j_single].n;
			q++;
		} else {
			while (q < n_pairs && pair_i[order[(size_t)q]] == i && bt.nseg < kMaxSegs) {
				const int j = pair_j[order[(size_t)q]];
				const Scan &B = c->scans[(size_t)j];
				const size_t padded = ((size_t)B.n + kSegChunk - 1) / kSegChunk * kSegChunk;
				if (bt.nseg > 0 && bt.total + padded > kBatchQueries) break;
				SweepSeg sg;
				sg.sx = B.sx; sg.sn = B.sn; sg.n = B.n; sg.off = (int)bt.total; sg.pose = j; sg.pad = 0;
				all_segs.push_back(sg);
				bt.total += padded;
				bt.nseg++;
				q++;
			}
			if (bt.total > 0x7fffffffu) return M3DREG_E_INVALID_ARG;
		}
		if (bt.total > max_second) max_second = bt.total;
		batches.push_back(bt);
	}
	if (c->sweep_reserve_second > max_second) max_second = c->sweep_reserve_second;
	if (c->sweep_reserve_first > max_first) max_first = c->sweep_reserve_first;
	if (c->sweep_reserve_cap > max_cap) max_cap = c->sweep_reserve_cap;
	const size_t seg_slots = all_segs.size() > c->sweep_reserve_segs ? all_segs.size() : c->sweep_reserve_segs;
	if ((e = c->d_poses1.ensure((size_t)n_scans * 16))) return e;
	if ((e = c->d_pose6.ensure((size_t)n_scans * 6))) return e;
	if (max_first) {
		if ((e = ensure_first(c, max_first))) return e;
		if ((e = ensure_candidates(c, max_first, prm->max_inner, prm->max_outer))) return e;
		if ((e = ensure_second(c, max_second))) return e;
		if ((e = ensure_buckets(c, (size_t)max_cap, ndt))) return e;
	}
	if (!ndt) {
		if ((e = c->d_segs.ensure(seg_slots + 1))) return e;
		if ((e = c->d_seg_counts.ensure(4 * kMaxSegs))) return e;
		if ((e = c->d_seg_of_chunk.ensure(max_second / kSegChunk + 1))) return e;
	}
	const size_t bytes_p1 = (size_t)n_scans * 16 * sizeof(float), bytes_p6 = (size_t)n_scans * 6 * sizeof(double),
			bytes_segs = all_segs.size() * sizeof(SweepSeg);
	if ((e = ensure_sweep_staging(c, bytes_p6 + bytes_p1 + seg_slots * sizeof(SweepSeg) + 64))) return e;
	/* the previous sweep's uploads have completed: every sweep ends with check_flags (one synchronisation per sweep) */
	double *h_p6 = static_cast<double *>(c->h_sweep);
	float *h_p1 = reinterpret_cast<float *>(static_cast<char *>(c->h_sweep) + bytes_p6);
	SweepSeg *h_segs = reinterpret_cast<SweepSeg *>(static_cast<char *>(c->h_sweep) + bytes_p6 + bytes_p1);
	/* Euler round trip of every (old) pose: gpu6DSLAM.cpp:440-441, 461-462 */
	for (int s = 0; s < n_scans; s++) host_roundtrip_pose(poses + 16 * (size_t)s, h_p1 + 16 * (size_t)s, h_p6 + 6 * (size_t)s);
	if (bytes_segs) memcpy(h_segs, all_segs.data(), bytes_segs);
	CK(cudaMemcpyAsync(c->d_poses1.p, h_p1, bytes_p1, cudaMemcpyHostToDevice, c->stream));
	CK(cudaMemcpyAsync(c->d_pose6.p, h_p6, bytes_p6, cudaMemcpyHostToDevice, c->stream));
	if (bytes_segs) CK(cudaMemcpyAsync(c->d_segs.p, h_segs, bytes_segs, cudaMemcpyHostToDevice, c->stream));
	CK(cudaMemsetAsync(c->label_counts, 0, 4 * sizeof(unsigned long long), c->stream));
	CK(cudaMemsetAsync(c->flags, 0, FLAG_COUNT * sizeof(int), c->stream));
	CK(cudaMemsetAsync(c->ticket, 0, sizeof(unsigned int), c->stream));
	if (!ndt) CK(cudaMemsetAsync(c->d_seg_counts.p, 0, 4 * kMaxSegs * sizeof(unsigned long long), c->stream));

	int cur_i = -1;
	for (const Batch &bt : batches) {
		const int i = bt.i;
		const Scan &A = c->scans[(size_t)i];
		const float *pose_i = c->d_poses1.p + 16 * (size_t)i;
		if (i != cur_i) {
			if (d_group_ms) LAUNCH(c, k_group_stamp, 1, 32, d_group_ms, cur_i, c->stamp_last);
			LAUNCH(c, k_reset_bounds, 1, 32, c->bounds);
			if (!c->grid_mega) {
				LAUNCH(c, k_transform_soa<true>, box_pass_blocks(c, A.n), 512, A.xyzl, A.nrm, A.n, pose_i, ndt ? c->g_xyzl.p : (float4 *)nullptr, (float4 *)nullptr, c->bounds);
				build_grid_legacy(c, A.xyzl, A.nrm, A.n, pose_i, prm, planned_sort_bits(A.bb_min, A.bb_max, h_p1 + 16 * (size_t)i, prm, 1));
			} else {
				build_grid_mega(c, A.xyzl, A.nrm, A.n, pose_i, prm, ndt);
			}
			if (ndt) {
				float mag = 1.0f;
				for (int k = 0; k < 3; k++) mag = fmaxf(mag, fmaxf(fabsf(A.bb_min[k]), fabsf(A.bb_max[k])));
				ndt_bucket_stats(c, A.xyzl, A.n, prm, mag);
			}
			cur_i = i;
		}
		FinalizeArgs fin = {};
		fin.ps = nullptr; fin.neq_out = d_neq + (size_t)i * kNeqCount; fin.accumulate = 1; fin.solve = 0; fin.dof = prm->dof;
		fin.obs_threshold = prm->obs_threshold; fin.pose6_in = c->d_pose6.p + 6 * (size_t)i; fin.bounds_reset = nullptr;
		if (ndt) {
			const Scan &B = c->scans[(size_t)bt.j_single];
			LAUNCH(c, k_transform_soa<false>, grid_for(c, B.n, 256), 256, B.sx, B.sn, B.n, c->d_poses1.p + 16 * (size_t)bt.j_single,
					c->q_xyzl.p, c->q_nrm.p, (uint32_t *)nullptr);
			fin.label_counts_reset = nullptr;
			ndt_queries_and_reduce(c, B.n, fin, true);
			c->last_n_first = A.n; c->last_n_second = B.n; c->last_valid = true; c->last_nn_valid = false;
			continue;
		}
		const int n_chunks = (int)(bt.total / kSegChunk);
		const SweepSeg *segs = c->d_segs.p + bt.seg0;
		LAUNCH(c, k_transform_segments, n_chunks, kSegChunk, segs, bt.nseg, c->d_seg_of_chunk.p, c->d_poses1.p, c->q_xyzl.p, c->q_nrm.p);
		const float res3[3] = {prm->bucket_size, prm->bucket_size, prm->bucket_size};
		{
			const NNTuning pair_tune = c->nn_tune;
			c->nn_tune = c->nn_tune_sweep;
			launch_nn(c, nullptr, (int)bt.total, c->vals[c->last_sorted].p, A.n, c->buckets.p, res3, prm->search_radius, prm->max_inner, prm->max_outer, 1,
					nullptr, c->obs_rec.p, A.xyzl, c->d_seg_counts.p, c->d_seg_of_chunk.p);
			c->nn_tune = pair_tune;
		}
		ObsFromRec src = {};
		src.rec = c->obs_rec.p; src.q_xyzl = c->q_xyzl.p; src.m = pose_i;
		src.label_counts = c->d_seg_counts.p; src.seg_of_chunk = c->d_seg_of_chunk.p; src.n_segs = bt.nseg;
		for (int k = 0; k < 4; k++) src.weight[k] = prm->weight[k];
		fin.label_counts_reset = c->d_seg_counts.p; fin.label_count_sets = bt.nseg;
		LAUNCH(c, k_normal_equations<ObsFromRec>, grid_for(c, (long long)bt.total, kNeqThreads, 2), kNeqThreads, src, (int)bt.total, c->partials.p, c->ticket, fin);
		c->last_n_first = A.n; c->last_n_second = all_segs[(size_t)(bt.seg0 + bt.nseg - 1)].n; c->last_valid = true; c->last_nn_valid = false; c->nn_pending = false;
	}
	if (d_group_ms && cur_i >= 0) LAUNCH(c, k_group_stamp, 1, 32, d_group_ms, cur_i, c->stamp_last);
	int f = check_flags(c);
	if (f) return f;
	return launch_status(c);
}

int m3dreg_sweep_solve(m3dreg_ctx *c, const double *d_neq, int n_scans, int scan_begin, int scan_end, float *poses,
		const m3dreg_reg_params *prm, int *status_out)
{
	if (!c || !d_neq || n_scans <= 0 || scan_begin < 0 || scan_end > n_scans || scan_begin > scan_end || !poses || !valid_params(prm))
		return M3DREG_E_INVALID_ARG;
	if (scan_begin == scan_end) return 0;
	CK(cudaSetDevice(c->dev));
	int e;
	if ((e = c->d_poses1.ensure((size_t)n_scans * 16))) return e;
	DevBuf<int> &st = c->d_sweep_status;      /* persistent arena member: nothing is allocated per sweep */
	if ((e = st.ensure((size_t)n_scans))) return e;
	CK(cudaMemcpyAsync(c->d_poses1.p, poses, (size_t)n_scans * 16 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
	int cnt = scan_end - scan_begin;
	LAUNCH(c, k_sweep_solve, (cnt + 63) / 64, 64, d_neq, scan_begin, scan_end, c->d_poses1.p, prm->dof, prm->obs_threshold, st.p);
	CK(cudaMemcpyAsync(poses + 16 * (size_t)scan_begin, c->d_poses1.p + 16 * (size_t)scan_begin, (size_t)cnt * 16 * sizeof(float),
			cudaMemcpyDeviceToHost, c->stream));
	if (status_out)
		CK(cudaMemcpyAsync(status_out + scan_begin, st.p + scan_begin, (size_t)cnt * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	return launch_status(c);
}

} /* extern "C" */

#include "slam_host.inl"
#include "preproc_host.inl"
#include "../../include/m3dreg_node.h"
#include "formats_host.inl"
#include "node_host.inl"
