/* preproc.cuh — the pre-registration trio on the registration path's own grid build:
 *   noise filter   CCudaWrapper::removeNoiseNaive (src/cudaWrapper.cpp:118-179) / cudaRemoveNoiseNaive (src/lesson_16.cu:749-787)
 *   downsampling   CCudaWrapper::downsampling     (src/cudaWrapper.cpp:181-262) / cudaDownSample       (src/lesson_16.cu:789-815)
 *   classification CCudaWrapper::classify         (src/cudaWrapper.cpp:264-342) / cudaSemanticLabelingPlaneEdges +
 *                  cudaSemanticLabelingFloorCeiling (src/lesson_16.cu:817-1239), 3x3 decomposition src/cuda_SVD.cu:187-380
 * and the yaw sweep of row N2 (findBestYaw, src/cudaWrapper.cpp:662-836, src/lesson_16.cu:1241-1384).
 *
 * The reference copies the cloud to the device, builds the grid, marks, copies a bool per point back and compacts on the
 * host; here the survivors are compacted on the device in their original order (block counts, one scan, scatter) and
 * only they come back.  Classification upstream is two thread-per-point kernels over the 27 neighbouring buckets, each an
 * uncoalesced two-level gather (hash[l] -> cloud[index]) per candidate and thread; here a warp owns 32 consecutive
 * SORTED positions — points of one bucket or of a few adjacent ones — stages the candidates of the home bucket's
 * neighbourhood in shared memory once per sweep for all its lanes, and the lanes run the reference's two accumulations
 * (float mean, then covariance about it) over the staged copies in the reference's visit order: same counts, same mean
 * bits, same covariance bits.  Only the 3x3 decomposition differs (Jacobi here, closed-form cubic upstream): normals to
 * round-off, labels except where lambda_mid / lambda_min sits on the threshold — tolerance parity, stated in DESIGN.md. */
#pragma once
#include "m3dreg_kernels.cuh"

namespace m3d {

/* ---- markers ------------------------------------------------------------------------------------------------------- */

/* kernel_setAllPointsToRemove + kernel_markPointsToRemain (lesson_16.cu:740-766): a point stays iff its bucket holds
 * more than `threshold` points (the dense table's count: the first-element quirk bucket counts 0 and loses its points,
 * as upstream). */
__global__ void k_mark_noise(const m3dreg_point *__restrict__ cloud, int n, const m3dreg_grid_params *__restrict__ gp,
		const m3dreg_bucket *__restrict__ buckets, int threshold, unsigned char *__restrict__ markers)
{
	pdl_enter();
	const float mnx = gp->bounding_box_min_X, mny = gp->bounding_box_min_Y, mnz = gp->bounding_box_min_Z;
	const float rx = gp->resolution_X, ry = gp->resolution_Y, rz = gp->resolution_Z;
	const int nby = gp->number_of_buckets_Y, nbz = gp->number_of_buckets_Z;
	const long long nb = gp->number_of_buckets;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint2 *p = reinterpret_cast<const uint2 *>(cloud + i);
		const uint2 a = __ldg(p), b = __ldg(p + 1);
		const int ix = cell_of(__uint_as_float(a.x), mnx, rx), iy = cell_of(__uint_as_float(a.y), mny, ry), iz = cell_of(__uint_as_float(b.x), mnz, rz);
		const int key = ix * nby * nbz + iy * nbz + iz;
		unsigned char keep = 0;
		if (key >= 0 && (long long)key < nb) keep = __ldg(reinterpret_cast<const int *>(buckets + key) + 2) > threshold ? 1 : 0;
		markers[i] = keep;
	}
}

__global__ void k_zero_u8(unsigned char *__restrict__ p, int n)
{
	pdl_enter();
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0;
}

/* kernel_markFirstPointInBuckets (lesson_16.cu:789-801): the first sorted point of every bucket that has an index_begin */
__global__ void k_mark_first_in_bucket(const m3dreg_bucket *__restrict__ buckets, const m3dreg_grid_params *__restrict__ gp,
		const uint32_t *__restrict__ vals, unsigned char *__restrict__ markers)
{
	pdl_enter();
	const long long nb = gp->number_of_buckets;
	for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (long long)gridDim.x * blockDim.x) {
		const int begin = __ldg(reinterpret_cast<const int *>(buckets + b));
		if (begin != -1) markers[__ldg(vals + begin)] = 1;
	}
}

/* ---- order-preserving compaction of the marked points --------------------------------------------------------------- */
constexpr int kCompactTile = 1024;

__global__ void __launch_bounds__(256) k_compact_count(const unsigned char *__restrict__ markers, int n, int *__restrict__ tile_count)
{
	pdl_enter();
	__shared__ int s[8];
	const int base = blockIdx.x * kCompactTile;
	int cnt = 0;
#pragma unroll
	for (int k = 0; k < kCompactTile / 256; k++) {
		const int i = base + k * 256 + threadIdx.x;
		cnt += (i < n && markers[i]) ? 1 : 0;
	}
	cnt = __reduce_add_sync(0xffffffffu, cnt);
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
	__syncthreads();
	if (threadIdx.x == 0) {
		int t = 0;
		for (int k = 0; k < 8; k++) t += s[k];
		tile_count[blockIdx.x] = t;
	}
}

/* exclusive scan of the tile counts by one block (tiles <= a few thousand); total -> *n_out */
__global__ void __launch_bounds__(1024) k_compact_scan(int *__restrict__ tile_count, int tiles, int *__restrict__ n_out)
{
	pdl_enter();
	__shared__ int s[33];
	__shared__ int carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (int base = 0; base < tiles; base += 1024) {
		const int i = base + threadIdx.x;
		const int v = i < tiles ? tile_count[i] : 0;
		int incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		if (lane == 31) s[w] = incl;
		__syncthreads();
		if (w == 0) {
			int x = s[lane], xi = x;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const int t = __shfl_up_sync(0xffffffffu, xi, o);
				if (lane >= o) xi += t;
			}
			s[lane] = xi - x;
			if (lane == 31) s[32] = xi;
		}
		__syncthreads();
		if (i < tiles) tile_count[i] = carry + s[w] + incl - v;
		__syncthreads();
		if (threadIdx.x == 0) carry += s[32];
		__syncthreads();
	}
	if (threadIdx.x == 0) *n_out = carry;
}

/* one block per tile, threads in index order: a marked point's slot = tile offset + marked points before it in the tile */
__global__ void __launch_bounds__(kCompactTile) k_compact_scatter(const m3dreg_point *__restrict__ in, const unsigned char *__restrict__ markers, int n,
		const int *__restrict__ tile_offset, m3dreg_point *__restrict__ out)
{
	pdl_enter();
	__shared__ int s[33];
	const int i = blockIdx.x * kCompactTile + threadIdx.x;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const bool keep = i < n && markers[i];
	const unsigned m = __ballot_sync(0xffffffffu, keep);
	if (lane == 0) s[w] = __popc(m);
	__syncthreads();
	if (w == 0) {
		int x = s[lane], xi = x;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, xi, o);
			if (lane >= o) xi += t;
		}
		s[lane] = xi - x;
	}
	__syncthreads();
	if (keep) {
		const int slot = tile_offset[blockIdx.x] + s[w] + __popc(m & ((1u << lane) - 1u));
		const uint2 *src = reinterpret_cast<const uint2 *>(in + i);
		uint2 *dst = reinterpret_cast<uint2 *>(out + slot);
#pragma unroll
		for (int k = 0; k < 5; k++) dst[k] = __ldg(src + k);
	}
}

/* ---- classification -------------------------------------------------------------------------------------------------- */

/* Eigen-decomposition of a symmetric 3x3 matrix (fp64, cyclic Jacobi, 8 sweeps: converged to round-off for any input);
 * eigenvalues in descending order in l[], the unit eigenvector of the SMALLEST one in n[]. */
__device__ __host__ __forceinline__ void sym3_smallest_eigenvector(const double c[6] /* xx xy xz yy yz zz */, double l[3], double n[3])
{
	double a[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
	double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
	for (int sweep = 0; sweep < 8; sweep++) {
#pragma unroll
		for (int p = 0; p < 2; p++)
#pragma unroll
			for (int q = p + 1; q < 3; q++) {
				const double apq = a[p][q];
				if (apq == 0.0) continue;
				const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
				const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
				const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
#pragma unroll
				for (int k = 0; k < 3; k++) {
					const double akp = a[k][p], akq = a[k][q];
					a[k][p] = cs * akp - sn * akq;
					a[k][q] = sn * akp + cs * akq;
				}
#pragma unroll
				for (int k = 0; k < 3; k++) {
					const double apk = a[p][k], aqk = a[q][k];
					a[p][k] = cs * apk - sn * aqk;
					a[q][k] = sn * apk + cs * aqk;
				}
#pragma unroll
				for (int k = 0; k < 3; k++) {
					const double vkp = v[k][p], vkq = v[k][q];
					v[k][p] = cs * vkp - sn * vkq;
					v[k][q] = sn * vkp + cs * vkq;
				}
			}
	}
	/* eigenvalues in descending order, the eigenvector of the smallest (no dynamic indexing: everything stays in registers) */
	const double e0 = a[0][0], e1 = a[1][1], e2 = a[2][2];
	const int imin = (e0 <= e1 && e0 <= e2) ? 0 : (e1 <= e2 ? 1 : 2);
	const double lo = imin == 0 ? e0 : (imin == 1 ? e1 : e2);
	const double o1 = imin == 0 ? e1 : e0, o2 = imin == 2 ? e1 : e2;      /* the other two */
	l[0] = o1 > o2 ? o1 : o2; l[1] = o1 > o2 ? o2 : o1; l[2] = lo;
	n[0] = imin == 0 ? v[0][0] : (imin == 1 ? v[0][1] : v[0][2]);
	n[1] = imin == 0 ? v[1][0] : (imin == 1 ? v[1][1] : v[1][2]);
	n[2] = imin == 0 ? v[2][0] : (imin == 1 ? v[2][1] : v[2][2]);
}

struct ClassifyParams {
	float radius;                 /* normal_vectors_search_radius (= bucket size of this grid) */
	float curvature_threshold;
	float ground_z_threshold;
	int   plane_points_threshold;
	int   max_inner, max_outer;
	float vx, vy, vz;             /* viewpoint */
};

constexpr int kClsWarps = 4;
constexpr int kClsStage = 256;    /* candidates staged per batch and warp (float4) */

/* One warp = 32 consecutive sorted positions.  For every distinct home bucket among its lanes, two sweeps over the 27
 * neighbouring buckets in the reference's (i, j, k) order, their strided candidates (begin, begin + s, ...; s = n / cap,
 * INNER cap for the home bucket) staged batch by batch in shared memory, the lanes of that bucket accumulating over the
 * staged candidates in order:
 *   sweep 0 = kernel_normalvectorcomputation_step1_fast (lesson_16.cu:817-957): float coordinate sums of the neighbours
 *             within the radius, in visit order, and their count -> the float mean (zero below three neighbours);
 *   sweep 1 = kernel_normalvectorcomputation_step2_fast_with_classification (lesson_16.cu:959-1110): covariance about
 *             that mean — float differences, float products, double accumulation, exactly the reference's operations —
 *             so the matrix handed to the decomposition has the reference's bits.
 * The decomposition is a cyclic Jacobi iteration on the symmetric matrix (upstream: a closed-form cubic on A^T A,
 * src/cuda_SVD.cu:187-380, third party): the normal is the eigenvector of the smallest eigenvalue (= upstream's cross
 * product of the two dominant singular vectors), the plane test lambda_mid / lambda_min (= upstream's SS[4] / SS[8]). */
__global__ void __launch_bounds__(kClsWarps * 32) k_classify(m3dreg_point *__restrict__ cloud, int n, const uint32_t *__restrict__ keys,
		const uint32_t *__restrict__ vals, const m3dreg_bucket *__restrict__ buckets, const m3dreg_grid_params *__restrict__ gp, ClassifyParams prm,
		float *__restrict__ mean_out /* 3 floats per sorted position (parity export), may be 0 */)
{
	pdl_enter();
	__shared__ float4 s_cand[kClsWarps][kClsStage];
	const unsigned full = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	float4 *stage = s_cand[threadIdx.x >> 5];
	const int nbx = gp->number_of_buckets_X, nby = gp->number_of_buckets_Y, nbz = gp->number_of_buckets_Z;
	const long long nb = gp->number_of_buckets;
	const int nwarps = gridDim.x * kClsWarps;
	const int nchunks = (n + 31) >> 5;
	for (int chunk = blockIdx.x * kClsWarps + (threadIdx.x >> 5); chunk < nchunks; chunk += nwarps) {
		const int l = (chunk << 5) + lane;
		const bool valid = l < n;
		int key = -1;
		uint32_t idx = 0;
		float x = 0.0f, y = 0.0f, z = 0.0f;
		if (valid) {
			key = (int)__ldg(keys + l);
			idx = __ldg(vals + l);
			const uint2 *p = reinterpret_cast<const uint2 *>(cloud + idx);
			const uint2 a = p[0], b = p[1];
			x = __uint_as_float(a.x); y = __uint_as_float(a.y); z = __uint_as_float(b.x);
		}
		const bool active0 = valid && key >= 0 && (long long)key < nb && idx < (uint32_t)n;
		float mx = 0.0f, my = 0.0f, mz = 0.0f;      /* sweep 0: float sums in visit order; afterwards the float mean */
		double cv[6] = {0, 0, 0, 0, 0, 0};          /* sweep 1: xx xy xz yy yz zz */
		int cnt = 0, cnt2 = 0;
		bool have_mean = false;
		unsigned todo = __ballot_sync(full, active0);
		while (todo) {
			const int home = __shfl_sync(full, key, __ffs(todo) - 1);
			const bool mine = active0 && key == home;
			todo &= ~__ballot_sync(full, mine);
			const int ix = home / (nby * nbz), iy = (home % (nby * nbz)) / nbz, iz = (home % (nby * nbz)) % nbz;
			const int sx = ix == 0 ? 0 : -1, sy = iy == 0 ? 0 : -1, sz = iz == 0 ? 0 : -1;
			const int stx = ix == nbx - 1 ? 1 : 2, sty = iy == nby - 1 ? 1 : 2, stz = iz == nbz - 1 ? 1 : 2;
			for (int sweep = 0; sweep < 2; sweep++) {
				if (sweep == 1) {
					if (mine) {
						if (cnt >= 3) { mx = f_div(mx, (float)cnt); my = f_div(my, (float)cnt); mz = f_div(mz, (float)cnt); }
						else { mx = 0.0f; my = 0.0f; mz = 0.0f; }
						have_mean = mx != 0.0f && my != 0.0f && mz != 0.0f;      /* lesson_16.cu:979 */
					}
					if (!__any_sync(full, mine && have_mean)) break;
				}
				const bool acc = mine && (sweep == 0 || have_mean);
				for (int i = sx; i < stx; i++)
				for (int j = sy; j < sty; j++)
				for (int k = sz; k < stz; k++) {
					const int nbk = home + i * nby * nbz + j * nbz + k;
					if (nbk < 0 || (long long)nbk >= nb) continue;
					const int *rec = reinterpret_cast<const int *>(buckets + nbk);
					const int npts = __ldg(rec + 2);
					if (npts <= 0) continue;
					const int cap = nbk == home ? prm.max_inner : prm.max_outer;
					if (cap <= 0) continue;
					const int iter = candidate_stride(npts, cap);
					const int begin = __ldg(rec), end = __ldg(rec + 1);
					const int ncand = end > begin ? (end - begin + iter - 1) / iter : 0;
					for (int c0 = 0; c0 < ncand; c0 += kClsStage) {
						const int nst = min(kClsStage, ncand - c0);
						__syncwarp();
						for (int t = lane; t < nst; t += 32) {
							const int pos = begin + (c0 + t) * iter;
							float4 cand = make_float4(INFINITY, INFINITY, INFINITY, 0.0f);      /* l outside [0, n): skipped upstream, never inside the radius here */
							if (pos >= 0 && pos < n) {
								const uint2 *p = reinterpret_cast<const uint2 *>(cloud + __ldg(vals + pos));
								const uint2 a = p[0], b = p[1];
								cand = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(b.x), 0.0f);
							}
							stage[t] = cand;
						}
						__syncwarp();
						if (acc) {
							for (int t = 0; t < nst; t++) {
								const float4 q = stage[t];
								const float dx = f_sub(x, q.x), dy = f_sub(y, q.y), dz = f_sub(z, q.z);
								const float dist = sqrtf(f_fma(dz, dz, f_fma(dx, dx, f_mul(dy, dy))));      /* SASS-verified association */
								if (dist <= prm.radius) {
									if (sweep == 0) {
										mx = f_add(mx, q.x); my = f_add(my, q.y); mz = f_add(mz, q.z);
										cnt++;
									} else {
										const float ex = f_sub(mx, q.x), ey = f_sub(my, q.y), ez = f_sub(mz, q.z);
										cv[0] += (double)f_mul(ex, ex); cv[1] += (double)f_mul(ex, ey); cv[2] += (double)f_mul(ex, ez);
										cv[3] += (double)f_mul(ey, ey); cv[4] += (double)f_mul(ey, ez); cv[5] += (double)f_mul(ez, ez);
										cnt2++;
									}
								}
							}
						}
					}
				}
			}
		}
		if (!valid) continue;
		if (mean_out) { mean_out[3 * (size_t)l] = active0 ? mx : 0.0f; mean_out[3 * (size_t)l + 1] = active0 ? my : 0.0f; mean_out[3 * (size_t)l + 2] = active0 ? mz : 0.0f; }
		float nx = 0.0f, ny = 0.0f, nz = 0.0f;
		int label = M3DREG_LABEL_EDGE;
		if (have_mean && cnt2 >= prm.plane_points_threshold) {
			const double dn = (double)cnt2;
#pragma unroll
			for (int k = 0; k < 6; k++) cv[k] /= dn;                           /* lesson_16.cu:1064-1072 */
			double ev[3], nv[3];
			sym3_smallest_eigenvector(cv, ev, nv);
			/* upstream rounds the cross product to float before normalising it in double (lesson_16.cu:1077-1081) */
			const double fx = (double)(float)nv[0], fy = (double)(float)nv[1], fz = (double)(float)nv[2];
			const double len = sqrt(fx * fx + fy * fy + fz * fz);
			if (len != 0.0) {
				nx = (float)(fx / len); ny = (float)(fy / len); nz = (float)(fz / len);
				if (ev[1] / ev[2] > (double)prm.curvature_threshold) label = M3DREG_LABEL_PLANE;      /* SS[4] / SS[8] */
			}
		}
		/* kernel_flipNormalsTowardsViewpoint (lesson_16.cu:1112-1133) */
		if (nx * (prm.vx - x) + ny * (prm.vy - y) + nz * (prm.vz - z) < 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
		/* kernel_semanticLabelingFloorCeiling (lesson_16.cu:1196-1219) */
		if (label == M3DREG_LABEL_PLANE && ((double)nz > 0.7 || (double)nz < -0.7)) label = z < prm.ground_z_threshold ? M3DREG_LABEL_GROUND : M3DREG_LABEL_CEILING;
		if (active0) {
			m3dreg_point *o = cloud + idx;
			o->normal_x = nx; o->normal_y = ny; o->normal_z = nz;
			o->label = label;
		}
	}
}

/* sorted (key, value) arrays -> the reference's hashElement records {index_of_point, index_of_bucket} (parity export) */
__global__ void k_join_table(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int n, m3dreg_hash_element *__restrict__ table)
{
	pdl_enter();
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		m3dreg_hash_element h;
		h.index_of_point = (int)__ldg(vals + i); h.index_of_bucket = (int)__ldg(keys + i);
		table[i] = h;
	}
}

/* ---- yaw sweep (findBestYaw): matched queries of a correspondence array ------------------------------------------- */
__global__ void k_count_matches(const int *__restrict__ nn, int n, unsigned int *__restrict__ count)
{
	pdl_enter();
	unsigned int c = 0;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += __ldg(nn + i) >= 0 ? 1u : 0u;
	c = __reduce_add_sync(0xffffffffu, c);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

} /* namespace m3d */
