/* oracle/m3d_oracle.c — CPU ORACLE: TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See m3d_oracle.h.
 *
 * Compile with -ffp-contract=off: every fused multiply-add below is an explicit fmaf() placed
 * where nvcc 12.9 contracts the reference's expressions (PTX of lesson_16.cu inspected, see
 * DESIGN.md "FP contraction"), everything else is a separately rounded IEEE operation.
 * Citations "L16" = src/lesson_16.cu, "CW" = src/cudaWrapper.cpp, "SL" = src/gpu6DSLAM.cpp,
 * "AXB" = src/CCUDAAXBSolverWrapper.cpp of the reference.
 */
#include "m3d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 0;

int orc_num_threads(void)
{
#ifdef _OPENMP
	return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
	return 1;
#endif
}

void orc_set_num_threads(int n)
{
	g_threads = n;
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
#endif
}

static float f32_from_bits(uint32_t u)
{
	float f;
	memcpy(&f, &u, 4);
	return f;
}

/* ---------------------------------------------------------------------------------------------
 * Grid parameters — L16:23-106 (cudaCalculateGridParams).
 * thrust::minmax_element per axis (L16:34-45), then on the host (L16:64-91):
 *   max += ext; min -= ext; nb = int((max-min)/res + 1)  [float arithmetic, truncation]
 *   number_of_buckets = nbX*nbY*nbZ evaluated in int32 and widened (L16:80).
 * --------------------------------------------------------------------------------------------- */
void orc_grid_params_compute(const orc_point *c, int n, float rx, float ry, float rz, float ext, orc_grid_params *out)
{
	float mnx = c[0].x, mxx = c[0].x, mny = c[0].y, mxy = c[0].y, mnz = c[0].z, mxz = c[0].z;
	for (int i = 1; i < n; i++) {
		if (c[i].x < mnx) mnx = c[i].x;
		if (mxx < c[i].x) mxx = c[i].x;
		if (c[i].y < mny) mny = c[i].y;
		if (mxy < c[i].y) mxy = c[i].y;
		if (c[i].z < mnz) mnz = c[i].z;
		if (mxz < c[i].z) mxz = c[i].z;
	}
	mxx += ext; mnx -= ext;
	mxy += ext; mny -= ext;
	mxz += ext; mnz -= ext;
	int nbx = (int)(((mxx - mnx) / rx) + 1);
	int nby = (int)(((mxy - mny) / ry) + 1);
	int nbz = (int)(((mxz - mnz) / rz) + 1);
	memset(out, 0, sizeof(*out));
	out->number_of_buckets_X = nbx;
	out->number_of_buckets_Y = nby;
	out->number_of_buckets_Z = nbz;
	/* int32 product with wrap-around, as the reference's `int*int*int` behaves on the GPU host */
	out->number_of_buckets = (int64_t)(int32_t)((uint32_t)nbx * (uint32_t)nby * (uint32_t)nbz);
	out->bounding_box_max_X = mxx; out->bounding_box_min_X = mnx;
	out->bounding_box_max_Y = mxy; out->bounding_box_min_Y = mny;
	out->bounding_box_max_Z = mxz; out->bounding_box_min_Z = mnz;
	out->resolution_X = rx; out->resolution_Y = ry; out->resolution_Z = rz;
}

/* Bucket key of one coordinate triple — L16:124-127 (and L16:578-583 for queries):
 * sub.f32, div.rn.f32, cvt.rzi.s32.f32, then int32 mul-add. */
static inline int32_t bucket_key(float x, float y, float z, const orc_grid_params *p, int *ix_o, int *iy_o, int *iz_o)
{
	int ix = (int)((x - p->bounding_box_min_X) / p->resolution_X);
	int iy = (int)((y - p->bounding_box_min_Y) / p->resolution_Y);
	int iz = (int)((z - p->bounding_box_min_Z) / p->resolution_Z);
	if (ix_o) { *ix_o = ix; *iy_o = iy; *iz_o = iz; }
	return (int32_t)((uint32_t)ix * (uint32_t)p->number_of_buckets_Y * (uint32_t)p->number_of_buckets_Z +
			(uint32_t)iy * (uint32_t)p->number_of_buckets_Z + (uint32_t)iz);
}

void orc_bucket_keys(const orc_point *c, int n, const orc_grid_params *p, int32_t *keys)
{
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; i++) keys[i] = bucket_key(c[i].x, c[i].y, c[i].z, p, 0, 0, 0);
}

/* Stable LSD radix sort of (key, index) by key: the permutation thrust::sort produces for the
 * key-only comparator compareHashElements (lesson_16.cuh:6-13) — CUB merge sort, stable, so
 * equal buckets keep ascending original index (L16:213-227). */
static void stable_sort_by_key(orc_hash_element *t, int n)
{
	orc_hash_element *tmp = (orc_hash_element *)malloc((size_t)(n > 0 ? n : 1) * sizeof(*tmp));
	orc_hash_element *src = t, *dst = tmp;
	for (int pass = 0; pass < 4; pass++) {
		size_t cnt[257];
		memset(cnt, 0, sizeof(cnt));
		int shift = pass * 8;
		for (int i = 0; i < n; i++) {
			uint32_t k = (uint32_t)src[i].index_of_bucket ^ 0x80000000u;
			cnt[((k >> shift) & 255u) + 1]++;
		}
		for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
		for (int i = 0; i < n; i++) {
			uint32_t k = (uint32_t)src[i].index_of_bucket ^ 0x80000000u;
			dst[cnt[(k >> shift) & 255u]++] = src[i];
		}
		orc_hash_element *s = src; src = dst; dst = s;
	}
	/* 4 passes: result is back in t */
	free(tmp);
}

/* cudaCalculateGrid — L16:200-243.
 *  initializeIndByKey (L16:109) + getIndexOfBucketForPoints (L16:119) -> table
 *  thrust::sort by bucket (L16:216)
 *  initializeBuckets (L16:131): {-1,-1,0}
 *  updateBuckets (L16:142-174): run boundaries, INCLUDING the ind==0 behaviour: when element 0 is
 *     alone in its bucket the next bucket gets index_end=1 instead of index_begin=1 (L16:154-158),
 *     so that bucket keeps index_begin=-1.  Its index_end is a write race in the reference
 *     (thread 0 writes 1, the thread at the end of that run writes the run end); this restatement
 *     applies the writes in ascending thread order, so the run end wins.  number_of_points stays 0
 *     for it either way, which is all the NN kernel reads before skipping (L16:615-616).
 *  countNumberOfPointsForBuckets (L16:176-189): n = end-begin iff both != -1
 *  copyKeys (L16:191): table out. */
void orc_build_grid(const orc_point *c, int n, const orc_grid_params *p, orc_bucket *buckets, orc_hash_element *table)
{
	int64_t nb = p->number_of_buckets;
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; i++) {
		table[i].index_of_point = i;
		table[i].index_of_bucket = bucket_key(c[i].x, c[i].y, c[i].z, p, 0, 0, 0);
	}
	stable_sort_by_key(table, n);
#pragma omp parallel for schedule(static)
	for (int64_t b = 0; b < nb; b++) {
		buckets[b].index_begin = -1;
		buckets[b].index_end = -1;
		buckets[b].number_of_points = 0;
	}
	for (int ind = 0; ind < n; ind++) {
		if (ind == 0) {
			int b0 = table[0].index_of_bucket;
			buckets[b0].index_begin = 0;
			if (n > 1) { /* the reference reads table[1] unconditionally (out of bounds for n==1) */
				int b1 = table[1].index_of_bucket;
				if (b0 != b1) {
					buckets[b0].index_end = 1;
					buckets[b1].index_end = 1;
				}
			}
			if (n == 1) buckets[b0].index_end = 1; /* the `else if (ind == n-1)` branch is shadowed; n==1 is UB upstream, see DESIGN.md */
		} else if (ind == n - 1) {
			buckets[table[ind].index_of_bucket].index_end = ind + 1;
		} else {
			int b0 = table[ind].index_of_bucket;
			int b1 = table[ind + 1].index_of_bucket;
			if (b0 != b1) {
				buckets[b0].index_end = ind + 1;
				buckets[b1].index_begin = ind + 1;
			}
		}
	}
#pragma omp parallel for schedule(static)
	for (int64_t b = 0; b < nb; b++) {
		if (buckets[b].index_begin != -1 && buckets[b].index_end != -1)
			buckets[b].number_of_points = buckets[b].index_end - buckets[b].index_begin;
	}
}

/* ---------------------------------------------------------------------------------------------
 * Angle gate — L16:666-676:  angle = acos(dot) [float overload -> acosf], angled = angle*180.0f/M_PI
 * (f32 multiply, f64 divide, rounded back to f32), |angled| < 90.0f.
 *
 * acosf here is the CUDA device function.  Its code as inlined by nvcc 12.9 (PTX of the NN kernel):
 *   a = |d|;   if a > 0.56 (0x3F0F5C29): t = sqrt((1-a)/2) via rsqrt.approx + one Newton step, sign of d
 *              else                      t = d
 *   s = t + t*(t2*poly(t2))   (asin polynomial, 6 fmaf)
 *   a <= 0.56 : acos = fma(0x3F6EE581, 0x3FD774EB, -s)            (pi/2 - asin d)
 *   d >  0.56 : acos = 2*s                      in [0, 0.98 rad]   -> angled <= 56.1 deg -> gate TRUE
 *   d < -0.56 : acos = 2*fma(c1,c2,s) = pi-2|s| in [2.16, pi]      -> angled >= 123 deg -> gate FALSE
 *   |d| > 1 or NaN: rsqrt of a negative -> NaN -> every compare false -> gate FALSE
 * Only the |d| <= 0.56 branch decides anything at the bit level, and it contains no approximate
 * instruction, so it is restated exactly with fmaf.  tests/test_gpu_gate.py proves this function
 * equal to the reference expression for ALL 2^32 float inputs on the device.
 * --------------------------------------------------------------------------------------------- */
int orc_angle_gate(float d)
{
	float a = fabsf(d);
	if (!(a <= 1.0f)) return 0;                     /* NaN or |d| > 1 */
	if (a > f32_from_bits(0x3F0F5C29u)) return d > 0.0f;
	float t2 = d * d;
	float p = fmaf(t2, f32_from_bits(0x3D10ECEFu), f32_from_bits(0x3C8B1ABBu));
	p = fmaf(p, t2, f32_from_bits(0x3CFC028Cu));
	p = fmaf(p, t2, f32_from_bits(0x3D372139u));
	p = fmaf(p, t2, f32_from_bits(0x3D9993DBu));
	p = fmaf(p, t2, f32_from_bits(0x3E2AAAC6u));
	float q = t2 * p;
	float s = fmaf(q, d, d);
	float ang = fmaf(f32_from_bits(0x3F6EE581u), f32_from_bits(0x3FD774EBu), -s);
	float deg = ang * 180.0f;
	float angled = (float)((double)deg / 3.14159265358979323846);
	if (angled < 0) angled = -angled;
	return angled < 90.0f;
}

/* ---------------------------------------------------------------------------------------------
 * Semantic nearest neighbour — L16:531-703 (kernel_semanticNearestNeighborSearch), one query per
 * loop iteration instead of per thread.  FP pattern (PTX): dist = fma(dz,dz, fma(dx,dx, dy*dy)),
 * dot = fma(nz,nnz, fma(nx,nnx, ny*nny)), r2 = r*r.
 * --------------------------------------------------------------------------------------------- */
static inline int32_t nn_one_query(const orc_point *first, int n_first, const orc_point *q,
		const orc_hash_element *table, const orc_bucket *buckets, const orc_grid_params *p,
		float r2, int max_inner, int max_outer, int64_t *evals)
{
	float x = q->x, y = q->y, z = q->z;
	float nx = q->normal_x, ny = q->normal_y, nz = q->normal_z;
	int label = q->label;
	if (x < p->bounding_box_min_X || x > p->bounding_box_max_X) return -1; /* L16:562-576 */
	if (y < p->bounding_box_min_Y || y > p->bounding_box_max_Y) return -1;
	if (z < p->bounding_box_min_Z || z > p->bounding_box_max_Z) return -1;
	int ix, iy, iz;
	int32_t index_bucket = bucket_key(x, y, z, p, &ix, &iy, &iz);           /* L16:578-583 */
	int32_t nn_index = -1;
	int isok = 0;
	if (index_bucket >= 0 && (int64_t)index_bucket < p->number_of_buckets) {
		int nbY = p->number_of_buckets_Y, nbZ = p->number_of_buckets_Z;
		int sx = ix == 0 ? 0 : -1, sy = iy == 0 ? 0 : -1, sz = iz == 0 ? 0 : -1;    /* L16:588-595 */
		int stx = ix == p->number_of_buckets_X - 1 ? 1 : 2;
		int sty = iy == nbY - 1 ? 1 : 2;
		int stz = iz == nbZ - 1 ? 1 : 2;
		float best = 100000000.0f;                                                   /* L16:597 */
		for (int i = sx; i < stx; i++)
		for (int j = sy; j < sty; j++)
		for (int k = sz; k < stz; k++) {
			int32_t nb = (int32_t)((uint32_t)index_bucket + (uint32_t)(i * nbY * nbZ) + (uint32_t)(j * nbZ) + (uint32_t)k);
			if (!(nb >= 0 && (int64_t)nb < p->number_of_buckets)) continue;
			int npts = buckets[nb].number_of_points;
			if (npts <= 0) continue;                                                 /* L16:615-616 */
			int cap = nb == index_bucket ? max_inner : max_outer;                    /* L16:618-626 */
			if (cap <= 0) continue;
			int iter;
			if (cap >= npts) iter = 1;
			else { iter = npts / cap; if (iter <= 0) iter = 1; }                     /* L16:628-635 */
			int lb = buckets[nb].index_begin, le = buckets[nb].index_end;
			for (int l = lb; l < le; l += iter) {
				if (!(l >= 0 && l < n_first)) continue;
				int hp = table[l].index_of_point;
				const orc_point *c = &first[hp];
				float dx = x - c->x, dy = y - c->y, dz = z - c->z;
				float dist = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
				float dot = fmaf(nz, c->normal_z, fmaf(nx, c->normal_x, ny * c->normal_y));
				if (evals) (*evals)++;
				if (c->label == label && orc_angle_gate(dot) && dist <= r2 && dist < best) { /* L16:674-686 */
					isok = 1;
					nn_index = hp;
					best = dist;
				}
			}
		}
	}
	return isok ? nn_index : -1;
}

void orc_nn_search(const orc_point *first, int n_first, const orc_point *second, int n_second,
		const orc_hash_element *table, const orc_bucket *buckets, const orc_grid_params *p,
		float search_radius, int max_inner, int max_outer, int32_t *nn_out)
{
	float r2 = search_radius * search_radius;
#pragma omp parallel for schedule(dynamic, 1024)
	for (int q = 0; q < n_second; q++)
		nn_out[q] = nn_one_query(first, n_first, &second[q], table, buckets, p, r2, max_inner, max_outer, 0);
}

int64_t orc_nn_count_evaluations(const orc_point *second, int n_second, const orc_bucket *buckets,
		const orc_grid_params *p, int max_inner, int max_outer)
{
	int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+:total)
	for (int qi = 0; qi < n_second; qi++) {
		const orc_point *q = &second[qi];
		if (q->x < p->bounding_box_min_X || q->x > p->bounding_box_max_X) continue;
		if (q->y < p->bounding_box_min_Y || q->y > p->bounding_box_max_Y) continue;
		if (q->z < p->bounding_box_min_Z || q->z > p->bounding_box_max_Z) continue;
		int ix, iy, iz;
		int32_t ib = bucket_key(q->x, q->y, q->z, p, &ix, &iy, &iz);
		if (!(ib >= 0 && (int64_t)ib < p->number_of_buckets)) continue;
		int nbY = p->number_of_buckets_Y, nbZ = p->number_of_buckets_Z;
		for (int i = (ix == 0 ? 0 : -1); i < (ix == p->number_of_buckets_X - 1 ? 1 : 2); i++)
		for (int j = (iy == 0 ? 0 : -1); j < (iy == nbY - 1 ? 1 : 2); j++)
		for (int k = (iz == 0 ? 0 : -1); k < (iz == nbZ - 1 ? 1 : 2); k++) {
			int32_t nb = ib + i * nbY * nbZ + j * nbZ + k;
			if (!(nb >= 0 && (int64_t)nb < p->number_of_buckets)) continue;
			int npts = buckets[nb].number_of_points;
			if (npts <= 0) continue;
			int cap = nb == ib ? max_inner : max_outer;
			if (cap <= 0) continue;
			int iter = cap >= npts ? 1 : npts / cap;
			if (iter <= 0) iter = 1;
			total += (npts + iter - 1) / iter;
		}
	}
	return total;
}

/* ---------------------------------------------------------------------------------------------
 * Host observation assembly — SL:323-398 (duplicated at SL:490-565).
 * counts per label of the QUERY point (second cloud) over matched queries; obs = {p1-p2 in float
 * from the TRANSFORMED clouds, x0y0z0 from the UNTRANSFORMED first cloud, P = weight/count}.
 * Labels outside 0..3 leave P uninitialised upstream; here P = 0 (documented deviation).
 * --------------------------------------------------------------------------------------------- */
int orc_build_observations(const orc_point *first_global, const orc_point *first_local,
		const orc_point *second_global, int n_second, const int32_t *nn, const float *weight4,
		orc_obs_nn *obs_out)
{
	int cnt[4] = {0, 0, 0, 0};
	for (int i = 0; i < n_second; i++)
		if (nn[i] != -1) {
			int l = second_global[i].label;
			if (l >= 0 && l < 4) cnt[l]++;
		}
	int n = 0;
	for (int i = 0; i < n_second; i++) {
		if (nn[i] == -1) continue;
		const orc_point *p1 = &first_global[nn[i]];
		const orc_point *p2 = &second_global[i];
		orc_obs_nn o;
		o.x0 = first_local[nn[i]].x;
		o.y0 = first_local[nn[i]].y;
		o.z0 = first_local[nn[i]].z;
		o.x_diff = p1->x - p2->x;
		o.y_diff = p1->y - p2->y;
		o.z_diff = p1->z - p2->z;
		int l = p2->label;
		o.P = (l >= 0 && l < 4) ? weight4[l] / (float)cnt[l] : 0.0f;
		obs_out[n++] = o;
	}
	return n;
}

/* ---------------------------------------------------------------------------------------------
 * Normal equations — fill_A_l_cuda (L16:355-428) / fill_A_l_4DOFcuda (L16:441-518), computeR and
 * compute_a1x..a3x (L16:278-353), cudaCompute_AtP (L16:245-255: AtP = A^T * diag(P)), then
 * AtPA = AtP*A and AtPl = AtP*l (AXB:420,427, cuBLAS DGEMM).  Here: the same per-observation 3xdof
 * block, accumulated in long double (the reference's cuBLAS summation order is not specified, so
 * parity downstream of this point is tolerance-based).  Output column-major dof x dof, like cuBLAS.
 * --------------------------------------------------------------------------------------------- */
void orc_normal_equations(const orc_obs_nn *obs, int n_obs, const double *pose6, int dof,
		double *AtPA_out, double *AtPl_out)
{
	double om = pose6[3], fi = pose6[4], ka = pose6[5];
	double r[9];
	r[0] = cos(fi) * cos(ka);
	r[1] = -cos(fi) * sin(ka);
	r[2] = sin(fi);
	r[3] = cos(om) * sin(ka) + sin(om) * sin(fi) * cos(ka);
	r[4] = cos(om) * cos(ka) - sin(om) * sin(fi) * sin(ka);
	r[5] = -sin(om) * cos(fi);
	r[6] = sin(om) * sin(ka) - cos(om) * sin(fi) * cos(ka);
	r[7] = sin(om) * cos(ka) + cos(om) * sin(fi) * sin(ka);
	r[8] = cos(om) * cos(fi);
	long double N[36], b[6];
	for (int i = 0; i < 36; i++) N[i] = 0;
	for (int i = 0; i < 6; i++) b[i] = 0;
	for (int k = 0; k < n_obs; k++) {
		double x0 = obs[k].x0, y0 = obs[k].y0, z0 = obs[k].z0;
		double a14 = 0.0;
		double a15 = (-sin(fi) * cos(ka) * x0 + sin(fi) * sin(ka) * y0 + cos(fi) * z0);
		double a16 = (r[1] * x0 - r[0] * y0);
		double a24 = (-r[6] * x0 - r[7] * y0 - r[8] * z0);
		double a25 = (sin(om) * cos(fi) * cos(ka) * x0 - sin(om) * cos(fi) * sin(ka) * y0 + sin(om) * sin(fi) * z0);
		double a26 = (r[4] * x0 - r[3] * y0);
		double a34 = (r[3] * x0 + r[4] * y0 + r[5] * z0);
		double a35 = (-cos(om) * cos(fi) * cos(ka) * x0 + cos(om) * cos(fi) * sin(ka) * y0 - cos(om) * sin(fi) * z0);
		double a36 = (r[7] * x0 - r[6] * y0);
		double A[3][6];
		if (dof == 6) {
			double t[3][6] = {{-1, 0, 0, -a14, -a15, -a16}, {0, -1, 0, -a24, -a25, -a26}, {0, 0, -1, -a34, -a35, -a36}};
			memcpy(A, t, sizeof(t));
		} else {
			double t[3][6] = {{-1, 0, 0, -a16, 0, 0}, {0, -1, 0, -a26, 0, 0}, {0, 0, -1, -a36, 0, 0}};
			memcpy(A, t, sizeof(t));
		}
		double P = obs[k].P;
		double l[3] = {obs[k].x_diff, obs[k].y_diff, obs[k].z_diff};
		for (int rr = 0; rr < 3; rr++)
			for (int i = 0; i < dof; i++) {
				double atp = A[rr][i] * P;            /* L16:253 */
				b[i] += (long double)atp * l[rr];
				for (int j = 0; j < dof; j++) N[i + j * dof] += (long double)atp * A[rr][j];
			}
	}
	for (int i = 0; i < dof * dof; i++) AtPA_out[i] = (double)N[i];
	for (int i = 0; i < dof; i++) AtPl_out[i] = (double)b[i];
}

/* linearSolverCHOL — AXB:484-539: Dpotrf(LOWER) + Dpotrs.  Returns 0, or k>0 when the k-th leading
 * minor is not positive definite (cuSOLVER's devInfo convention). */
int orc_chol_solve(const double *A, const double *b, int n, double *x)
{
	double L[36];
	memset(L, 0, sizeof(L));
	for (int j = 0; j < n; j++) {
		double d = A[j + j * n];
		for (int k = 0; k < j; k++) d -= L[j + k * n] * L[j + k * n];
		if (!(d > 0.0)) return j + 1;
		d = sqrt(d);
		L[j + j * n] = d;
		for (int i = j + 1; i < n; i++) {
			double s = A[i + j * n];
			for (int k = 0; k < j; k++) s -= L[i + k * n] * L[j + k * n];
			L[i + j * n] = s / d;
		}
	}
	double y[6];
	for (int i = 0; i < n; i++) {
		double s = b[i];
		for (int k = 0; k < i; k++) s -= L[i + k * n] * y[k];
		y[i] = s / L[i + i * n];
	}
	for (int i = n - 1; i >= 0; i--) {
		double s = y[i];
		for (int k = i + 1; k < n; k++) s -= L[k + i * n] * x[k];
		x[i] = s / L[i + i * n];
	}
	return 0;
}

/* registerLS (CW:516-581) / registerLS_4DOF (CW:583-648): pose6 = {tx,ty,tz,om,fi,ka}. */
int orc_register_ls(const orc_obs_nn *obs, int n_obs, double *pose6, int dof, double *x_out)
{
	double N[36], b[6], x[6] = {0, 0, 0, 0, 0, 0};
	orc_normal_equations(obs, n_obs, pose6, dof, N, b);
	int info = orc_chol_solve(N, b, dof, x);
	if (info != 0) return -3;
	pose6[0] += x[0];
	pose6[1] += x[1];
	pose6[2] += x[2];
	if (dof == 6) { pose6[3] += x[3]; pose6[4] += x[4]; pose6[5] += x[5]; }
	else pose6[5] += x[3];
	if (x_out) memcpy(x_out, x, sizeof(double) * (size_t)dof);
	return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Pose helpers.  Matrix4ToEuler(Affine3f) — CW:470-504; EulerToMatrix — CW:506-514.
 * m4x4 is row-major.  EulerToMatrix upstream is Eigen: AngleAxisf(x)*AngleAxisf(y)*AngleAxisf(z)
 * (a float quaternion product) converted to a rotation matrix; Eigen is a third-party dependency
 * absent from /root/reference (unpinned version, find_package(Eigen)), so the published formulae
 * (half-angle quaternions, Hamilton product, quaternion->matrix) are restated in float.  PARITY
 * UNPINNED at the last-ulp level; agreement is by tolerance.
 * --------------------------------------------------------------------------------------------- */
void orc_matrix4_to_euler(const float *m, float *omfika, float *xyz)
{
	double trX, trY;
	if (m[0] > 0.0) omfika[1] = (float)asin((double)m[2]);
	else omfika[1] = (float)(M_PI - asin((double)m[2]));
	double C = cos((double)omfika[1]);
	if (fabs(C) > 0.005) {
		trX = m[10] / C;
		trY = -m[6] / C;
		omfika[0] = (float)atan2(trY, trX);
		trX = m[0] / C;
		trY = -m[1] / C;
		omfika[2] = (float)atan2(trY, trX);
	} else {
		omfika[0] = 0.0f;
		trX = m[5];
		trY = m[4];
		omfika[2] = (float)atan2(trY, trX);
	}
	xyz[0] = m[3];
	xyz[1] = m[7];
	xyz[2] = m[11];
}

void orc_euler_to_matrix(const float *omfika, const float *xyz, float *m)
{
	float hx = 0.5f * omfika[0], hy = 0.5f * omfika[1], hz = 0.5f * omfika[2];
	/* half-angle sin/cos evaluated in double and rounded to float (Eigen uses sinf/cosf; libm and CUDA
	 * float versions differ in the last ulp, the double ones rounded to float practically never do) */
	float ax = (float)sin((double)hx), aw = (float)cos((double)hx);        /* qx = (aw; ax,0,0) */
	float by = (float)sin((double)hy), bw = (float)cos((double)hy);        /* qy = (bw; 0,by,0) */
	float cz = (float)sin((double)hz), cw = (float)cos((double)hz);        /* qz = (cw; 0,0,cz) */
	/* q1 = qx*qy */
	float w1 = aw * bw, x1 = ax * bw, y1 = aw * by, z1 = ax * by;
	/* q = q1*qz */
	float w = w1 * cw - z1 * cz;
	float x = x1 * cw + y1 * cz;
	float y = y1 * cw - x1 * cz;
	float z = w1 * cz + z1 * cw;
	float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
	float twx = tx * w, twy = ty * w, twz = tz * w;
	float txx = tx * x, txy = ty * x, txz = tz * x;
	float tyy = ty * y, tyz = tz * y, tzz = tz * z;
	m[0] = 1.0f - (tyy + tzz); m[1] = txy - twz;          m[2] = txz + twy;           m[3] = xyz[0];
	m[4] = txy + twz;          m[5] = 1.0f - (txx + tzz); m[6] = tyz - twx;           m[7] = xyz[1];
	m[8] = txz - twy;          m[9] = tyz + twx;          m[10] = 1.0f - (txx + tyy); m[11] = xyz[2];
	m[12] = 0.0f; m[13] = 0.0f; m[14] = 0.0f; m[15] = 1.0f;
}

/* Rigid transform of xyz and normals.  Upstream the ICP loop does this on the CPU with Eigen
 * (SL:635-663, rounding not reproducible without Eigen); the only transform whose arithmetic can be
 * pinned is the reference's DEVICE kernel (L16:1341-1367), whose PTX is
 *   v = fma(r02,z, fma(r00,x, r01*y)) + t     and the same without t for normals,
 * so that is what is restated here (and what the product's transform kernel implements). */
void orc_transform_cloud(const orc_point *in, orc_point *out, int n, const float *m)
{
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; i++) {
		orc_point p = in[i];
		float x = p.x, y = p.y, z = p.z;
		float nx = p.normal_x, ny = p.normal_y, nz = p.normal_z;
		p.x = m[3] + fmaf(m[2], z, fmaf(m[0], x, m[1] * y));
		p.y = m[7] + fmaf(m[6], z, fmaf(m[4], x, m[5] * y));
		p.z = m[11] + fmaf(m[10], z, fmaf(m[8], x, m[9] * y));
		p.normal_x = fmaf(m[2], nz, fmaf(m[0], nx, m[1] * ny));
		p.normal_y = fmaf(m[6], nz, fmaf(m[4], nx, m[5] * ny));
		p.normal_z = fmaf(m[10], nz, fmaf(m[8], nx, m[9] * ny));
		out[i] = p;
	}
}

/* ---------------------------------------------------------------------------------------------
 * NDT, point-to-distribution.  The reference has no NDT (only an unused #include, include/gpu6DSLAM.h:17), so this
 * is the DEFINITION the CUDA path is checked against (tolerance; parity unpinned), built from the reference's own
 * pieces: the same grid (L16:23-243), the same pose parametrisation and Jacobian (L16:278-353), the same solve.
 *   bucket b with n_b >= 5 points of the (transformed) first cloud:
 *        mu_g, Sigma_g  mean / sample covariance of the GLOBAL coordinates (accumulated relative to the cell centre)
 *        mu_l           mean of the same points' LOCAL coordinates
 *        W_b = (Sigma_g + eps I)^-1,  eps = (0.05 * resolution)^2
 *   query q (second cloud, global) inside the box whose home bucket is such a b:
 *        l = mu_g - q,  A = -[ I | J(mu_l) ]   (3x6, J = dR/d(om,fi,ka) mu_l),  weight matrix W_b
 *   N = sum A^T W A,  rhs = sum A^T W l,  n_obs = number of such queries.
 * --------------------------------------------------------------------------------------------- */
#define ORC_NDT_MIN_POINTS 5
#define ORC_NDT_REG_REL 0.05

static void ndt_jacobian_coeffs(double om, double fi, double ka, double C[3][3][3])
{
	double so = sin(om), co = cos(om), sf = sin(fi), cf = cos(fi), sk = sin(ka), ck = cos(ka);
	double R11 = cf * ck, R12 = -cf * sk;
	double R21 = co * sk + so * sf * ck, R22 = co * ck - so * sf * sk, R23 = -so * cf;
	double R31 = so * sk - co * sf * ck, R32 = so * ck + co * sf * sk, R33 = co * cf;
	double T[3][3][3] = {
		{{0, 0, 0}, {-sf * ck, sf * sk, cf}, {R12, -R11, 0}},
		{{-R31, -R32, -R33}, {so * cf * ck, -so * cf * sk, so * sf}, {R22, -R21, 0}},
		{{R21, R22, R23}, {-co * cf * ck, co * cf * sk, -co * sf}, {R32, -R31, 0}}};
	memcpy(C, T, sizeof(T));
}

static int sym3_inverse(const double *S, double *W)   /* S, W = {xx,xy,xz,yy,yz,zz} */
{
	double a = S[0], b = S[1], c = S[2], d = S[3], e = S[4], f = S[5];
	double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
	double det = a * c00 + b * c01 + c * c02;
	if (!(det > 0.0)) return 0;
	double id = 1.0 / det;
	W[0] = c00 * id; W[1] = c01 * id; W[2] = c02 * id;
	W[3] = (a * f - c * c) * id; W[4] = (b * c - a * e) * id; W[5] = (a * d - b * b) * id;
	return 1;
}

int64_t orc_ndt_normal_equations(const orc_point *fg, const orc_point *fl, int n_first,
		const orc_point *second, int n_second, const orc_hash_element *table, const orc_bucket *buckets,
		const orc_grid_params *p, const double *pose6, double *neq28)
{
	(void)n_first;
	int64_t nb = p->number_of_buckets;
	int nbY = p->number_of_buckets_Y, nbZ = p->number_of_buckets_Z;
	double res = p->resolution_X;
	double eps = (ORC_NDT_REG_REL * res) * (ORC_NDT_REG_REL * res);
	double *st = (double *)calloc((size_t)nb * 12, sizeof(double));   /* mu_g[3], mu_l[3], W[6] */
	unsigned char *valid = (unsigned char *)calloc((size_t)nb, 1);
#pragma omp parallel for schedule(dynamic, 64)
	for (int64_t b = 0; b < nb; b++) {
		int n = buckets[b].number_of_points;
		if (n < ORC_NDT_MIN_POINTS) continue;
		int ix = (int)(b / ((int64_t)nbY * nbZ)), iy = (int)((b / nbZ) % nbY), iz = (int)(b % nbZ);
		double cx = (double)p->bounding_box_min_X + ((double)ix + 0.5) * (double)p->resolution_X;
		double cy = (double)p->bounding_box_min_Y + ((double)iy + 0.5) * (double)p->resolution_Y;
		double cz = (double)p->bounding_box_min_Z + ((double)iz + 0.5) * (double)p->resolution_Z;
		double s[3] = {0, 0, 0}, ss[6] = {0, 0, 0, 0, 0, 0}, sl[3] = {0, 0, 0};
		for (int l = buckets[b].index_begin; l < buckets[b].index_end; l++) {
			int i = table[l].index_of_point;
			double x = (double)fg[i].x - cx, y = (double)fg[i].y - cy, z = (double)fg[i].z - cz;
			s[0] += x; s[1] += y; s[2] += z;
			ss[0] += x * x; ss[1] += x * y; ss[2] += x * z; ss[3] += y * y; ss[4] += y * z; ss[5] += z * z;
			sl[0] += fl[i].x; sl[1] += fl[i].y; sl[2] += fl[i].z;
		}
		double inv = 1.0 / n, m[3] = {s[0] * inv, s[1] * inv, s[2] * inv};
		double S[6];
		double d = 1.0 / (n - 1);
		S[0] = (ss[0] - n * m[0] * m[0]) * d + eps; S[1] = (ss[1] - n * m[0] * m[1]) * d; S[2] = (ss[2] - n * m[0] * m[2]) * d;
		S[3] = (ss[3] - n * m[1] * m[1]) * d + eps; S[4] = (ss[4] - n * m[1] * m[2]) * d; S[5] = (ss[5] - n * m[2] * m[2]) * d + eps;
		double *o = st + 12 * b;
		if (!sym3_inverse(S, o + 6)) continue;
		o[0] = m[0] + cx; o[1] = m[1] + cy; o[2] = m[2] + cz;
		o[3] = sl[0] * inv; o[4] = sl[1] * inv; o[5] = sl[2] * inv;
		valid[b] = 1;
	}
	double C[3][3][3];
	ndt_jacobian_coeffs(pose6[3], pose6[4], pose6[5], C);
	long double N[6][6], rhs[6];
	for (int i = 0; i < 6; i++) { rhs[i] = 0; for (int j = 0; j < 6; j++) N[i][j] = 0; }
	int64_t n_obs = 0;
	for (int qi = 0; qi < n_second; qi++) {
		const orc_point *q = &second[qi];
		if (q->x < p->bounding_box_min_X || q->x > p->bounding_box_max_X) continue;
		if (q->y < p->bounding_box_min_Y || q->y > p->bounding_box_max_Y) continue;
		if (q->z < p->bounding_box_min_Z || q->z > p->bounding_box_max_Z) continue;
		int32_t b = bucket_key(q->x, q->y, q->z, p, 0, 0, 0);
		if (!(b >= 0 && (int64_t)b < nb) || !valid[b]) continue;
		const double *o = st + 12 * (int64_t)b;
		double l[3] = {o[0] - (double)q->x, o[1] - (double)q->y, o[2] - (double)q->z};
		double W[3][3] = {{o[6], o[7], o[8]}, {o[7], o[9], o[10]}, {o[8], o[10], o[11]}};
		double A[3][6];
		for (int r = 0; r < 3; r++) {
			for (int c = 0; c < 3; c++) {
				A[r][c] = (r == c) ? -1.0 : 0.0;
				A[r][3 + c] = -(C[r][c][0] * o[3] + C[r][c][1] * o[4] + C[r][c][2] * o[5]);
			}
		}
		for (int i = 0; i < 6; i++) {
			double wa[3];   /* (A^T W) row i */
			for (int r = 0; r < 3; r++) wa[r] = A[0][i] * W[0][r] + A[1][i] * W[1][r] + A[2][i] * W[2][r];
			rhs[i] += (long double)(wa[0] * l[0] + wa[1] * l[1] + wa[2] * l[2]);
			for (int j = 0; j < 6; j++) N[i][j] += (long double)(wa[0] * A[0][j] + wa[1] * A[1][j] + wa[2] * A[2][j]);
		}
		n_obs++;
	}
	int k = 0;
	for (int i = 0; i < 6; i++)
		for (int j = i; j < 6; j++) neq28[k++] = (double)N[i][j];
	for (int i = 0; i < 6; i++) neq28[k++] = (double)rhs[i];
	neq28[k] = (double)n_obs;
	free(st); free(valid);
	return n_obs;
}

/* dof-subsystem of a packed 28-double system solved by Cholesky (dof 4 keeps tx,ty,tz,ka). 0 ok, -3 not SPD. */
int orc_solve_packed(const double *neq, int dof, double *x)
{
	const int sel6[6] = {0, 1, 2, 3, 4, 5}, sel4[4] = {0, 1, 2, 5};
	const int *sel = dof == 6 ? sel6 : sel4;
	double full[6][6], A[36], b[6];
	int k = 0;
	for (int i = 0; i < 6; i++)
		for (int j = i; j < 6; j++) { full[i][j] = neq[k]; full[j][i] = neq[k]; k++; }
	for (int i = 0; i < dof; i++) {
		b[i] = neq[21 + sel[i]];
		for (int j = 0; j < dof; j++) A[i + j * dof] = full[sel[i]][sel[j]];
	}
	return orc_chol_solve(A, b, dof, x) == 0 ? 0 : -3;
}

/* ---------------------------------------------------------------------------------------------
 * One iteration of registerLastArrivedScan on pair (i=first, j=second) — SL:264-422.
 *   SL:276-277  Euler round trip of the stored float pose, SL:279-280 transform of cloud i,
 *   SL:313      semantic NN (grid on cloud i, queries = cloud j), SL:323-398 observations,
 *   SL:402-413  if n_obs > threshold: registerLS / registerLS_4DOF, pose = EulerToMatrix(float(obs)).
 * --------------------------------------------------------------------------------------------- */
int orc_icp_iteration(const orc_point *first_local, int n_first, const orc_point *second_global, int n_second,
		float *pose, const orc_reg_params *prm, orc_point *scratch_first, int32_t *nn_out,
		int64_t *n_obs_out, double *x_out)
{
	float omfika[3], xyz[3], pose1[16];
	orc_matrix4_to_euler(pose, omfika, xyz);
	orc_euler_to_matrix(omfika, xyz, pose1);
	orc_transform_cloud(first_local, scratch_first, n_first, pose1);

	orc_grid_params gp;
	orc_grid_params_compute(scratch_first, n_first, prm->bucket_size, prm->bucket_size, prm->bucket_size, prm->bbox_extension, &gp);
	orc_hash_element *table = (orc_hash_element *)malloc((size_t)n_first * sizeof(*table));
	orc_bucket *buckets = (orc_bucket *)malloc((size_t)gp.number_of_buckets * sizeof(*buckets));
	int32_t *nn = nn_out ? nn_out : (int32_t *)malloc((size_t)n_second * sizeof(int32_t));
	orc_obs_nn *obs = (orc_obs_nn *)malloc((size_t)(n_second > 0 ? n_second : 1) * sizeof(*obs));
	orc_build_grid(scratch_first, n_first, &gp, buckets, table);
	int status = -4;
	if (prm->mode == 1) {   /* NDT */
		double pose6[6] = {xyz[0], xyz[1], xyz[2], omfika[0], omfika[1], omfika[2]};
		double neq[28], x[6] = {0, 0, 0, 0, 0, 0};
		int64_t n_ndt = orc_ndt_normal_equations(scratch_first, first_local, n_first, second_global, n_second, table, buckets, &gp, pose6, neq);
		if (n_obs_out) *n_obs_out = n_ndt;
		if (nn_out) for (int i = 0; i < n_second; i++) nn_out[i] = -1;
		if (n_ndt > prm->obs_threshold) {
			status = orc_solve_packed(neq, prm->dof, x);
			if (status == 0) {
				pose6[0] += x[0]; pose6[1] += x[1]; pose6[2] += x[2];
				if (prm->dof == 6) { pose6[3] += x[3]; pose6[4] += x[4]; pose6[5] += x[5]; }
				else pose6[5] += x[3];
				if (x_out) memcpy(x_out, x, sizeof(double) * (size_t)prm->dof);
				float of[3] = {(float)pose6[3], (float)pose6[4], (float)pose6[5]};
				float tf[3] = {(float)pose6[0], (float)pose6[1], (float)pose6[2]};
				orc_euler_to_matrix(of, tf, pose);
			}
		}
		free(table); free(buckets); free(obs);
		if (!nn_out) free(nn);
		return status;
	}
	orc_nn_search(scratch_first, n_first, second_global, n_second, table, buckets, &gp,
			prm->search_radius, prm->max_inner, prm->max_outer, nn);
	int n_obs = orc_build_observations(scratch_first, first_local, second_global, n_second, nn, prm->weight, obs);
	if (n_obs_out) *n_obs_out = n_obs;
	if (n_obs > prm->obs_threshold) {
		double pose6[6] = {xyz[0], xyz[1], xyz[2], omfika[0], omfika[1], omfika[2]};
		status = orc_register_ls(obs, n_obs, pose6, prm->dof, x_out);
		if (status == 0) {
			float of[3] = {(float)pose6[3], (float)pose6[4], (float)pose6[5]};
			float tf[3] = {(float)pose6[0], (float)pose6[1], (float)pose6[2]};
			orc_euler_to_matrix(of, tf, pose);
		}
	}
	free(table); free(buckets); free(obs);
	if (!nn_out) free(nn);
	return status;
}

/* 28-double packing of a 6-DOF system: 21 upper-triangular AtPA entries row by row, 6 AtPl, count. */
static void pack_neq(const double *N6, const double *b6, double count, double *out28)
{
	int k = 0;
	for (int i = 0; i < 6; i++)
		for (int j = i; j < 6; j++) out28[k++] = N6[i + j * 6];
	for (int i = 0; i < 6; i++) out28[k++] = b6[i];
	out28[k] = count;
}

/* ---------------------------------------------------------------------------------------------
 * One Jacobi sweep of registerAll over all scans — SL:424-597 with number_of_last_EOZ = n_scans
 * (the service path, SL:599-633).  Every i uses the OLD poses (SL:596); when the observation gate
 * fails the pose is still replaced by its Euler round trip (SL:586-593).  The 4-DOF system is the
 * {tx,ty,tz,ka} sub-system of the 6-DOF one, so neq_out always carries the 6-DOF equations.
 * --------------------------------------------------------------------------------------------- */
int orc_register_all_sweep(const orc_point *scans, const int64_t *off, int n_scans,
		float *poses, const orc_reg_params *prm, float pair_thr, double *neq_out, int32_t *status_out)
{
	return orc_register_all_sweep_last(scans, off, n_scans, poses, prm, pair_thr, 0, neq_out, status_out);
}

/* registerAll(cudaWrapper, radius, bucket, number_of_last_EOZ) (SL:424-597): only scans first_optimised .. n_scans-1
 * (= the last number_of_last_EOZ) are optimised, the earlier ones keep their pose and serve as neighbours (SL:430-432). */
int orc_register_all_sweep_last(const orc_point *scans, const int64_t *off, int n_scans,
		float *poses, const orc_reg_params *prm, float pair_thr, int first_optimised, double *neq_out, int32_t *status_out)
{
	float *newp = (float *)malloc((size_t)n_scans * 16 * sizeof(float));
	memcpy(newp, poses, (size_t)n_scans * 16 * sizeof(float));
	if (first_optimised < 0) first_optimised = 0;
	for (int i = 0; i < first_optimised && i < n_scans; i++) {
		if (status_out) status_out[i] = 0;
		if (neq_out) memset(neq_out + 28 * (size_t)i, 0, 28 * sizeof(double));
	}
	int64_t maxn = 0;
	for (int i = 0; i < n_scans; i++) if (off[i + 1] - off[i] > maxn) maxn = off[i + 1] - off[i];
	orc_point *pc1 = (orc_point *)malloc((size_t)maxn * sizeof(orc_point));
	orc_point *pc2 = (orc_point *)malloc((size_t)maxn * sizeof(orc_point));
	int32_t *nn = (int32_t *)malloc((size_t)maxn * sizeof(int32_t));
	for (int i = first_optimised; i < n_scans; i++) {
		int n1 = (int)(off[i + 1] - off[i]);
		float of1[3], t1[3], pose1[16];
		orc_matrix4_to_euler(poses + 16 * i, of1, t1);
		orc_euler_to_matrix(of1, t1, pose1);
		orc_transform_cloud(scans + off[i], pc1, n1, pose1);
		orc_grid_params gp;
		orc_grid_params_compute(pc1, n1, prm->bucket_size, prm->bucket_size, prm->bucket_size, prm->bbox_extension, &gp);
		orc_hash_element *table = (orc_hash_element *)malloc((size_t)n1 * sizeof(*table));
		orc_bucket *buckets = (orc_bucket *)malloc((size_t)gp.number_of_buckets * sizeof(*buckets));
		orc_build_grid(pc1, n1, &gp, buckets, table);   /* upstream rebuilds this per j (SL:478); same result */
		double pose6[6] = {t1[0], t1[1], t1[2], of1[0], of1[1], of1[2]};
		double acc[28];
		for (int k = 0; k < 28; k++) acc[k] = 0.0;
		orc_obs_nn *obs = (orc_obs_nn *)malloc((size_t)(maxn > 0 ? maxn : 1) * sizeof(*obs));
		for (int j = 0; j < n_scans; j++) {
			if (j == i) continue;
			int n2 = (int)(off[j + 1] - off[j]);
			float of2[3], t2[3], pose2[16];
			orc_matrix4_to_euler(poses + 16 * j, of2, t2);
			orc_euler_to_matrix(of2, t2, pose2);
			float dist = sqrtf((t1[0] - t2[0]) * (t1[0] - t2[0]) + (t1[1] - t2[1]) * (t1[1] - t2[1]) + (t1[2] - t2[2]) * (t1[2] - t2[2]));
			if (!(dist < pair_thr)) continue;                                     /* SL:469 */
			orc_transform_cloud(scans + off[j], pc2, n2, pose2);
			double pair[28];
			if (prm->mode == 1) {
				orc_ndt_normal_equations(pc1, scans + off[i], n1, pc2, n2, table, buckets, &gp, pose6, pair);
			} else {
				orc_nn_search(pc1, n1, pc2, n2, table, buckets, &gp, prm->search_radius, prm->max_inner, prm->max_outer, nn);
				int n_pair = orc_build_observations(pc1, scans + off[i], pc2, n2, nn, prm->weight, obs);   /* per-pair label counts, SL:490-562 */
				double N6[36], b6[6];
				orc_normal_equations(obs, n_pair, pose6, 6, N6, b6);
				pack_neq(N6, b6, (double)n_pair, pair);
			}
			for (int k = 0; k < 28; k++) acc[k] += pair[k];
		}
		if (neq_out) memcpy(neq_out + 28 * (size_t)i, acc, sizeof(acc));
		int status = -4;
		if ((int64_t)(acc[27] + 0.5) > prm->obs_threshold) {                      /* SL:572 */
			double x[6] = {0, 0, 0, 0, 0, 0};
			status = orc_solve_packed(acc, prm->dof, x);
			if (status == 0) {
				pose6[0] += x[0]; pose6[1] += x[1]; pose6[2] += x[2];
				if (prm->dof == 6) { pose6[3] += x[3]; pose6[4] += x[4]; pose6[5] += x[5]; }
				else pose6[5] += x[3];
			}
		}
		if (status_out) status_out[i] = status;
		float of[3] = {(float)pose6[3], (float)pose6[4], (float)pose6[5]};
		float tf[3] = {(float)pose6[0], (float)pose6[1], (float)pose6[2]};
		orc_euler_to_matrix(of, tf, newp + 16 * i);                               /* SL:586-593 */
		free(obs); free(table); free(buckets);
	}
	memcpy(poses, newp, (size_t)n_scans * 16 * sizeof(float));
	free(newp); free(pc1); free(pc2); free(nn);
	return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Pre-registration steps.  CW = src/cudaWrapper.cpp, L16 = src/lesson_16.cu,
 * SVD = src/cuda_SVD.cu.  Every step starts with the grid of the cloud itself with cubic buckets
 * (cudaCalculateGridParams + cudaCalculateGrid, e.g. CW:131-141).
 * --------------------------------------------------------------------------------------------- */

static void grid_of(const orc_point *c, int n, float res, float ext, orc_grid_params *gp, orc_bucket **buckets, orc_hash_element **table)
{
	orc_grid_params_compute(c, n, res, res, res, ext, gp);
	int64_t nb = gp->number_of_buckets > 0 ? gp->number_of_buckets : 1;
	*buckets = (orc_bucket *)malloc((size_t)nb * sizeof(orc_bucket));
	*table = (orc_hash_element *)malloc((size_t)(n > 0 ? n : 1) * sizeof(orc_hash_element));
	orc_build_grid(c, n, gp, *buckets, *table);
}

/* CCudaWrapper::removeNoiseNaive (CW:118-179): kernel_setAllPointsToRemove + kernel_markPointsToRemain (L16:740-766) —
 * a point stays iff the bucket of ITS coordinates holds more than `threshold` points (the dense table's count, so the
 * points of the first-element quirk bucket, whose count stays 0, are dropped). */
void orc_remove_noise_markers(const orc_point *c, int n, float res, float ext, int threshold, uint8_t *markers)
{
	orc_grid_params gp; orc_bucket *b; orc_hash_element *t;
	grid_of(c, n, res, ext, &gp, &b, &t);
#pragma omp parallel for schedule(static)
	for (int i = 0; i < n; i++) {
		int32_t key = bucket_key(c[i].x, c[i].y, c[i].z, &gp, 0, 0, 0);
		markers[i] = (key >= 0 && (int64_t)key < gp.number_of_buckets && b[key].number_of_points > threshold) ? 1 : 0;
	}
	free(b); free(t);
}

/* CCudaWrapper::downsampling (CW:181-262): kernel_markFirstPointInBuckets (L16:789-801) — the first sorted point of
 * every bucket with an index_begin. */
void orc_downsample_markers(const orc_point *c, int n, float res, float ext, uint8_t *markers)
{
	orc_grid_params gp; orc_bucket *b; orc_hash_element *t;
	grid_of(c, n, res, ext, &gp, &b, &t);
	memset(markers, 0, (size_t)n);
	for (int64_t k = 0; k < gp.number_of_buckets; k++)
		if (b[k].index_begin != -1) markers[t[b[k].index_begin].index_of_point] = 1;
	free(b); free(t);
}

/* ---- the 3x3 decomposition upstream uses (SVD:29-380, Nathan Lay's closed-form SVD): eigen-decomposition of A^T A by
 * the trigonometric cubic solution, null vectors by a pivoted LDU, U = A V.  Restated for the oracle with the same
 * storage convention: matrices are 9 doubles, element (row r, column c) at [3*c + r]; U + 3*k is the k-th vector. ---- */
static void svd_ata3(double *AA, const double *A)      /* SVD:29-42 */
{
	AA[0] = A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
	AA[3] = A[0] * A[3] + A[1] * A[4] + A[2] * A[5];
	AA[6] = A[0] * A[6] + A[1] * A[7] + A[2] * A[8];
	AA[1] = AA[3];
	AA[4] = A[3] * A[3] + A[4] * A[4] + A[5] * A[5];
	AA[7] = A[3] * A[6] + A[4] * A[7] + A[5] * A[8];
	AA[2] = AA[6];
	AA[5] = AA[7];
	AA[8] = A[6] * A[6] + A[7] * A[7] + A[8] * A[8];
}

static void svd_solvecubic(double *c)                  /* SVD:44-82: roots of x^3 + c2 x^2 + c1 x + c0 */
{
	const double sq3d2 = 0.86602540378443864676, c2d3 = c[2] / 3, c2sq = c[2] * c[2], Q = (3 * c[1] - c2sq) / 9,
			R = (c[2] * (9 * c[1] - 2 * c2sq) - 27 * c[0]) / 54;
	if (Q < 0) {
		double tmp = 2 * sqrt(-Q), t = acos(R / sqrt(-Q * Q * Q)) / 3, cost = tmp * cos(t), sint = tmp * sin(t);
		c[0] = cost - c2d3;
		cost = -0.5 * cost - c2d3;
		sint = sq3d2 * sint;
		c[1] = cost - sint;
		c[2] = cost + sint;
	} else {
		double tmp = cbrt(R);
		c[0] = -c2d3 + 2 * tmp;
		c[1] = c[2] = -c2d3 - tmp;
	}
}

static void svd_sort3(double *x)                       /* SVD:84-106: descending */
{
	double tmp;
	if (x[0] < x[1]) { tmp = x[0]; x[0] = x[1]; x[1] = tmp; }
	if (x[1] < x[2]) {
		if (x[0] < x[2]) { tmp = x[2]; x[2] = x[1]; x[1] = x[0]; x[0] = tmp; }
		else { tmp = x[1]; x[1] = x[2]; x[2] = tmp; }
	}
}

static void svd_ldu3(double *A, int *P)                /* SVD:108-141: row-pivoted LDU in place */
{
	P[1] = 1; P[2] = 2;
	P[0] = fabs(A[3]) > fabs(A[0]) ? (fabs(A[6]) > fabs(A[3]) ? 2 : 1) : (fabs(A[6]) > fabs(A[0]) ? 2 : 0);
	P[P[0]] = 0;
	if (fabs(A[3 * P[2] + 1]) > fabs(A[3 * P[1] + 1])) { int tmp = P[1]; P[1] = P[2]; P[2] = tmp; }
	if (A[3 * P[0]] != 0) {
		A[3 * P[1]] = A[3 * P[1]] / A[3 * P[0]];
		A[3 * P[2]] = A[3 * P[2]] / A[3 * P[0]];
		A[3 * P[0] + 1] = A[3 * P[0] + 1] / A[3 * P[0]];
		A[3 * P[0] + 2] = A[3 * P[0] + 2] / A[3 * P[0]];
	}
	A[3 * P[1] + 1] = A[3 * P[1] + 1] - A[3 * P[0] + 1] * A[3 * P[1]] * A[3 * P[0]];
	if (A[3 * P[1] + 1] != 0) {
		A[3 * P[2] + 1] = (A[3 * P[2] + 1] - A[3 * P[0] + 1] * A[3 * P[2]] * A[3 * P[0]]) / A[3 * P[1] + 1];
		A[3 * P[1] + 2] = (A[3 * P[1] + 2] - A[3 * P[0] + 2] * A[3 * P[1]] * A[3 * P[0]]) / A[3 * P[1] + 1];
	}
	A[3 * P[2] + 2] = A[3 * P[2] + 2] - A[3 * P[0] + 2] * A[3 * P[2]] * A[3 * P[0]] - A[3 * P[1] + 2] * A[3 * P[2] + 1] * A[3 * P[1] + 1];
}

static void svd_ldubsolve3(double *x, const double *y, const double *LDU, const int *P)      /* SVD:143-148 */
{
	x[P[2]] = y[2];
	x[P[1]] = y[1] - LDU[3 * P[2] + 1] * x[P[2]];
	x[P[0]] = y[0] - LDU[3 * P[2]] * x[P[2]] - LDU[3 * P[1]] * x[P[1]];
}

static void svd_cross(double *z, const double *x, const double *y)                          /* SVD:150-155 */
{
	z[0] = x[1] * y[2] - x[2] * y[1];
	z[1] = -(x[0] * y[2] - x[2] * y[0]);
	z[2] = x[0] * y[1] - x[1] * y[0];
}

static void svd_matvec3(double *y, const double *A, const double *x)                        /* SVD:157-162 */
{
	y[0] = A[0] * x[0] + A[3] * x[1] + A[6] * x[2];
	y[1] = A[1] * x[0] + A[4] * x[1] + A[7] * x[2];
	y[2] = A[2] * x[0] + A[5] * x[1] + A[8] * x[2];
}

static void svd_unit3(double *x)                                                             /* SVD:179-185 */
{
	double tmp = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
	x[0] /= tmp; x[1] /= tmp; x[2] /= tmp;
}

static int svd_nearest_zero(const double *LDU, const int *P, int first_system)              /* SVD:253-256, 283-286 */
{
	double d0 = fabs(LDU[3 * P[0]]), d1 = fabs(LDU[3 * P[1] + 1]), d2 = fabs(LDU[3 * P[2] + 2]);
	if (first_system) return d1 < d0 ? (d2 < d1 ? 2 : 1) : (d2 < d0 ? 2 : 0);
	return d0 < d2 ? (d0 < d1 ? 0 : 1) : (d1 < d2 ? 1 : 2);
}

/* gpuSVD (SVD:187-380).  NOTE the index expressions LDU[P[k]][k] upstream address element [3*P[k] + k] of the flat
 * array, restated as such. */
static void svd3(const double *A, double *U, double *S9, double *V)
{
	const double thr = 1e-10;
	int P[3], k;
	double y[3], AA[9], LDU[9], S[3];
	svd_ata3(AA, A);
	S[2] = -AA[0] - AA[4] - AA[8];
	S[1] = AA[0] * AA[4] + AA[8] * AA[0] + AA[8] * AA[4] - AA[7] * AA[5] - AA[6] * AA[2] - AA[3] * AA[1];
	S[0] = AA[7] * AA[5] * AA[0] + AA[6] * AA[2] * AA[4] + AA[3] * AA[1] * AA[8] - AA[0] * AA[4] * AA[8] - AA[3] * AA[7] * AA[2] -
			AA[6] * AA[1] * AA[5];
	svd_solvecubic(S);
	if (S[0] < 0) S[0] = 0;
	if (S[1] < 0) S[1] = 0;
	if (S[2] < 0) S[2] = 0;
	svd_sort3(S);
	memcpy(LDU, AA, sizeof(LDU));
	LDU[0] -= S[0]; LDU[4] -= S[0]; LDU[8] -= S[0];
	svd_ldu3(LDU, P);
	y[0] = y[1] = y[2] = 0;
	y[svd_nearest_zero(LDU, P, 1)] = 1;
	svd_ldubsolve3(V, y, LDU, P);
	memcpy(LDU, AA, sizeof(LDU));
	LDU[0] -= S[2]; LDU[4] -= S[2]; LDU[8] -= S[2];
	svd_ldu3(LDU, P);
	y[0] = y[1] = y[2] = 0;
	y[svd_nearest_zero(LDU, P, 0)] = 1;
	svd_ldubsolve3(V + 6, y, LDU, P);
	svd_cross(V + 3, V + 6, V);
	k = (S[0] > thr) + (S[1] > thr) + (S[2] > thr);
	switch (k) {
	case 0:
		memcpy(U, V, 9 * sizeof(double));
		break;
	case 1:
		svd_matvec3(U, A, V);
		y[0] = y[1] = y[2] = 0;
		k = fabs(U[0]) < fabs(U[2]) ? (fabs(U[0]) < fabs(U[1]) ? 0 : 1) : (fabs(U[1]) < fabs(U[2]) ? 1 : 2);
		y[k] = 1;
		svd_cross(U + 3, y, U);
		svd_cross(U + 6, U, U + 3);
		break;
	case 2:
		svd_matvec3(U, A, V);
		svd_matvec3(U + 3, A, V + 3);
		svd_cross(U + 6, U, U + 3);
		break;
	default:
		svd_matvec3(U, A, V);
		svd_matvec3(U + 3, A, V + 3);
		svd_matvec3(U + 6, A, V + 6);
		break;
	}
	svd_unit3(V); svd_unit3(V + 3); svd_unit3(V + 6);
	svd_unit3(U); svd_unit3(U + 3); svd_unit3(U + 6);
	memset(S9, 0, 9 * sizeof(double));
	S9[0] = sqrt(S[0]); S9[4] = sqrt(S[1]); S9[8] = sqrt(S[2]);
}

/* one sweep over the 27-neighbourhood of sorted position `pos` (L16:858-941 and L16:989-1060): sweep 0 accumulates the
 * float coordinate sums, sweep 1 the covariance about `mean` (float differences, float products, double sums). */
static int classify_sweep(const orc_point *c, int n, const orc_hash_element *t, const orc_bucket *b, const orc_grid_params *gp,
		int key, float x, float y, float z, float radius, int max_in, int max_out, int sweep, float *sum3, const float *mean, double *cov9)
{
	const int nby = gp->number_of_buckets_Y, nbz = gp->number_of_buckets_Z, nbx = gp->number_of_buckets_X;
	const int ix = key / (nby * nbz), iy = (key % (nby * nbz)) / nbz, iz = (key % (nby * nbz)) % nbz;
	const int sx = ix == 0 ? 0 : -1, sy = iy == 0 ? 0 : -1, sz = iz == 0 ? 0 : -1;
	const int stx = ix == nbx - 1 ? 1 : 2, sty = iy == nby - 1 ? 1 : 2, stz = iz == nbz - 1 ? 1 : 2;
	int cnt = 0;
	for (int i = sx; i < stx; i++)
	for (int j = sy; j < sty; j++)
	for (int k = sz; k < stz; k++) {
		const int nbk = key + i * nby * nbz + j * nbz + k;
		if (nbk < 0 || (int64_t)nbk >= gp->number_of_buckets) continue;
		const int npts = b[nbk].number_of_points;
		if (npts <= 0) continue;
		const int cap = nbk == key ? max_in : max_out;
		if (cap <= 0) continue;
		int iter = 1;
		if (cap < npts) { iter = npts / cap; if (iter <= 0) iter = 1; }
		for (int l = b[nbk].index_begin; l < b[nbk].index_end; l += iter) {
			if (l < 0 || l >= n) continue;
			const orc_point *q = &c[t[l].index_of_point];
			const float dx = x - q->x, dy = y - q->y, dz = z - q->z;
			const float dist = sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));      /* SASS of the reference build: (dy*dy) then fma dx, fma dz */
			if (!(dist <= radius)) continue;
			if (sweep == 0) { sum3[0] += q->x; sum3[1] += q->y; sum3[2] += q->z; }
			else {
				const float ex = mean[0] - q->x, ey = mean[1] - q->y, ez = mean[2] - q->z;
				cov9[0] += (double)(ex * ex); cov9[1] += (double)(ex * ey); cov9[2] += (double)(ex * ez);
				cov9[3] += (double)(ey * ex); cov9[4] += (double)(ey * ey); cov9[5] += (double)(ey * ez);
				cov9[6] += (double)(ez * ex); cov9[7] += (double)(ez * ey); cov9[8] += (double)(ez * ez);
			}
			cnt++;
		}
	}
	return cnt;
}

/* CCudaWrapper::classify (CW:264-342): grid with cubic buckets of the search radius, cudaSemanticLabelingPlaneEdges
 * (L16:817-1194: step1 mean, step2 covariance + SVD + plane test, flip towards the viewpoint), then
 * cudaSemanticLabelingFloorCeiling (L16:1196-1239).  In place.  mean_out: 3 floats per SORTED position (d_mean). */
void orc_classify(orc_point *c, int n, float radius, float curvature_threshold, float ground_z, int plane_points, float ext,
		int max_in, int max_out, float vx, float vy, float vz, float *mean_out, orc_hash_element *table_out)
{
	orc_grid_params gp; orc_bucket *b; orc_hash_element *t;
	grid_of(c, n, radius, ext, &gp, &b, &t);
	float *mean = (float *)calloc((size_t)(n > 0 ? n : 1) * 3, sizeof(float));
	for (int i = 0; i < n; i++) { c[i].normal_x = 0.0f; c[i].normal_y = 0.0f; c[i].normal_z = 0.0f; }      /* L16:831-833 */
#pragma omp parallel for schedule(dynamic, 256)
	for (int pos = 0; pos < n; pos++) {                                       /* step 1, L16:817-957 */
		const int key = t[pos].index_of_bucket, idx = t[pos].index_of_point;
		if (!(key >= 0 && (int64_t)key < gp.number_of_buckets) || !(idx >= 0 && idx < n)) continue;
		float s[3] = {0.0f, 0.0f, 0.0f};
		const int cnt = classify_sweep(c, n, t, b, &gp, key, c[idx].x, c[idx].y, c[idx].z, radius, max_in, max_out, 0, s, 0, 0);
		if (cnt >= 3) { mean[3 * pos] = s[0] / (float)cnt; mean[3 * pos + 1] = s[1] / (float)cnt; mean[3 * pos + 2] = s[2] / (float)cnt; }
	}
#pragma omp parallel for schedule(dynamic, 256)
	for (int pos = 0; pos < n; pos++) {                                       /* step 2, L16:959-1110 */
		const int key = t[pos].index_of_bucket, idx = t[pos].index_of_point;
		if (!(key >= 0 && (int64_t)key < gp.number_of_buckets) || !(idx >= 0 && idx < n)) continue;
		orc_point *p = &c[idx];
		p->label = 1;                                                          /* LABEL_EDGE */
		const float *m = mean + 3 * pos;
		if (!(m[0] != 0.0f && m[1] != 0.0f && m[2] != 0.0f)) continue;
		double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
		const int cnt = classify_sweep(c, n, t, b, &gp, key, p->x, p->y, p->z, radius, max_in, max_out, 1, 0, m, cov);
		if (cnt >= plane_points) {
			for (int k = 0; k < 9; k++) cov[k] /= cnt;
			double U[9], V[9], SS[9];
			svd3(cov, U, SS, V);
			const double nx = (float)(U[1] * U[5] - U[2] * U[4]), ny = (float)(-(U[0] * U[5] - U[2] * U[3])), nz = (float)(U[0] * U[4] - U[1] * U[3]);
			const double len = sqrt(nx * nx + ny * ny + nz * nz);
			if (len != 0) {
				p->normal_x = (float)(nx / len); p->normal_y = (float)(ny / len); p->normal_z = (float)(nz / len);
				if (SS[4] / SS[8] > (double)curvature_threshold) p->label = 0;    /* LABEL_PLANE */
			}
		}
	}
	for (int i = 0; i < n; i++) {
		orc_point *p = &c[i];
		/* kernel_flipNormalsTowardsViewpoint (L16:1112-1133); float expression, nvcc contracts it as fma(nz,dz, fma(ny,dy, nx*dx)) */
		const float d = fmaf(p->normal_z, vz - p->z, fmaf(p->normal_y, vy - p->y, p->normal_x * (vx - p->x)));
		if ((double)d < 0.0) { p->normal_x = -p->normal_x; p->normal_y = -p->normal_y; p->normal_z = -p->normal_z; }
		/* kernel_semanticLabelingFloorCeiling (L16:1196-1219) */
		if (p->label == 0 && ((double)p->normal_z > 0.7 || (double)p->normal_z < -0.7)) p->label = p->z < ground_z ? 3 : 2;
	}
	if (mean_out) memcpy(mean_out, mean, (size_t)n * 3 * sizeof(float));
	if (table_out) memcpy(table_out, t, (size_t)n * sizeof(orc_hash_element));
	free(mean); free(b); free(t);
}

/* CCudaWrapper::findBestYaw (CW:662-836): second cloud into the first one's frame (two device transforms), grid of the
 * first cloud, then per angle: rotate about Z out of place, semantic NN, count the matched queries
 * (cudaCountNumberOfSemanticNearestNeighbours, L16:1241-1298); strictly more matches win.  Matrices row-major 3x4
 * (rows of a 4x4 also fine: 12 floats are read); NULL = identity.  Returns the index of the best angle or -1. */
int orc_find_best_yaw(const orc_point *first, int n1, const orc_point *second, int n2, const float *second_m, const float *first_inv_m,
		float bucket, float ext, float radius, int max_in, int max_out, const float *yaw_mats, int n_angles, int32_t *counts_out)
{
	orc_point *s = (orc_point *)malloc((size_t)n2 * sizeof(orc_point)), *r = (orc_point *)malloc((size_t)n2 * sizeof(orc_point));
	int32_t *nn = (int32_t *)malloc((size_t)n2 * sizeof(int32_t));
	memcpy(s, second, (size_t)n2 * sizeof(orc_point));
	for (int k = 0; k < 2; k++) {
		const float *m = k == 0 ? second_m : first_inv_m;
		if (m) { orc_transform_cloud(s, r, n2, m); memcpy(s, r, (size_t)n2 * sizeof(orc_point)); }
	}
	orc_grid_params gp; orc_bucket *b; orc_hash_element *t;
	grid_of(first, n1, bucket, ext, &gp, &b, &t);
	int best = -1, best_n = 0;
	for (int a = 0; a < n_angles; a++) {
		orc_transform_cloud(s, r, n2, yaw_mats + 12 * (size_t)a);
		orc_nn_search(first, n1, r, n2, t, b, &gp, radius, max_in, max_out, nn);
		int cnt = 0;
		for (int i = 0; i < n2; i++) cnt += nn[i] >= 0 ? 1 : 0;
		if (counts_out) counts_out[a] = cnt;
		if (cnt > best_n) { best_n = cnt; best = a; }
	}
	free(s); free(r); free(nn); free(b); free(t);
	return best;
}
