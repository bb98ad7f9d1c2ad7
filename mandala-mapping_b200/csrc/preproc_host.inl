/* preproc_host.inl — C ABI of the pre-registration steps, part of m3dreg.cu.
 * Kernels: preproc.cuh; the grid is the registration path's own (bounds -> keys -> stable sort -> dense table). */

namespace {

/* Grid of an AoS cloud on the device with cubic buckets (cudaCalculateGridParams + cudaCalculateGrid as every
 * pre-registration method of CCudaWrapper starts, e.g. src/cudaWrapper.cpp:131-141): parameters to c->gp / *gp_host,
 * sorted keys / values in c->keys[cur] / c->vals[cur], dense table in c->buckets.  One read-back (the bounds). */
int grid_of_aos(m3dreg_ctx *c, const m3dreg_point *d_cloud, int n, float res, float ext, m3dreg_grid_params *gp_host, int *cur_out)
{
	int e;
	if ((e = ensure_first(c, (size_t)n))) return e;
	LAUNCH(c, k_reset_bounds, 1, 32, c->bounds);
	LAUNCH(c, k_bounds_aos, grid_for(c, n, 256), 256, d_cloud, n, c->bounds);
	CK(cudaMemcpyAsync(c->h->bounds, c->bounds, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	float mn[3], mx[3];
	for (int k = 0; k < 3; k++) { mn[k] = o2f_host(c->h->bounds[k]); mx[k] = o2f_host(c->h->bounds[3 + k]); }
	int st = grid_params_from_bounds(mn, mx, res, res, res, ext, gp_host);
	if (st) return st;
	if ((e = ensure_buckets(c, (size_t)gp_host->number_of_buckets, false))) return e;
	c->h->gp = *gp_host;
	CK(cudaMemcpyAsync(c->gp, &c->h->gp, sizeof(m3dreg_grid_params), cudaMemcpyHostToDevice, c->stream));
	LAUNCH(c, k_keys_aos, grid_for(c, n, 256), 256, d_cloud, n, c->gp, c->keys[0].p, c->vals[0].p);
	const int cur = sort_by_bucket(c, n, bits_for(gp_host->number_of_buckets), nullptr);
	LAUNCH(c, k_init_buckets, grid_for(c, gp_host->number_of_buckets * 3, 256), 256, c->buckets.p, (const m3dreg_grid_params *)nullptr,
			(long long)gp_host->number_of_buckets);
	LAUNCH(c, k_finalize_grid, grid_for(c, n, 256), 256, c->keys[cur].p, c->vals[cur].p, n, (const m3dreg_grid_params *)nullptr, c->buckets.p,
			(m3dreg_hash_element *)nullptr, (uint32_t *)nullptr, (unsigned int *)nullptr);
	*cur_out = cur;
	c->last_valid = false;
	c->active = false;
	return 0;
}

/* survivors of d_in (markers on the device) in their original order -> host `out`; count -> *n_out */
int compact_to_host(m3dreg_ctx *c, const m3dreg_point *d_in, int n, m3dreg_point *out, int *n_out, unsigned char *markers_out)
{
	int e;
	const int tiles = (n + kCompactTile - 1) / kCompactTile;
	if ((e = c->pp_tiles.ensure((size_t)tiles + 4))) return e;
	if ((e = c->aos_b.ensure((size_t)n))) return e;
	int *d_total = c->pp_tiles.p + tiles;
	LAUNCH(c, k_compact_count, tiles, 256, c->pp_markers.p, n, c->pp_tiles.p);
	LAUNCH(c, k_compact_scan, 1, 1024, c->pp_tiles.p, tiles, d_total);
	LAUNCH(c, k_compact_scatter, tiles, kCompactTile, d_in, c->pp_markers.p, n, c->pp_tiles.p, c->aos_b.p);
	CK(cudaMemcpyAsync(c->h->flags, d_total, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	if (markers_out) CK(cudaMemcpyAsync(markers_out, c->pp_markers.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	const int kept = c->h->flags[0];
	c->h->flags[0] = 0;
	if (kept > 0 && out) CK(cudaMemcpyAsync(out, c->aos_b.p, (size_t)kept * sizeof(m3dreg_point), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	if (n_out) *n_out = kept;
	return launch_status(c);
}

} /* namespace */

extern "C" {

int m3dreg_remove_noise_host(m3dreg_ctx *c, const m3dreg_point *cloud, int n, float resolution, float bounding_box_extension,
		int number_of_points_in_bucket_threshold, m3dreg_point *out, int *n_out, unsigned char *markers_out)
{
	if (!c || !cloud || n <= 0 || !(resolution > 0.0f) || (!out && !markers_out)) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	int e, cur = 0;
	if ((e = c->aos_a.ensure((size_t)n))) return e;
	if ((e = c->pp_markers.ensure((size_t)n))) return e;
	CK(cudaMemcpyAsync(c->aos_a.p, cloud, (size_t)n * sizeof(m3dreg_point), cudaMemcpyHostToDevice, c->stream));
	m3dreg_grid_params gp;
	if ((e = grid_of_aos(c, c->aos_a.p, n, resolution, bounding_box_extension, &gp, &cur))) return e;
	LAUNCH(c, k_mark_noise, grid_for(c, n, 256), 256, c->aos_a.p, n, c->gp, c->buckets.p, number_of_points_in_bucket_threshold, c->pp_markers.p);
	return compact_to_host(c, c->aos_a.p, n, out, n_out, markers_out);
}

int m3dreg_downsample_host(m3dreg_ctx *c, const m3dreg_point *cloud, int n, float resolution, float bounding_box_extension,
		m3dreg_point *out, int *n_out, unsigned char *markers_out)
{
	if (!c || !cloud || n <= 0 || !(resolution > 0.0f) || (!out && !markers_out)) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	int e, cur = 0;
	if ((e = c->aos_a.ensure((size_t)n))) return e;
	if ((e = c->pp_markers.ensure((size_t)n))) return e;
	CK(cudaMemcpyAsync(c->aos_a.p, cloud, (size_t)n * sizeof(m3dreg_point), cudaMemcpyHostToDevice, c->stream));
	m3dreg_grid_params gp;
	if ((e = grid_of_aos(c, c->aos_a.p, n, resolution, bounding_box_extension, &gp, &cur))) return e;
	LAUNCH(c, k_zero_u8, grid_for(c, n, 256), 256, c->pp_markers.p, n);
	LAUNCH(c, k_mark_first_in_bucket, grid_for(c, gp.number_of_buckets, 256), 256, c->buckets.p, c->gp, c->vals[cur].p, c->pp_markers.p);
	return compact_to_host(c, c->aos_a.p, n, out, n_out, markers_out);
}

int m3dreg_classify_host(m3dreg_ctx *c, m3dreg_point *cloud, int n, float normal_vectors_search_radius, float curvature_threshold,
		float ground_Z_coordinate_threshold, int number_of_points_needed_for_plane_threshold, float bounding_box_extension,
		int max_number_considered_in_INNER_bucket, int max_number_considered_in_OUTER_bucket,
		float viewpointX, float viewpointY, float viewpointZ, float *mean_out, m3dreg_hash_element *table_out)
{
	if (!c || !cloud || n <= 0 || !(normal_vectors_search_radius > 0.0f)) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	int e, cur = 0;
	if ((e = c->aos_a.ensure((size_t)n))) return e;
	CK(cudaMemcpyAsync(c->aos_a.p, cloud, (size_t)n * sizeof(m3dreg_point), cudaMemcpyHostToDevice, c->stream));
	m3dreg_grid_params gp;
	if ((e = grid_of_aos(c, c->aos_a.p, n, normal_vectors_search_radius, bounding_box_extension, &gp, &cur))) return e;
	ClassifyParams p;
	p.radius = normal_vectors_search_radius; p.curvature_threshold = curvature_threshold; p.ground_z_threshold = ground_Z_coordinate_threshold;
	p.plane_points_threshold = number_of_points_needed_for_plane_threshold;
	p.max_inner = max_number_considered_in_INNER_bucket; p.max_outer = max_number_considered_in_OUTER_bucket;
	p.vx = viewpointX; p.vy = viewpointY; p.vz = viewpointZ;
	float *d_mean = nullptr;
	if (mean_out) {      /* parity export only: the reference's d_mean, 3 floats per sorted position */
		if ((e = c->obs_rec.ensure((size_t)n))) return e;
		d_mean = reinterpret_cast<float *>(c->obs_rec.p);
	}
	LAUNCH(c, k_classify, c->sm_count * 8, kClsWarps * 32, c->aos_a.p, n, c->keys[cur].p, c->vals[cur].p, c->buckets.p, c->gp, p, d_mean);
	CK(cudaMemcpyAsync(cloud, c->aos_a.p, (size_t)n * sizeof(m3dreg_point), cudaMemcpyDeviceToHost, c->stream));
	if (mean_out) CK(cudaMemcpyAsync(mean_out, d_mean, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
	if (table_out) {
		if ((e = c->table.ensure((size_t)n))) return e;
		LAUNCH(c, k_join_table, grid_for(c, n, 256), 256, c->keys[cur].p, c->vals[cur].p, n, c->table.p);
		CK(cudaMemcpyAsync(table_out, c->table.p, (size_t)n * sizeof(m3dreg_hash_element), cudaMemcpyDeviceToHost, c->stream));
	}
	CK(cudaStreamSynchronize(c->stream));
	return launch_status(c);
}

int m3dreg_find_best_yaw_host(m3dreg_ctx *c, const m3dreg_point *first, int n1, const m3dreg_point *second, int n2,
		const float *second_transform3x4, const float *first_transform_inverse3x4,
		float bucket_size, float bounding_box_extension, float search_radius, int max_inner, int max_outer,
		float angle_start, float angle_finish, float angle_step, float *best_angle_out, int *best_count_out, int *counts_out, int counts_cap)
{
	if (!c || !first || !second || n1 <= 0 || n2 <= 0 || !(bucket_size > 0.0f) || !(angle_step > 0.0f) || !best_angle_out) return M3DREG_E_INVALID_ARG;
	CK(cudaSetDevice(c->dev));
	int e, cur = 0;
	if ((e = c->aos_a.ensure((size_t)n1))) return e;
	if ((e = c->aos_b.ensure((size_t)n2))) return e;
	if ((e = c->pp_aos.ensure((size_t)n2))) return e;
	if ((e = ensure_second(c, (size_t)n2))) return e;
	CK(cudaMemcpyAsync(c->aos_a.p, first, (size_t)n1 * sizeof(m3dreg_point), cudaMemcpyHostToDevice, c->stream));
	CK(cudaMemcpyAsync(c->aos_b.p, second, (size_t)n2 * sizeof(m3dreg_point), cudaMemcpyHostToDevice, c->stream));
	/* second cloud into the first one's frame: two in-place device transforms (src/cudaWrapper.cpp:695-734) */
	const float *ms[2] = {second_transform3x4, first_transform_inverse3x4};
	for (int k = 0; k < 2; k++) {
		const float *m = ms[k];
		if (m) LAUNCH(c, k_transform_aos, (n2 + 255) / 256, 256, c->aos_b.p, c->aos_b.p, n2, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11]);
	}
	/* grid + candidate sets of the FIRST cloud, once (src/cudaWrapper.cpp:740-756) */
	m3dreg_grid_params gp;
	if ((e = grid_of_aos(c, c->aos_a.p, n1, bucket_size, bounding_box_extension, &gp, &cur))) return e;
	if (!nn_columns_usable(gp.number_of_buckets_X, gp.number_of_buckets_Y, gp.number_of_buckets_Z)) return M3DREG_E_TOO_MANY_BUCKETS;
	if ((e = ensure_candidates(c, (size_t)n1, max_inner, max_outer))) return e;
	LAUNCH(c, k_unpack_points, (n1 + 255) / 256, 256, c->aos_a.p, n1, c->g_xyzl.p, c->g_nrm.p);
	CK(cudaMemsetAsync(c->cell_count, 0, sizeof(unsigned int), c->stream));
	LAUNCH(c, k_list_cells, grid_for(c, n1, 256), 256, c->keys[cur].p, n1, c->buckets.p, c->cell_list.p, c->cell_count);
	build_candidates(c, c->vals[cur].p, c->buckets.p, c->cell_list.p, c->g_xyzl.p, c->g_nrm.p, (const float4 *)nullptr, (const float *)nullptr, false, max_inner, max_outer);
	const float res3[3] = {bucket_size, bucket_size, bucket_size};
	int best_count = 0, k = 0;
	float best_angle = angle_start;
	std::vector<unsigned int> h_counts;
	/* every angle: rotate the second cloud about Z (out of place), search, count the matched queries
	 * (src/cudaWrapper.cpp:761-812); the counts come back in one copy at the end */
	int n_angles = 0;
	for (float a = angle_start; a <= angle_finish; a += angle_step) n_angles++;
	if (n_angles <= 0) return M3DREG_E_INVALID_ARG;
	if ((e = c->pp_tiles.ensure((size_t)n_angles + 4))) return e;
	CK(cudaMemsetAsync(c->pp_tiles.p, 0, (size_t)n_angles * sizeof(int), c->stream));
	for (float a = angle_start; a <= angle_finish; a += angle_step, k++) {
		const float rad = (float)((double)a * 3.14159265358979323846 / 180.0);      /* float anglaRad = i * M_PI / 180.0 */
		/* AngleAxis(0, X) * AngleAxis(0, Y) * AngleAxis(rad, Z) goes through a quaternion product upstream, the path
		 * euler_to_matrix() follows (cudaWrapper.cpp:506-514 has the same form) */
		const float of[3] = {0.0f, 0.0f, rad}, tz[3] = {0.0f, 0.0f, 0.0f};
		float ym[16];
		euler_to_matrix(of, tz, ym);
		LAUNCH(c, k_transform_aos, (n2 + 255) / 256, 256, c->aos_b.p, c->pp_aos.p, n2, ym[0], ym[1], ym[2], ym[3], ym[4], ym[5], ym[6], ym[7],
				ym[8], ym[9], ym[10], ym[11]);
		LAUNCH(c, k_unpack_points, (n2 + 255) / 256, 256, c->pp_aos.p, n2, c->q_xyzl.p, c->q_nrm.p);
		launch_nn(c, nullptr, n2, c->vals[cur].p, n1, c->buckets.p, res3, search_radius, max_inner, max_outer, c->prune, c->nn.p, (float4 *)nullptr, c->g_xyzl.p, nullptr);
		LAUNCH(c, k_count_matches, grid_for(c, n2, 256), 256, c->nn.p, n2, reinterpret_cast<unsigned int *>(c->pp_tiles.p) + k);
	}
	h_counts.resize((size_t)n_angles);
	CK(cudaMemcpyAsync(h_counts.data(), c->pp_tiles.p, (size_t)n_angles * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
	CK(cudaStreamSynchronize(c->stream));
	k = 0;
	for (float a = angle_start; a <= angle_finish; a += angle_step, k++) {
		const int cnt = (int)h_counts[(size_t)k];
		if (counts_out && k < counts_cap) counts_out[k] = cnt;
		if (cnt > best_count) { best_count = cnt; best_angle = a; }      /* strict >: the first maximum wins (cudaWrapper.cpp:806-811) */
	}
	*best_angle_out = best_angle;
	if (best_count_out) *best_count_out = best_count;
	c->last_valid = false;
	return launch_status(c);
}

} /* extern "C" */
