/* node_host.inl — ROS-free replay of the reference's per-scan driver, class gpu6DSLAM, part of
 * m3dreg.cu.  ref: include/gpu6DSLAM.h:28-262, src/gpu6DSLAM.cpp.
 *
 * What differs from upstream, by design: the scan store is device-resident (every processed scan is uploaded ONCE,
 * m3dreg_scan_upload; upstream re-uploads both clouds for every one of the ~120 iterations of a scan's schedule), the
 * 30 + 30 + 30 iterations of registerLastArrivedScan run as three fused device loops (m3dreg_icp_pair) and every
 * registerAll as one m3dreg_slam_sweep; the pre-registration steps run on the device (preproc_host.inl).
 * Pose algebra (odometry increment, re-anchoring) is plain float 4x4 arithmetic; upstream's is Eigen::Affine3f — unpinned
 * (Eigen absent), tolerance parity only. */
#include <chrono>

namespace {

struct Mat4 {
	float m[16];
	Mat4() { for (int k = 0; k < 16; k++) m[k] = (k % 5 == 0) ? 1.0f : 0.0f; }
	explicit Mat4(const float *p) { memcpy(m, p, sizeof(m)); }
};

Mat4 mat_mul(const Mat4 &a, const Mat4 &b)
{
	Mat4 r;
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++) {
			float s = 0.0f;
			for (int k = 0; k < 4; k++) s += a.m[i * 4 + k] * b.m[k * 4 + j];
			r.m[i * 4 + j] = s;
		}
	return r;
}

/* Eigen::Affine3f::inverse() (Affine mode): inverse of the 3x3 linear part by cofactors, translation = -L^-1 t */
Mat4 mat_affine_inverse(const Mat4 &a)
{
	const float *m = a.m;
	const float c00 = m[5] * m[10] - m[6] * m[9], c01 = m[6] * m[8] - m[4] * m[10], c02 = m[4] * m[9] - m[5] * m[8];
	const float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
	const float id = 1.0f / det;
	Mat4 r;
	r.m[0] = c00 * id; r.m[1] = (m[2] * m[9] - m[1] * m[10]) * id; r.m[2] = (m[1] * m[6] - m[2] * m[5]) * id;
	r.m[4] = c01 * id; r.m[5] = (m[0] * m[10] - m[2] * m[8]) * id; r.m[6] = (m[2] * m[4] - m[0] * m[6]) * id;
	r.m[8] = c02 * id; r.m[9] = (m[1] * m[8] - m[0] * m[9]) * id; r.m[10] = (m[0] * m[5] - m[1] * m[4]) * id;
	for (int i = 0; i < 3; i++) r.m[i * 4 + 3] = -(r.m[i * 4] * m[3] + r.m[i * 4 + 1] * m[7] + r.m[i * 4 + 2] * m[11]);
	r.m[12] = 0.0f; r.m[13] = 0.0f; r.m[14] = 0.0f; r.m[15] = 1.0f;
	return r;
}

void make_dir(const std::string &p)
{
	if (p.empty()) return;
	std::string cur;
	for (size_t i = 0; i <= p.size(); i++) {
		if (i == p.size() || p[i] == '/') {
			if (!cur.empty()) mkdir(cur.c_str(), 0777);      /* boost::filesystem::create_directories */
		}
		if (i < p.size()) cur += p[i];
	}
}

} /* namespace */

struct m3dreg_node {
	m3dreg_ctx *ctx = nullptr;
	m3dreg_node_params prm;
	std::string root, raw_dir, processed_dir;
	bool files = false;
	std::vector<std::vector<m3dreg_point>> vpc;      /* processed scans, local frame (gpu6DSLAM::vpc) */
	std::vector<Mat4> vmtf, vmregistered;
	std::vector<std::string> cloud_ids;
	m3dreg_model *tf_model = nullptr, *processed_model = nullptr, *registered_model = nullptr;
	bool have_last_mtf = false;
	Mat4 last_mtf;

	m3dreg_reg_params reg_params(float radius, float bucket) const
	{
		m3dreg_reg_params r;
		memset(&r, 0, sizeof(r));
		r.search_radius = radius; r.bucket_size = bucket; r.bbox_extension = prm.slam_bounding_box_extension;
		r.max_inner = prm.slam_max_number_considered_in_INNER_bucket; r.max_outer = prm.slam_max_number_considered_in_OUTER_bucket;
		r.obs_threshold = prm.slam_number_of_observations_threshold;
		for (int k = 0; k < 4; k++) r.weight[k] = prm.slam_observation_weight[k];
		r.dof = prm.dof; r.mode = M3DREG_MODE_ICP;
		return r;
	}

	/* gpu6DSLAM::registerAll(cudaWrapper, radius, bucket, number_of_last_EOZ) (src/gpu6DSLAM.cpp:424-597) */
	int sweep(float radius, float bucket, size_t last_eoz, int *solved)
	{
		const int n = (int)vpc.size();
		if (solved) *solved = -1;                                               /* -1: the sweep did not run */
		if ((size_t)n < last_eoz) return 0;                                    /* :428 */
		if (solved) *solved = 0;
		std::vector<float> flat((size_t)n * 16);
		for (int k = 0; k < n; k++) memcpy(&flat[(size_t)k * 16], vmregistered[(size_t)k].m, 64);
		m3dreg_slam_params sp;
		sp.reg = reg_params(radius, bucket);
		sp.distance_threshold = prm.slam_registerAll_distance_threshold;
		sp.first_optimised = n - (int)last_eoz;
		std::vector<int> st((size_t)n, 0);
		int e = m3dreg_slam_sweep(ctx, n, flat.data(), &sp, st.data(), nullptr);
		if (e) return e;
		for (int k = sp.first_optimised; k < n; k++) {
			memcpy(vmregistered[(size_t)k].m, &flat[(size_t)k * 16], 64);
			if (solved && st[(size_t)k] == 0) (*solved)++;
		}
		return 0;
	}
};

extern "C" {

void m3dreg_node_default_params(m3dreg_node_params *p)
{
	if (!p) return;
	memset(p, 0, sizeof(*p));
	/* include/gpu6DSLAM.h:163-223 */
	p->noise_removal_resolution = 0.5f; p->noise_removal_number_of_points_in_bucket_threshold = 3; p->noise_removal_bounding_box_extension = 1.0f;
	p->downsampling_resolution = 0.3f;
	p->semantic_classification_normal_vectors_search_radius = 1.0f; p->semantic_classification_curvature_threshold = 10.0f;
	p->semantic_classification_ground_Z_coordinate_threshold = 1.0f; p->semantic_classification_number_of_points_needed_for_plane_threshold = 15;
	p->semantic_classification_max_number_considered_in_INNER_bucket = 100; p->semantic_classification_max_number_considered_in_OUTER_bucket = 100;
	p->semantic_classification_bounding_box_extension = 1.0f;
	p->slam_registerLastArrivedScan_distance_threshold = 100.0f; p->slam_registerAll_distance_threshold = 10.0f;
	p->slam_number_of_observations_threshold = 100;
	const float rs[3] = {2.5f, 2.0f, 1.0f};
	for (int k = 0; k < 3; k++) {
		p->slam_search_radius_step[k] = rs[k]; p->slam_bucket_size_step[k] = rs[k];
		p->slam_registerLastArrivedScan_number_of_iterations_step[k] = 30; p->slam_registerAll_number_of_iterations_step[k] = 10;
	}
	p->slam_search_radius_register_all = 0.5f; p->slam_bucket_size_step_register_all = 0.5f;
	p->slam_bounding_box_extension = 1.0f;
	p->slam_max_number_considered_in_INNER_bucket = 100; p->slam_max_number_considered_in_OUTER_bucket = 100;
	p->slam_observation_weight[0] = 10.0f; p->slam_observation_weight[1] = 1.0f; p->slam_observation_weight[2] = 10.0f; p->slam_observation_weight[3] = 10.0f;
	p->findBestYaw_start_angle = -30.0f; p->findBestYaw_finish_angle = 30.0f; p->findBestYaw_step_angle = 0.5f;
	p->findBestYaw_bucket_size = 1.0f; p->findBestYaw_bounding_box_extension = 1.0f; p->findBestYaw_search_radius = 0.3f;
	p->findBestYaw_max_number_considered_in_INNER_bucket = 50; p->findBestYaw_max_number_considered_in_OUTER_bucket = 50;
	p->viewpoint[0] = 0.0f; p->viewpoint[1] = 0.0f; p->viewpoint[2] = 2.0f;
	p->cutoff_z_min = -1.0f; p->cutoff_z_max = 15.0f; p->cutoff_xy2_min = 1.5f;      /* src/gpu6DSLAM.cpp:52 */
	p->number_of_last_scans_in_sweeps = 3;
	p->dof = 4;
	p->use_find_best_yaw = 0;
	p->write_files = 1;
}

int m3dreg_node_create(m3dreg_node **out, m3dreg_ctx *ctx, const m3dreg_node_params *params, const char *root_folder)
{
	if (!out || !ctx) return M3DREG_E_INVALID_ARG;
	*out = nullptr;
	m3dreg_node *nd = new (std::nothrow) m3dreg_node();
	if (!nd) return (int)cudaErrorMemoryAllocation;
	nd->ctx = ctx;
	if (params) nd->prm = *params; else m3dreg_node_default_params(&nd->prm);
	if (nd->prm.dof != 4 && nd->prm.dof != 6) { delete nd; return M3DREG_E_INVALID_ARG; }
	nd->files = root_folder && root_folder[0] && nd->prm.write_files;
	if (nd->files) {      /* include/gpu6DSLAM.h:111-153 */
		nd->root = root_folder; nd->raw_dir = nd->root + "/rawData"; nd->processed_dir = nd->root + "/processedData";
		make_dir(nd->root); make_dir(nd->raw_dir); make_dir(nd->processed_dir);
	}
	nd->tf_model = m3dreg_model_create(); nd->processed_model = m3dreg_model_create(); nd->registered_model = m3dreg_model_create();
	/* include/gpu6DSLAM.h:155-162 */
	m3dreg_model_set_algorithm_name(nd->tf_model, "localisation from tf"); m3dreg_model_set_dataset_path(nd->tf_model, "rawData");
	m3dreg_model_set_algorithm_name(nd->processed_model, "processed data: 1: noise removal, 2: downsampling, 3: semantic classification");
	m3dreg_model_set_dataset_path(nd->processed_model, "processedData");
	m3dreg_model_set_algorithm_name(nd->registered_model, "registration: semantic point to point"); m3dreg_model_set_dataset_path(nd->registered_model, "processedData");
	*out = nd;
	return 0;
}

void m3dreg_node_destroy(m3dreg_node *nd)
{
	if (!nd) return;
	m3dreg_model_destroy(nd->tf_model); m3dreg_model_destroy(nd->processed_model); m3dreg_model_destroy(nd->registered_model);
	delete nd;
}

int m3dreg_node_register_single_scan(m3dreg_node *nd, const m3dreg_point *cloud, int n, const float *mtf_in, const char *iso_time_str,
		m3dreg_node_scan_stats *stats)
{
	if (!nd || !cloud || n <= 0 || !mtf_in || !iso_time_str) return M3DREG_E_INVALID_ARG;
	m3dreg_node_scan_stats st;
	memset(&st, 0, sizeof(st));
	const auto t0 = std::chrono::steady_clock::now();
	const Mat4 mtf(mtf_in);
	if (!nd->have_last_mtf) { nd->last_mtf = mtf; nd->have_last_mtf = true; }      /* static last_mtf = mtf (:6) */
	const Mat4 odometry_increment = mat_mul(mat_affine_inverse(nd->last_mtf), mtf); /* :8 */
	const std::string scan_name = std::string("scan_") + iso_time_str, pcd_name = scan_name + ".pcd";
	int e;
	st.n_raw = n;
	if (nd->files && (e = m3dreg_pcd_write_binary((nd->raw_dir + "/" + pcd_name).c_str(), cloud, n))) return e;      /* :41 */

	/* cut off (:47-58) */
	std::vector<m3dreg_point> pc;
	pc.reserve((size_t)n);
	for (int i = 0; i < n; i++) {
		const m3dreg_point &p = cloud[i];
		if ((p.z < nd->prm.cutoff_z_max && p.z > nd->prm.cutoff_z_min) && (p.x * p.x + p.y * p.y > nd->prm.cutoff_xy2_min)) pc.push_back(p);
	}
	st.n_after_cutoff = (int)pc.size();
	/* noise removal, downsampling, classification (:60-85) */
	int kept = 0;
	if (!pc.empty()) {
		if ((e = m3dreg_remove_noise_host(nd->ctx, pc.data(), (int)pc.size(), nd->prm.noise_removal_resolution, nd->prm.noise_removal_bounding_box_extension,
				nd->prm.noise_removal_number_of_points_in_bucket_threshold, pc.data(), &kept, nullptr))) return e;
		pc.resize((size_t)kept);
	}
	st.n_after_noise_removal = (int)pc.size();
	if (!pc.empty()) {
		if ((e = m3dreg_downsample_host(nd->ctx, pc.data(), (int)pc.size(), nd->prm.downsampling_resolution, nd->prm.downsampling_resolution,
				pc.data(), &kept, nullptr))) return e;
		pc.resize((size_t)kept);
	}
	st.n_after_downsampling = (int)pc.size();
	if (pc.empty()) return M3DREG_E_SIZE_MISMATCH;      /* nothing left to register (upstream would push an empty cloud and fail later) */
	if ((e = m3dreg_classify_host(nd->ctx, pc.data(), (int)pc.size(), nd->prm.semantic_classification_normal_vectors_search_radius,
			nd->prm.semantic_classification_curvature_threshold, nd->prm.semantic_classification_ground_Z_coordinate_threshold,
			nd->prm.semantic_classification_number_of_points_needed_for_plane_threshold, nd->prm.semantic_classification_bounding_box_extension,
			nd->prm.semantic_classification_max_number_considered_in_INNER_bucket, nd->prm.semantic_classification_max_number_considered_in_OUTER_bucket,
			nd->prm.viewpoint[0], nd->prm.viewpoint[1], nd->prm.viewpoint[2], nullptr, nullptr))) return e;
	if (nd->files && (e = m3dreg_pcd_write_binary((nd->processed_dir + "/" + pcd_name).c_str(), pc.data(), (int)pc.size()))) return e;      /* :88 */
	const auto t1 = std::chrono::steady_clock::now();
	st.preprocess_ms = std::chrono::duration<float, std::milli>(t1 - t0).count();

	/* scan store: host copy + resident device copy */
	const int slot = (int)nd->vpc.size();
	if ((e = m3dreg_scan_upload(nd->ctx, slot, pc.data(), (int)pc.size(), 0))) return e;
	nd->vpc.push_back(std::move(pc));
	nd->cloud_ids.push_back(scan_name);
	if (slot == 0) {      /* :93-98 */
		nd->vmtf.push_back(mtf);
		nd->vmregistered.push_back(mtf);
	} else {
		nd->vmtf.push_back(mat_mul(nd->vmtf.back(), odometry_increment));                       /* :103-104 */
		nd->vmregistered.push_back(mat_mul(nd->vmregistered.back(), odometry_increment));       /* :106-107 */
		const Mat4 last_inv = mat_affine_inverse(nd->vmregistered.back());                      /* :110-119: re-anchor on the new scan's tf pose */
		for (auto &m : nd->vmregistered) m = mat_mul(last_inv, m);
		for (auto &m : nd->vmregistered) m = mat_mul(mtf, m);
		const int i = slot, j = slot - 1;
		if (nd->prm.use_find_best_yaw) {      /* :135-155 (commented out upstream) */
			const Mat4 inv_prev = mat_affine_inverse(nd->vmregistered[(size_t)j]);
			float best = 0.0f;
			int best_n = 0;
			if ((e = m3dreg_find_best_yaw_host(nd->ctx, nd->vpc[(size_t)j].data(), (int)nd->vpc[(size_t)j].size(), nd->vpc[(size_t)i].data(), (int)nd->vpc[(size_t)i].size(),
					nd->vmregistered[(size_t)i].m, inv_prev.m, nd->prm.findBestYaw_bucket_size, nd->prm.findBestYaw_bounding_box_extension,
					nd->prm.findBestYaw_search_radius, nd->prm.findBestYaw_max_number_considered_in_INNER_bucket,
					nd->prm.findBestYaw_max_number_considered_in_OUTER_bucket, nd->prm.findBestYaw_start_angle, nd->prm.findBestYaw_finish_angle,
					nd->prm.findBestYaw_step_angle, &best, &best_n, nullptr, 0))) return e;
			if (best_n > 0) {
				const float of[3] = {0.0f, 0.0f, (float)((double)best * 3.14159265358979323846 / 180.0)}, tz[3] = {0.0f, 0.0f, 0.0f};
				Mat4 yaw;
				euler_to_matrix(of, tz, yaw.m);
				nd->vmregistered[(size_t)i] = mat_mul(nd->vmregistered[(size_t)i], yaw);
				st.yaw_deg = best;
			}
		}
		/* registerLastArrivedScan x (30, 30, 30) (:159-172): the last scan against its predecessor (j = i - 1, :289), gated
		 * on the pose distance (:304), the pose only replaced by a successful solve (:405-415) — the fused loop's rules */
		for (int step = 0; step < 3; step++) {
			const int iters = nd->prm.slam_registerLastArrivedScan_number_of_iterations_step[step];
			if (iters <= 0) continue;
			const float *a = nd->vmregistered[(size_t)i].m, *b = nd->vmregistered[(size_t)j].m;
			const float dx = a[3] - b[3], dy = a[7] - b[7], dz = a[11] - b[11];
			if (!(sqrtf(dx * dx + dy * dy + dz * dz) < nd->prm.slam_registerLastArrivedScan_distance_threshold)) continue;
			const m3dreg_reg_params rp = nd->reg_params(nd->prm.slam_search_radius_step[step], nd->prm.slam_bucket_size_step[step]);
			m3dreg_icp_stats is;
			if ((e = m3dreg_icp_pair(nd->ctx, i, j, nd->vmregistered[(size_t)i].m, nd->vmregistered[(size_t)j].m, &rp, iters, &is))) return e;
			st.pair_iterations += is.iterations_run;
			st.pair_last_status = is.last_status;
		}
		/* registerAll(..., 3) x (10, 10, 10) (:173-187) */
		for (int step = 0; step < 3; step++)
			for (int it = 0; it < nd->prm.slam_registerAll_number_of_iterations_step[step]; it++) {
				int solved = 0;
				if ((e = nd->sweep(nd->prm.slam_search_radius_step[step], nd->prm.slam_bucket_size_step[step], (size_t)nd->prm.number_of_last_scans_in_sweeps, &solved))) return e;
				if (solved < 0) continue;      /* fewer scans than the sweep optimises: upstream returns at once (:428) */
				st.sweeps++;
				st.sweep_solved_last = solved;
			}
	}
	/* the three models (:201-215) */
	const char *id = nd->cloud_ids.back().c_str();
	m3dreg_model_set_affine(nd->tf_model, id, mtf.m); m3dreg_model_set_cloud_name(nd->tf_model, id, pcd_name.c_str());
	m3dreg_model_set_affine(nd->processed_model, id, mtf.m); m3dreg_model_set_cloud_name(nd->processed_model, id, pcd_name.c_str());
	m3dreg_model_set_cloud_name(nd->registered_model, id, pcd_name.c_str());
	for (size_t k = 0; k < nd->vmregistered.size(); k++) m3dreg_model_set_affine(nd->registered_model, nd->cloud_ids[k].c_str(), nd->vmregistered[k].m);
	if (nd->files) {
		const std::string t = iso_time_str;
		if ((e = m3dreg_model_save(nd->tf_model, (nd->root + "/tfModel_" + t + ".xml").c_str()))) return e;
		if ((e = m3dreg_model_save(nd->processed_model, (nd->root + "/tfModelProcessedData_" + t + ".xml").c_str()))) return e;
		if ((e = m3dreg_model_save(nd->registered_model, (nd->root + "/registeredData_" + t + ".xml").c_str()))) return e;
	}
	nd->last_mtf = mtf;      /* :218 */
	st.register_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t1).count();
	if (stats) *stats = st;
	return 0;
}

int m3dreg_node_scan_count(const m3dreg_node *nd) { return nd ? (int)nd->vpc.size() : M3DREG_E_INVALID_ARG; }

int m3dreg_node_get_pose(const m3dreg_node *nd, int i, float *registered, float *tf)
{
	if (!nd) return M3DREG_E_INVALID_ARG;
	if (i < 0 || (size_t)i >= nd->vpc.size()) return M3DREG_E_BAD_SLOT;
	if (registered) memcpy(registered, nd->vmregistered[(size_t)i].m, 64);
	if (tf) memcpy(tf, nd->vmtf[(size_t)i].m, 64);
	return 0;
}

int m3dreg_node_scan_size(const m3dreg_node *nd, int i)
{
	if (!nd) return M3DREG_E_INVALID_ARG;
	if (i < 0 || (size_t)i >= nd->vpc.size()) return M3DREG_E_BAD_SLOT;
	return (int)nd->vpc[(size_t)i].size();
}

int m3dreg_node_get_scan(const m3dreg_node *nd, int i, m3dreg_point *out, int cap)
{
	if (!nd || !out) return M3DREG_E_INVALID_ARG;
	if (i < 0 || (size_t)i >= nd->vpc.size()) return M3DREG_E_BAD_SLOT;
	if ((size_t)cap < nd->vpc[(size_t)i].size()) return M3DREG_E_SIZE_MISMATCH;
	memcpy(out, nd->vpc[(size_t)i].data(), nd->vpc[(size_t)i].size() * sizeof(m3dreg_point));
	return 0;
}

int m3dreg_node_scan_id(const m3dreg_node *nd, int i, char *out, int cap)
{
	if (!nd) return M3DREG_E_INVALID_ARG;
	if (i < 0 || (size_t)i >= nd->cloud_ids.size()) return M3DREG_E_BAD_SLOT;
	return copy_out(nd->cloud_ids[(size_t)i], out, cap);
}

int m3dreg_node_metascan(m3dreg_node *nd, m3dreg_point *out, int cap, int *n_out)
{
	if (!nd || !n_out) return M3DREG_E_INVALID_ARG;
	size_t total = 0;
	for (auto &v : nd->vpc) total += v.size();
	*n_out = (int)total;
	if (!out) return 0;
	if ((size_t)cap < total) return M3DREG_E_SIZE_MISMATCH;
	/* every scan through the device transform (the arithmetic the registration itself uses), in scan order (:240-245) */
	m3dreg_ctx *c = nd->ctx;
	CK(cudaSetDevice(c->dev));
	size_t off = 0;
	for (size_t k = 0; k < nd->vpc.size(); k++) {
		const int n = (int)nd->vpc[k].size();
		int e;
		if ((e = c->aos_a.ensure((size_t)n))) return e;
		if ((e = c->aos_b.ensure((size_t)n))) return e;
		CK(cudaMemcpyAsync(c->aos_a.p, nd->vpc[k].data(), (size_t)n * sizeof(m3dreg_point), cudaMemcpyHostToDevice, c->stream));
		if ((e = m3dreg_transform(c, c->aos_a.p, c->aos_b.p, n, nd->vmregistered[k].m))) return e;
		CK(cudaMemcpyAsync(out + off, c->aos_b.p, (size_t)n * sizeof(m3dreg_point), cudaMemcpyDeviceToHost, c->stream));
		CK(cudaStreamSynchronize(c->stream));
		off += (size_t)n;
	}
	return 0;
}

int m3dreg_node_register_all(m3dreg_node *nd, int *solved_out)
{
	if (!nd) return M3DREG_E_INVALID_ARG;
	if (solved_out) *solved_out = 0;
	if (nd->vpc.size() <= 1) return 0;                                         /* :625 */
	int solved = 0;
	const int e = nd->sweep(nd->prm.slam_search_radius_register_all, nd->prm.slam_bucket_size_step_register_all, nd->vpc.size(), &solved);
	if (solved_out) *solved_out = solved < 0 ? 0 : solved;
	return e;
}

int m3dreg_node_load_map(m3dreg_node *nd, const char *xml_path)
{
	if (!nd || !xml_path) return M3DREG_E_INVALID_ARG;
	m3dreg_model *mdl = m3dreg_model_create();
	int e = m3dreg_model_load(mdl, xml_path);
	std::vector<std::vector<m3dreg_point>> vpc;
	std::vector<Mat4> vt;
	std::vector<std::string> ids;
	if (e == 0) {
		const int n = m3dreg_model_scan_count(mdl);
		for (int k = 0; k < n && e == 0; k++) {
			char id[512], path[4096];
			m3dreg_model_scan_id(mdl, k, id, sizeof(id));
			Mat4 t;
			const bool ok_tr = m3dreg_model_get_affine(mdl, id, t.m) == 0;
			ids.push_back(id);
			vt.push_back(t);
			if (!ok_tr) continue;                                               /* :690 */
			if (m3dreg_model_full_cloud_path(mdl, id, path, sizeof(path)) < 0) { e = M3DREG_E_IO; break; }
			int cnt = 0;
			if ((e = m3dreg_pcd_read(path, nullptr, 0, &cnt))) break;
			std::vector<m3dreg_point> pc((size_t)cnt);
			if (cnt > 0 && (e = m3dreg_pcd_read(path, pc.data(), cnt, &cnt))) break;
			vpc.push_back(std::move(pc));
		}
	}
	m3dreg_model_destroy(mdl);
	if (e) return e;
	if (vpc.size() != vt.size()) return M3DREG_E_IO;
	m3dreg_scan_clear(nd->ctx);
	for (size_t k = 0; k < vpc.size(); k++)
		if (!vpc[k].empty() && (e = m3dreg_scan_upload(nd->ctx, (int)k, vpc[k].data(), (int)vpc[k].size(), 0))) return e;
	nd->vpc = std::move(vpc);                                                  /* :705-715 */
	nd->vmtf = vt;
	nd->vmregistered = vt;
	nd->cloud_ids = ids;
	return 0;
}

int m3dreg_node_set_initial_pose(m3dreg_node *nd, const float *initial_pose)
{
	if (!nd || !initial_pose) return M3DREG_E_INVALID_ARG;
	const Mat4 ip(initial_pose);
	float min_dist = 10000000.0f;
	Mat4 m;
	for (auto &r : nd->vmregistered) {
		const float dx = ip.m[3] - r.m[3], dy = ip.m[7] - r.m[7], dz = ip.m[11] - r.m[11];
		const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
		if (dist < min_dist) { min_dist = dist; m = r; }
	}
	const Mat4 minv = mat_affine_inverse(m);
	for (auto &r : nd->vmregistered) r = mat_mul(r, minv);                     /* :741-743 */
	for (auto &r : nd->vmregistered) r = mat_mul(r, ip);                       /* :745-747 */
	return 0;
}

} /* extern "C" */
