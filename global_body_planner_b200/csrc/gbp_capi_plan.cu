// extern "C" layer, part 3: extend / connect through the host call and the device-resident batch planner.
#include "gbp_host.h"
#include "gbp_planner.cuh"

// the pipelined form lives in gbp_capi_pipeline.cu
bool gbp_plan_pipe_applies(const TerrainView &Tv, const gbp_plan_params &P, int64_t nq);
int gbp_plan_pipe_launch(const TerrainView &Tv, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
						 const gbp_plan_params &P, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap, cudaStream_t st,
						 const PlanTreeDump &dump, std::string &err);
bool gbp_plan_wide_applies(const gbp_plan_params &P, int64_t nq);
int gbp_plan_wide_launch(const TerrainView &Tv, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
						 const gbp_plan_params &P, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap, cudaStream_t st,
						 const PlanTreeDump &dump, std::string &err);

extern "C" {

// -------------------------------------------------------------------------- extend / connect
static int ensure_k(gbp_tree *T, int K) {
	if (K <= T->k_cap) return GBP_OK;
	cudaFree(T->S.valid); cudaFree(T->S.dist); cudaFree(T->S.s_test);
	T->S.valid = nullptr; T->S.dist = nullptr; T->S.s_test = nullptr; T->k_cap = 0;
	CU(cudaMalloc(&T->S.valid, K));
	CU(cudaMalloc(&T->S.dist, sizeof(double) * K));
	CU(cudaMalloc(&T->S.s_test, sizeof(double) * 8 * K));
	T->k_cap = K;
	return GBP_OK;
}
int gbp_extend(gbp_tree *T, const gbp_terrain *t, const double *target, int direction, int K, int best_of_k, int adaptive,
			   double dir_thresh, uint64_t seed, uint64_t stream, uint64_t idx0, int *status, int *new_id, int64_t *pair_checks) {
	if (!T || !t || !target || K < 1 || (direction != GBP_FORWARD && direction != GBP_REVERSE)) return fail(GBP_E_INVALID, "bad arguments");
	int rc;
	if ((rc = ensure_k(T, K))) return rc;
	cudaStream_t st = lib_stream();
	Target8 tg;
	memcpy(tg.v, target, sizeof tg.v);
	int *dres = nullptr;
	CU(cudaHostGetDevicePointer((void **) &dres, T->h_result, 0));
	if (adaptive) {
		GBP_DISPATCH(t->view, k_extend_fused_adaptive, (blocks_for(K, 128), 128), st, t->view, T->view, tg, direction, K, best_of_k, 1, seed, stream, idx0,
					 dir_thresh, T->S, T->d_done, dres);
	} else {
		// lanes per candidate: as many as keep the launch inside one wave of resident CTAs (2 x 128 threads per SM at ~250 registers)
		int S = 32;
		while (S > 1 && (int64_t) K * S > (int64_t) sm_count() * 2 * 128) S >>= 1;
		const int per_block = 4 * (32 / S);
		GBP_DISPATCH(t->view, k_extend_fused, (blocks_for(K, per_block), 128), st, t->view, T->view, tg, direction, K, best_of_k, S, seed, stream, idx0,
					 dir_thresh, T->S, T->d_done, dres);
	}
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(st));
	const int res[3] = {T->h_result[0], T->h_result[1], T->h_result[2]};
	if (pair_checks) *pair_checks = res[2];
	if (res[0] < 0) return fail(GBP_E_CAPACITY, "tree full: the accepted vertex was not appended");
	if (status) *status = res[0];
	if (new_id) *new_id = res[1];
	return GBP_OK;
}
int gbp_attempt_connect_ts(const gbp_terrain *t, int64_t n, const double *s_existing, const double *s, const double *t_s,
						   const uint8_t *direction, int adaptive, int *status, double *s_new, double *a_new, uint8_t *flags) {
	if (!t || n < 0 || (n && (!s_existing || !s || !direction || !status || !s_new || !a_new))) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	cudaStream_t st = lib_stream();
	Dev de(st), ds(st), dts(st), dd(st), dst(st), dsn(st), dan(st), dfl(st);
	int rc;
	if ((rc = upload(de, s_existing, (size_t) 8 * n, st)) || (rc = upload(ds, s, (size_t) 8 * n, st)) || (rc = upload(dd, direction, (size_t) n, st))) return rc;
	if (t_s && (rc = upload(dts, t_s, (size_t) n, st))) return rc;
	CU(dst.alloc(sizeof(int) * n));
	CU(dsn.alloc(sizeof(double) * 8 * n));
	CU(dan.alloc(sizeof(double) * 10 * n));
	if (flags) CU(dfl.alloc(n));
	GBP_DISPATCH(t->view, k_attempt_connect, (blocks_for(n, 128), 128), st, t->view, n, de.as<double>(), ds.as<double>(),
				 t_s ? dts.as<double>() : nullptr, dd.as<uint8_t>(), adaptive, dst.as<int>(), dsn.as<double>(), dan.as<double>(), dfl.as<uint8_t>());
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(status, dst.p, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(s_new, dsn.p, sizeof(double) * 8 * n, cudaMemcpyDeviceToHost, st));
	CU(cudaMemcpyAsync(a_new, dan.p, sizeof(double) * 10 * n, cudaMemcpyDeviceToHost, st));
	if (flags) CU(cudaMemcpyAsync(flags, dfl.p, n, cudaMemcpyDeviceToHost, st));
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
int gbp_attempt_connect(const gbp_terrain *t, int64_t n, const double *s_existing, const double *s, const uint8_t *direction,
						int adaptive, int *status, double *s_new, double *a_new, uint8_t *flags) {
	return gbp_attempt_connect_ts(t, n, s_existing, s, nullptr, direction, adaptive, status, s_new, a_new, flags);
}
int gbp_new_config(const gbp_terrain *t, const double *target, const double *s_near, int direction, int K, int best_of_k, int adaptive,
				   double dir_thresh, uint64_t seed, uint64_t stream, uint64_t idx0, int *found, double *s_new, double *a_new, int64_t *pair_checks) {
	if (!t || !target || !s_near || !found || !s_new || !a_new) return fail(GBP_E_INVALID, "bad arguments");
	static thread_local gbp_tree *scratch = nullptr;  // a one-vertex tree whose nearest neighbour is s_near
	int rc;
	if (!scratch && (rc = gbp_tree_create(2, &scratch))) return rc;
	if ((rc = gbp_tree_init(scratch, s_near))) return rc;
	int status = GBP_TRAPPED, id = -1;
	if ((rc = gbp_extend(scratch, t, target, direction, K, best_of_k, adaptive, dir_thresh, seed, stream, idx0, &status, &id, pair_checks))) return rc;
	*found = status != GBP_TRAPPED;
	if (*found) return gbp_tree_read(scratch, id, 1, s_new, a_new, nullptr, nullptr, nullptr);
	return GBP_OK;
}
int gbp_connect(gbp_tree *T, const gbp_terrain *t, const double *target, int direction, int adaptive, int *status, int *new_id) {
	if (!T || !t || !target || (direction != GBP_FORWARD && direction != GBP_REVERSE)) return fail(GBP_E_INVALID, "bad arguments");
	cudaStream_t st = lib_stream();
	Target8 tg;
	memcpy(tg.v, target, sizeof tg.v);
	int *dres = nullptr;
	CU(cudaHostGetDevicePointer((void **) &dres, T->h_result, 0));
	GBP_DISPATCH(t->view, k_connect, (1, 32), st, t->view, T->view, tg, direction, adaptive, T->S, dres);
	CU(cudaGetLastError());
	CU(cudaStreamSynchronize(st));
	const int res[3] = {T->h_result[0], T->h_result[1], T->h_result[2]};
	if (res[0] < 0) return fail(GBP_E_CAPACITY, "tree full: the connect vertex was not appended");
	if (status) *status = res[0];
	if (new_id) *new_id = res[1];
	return GBP_OK;
}

// ------------------------------------------------------------------------------ batch planner
static int plan_check(const gbp_terrain *t, int64_t nq, const double *starts, const double *goals, const gbp_plan_params *p, const gbp_plan_stats *stats) {
	if (!t || nq < 0 || !p || (nq && (!starts || !goals || !stats))) return fail(GBP_E_INVALID, "bad arguments");
	if (p->k_candidates < 1 || p->max_iters < 0 || p->max_vertices < 2) return fail(GBP_E_INVALID, "bad plan parameters");
	return GBP_OK;
}
static int plan_dev(const gbp_terrain *t, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
					const gbp_plan_params *p, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap,
					const PlanTreeDump &dump, cudaStream_t st) {
	keep_pool();
	std::string err;
	const int rc = gbp_plan_wide_applies(*p, nq)
		? gbp_plan_wide_launch(t->view, nq, starts, goals, seed, query0, *p, stats, path_states, path_actions, path_cap, st, dump, err)
		: gbp_plan_pipe_applies(t->view, *p, nq)
		? gbp_plan_pipe_launch(t->view, nq, starts, goals, seed, query0, *p, stats, path_states, path_actions, path_cap, st, dump, err)
		: plan_step_applies(*p, nq)
		? plan_step_launch(t->view, nq, starts, goals, seed, query0, *p, stats, path_states, path_actions, path_cap, st, dump, err)
		: plan_batch_launch(t->view, nq, starts, goals, seed, query0, *p, stats, path_states, path_actions, path_cap, st, dump, err);
	if (rc) return fail(rc, err);
	return GBP_OK;
}
int gbp_plan_batch_form(const gbp_terrain *t, const gbp_plan_params *p, int64_t nq, int *form) {
	if (!t || !p || !form) return fail(GBP_E_INVALID, "bad arguments");
	*form = gbp_plan_wide_applies(*p, nq) ? 3 : gbp_plan_pipe_applies(t->view, *p, nq) ? 1 : (plan_step_applies(*p, nq) ? 2 : 0);
	return GBP_OK;
}
int gbp_plan_batch_dev(const gbp_terrain *t, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
					   const gbp_plan_params *p, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap,
					   void *stream) {
	int rc;
	if ((rc = plan_check(t, nq, starts, goals, p, stats))) return rc;
	if (nq == 0) return GBP_OK;
	const PlanTreeDump none = {0, nullptr, nullptr, nullptr, nullptr, nullptr};
	return plan_dev(t, nq, starts, goals, seed, query0, p, stats, path_states, path_actions, path_cap, none, (cudaStream_t) stream);
}
int gbp_plan_batch_trees(const gbp_terrain *t, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
						 const gbp_plan_params *p, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap, int tree_cap,
						 double *tree_states, double *tree_actions, int *tree_parent, double *tree_g, double *tree_yaw) {
	int rc;
	if ((rc = plan_check(t, nq, starts, goals, p, stats))) return rc;
	if (nq == 0) return GBP_OK;
	const bool want_trees = tree_cap > 0 && tree_states && tree_actions && tree_parent && tree_g && tree_yaw;
	if (tree_cap > 0 && !want_trees) return fail(GBP_E_INVALID, "tree_cap > 0 needs all five tree arrays");
	cudaStream_t st = lib_stream();
	Dev ds(st), dg(st), dst(st), dps(st), dpa(st), dts(st), dta(st), dtp(st), dtg(st), dty(st);
	if ((rc = upload(ds, starts, (size_t) 8 * nq, st)) || (rc = upload(dg, goals, (size_t) 8 * nq, st))) return rc;
	CU(dst.alloc(sizeof(gbp_plan_stats) * nq));
	const bool want_paths = path_states && path_actions && path_cap > 0;
	if (want_paths) {
		CU(dps.alloc(sizeof(double) * 8 * (size_t) path_cap * nq));
		CU(dpa.alloc(sizeof(double) * 10 * (size_t) path_cap * nq));
	}
	PlanTreeDump dump = {0, nullptr, nullptr, nullptr, nullptr, nullptr};
	const size_t rows = want_trees ? (size_t) nq * 2 * tree_cap : 0;
	if (want_trees) {
		CU(dts.alloc(rows * 64)); CU(dta.alloc(rows * 80)); CU(dtp.alloc(rows * 4)); CU(dtg.alloc(rows * 8)); CU(dty.alloc(rows * 8));
		dump.cap = tree_cap; dump.states = dts.as<double>(); dump.actions = dta.as<double>(); dump.parent = dtp.as<int>();
		dump.g = dtg.as<double>(); dump.y = dty.as<double>();
	}
	if ((rc = plan_dev(t, nq, ds.as<double>(), dg.as<double>(), seed, query0, p, dst.as<gbp_plan_stats>(), want_paths ? dps.as<double>() : nullptr,
					   want_paths ? dpa.as<double>() : nullptr, path_cap, dump, st)))
		return rc;
	CU(cudaMemcpyAsync(stats, dst.p, sizeof(gbp_plan_stats) * nq, cudaMemcpyDeviceToHost, st));
	if (want_paths) {
		CU(cudaMemcpyAsync(path_states, dps.p, sizeof(double) * 8 * (size_t) path_cap * nq, cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(path_actions, dpa.p, sizeof(double) * 10 * (size_t) path_cap * nq, cudaMemcpyDeviceToHost, st));
	}
	if (want_trees) {
		CU(cudaMemcpyAsync(tree_states, dts.p, rows * 64, cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(tree_actions, dta.p, rows * 80, cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(tree_parent, dtp.p, rows * 4, cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(tree_g, dtg.p, rows * 8, cudaMemcpyDeviceToHost, st));
		CU(cudaMemcpyAsync(tree_yaw, dty.p, rows * 8, cudaMemcpyDeviceToHost, st));
	}
	CU(cudaStreamSynchronize(st));
	return GBP_OK;
}
int gbp_plan_batch(const gbp_terrain *t, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
				   const gbp_plan_params *p, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap) {
	return gbp_plan_batch_trees(t, nq, starts, goals, seed, query0, p, stats, path_states, path_actions, path_cap, 0, nullptr, nullptr, nullptr,
								nullptr, nullptr);
}

}  // extern "C"
