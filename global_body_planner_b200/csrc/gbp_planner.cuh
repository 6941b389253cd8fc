// Device-resident batch planner: one WARP per planning query runs the whole bidirectional loop
// (runRRTConnect, rrt_connect.cpp:230-314) with an iteration budget instead of the wall clock.
// Queries are independent, so there is no inter-warp or inter-GPU traffic.
//   nearest neighbour : lanes stride over the query's SoA tree, warp-shuffle argmin
//   newConfig         : lane j validates candidate j (ACTION cell (2*iter+half)*K + j), ballot /
//                       argmin selection (first valid in stream order, or closest valid)
//   connect           : closed-form action + pair check, evaluated warp-uniformly
// Tree arenas live in HBM, one slot per resident warp, reused across the queries a warp processes.
#pragma once
#include <string>

#include "gbp_kernels.cuh"

namespace gbp {

struct PlanArena {
	double *v, *act, *g, *y;  // per slot: 2 trees x cap x {8, 10, 1, 1} doubles
	int *parent;              // per slot: 2 trees x cap
	int cap;
};

__device__ __forceinline__ TreeView arena_tree(const PlanArena &A, int slot, int which, int *n_ptr) {
	TreeView T;
	const size_t t = (size_t) slot * 2 + which;
	T.cap = A.cap;
	T.n = n_ptr;
	T.v = A.v + t * 8 * A.cap;
	T.act = A.act + t * 10 * A.cap;
	T.parent = A.parent + t * A.cap;
	T.g = A.g + t * A.cap;
	T.y = A.y + t * A.cap;
	return T;
}

__device__ __forceinline__ int warp_nearest(const TreeView &T, int nv, const double q[8], int lane) {
	double bd = INFINITY;
	int bi = 0x7fffffff;
	for (int j = lane; j < nv; j += 32) argmin_combine(bd, bi, vertex_distance(T, j, q), j);
	warp_argmin(bd, bi);
	return bi == 0x7fffffff ? 0 : bi;
}

// extend (rrt.cpp:77-102) on tree T toward s; returns status, appends on success.
template <typename M>
__device__ int warp_extend(const TerrainView &Tv, TreeView &T, int &nv, const double s[8], int dir, uint64_t seed, uint64_t query,
						   uint64_t cell, const gbp_plan_params &P, int lane, long long &pair_checks) {
	const int near = warp_nearest(T, nv, s, lane);
	double s_near[8], nn[3], R[9];
	tree_get(T, near, s_near);
	const double best0 = state_distance(s_near, s);
	unsigned fl = 0;
	surface_normal(Tv, s[0], s[1], nn, fl);  // rrt.cpp:25
	grf_rotation(nn, R);
	const int K = P.k_candidates;
	double my_d = INFINITY, my_sn[8], my_a[10];
	int my_j = 0x7fffffff;
	int first = 0x7fffffff;
	for (int base = 0; base < K; base += 32) {
		const int j = base + lane;
		bool ok = false;
		double a[10], sn[8], tn;
		if (j < K) {
			sample_action(seed, query, cell * (uint64_t) K + (uint64_t) j, R, false, 0.0, nullptr, nullptr, a);
			Counters c = {0, 0, 0, 0};
			ok = validate_pair_seq<M>(Tv, s_near, a, dir, P.adaptive != 0, sn, tn, c);
		}
		if (ok) {
			const double d = state_distance(sn, s);
			if (P.best_of_k ? (d < my_d) : (my_j == 0x7fffffff)) {
				my_d = d; my_j = j;
#pragma unroll
				for (int i = 0; i < 8; ++i) my_sn[i] = sn[i];
#pragma unroll
				for (int i = 0; i < 10; ++i) my_a[i] = a[i];
			}
		}
		if (!P.best_of_k) {
			const unsigned m = __ballot_sync(FULL, ok);
			if (m) { first = base + __ffs(m) - 1; break; }  // first valid action decides (rrt.cpp:44-47)
		}
	}
	int src_lane;
	double d_sel;
	if (P.best_of_k) {
		pair_checks += K;
		double bd = my_d;
		int bj = my_j;
		warp_argmin(bd, bj);
		if (bj == 0x7fffffff) return GBP_TRAPPED;
		src_lane = bj & 31;
		d_sel = bd;
	} else {
		pair_checks += (first == 0x7fffffff) ? K : first + 1;
		if (first == 0x7fffffff) return GBP_TRAPPED;
		src_lane = first & 31;
		d_sel = __shfl_sync(FULL, my_d, src_lane);
	}
	if (!(d_sel < best0)) return GBP_TRAPPED;  // rrt.cpp:55-66
	double sn[8], a[10];
#pragma unroll
	for (int i = 0; i < 8; ++i) sn[i] = __shfl_sync(FULL, my_sn[i], src_lane);
#pragma unroll
	for (int i = 0; i < 10; ++i) a[i] = __shfl_sync(FULL, my_a[i], src_lane);
	if (lane == 0) tree_push(T, near, sn, a);
	__syncwarp();
	nv += 1;
	return state_distance(sn, s) <= GOAL_BOUNDS ? GBP_REACHED : GBP_ADVANCED;
}

// connect (rrt_connect.cpp:98-120)
template <typename M>
__device__ int warp_connect(const TerrainView &Tv, TreeView &T, int &nv, const double s[8], int dir, const gbp_plan_params &P,
							int lane, long long &pair_checks) {
	const int near = warp_nearest(T, nv, s, lane);
	double s_near[8], sn[8], an[10];
	tree_get(T, near, s_near);
	Counters c = {0, 0, 0, 0};
	unsigned checks = 0;
	const int r = attempt_connect<M>(Tv, s_near, s, dir, P.adaptive != 0, sn, an, c, checks);
	pair_checks += checks;
	if (r != GBP_TRAPPED) {
		if (lane == 0) tree_push(T, near, sn, an);
		__syncwarp();
		nv += 1;
	}
	return r;
}

template <typename M>
__global__ void __launch_bounds__(128) k_plan_batch(TerrainView Tv, int64_t nq, const double *__restrict__ starts,
													 const double *__restrict__ goals, uint64_t seed, uint64_t query0,
													 gbp_plan_params P, PlanArena A, int *__restrict__ counts,
													 gbp_plan_stats *__restrict__ stats, double *__restrict__ path_states,
													 double *__restrict__ path_actions, int path_cap) {
	const int lane = threadIdx.x & 31;
	const int64_t slot = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
	const int64_t nslots = ((int64_t) gridDim.x * blockDim.x) >> 5;
	for (int64_t qi = slot; qi < nq; qi += nslots) {
		TreeView Ta = arena_tree(A, (int) slot, 0, counts + 2 * slot), Tb = arena_tree(A, (int) slot, 1, counts + 2 * slot + 1);
		double start[8], goal[8];
#pragma unroll
		for (int d = 0; d < 8; ++d) { start[d] = starts[8 * qi + d]; goal[d] = goals[8 * qi + d]; }
		if (lane == 0) {  // GraphClass::init (graph_class.cpp:140-152)
			*Ta.n = 1; *Tb.n = 1;
			for (int d = 0; d < 8; ++d) { Ta.v[(size_t) d * Ta.cap] = start[d]; Tb.v[(size_t) d * Tb.cap] = goal[d]; }
			for (int d = 0; d < 10; ++d) { Ta.act[(size_t) d * Ta.cap] = 0; Tb.act[(size_t) d * Tb.cap] = 0; }
			Ta.parent[0] = -1; Tb.parent[0] = -1; Ta.g[0] = 0; Tb.g[0] = 0; Ta.y[0] = 0; Tb.y[0] = 0;
		}
		__syncwarp();
		int na = 1, nb = 1, it = 0;
		bool solved = false, full = false;
		long long pair_checks = 0, nn_queries = 0;
		const uint64_t query = query0 + (uint64_t) qi;
		for (; it < P.max_iters && !solved && !full; ++it) {
			for (int half = 0; half < 2 && !solved; ++half) {
				TreeView &Tx = half == 0 ? Ta : Tb, &Ty = half == 0 ? Tb : Ta;
				int &nx = half == 0 ? na : nb, &ny = half == 0 ? nb : na;
				const int dir_ext = half == 0 ? GBP_FORWARD : GBP_REVERSE, dir_con = half == 0 ? GBP_REVERSE : GBP_FORWARD;
				if (nx >= A.cap || ny >= A.cap) { full = true; break; }
				const uint64_t cell = 2 * (uint64_t) it + (uint64_t) half;
				double s_rand[8];
				sample_state<M>(Tv, seed, query, cell, false, 0.0, false, nullptr, nullptr, s_rand);
				Counters c = {0, 0, 0, 0};
				if (!is_valid_state_auto<M>(Tv, pose6(s_rand), GBP_STANCE, c)) continue;  // rrt_connect.cpp:254
				++nn_queries;
				if (warp_extend<M>(Tv, Tx, nx, s_rand, dir_ext, seed, query, cell, P, lane, pair_checks) == GBP_TRAPPED) continue;
				double s_new[8];
				tree_get(Tx, nx - 1, s_new);
				++nn_queries;
				if (warp_connect<M>(Tv, Ty, ny, s_new, dir_con, P, lane, pair_checks) == GBP_REACHED) solved = true;
			}
		}
		// statistics + path (rrt_connect.cpp:269-270, :381-401, :463-466)
		if (lane == 0) {
			gbp_plan_stats st;
			st.solved = solved ? 1 : 0; st.iters = it; st.nv_a = na; st.nv_b = nb; st.path_states = 0; st.pad = 0;
			st.path_length = 0; st.path_yaw = 0; st.path_duration = 0; st.pair_checks = pair_checks; st.nn_queries = nn_queries;
			if (solved) {
				st.path_length = Ta.g[na - 1] + Tb.g[nb - 1];
				st.path_yaw = Ta.y[na - 1] + Tb.y[nb - 1];
				int la = 0, lb = 0;
				for (int i = na - 1; i != -1; i = Ta.parent[i]) ++la;
				for (int i = nb - 1; i != -1; i = Tb.parent[i]) ++lb;
				const int total = la + lb - 1;
				st.path_states = total;
				double dur = 0;
				double *ps = path_states ? path_states + (size_t) qi * path_cap * 8 : nullptr;
				double *pa = path_actions ? path_actions + (size_t) qi * path_cap * 10 : nullptr;
				// the duration sum follows the path order (a_0 .. a_{total-2}) so the fp64 sum matches the reference's
				int k = la - 1;
				for (int i = na - 1; i != -1; i = Ta.parent[i], --k) {
					if (ps && k < path_cap) for (int d = 0; d < 8; ++d) ps[8 * k + d] = Ta.v[(size_t) d * Ta.cap + i];
					if (pa && k > 0 && k - 1 < path_cap) for (int d = 0; d < 10; ++d) pa[10 * (k - 1) + d] = Ta.act[(size_t) d * Ta.cap + i];
				}
				// tree A actions in path order: walk again from the root side using the stored parents
				for (int step = 1; step < la; ++step) {
					int i = na - 1;
					for (int up = 0; up < la - 1 - step; ++up) i = Ta.parent[i];
					dur += Ta.act[(size_t) 6 * Ta.cap + i] + Ta.act[(size_t) 7 * Ta.cap + i];
				}
				k = la - 1;
				for (int i = nb - 1; Tb.parent[i] != -1; i = Tb.parent[i], ++k) {
					dur += Tb.act[(size_t) 6 * Tb.cap + i] + Tb.act[(size_t) 7 * Tb.cap + i];
					if (pa && k < path_cap) for (int d = 0; d < 10; ++d) pa[10 * k + d] = Tb.act[(size_t) d * Tb.cap + i];
					if (ps && k + 1 < path_cap) for (int d = 0; d < 8; ++d) ps[8 * (k + 1) + d] = Tb.v[(size_t) d * Tb.cap + Tb.parent[i]];
				}
				st.path_duration = dur;
			}
			stats[qi] = st;
		}
		__syncwarp();
	}
}

// host-side launcher: sizes the grid to the SM count, allocates the tree arena for the resident warps
inline int plan_batch_launch(const TerrainView &Tv, int64_t nq, const double *starts, const double *goals, uint64_t seed,
							 uint64_t query0, const gbp_plan_params &P, gbp_plan_stats *stats, double *path_states, double *path_actions,
							 int path_cap, cudaStream_t st, std::string &err) {
	if (P.rrt_star || P.post_process) { err = "rrt_star / post_process are not implemented in the device planner yet"; return GBP_E_INVALID; }
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int threads = 128, warps_per_block = threads / 32;
	int64_t slots = (int64_t) sms * 16;  // resident warps
	if (slots > nq) slots = ((nq + warps_per_block - 1) / warps_per_block) * warps_per_block;
	const unsigned grid = (unsigned) (slots / warps_per_block);
	PlanArena A;
	A.cap = P.max_vertices;
	const size_t per = (size_t) slots * 2 * A.cap;
	void *mem = nullptr;
	int *counts = nullptr;
	cudaError_t e;
	const size_t bytes = per * (8 + 10 + 1 + 1) * sizeof(double) + per * sizeof(int) + (size_t) slots * 2 * sizeof(int);
	if ((e = cudaMallocAsync(&mem, bytes, st)) != cudaSuccess) { err = std::string("plan arena: ") + cudaGetErrorString(e); return GBP_E_CUDA; }
	A.v = (double *) mem;
	A.act = A.v + per * 8;
	A.g = A.act + per * 10;
	A.y = A.g + per;
	A.parent = (int *) (A.y + per);
	counts = A.parent + per;
#define GBP_PLAN_(M) k_plan_batch<M><<<grid, threads, 0, st>>>(Tv, nq, starts, goals, seed, query0, P, A, counts, stats, path_states, path_actions, path_cap)
	if (Tv.cell_f32) { if (Tv.uniform) GBP_PLAN_(MapF32U); else GBP_PLAN_(MapF32N); }
	else { if (Tv.uniform) GBP_PLAN_(MapF64U); else GBP_PLAN_(MapF64N); }
#undef GBP_PLAN_
	e = cudaGetLastError();
	cudaFreeAsync(mem, st);
	if (e != cudaSuccess) { err = std::string("k_plan_batch: ") + cudaGetErrorString(e); return GBP_E_CUDA; }
	return GBP_OK;
}

}  // namespace gbp
