// Device-resident batch planner: one WARP per planning query runs the whole bidirectional loop
// (runRRTConnect, rrt_connect.cpp:230-314; RRT*-Connect main loop, rrt_star_connect.cpp:130-165) with an
// iteration budget instead of the wall clock.  Queries are independent: no inter-warp or
// inter-GPU traffic.
//   nearest neighbour : lanes stride over the query's SoA tree, warp-shuffle argmin
//   newConfig         : lane j validates candidate j (ACTION cell (2*iter+half)*K + j), ballot / argmin
//                       selection (first valid in stream order, or closest valid)
//   connect           : closed-form action + pair check, evaluated warp-uniformly
//   RRT* extend       : lane per near vertex runs both attemptConnect probes (geometry only), then lane 0
//                       replays parent choice and rewiring in ascending id order (rrt_star_connect.cpp:12-75)
//   postProcessPath   : lane per later path state probes attemptConnect, farthest REACHED wins
// Tree arenas live in HBM, one slot per resident warp, reused across the queries a warp processes.
#pragma once
#include <cstdlib>
#include <cstring>
#include <string>

#include "gbp_kernels.cuh"

namespace gbp {

struct PlanArena {
	int cap;
	double *v, *act, *g, *y;         // per slot: 2 trees x cap x {8, 10, 1, 1} doubles
	int *parent, *child, *sibling;   // per slot: 2 trees x cap (child / sibling lists replace GraphClass::successors)
	// scratch per slot: RRT* near set (cap entries) and the stitched path (2*cap entries)
	int *near_id, *near_in, *near_out;   // [cap]: vertex id, attemptConnect status near->new and new->near
	double *near_ain, *near_aout;        // [cap][10]: the two connect actions
	double *pstate, *paction;            // [2*cap][8], [2*cap][10]
};

struct PlanTree {
	TreeView t;
	int *child, *sibling;
};
// optional copy of every query's two trees (inspection / parity tests): AoS rows per (query, tree, vertex)
struct PlanTreeDump {
	int cap;           // vertices per tree in the dump
	double *states;    // [nq][2][cap][8]
	double *actions;   // [nq][2][cap][10]
	int *parent;       // [nq][2][cap]
	double *g, *y;     // [nq][2][cap]
};

__device__ __forceinline__ PlanTree arena_tree(const PlanArena &A, int slot, int which, int *n_ptr) {
	PlanTree T;
	const size_t t = (size_t) slot * 2 + which;
	T.t.cap = A.cap;
	T.t.n = n_ptr;
	T.t.v = A.v + t * 8 * A.cap;
	T.t.act = A.act + t * 10 * A.cap;
	T.t.parent = A.parent + t * A.cap;
	T.t.g = A.g + t * A.cap;
	T.t.y = A.y + t * A.cap;
	T.child = A.child + t * A.cap;
	T.sibling = A.sibling + t * A.cap;
	return T;
}
__device__ __forceinline__ void plan_tree_init(PlanTree &T, const double root[8]) {  // GraphClass::init (graph_class.cpp:140-152)
	*T.t.n = 1;
	for (int d = 0; d < 8; ++d) T.t.v[(size_t) d * T.t.cap] = root[d];
	for (int d = 0; d < 10; ++d) T.t.act[(size_t) d * T.t.cap] = 0;
	T.t.parent[0] = -1; T.child[0] = -1; T.sibling[0] = -1; T.t.g[0] = 0; T.t.y[0] = 0;
}
__device__ __forceinline__ void plan_link(PlanTree &T, int p, int c) { T.t.parent[c] = p; T.sibling[c] = T.child[p]; T.child[p] = c; }
__device__ __forceinline__ void plan_unlink(PlanTree &T, int p, int c) {  // graph_class.cpp:44-59
	int *it = &T.child[p];
	while (*it != -1 && *it != c) it = &T.sibling[*it];
	if (*it == c) *it = T.sibling[c];
	T.sibling[c] = -1;
}
__device__ __forceinline__ int plan_push(PlanTree &T, int parent, const double s[8], const double a[10]) {
	const int i = tree_push(T.t, parent, s, a);
	T.child[i] = -1; T.sibling[i] = -1;
	plan_link(T, parent, i);
	return i;
}
// updateGYValue (graph_class.cpp:131-138): set (g, y) of `root` and refresh its subtree, stack-free pre-order walk
static __device__ void plan_update_gy(PlanTree &T, int root, double g, double y) {
	T.t.g[root] = g;
	T.t.y[root] = y;
	int i = T.child[root];
	while (i != -1) {
		const int p = T.t.parent[i];
		double a[8], b[8];
		tree_get(T.t, p, a);
		tree_get(T.t, i, b);
		T.t.g[i] = T.t.g[p] + pose_distance(a, b);
		T.t.y[i] = T.t.y[p] + yaw_distance(a, b);
		if (T.child[i] != -1) { i = T.child[i]; continue; }
		while (i != root && T.sibling[i] == -1) i = T.t.parent[i];
		if (i == root) break;
		i = T.sibling[i];
	}
}

__device__ __forceinline__ int warp_nearest(const TreeView &T, int nv, const double q[8], int lane) {
	double bd = INFINITY;
	int bi = 0x7fffffff;
	for (int j = lane; j < nv; j += 32) argmin_combine(bd, bi, vertex_distance(T, j, q), j);
	warp_argmin(bd, bi);
	return bi == 0x7fffffff ? 0 : bi;
}

// Only the VERDICT of a candidate matters to newConfig (s_new / t_new of rejected actions are never used), so the
// fixed-step pair check can be speculated across sub-states: the 32 lanes are split into groups of S = 32 / K lanes
// (K = 6 -> 5 lanes per candidate, 30 of 32 lanes busy); lane r of a group evaluates sub-state r of the group's
// candidate (the reference's fp64-accumulated sample times, by walking r steps from the group's cursor), a ballot
// finds each group's first failing lane, groups whose S sub-states all passed move on by S sub-states.
// Returns, per group, whether its candidate is fully valid; exact s_test comes from finish_output afterwards.
template <typename M>
__device__ __forceinline__ bool group_validate(const TerrainView &Tv, const double s_near[8], const double a[10], int dir, int S, int r,
											   unsigned gmask, int gshift, bool has_candidate) {
	Cursor q;
#pragma unroll
	for (int i = 0; i < 8; ++i) q.s[i] = s_near[i];
#pragma unroll
	for (int i = 0; i < 10; ++i) q.a[i] = a[i];
	cursor_start(q, dir);
	const double ts = a[6], tf = a[7];
	int bph = q.phase;
	double bt = q.t;
	int state = has_candidate ? 0 : 1;  // 0 running, 1 invalid / no candidate, 2 valid
	while (__ballot_sync(FULL, state == 0)) {
		int ph = bph;
		double t = bt;
		for (int i = 0; i < r && ph != PH_DONE; ++i) walk_step(ph, t, ts, tf);
		const bool active = state == 0 && ph != PH_DONE;
		bool valid = true;
		if (active) {
			q.phase = ph;
			q.t = t;
			valid = cursor_check<M>(Tv, q);
		}
		const unsigned bad = (__ballot_sync(FULL, active && !valid) >> gshift) & gmask;
		const unsigned term = (__ballot_sync(FULL, active && (ph == PH_FWD_LAND || ph == PH_REV_START)) >> gshift) & gmask;
		// the cursor of the group's last lane, one step further, is the next round's base
		int lph = __shfl_sync(FULL, ph, gshift + S - 1);
		double lt = __shfl_sync(FULL, t, gshift + S - 1);
		if (state == 0) {
			if (bad) state = 1;            // some sub-state on the reference's path failed (earlier ones all passed or are moot)
			else if (term) state = 2;      // landing / exact start state reached with every sub-state valid
			else { walk_step(lph, lt, ts, tf); bph = lph; bt = lt; }
		}
	}
	return state == 2;
}

// How the 32 lanes split over the K candidates of an extend (fixed for a launch, computed once per kernel: the runtime
// integer divisions are out-of-line subroutines, and far jumps are what this kernel stalls on).
struct GroupMap { int S, G, g, r, gshift; unsigned gmask; };
__device__ __forceinline__ GroupMap make_group_map(int K, int lane) {
	GroupMap m;
	m.S = K >= 32 ? 1 : 32 / K;
	m.G = 32 / m.S;
	m.g = lane / m.S;
	m.r = lane - m.g * m.S;
	m.gshift = m.g * m.S;
	m.gmask = m.S == 32 ? FULL : ((1u << m.S) - 1u);
	return m;
}

// newConfig (rrt.cpp:20-70) generalised to K candidates; uniform outputs.  Returns found.
template <typename M, bool WIDE = false>
__device__ bool warp_new_config(const TerrainView &Tv, const double s[8], const double s_near[8], int dir, uint64_t seed, uint64_t query,
								uint64_t cell, const gbp_plan_params &P, const GroupMap &gm, int lane, long long &pair_checks, double s_new[8], double a_new[10]) {
	double nn[3], R[9];
	const double best0 = state_distance(s_near, s);
	unsigned fl = 0;
	surface_normal(Tv, s[0], s[1], nn, fl);  // rrt.cpp:25 — at the TARGET sample
	grf_rotation(nn, R);
	const int K = P.k_candidates;
	// getRandomAction(surf_norm, direction, flag, threshold, s, s_near) (rrt.cpp:34, :49): FORWARD samples from s_near
	// towards s, REVERSE from s towards s_near (planning_utils.cpp:385-388)
	const bool dirs = P.action_direction_sampling != 0;
	const double *a_from = dir == GBP_FORWARD ? s_near : s, *a_to = dir == GBP_FORWARD ? s : s_near;
	double my_d = INFINITY, my_sn[8], my_a[10];
	int my_j = 0x7fffffff, first = 0x7fffffff;
	if (WIDE) {
		// Closest valid of MANY candidates at the fixed step (configs[1]: K = 4096; the launcher picks this instantiation
		// for best_of_k, K > 32, no adaptive step): lane per candidate with warp-level REFILL, as in the walk kernels — a
		// lane whose candidate has failed (k = 2-4 sub-states on average) or landed takes the next candidate index instead
		// of idling until the slowest lane of a batch of 32 (up to 19 sub-states) is through.  Idle lanes are refilled once
		// a quarter of the warp is idle, so that sampling stays mostly convergent.  Every candidate is evaluated exactly as
		// in the batched form, and the (distance, index) argmin below does not depend on which lane evaluated which
		// candidate: results are identical.  A separate instantiation because inlined next to the K = 6 path it cost that
		// path 6.5 % (instruction fetch) and out of line 40 % (stack).
		Cursor q;
#pragma unroll
		for (int i = 0; i < 8; ++i) q.s[i] = s_near[i];
		int next = 0, j = -1;
		bool running = false;
		while (true) {
			const unsigned idle = __ballot_sync(FULL, !running);
			if (next < K && (__popc(idle) >= 8 || idle == FULL)) {
				if (!running) {
					const int rel = next + __popc(idle & ((1u << lane) - 1u));
					if (rel < K) {
						j = rel;
						sample_action(seed, query, cell * (uint64_t) K + (uint64_t) j, R, dirs, P.action_direction_threshold, a_from, a_to, q.a);
						cursor_start(q, dir);
						running = true;
					}
				}
				next = min(K, next + __popc(idle));
			}
			if (__ballot_sync(FULL, running) == 0) break;
			if (running) {
				const bool valid = cursor_check<M>(Tv, q);
				if (!valid) running = false;
				else if (q.phase == PH_FWD_LAND || q.phase == PH_REV_START) {  // every sub-state valid: exact end state, distance to the target
					double sn[8];
					finish_output(s_near, q.a, dir == GBP_FORWARD ? OUT_LAND : OUT_REV, 0.0, sn);
					const double d = state_distance(sn, s);
					if (d < my_d) {
						my_d = d; my_j = j;
#pragma unroll
						for (int i = 0; i < 8; ++i) my_sn[i] = sn[i];
#pragma unroll
						for (int i = 0; i < 10; ++i) my_a[i] = q.a[i];
					}
					running = false;
				} else walk_step(q.phase, q.t, q.a[6], q.a[7]);
			}
		}
	} else if (!P.adaptive) {
		// speculative form: G candidates per pass, S lanes each
		const int S = gm.S, G = gm.G, g = gm.g, r = gm.r, gshift = gm.gshift;
		const unsigned gmask = gm.gmask;
		for (int base = 0; base < K; base += G) {
			const int j = base + g;
			const bool has = g < G && j < K;
			double a[10];
			sample_action(seed, query, cell * (uint64_t) K + (uint64_t) (has ? j : 0), R, dirs, P.action_direction_threshold, a_from, a_to, a);
			const bool ok = group_validate<M>(Tv, s_near, a, dir, S, r, gmask, has ? gshift : 0, has);
			if (ok && r == 0) {  // one lane per candidate keeps the result
				double sn[8];
				finish_output(s_near, a, dir == GBP_FORWARD ? OUT_LAND : OUT_REV, 0.0, sn);
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
				const unsigned m = __ballot_sync(FULL, ok && r == 0);
				if (m) { first = base + __shfl_sync(FULL, g, __ffs(m) - 1); break; }  // first valid action decides (rrt.cpp:44-47): the group of the lowest set lane
			}
		}
	} else {
		for (int base = 0; base < K; base += 32) {
			const int j = base + lane;
			bool ok = false;
			double a[10], sn[8], tn;
			if (j < K) {
				sample_action(seed, query, cell * (uint64_t) K + (uint64_t) j, R, dirs, P.action_direction_threshold, a_from, a_to, a);
				Counters c = {0, 0, 0, 0};
				ok = validate_pair_seq<M>(Tv, s_near, a, dir, true, sn, tn, c);
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
				if (m) { first = base + __ffs(m) - 1; break; }
			}
		}
	}
	int src_lane;
	double d_sel;
	if (P.best_of_k) {
		pair_checks += K;
		double bd = my_d;
		int bj = my_j;
		// argmin over (distance, candidate index); remember which lane holds the winner
		int bl = lane;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const double od = __shfl_xor_sync(FULL, bd, o);
			const int oj = __shfl_xor_sync(FULL, bj, o), ol = __shfl_xor_sync(FULL, bl, o);
			if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; bl = ol; }
		}
		if (bj == 0x7fffffff) return false;
		src_lane = bl;
		d_sel = bd;
	} else {
		pair_checks += (first == 0x7fffffff) ? K : first + 1;
		if (first == 0x7fffffff) return false;
		// the lane that kept candidate `first`
		const unsigned holder = __ballot_sync(FULL, my_j == first);
		src_lane = __ffs(holder) - 1;
		d_sel = __shfl_sync(FULL, my_d, src_lane);
	}
	if (!(d_sel < best0)) return false;  // rrt.cpp:55-66
#pragma unroll
	for (int i = 0; i < 8; ++i) s_new[i] = __shfl_sync(FULL, my_sn[i], src_lane);
#pragma unroll
	for (int i = 0; i < 10; ++i) a_new[i] = __shfl_sync(FULL, my_a[i], src_lane);
	return true;
}

// RRTClass::extend (rrt.cpp:77-102)
template <typename M, bool WIDE = false>
__device__ int warp_extend(const TerrainView &Tv, PlanTree &T, int &nv, const double s[8], int dir, uint64_t seed, uint64_t query,
						   uint64_t cell, const gbp_plan_params &P, const GroupMap &gm, int lane, long long &pair_checks) {
	const int near = warp_nearest(T.t, nv, s, lane);
	double s_near[8], sn[8], a[10];
	tree_get(T.t, near, s_near);
	if (!warp_new_config<M, WIDE>(Tv, s, s_near, dir, seed, query, cell, P, gm, lane, pair_checks, sn, a)) return GBP_TRAPPED;
	if (lane == 0) plan_push(T, near, sn, a);
	__syncwarp();
	nv += 1;
	return state_distance(sn, s) <= GOAL_BOUNDS ? GBP_REACHED : GBP_ADVANCED;
}

// RRTStarConnectClass::extend (rrt_star_connect.cpp:12-75)
template <typename M>
__device__ int warp_extend_star(const TerrainView &Tv, PlanTree &T, int &nv, const double s[8], int dir, uint64_t seed, uint64_t query,
								uint64_t cell, const gbp_plan_params &P, const GroupMap &gm, const PlanArena &A, int slot, int lane, long long &pair_checks) {
	const int nearest = warp_nearest(T.t, nv, s, lane);
	double s_nearest[8], s_new[8], a_new[10];
	tree_get(T.t, nearest, s_nearest);
	if (!warp_new_config<M>(Tv, s, s_nearest, dir, seed, query, cell, P, gm, lane, pair_checks, s_new, a_new)) return GBP_TRAPPED;
	const int id = nv;
	if (lane == 0) {  // addVertex (:22)
		*T.t.n = id + 1;
		for (int d = 0; d < 8; ++d) T.t.v[(size_t) d * T.t.cap + id] = s_new[d];
		T.t.parent[id] = -1; T.child[id] = -1; T.sibling[id] = -1;
	}
	__syncwarp();
	nv += 1;
	// near set in ascending id (neighborhoodDist, planner_class.cpp:173-182); probe near -> new for every member
	int *nid = A.near_id + (size_t) slot * A.cap, *nin = A.near_in + (size_t) slot * A.cap, *nout = A.near_out + (size_t) slot * A.cap;
	double *ain = A.near_ain + (size_t) slot * A.cap * 10, *aout = A.near_aout + (size_t) slot * A.cap * 10;
	int count = 0;
	unsigned checks = 0;
	for (int j0 = 0; j0 < nv; j0 += 32) {
		const int j = j0 + lane;
		bool in = false;
		double sj[8];
		if (j < nv) {
			tree_get(T.t, j, sj);
			const double d = state_distance(s_new, sj);
			in = (d <= RRT_STAR_DELTA) && (d > 0);
		}
		const unsigned m = __ballot_sync(FULL, in);
		if (in) {
			const int pos = count + __popc(m & ((1u << lane) - 1));
			double dummy[8], ac[10];
			Counters c = {0, 0, 0, 0};
			nid[pos] = j;
			nin[pos] = attempt_connect<M>(Tv, sj, s_new, dir, P.adaptive != 0, dummy, ac, c, checks);
			for (int d = 0; d < 10; ++d) ain[(size_t) pos * 10 + d] = ac[d];
		}
		count += __popc(m);
	}
	__syncwarp();
	int parent = nearest;
	if (lane == 0) {  // choose parent (:30-44), link the new vertex
		double g_new = T.t.g[nearest] + pose_distance(s_new, s_nearest), y_new = T.t.y[nearest] + yaw_distance(s_new, s_nearest);
		for (int i = 0; i < count; ++i) {
			if (nin[i] != GBP_REACHED) continue;
			double sj[8];
			tree_get(T.t, nid[i], sj);
			const double g = T.t.g[nid[i]] + pose_distance(sj, s_new);
			if (g < g_new) {
				for (int d = 0; d < 10; ++d) a_new[d] = ain[(size_t) i * 10 + d];
				parent = nid[i];
				g_new = g;
				y_new = T.t.y[nid[i]] + yaw_distance(sj, s_new);
			}
		}
		plan_link(T, parent, id);
		plan_update_gy(T, id, g_new, y_new);
		for (int d = 0; d < 10; ++d) T.t.act[(size_t) d * T.t.cap + id] = a_new[d];
	}
	parent = __shfl_sync(FULL, parent, 0);
	// probe new -> near for every member but the parent (:51-53), then rewire in id order
	for (int i0 = 0; i0 < count; i0 += 32) {
		const int i = i0 + lane;
		if (i < count && nid[i] != parent) {
			double sj[8], dummy[8], ac[10];
			Counters c = {0, 0, 0, 0};
			tree_get(T.t, nid[i], sj);
			nout[i] = attempt_connect<M>(Tv, s_new, sj, dir, P.adaptive != 0, dummy, ac, c, checks);
			for (int d = 0; d < 10; ++d) aout[(size_t) i * 10 + d] = ac[d];
		}
	}
	pair_checks += __reduce_add_sync(FULL, checks);
	__syncwarp();
	if (lane == 0) {
		for (int i = 0; i < count; ++i) {  // rewire (:50-64)
			const int k = nid[i];
			if (k == parent || nout[i] != GBP_REACHED) continue;
			double sj[8];
			tree_get(T.t, k, sj);
			const double through = T.t.g[id] + pose_distance(sj, s_new);
			if (T.t.g[k] > through) {
				plan_unlink(T, T.t.parent[k], k);
				plan_link(T, id, k);
				plan_update_gy(T, k, through, T.t.y[id] + yaw_distance(sj, s_new));
				for (int d = 0; d < 10; ++d) T.t.act[(size_t) d * T.t.cap + k] = aout[(size_t) i * 10 + d];
			}
		}
	}
	__syncwarp();
	return state_distance(s_new, s) <= GOAL_BOUNDS ? GBP_REACHED : GBP_ADVANCED;
}

// connect (rrt_connect.cpp:98-120)
template <typename M>
__device__ int warp_connect(const TerrainView &Tv, PlanTree &T, int &nv, const double s[8], int dir, const gbp_plan_params &P, int lane,
							long long &pair_checks) {
	const int near = warp_nearest(T.t, nv, s, lane);
	double s_near[8], sn[8], an[10];
	tree_get(T.t, near, s_near);
	Counters c = {0, 0, 0, 0};
	unsigned checks = 0;
	// fixed step: the connect primitive's sub-states 32 at a time over the lanes instead of one after the other on all of them
	const int r = P.adaptive ? attempt_connect<M>(Tv, s_near, s, dir, true, sn, an, c, checks)
							 : attempt_connect_warp<M>(Tv, s_near, s, dir, sn, an, c, checks);
	pair_checks += checks;
	if (r != GBP_TRAPPED) {
		if (lane == 0) plan_push(T, near, sn, an);
		__syncwarp();
		nv += 1;
	}
	return r;
}

// postProcessPath (rrt_connect.cpp:139-227) on a stitched path in scratch, in place.  The reference probes
// attemptConnect(s, s_next) from the last state backwards and takes the first REACHED; lanes probe 32 later states
// at a time from the far end, the farthest REACHED one wins — the same choice.  Quirk kept: the fallback branch adds to
// path_cost only.  Returns the new number of states; stats3 = {length, yaw, cost}.
template <typename M>
__device__ int warp_post_process(const TerrainView &Tv, double *ps, double *pa, int ns, bool adaptive, int lane, const gbp_plan_params &P, double stats3[3]) {
	int m = 1, cur = 0;
	double len = 0, yaw = 0, cost = 0;
	while (cur < ns - 1) {
		double sc[8];
		for (int d = 0; d < 8; ++d) sc[d] = ps[8 * (size_t) cur + d];
		int pick = -1;
		double a_pick[10];
		for (int hi = ns - 1; hi > cur && pick < 0; hi -= 32) {
			const int j = hi - lane;
			int st = GBP_TRAPPED;
			double an[10];
			if (j > cur) {
				double sj[8], dummy[8];
				for (int d = 0; d < 8; ++d) sj[d] = ps[8 * (size_t) j + d];
				Counters c = {0, 0, 0, 0};
				unsigned checks = 0;
				st = attempt_connect<M>(Tv, sc, sj, GBP_FORWARD, adaptive, dummy, an, c, checks);
			}
			const unsigned reached = __ballot_sync(FULL, st == GBP_REACHED);
			if (reached) {
				const int src = __ffs(reached) - 1;  // lowest lane = farthest state
				pick = hi - src;
#pragma unroll
				for (int d = 0; d < 10; ++d) a_pick[d] = __shfl_sync(FULL, an[d], src);
			}
		}
		const int nxt = pick >= 0 ? pick : cur + 1;
		double sn[8];
		for (int d = 0; d < 8; ++d) sn[d] = ps[8 * (size_t) nxt + d];
		const double dl = pose_distance(sc, sn), dy = yaw_distance(sc, sn);
		if (pick >= 0) { len += dl; yaw += dy; }
		else { for (int d = 0; d < 10; ++d) a_pick[d] = pa[10 * (size_t) cur + d]; }  // the original action into state cur+1
		if (P.cost_add_yaw) cost += dl * P.cost_length_weight + dy * P.cost_yaw_weight; else cost += dl;  // rrt_connect.cpp:196, :212
		__syncwarp();
		if (lane == 0) {
			for (int d = 0; d < 8; ++d) ps[8 * (size_t) m + d] = sn[d];
			for (int d = 0; d < 10; ++d) pa[10 * (size_t) (m - 1) + d] = a_pick[d];
		}
		__syncwarp();
		++m;
		cur = nxt;
	}
	stats3[0] = len; stats3[1] = yaw; stats3[2] = cost;
	return m;
}

// End of a query: statistics (rrt_connect.cpp:269-270, :463-466), the stitched path (:381-401), postProcessPath, and the
// optional copies of the path and of both trees.  Shared by the megakernel and the stepped planner; all lanes call it.
template <typename M>
__device__ __forceinline__ void plan_finish(const TerrainView &Tv, const gbp_plan_params &P, const PlanArena &A, int64_t scratch_slot, const PlanTree &Ta,
											 const PlanTree &Tb, int na, int nb, bool solved, int it, long long pair_checks, long long nn_queries, int64_t qi,
											 gbp_plan_stats *__restrict__ stats, double *__restrict__ path_states, double *__restrict__ path_actions,
											 int path_cap, const PlanTreeDump &dump, int lane) {
		gbp_plan_stats st;
	st.solved = solved ? 1 : 0; st.iters = it; st.nv_a = na; st.nv_b = nb; st.path_states = 0; st.pad = 0;
	st.path_length = 0; st.path_yaw = 0; st.path_duration = 0; st.pair_checks = pair_checks; st.nn_queries = nn_queries;
	st.path_cost = 0; st.reserved = 0;
	if (solved) {
		double *ps = A.pstate + (size_t) scratch_slot * 2 * A.cap * 8, *pa = A.paction + (size_t) scratch_slot * 2 * A.cap * 10;
		int la = 0, lb = 0;
		for (int i = na - 1; i != -1; i = Ta.t.parent[i]) ++la;
		for (int i = nb - 1; i != -1; i = Tb.t.parent[i]) ++lb;
		int total = la + lb - 1;
		if (lane == 0) {  // stitch: start .. shared state (tree A), then tree B back to the goal
			int k = la - 1;
			for (int i = na - 1; i != -1; i = Ta.t.parent[i], --k) {
				for (int d = 0; d < 8; ++d) ps[8 * (size_t) k + d] = Ta.t.v[(size_t) d * Ta.t.cap + i];
				if (k > 0) for (int d = 0; d < 10; ++d) pa[10 * (size_t) (k - 1) + d] = Ta.t.act[(size_t) d * Ta.t.cap + i];
			}
			k = la - 1;  // Tb.last duplicates the shared state: its ACTION is kept, its STATE is dropped (:388-395)
			for (int i = nb - 1; Tb.t.parent[i] != -1; i = Tb.t.parent[i], ++k) {
				for (int d = 0; d < 10; ++d) pa[10 * (size_t) k + d] = Tb.t.act[(size_t) d * Tb.t.cap + i];
				for (int d = 0; d < 8; ++d) ps[8 * (size_t) (k + 1) + d] = Tb.t.v[(size_t) d * Tb.t.cap + Tb.t.parent[i]];
			}
		}
		__syncwarp();
		st.path_length = Ta.t.g[na - 1] + Tb.t.g[nb - 1];
		st.path_yaw = Ta.t.y[na - 1] + Tb.t.y[nb - 1];
		st.path_cost = P.cost_add_yaw ? st.path_length * P.cost_length_weight + st.path_yaw * P.cost_yaw_weight : st.path_length;  // :270-274
		if (P.post_process) {
			double s3[3];
			total = warp_post_process<M>(Tv, ps, pa, total, P.adaptive != 0, lane, P, s3);
			st.path_length = s3[0];
			st.path_yaw = s3[1];
			st.path_cost = s3[2];
		}
		st.path_states = total;
		double dur = 0;
		for (int i = 0; i + 1 < total; ++i) dur += pa[10 * (size_t) i + 6] + pa[10 * (size_t) i + 7];
		st.path_duration = dur;
		if (path_states && path_actions) {
			for (int i = lane; i < total && i < path_cap; i += 32)
				for (int d = 0; d < 8; ++d) path_states[((size_t) qi * path_cap + i) * 8 + d] = ps[8 * (size_t) i + d];
			for (int i = lane; i + 1 < total && i < path_cap; i += 32)
				for (int d = 0; d < 10; ++d) path_actions[((size_t) qi * path_cap + i) * 10 + d] = pa[10 * (size_t) i + d];
		}
	}
	if (dump.states) {  // both trees, vertex by vertex
		for (int w = 0; w < 2; ++w) {
			const PlanTree &Tw = w == 0 ? Ta : Tb;
			const int nw = w == 0 ? na : nb;
			const size_t base = ((size_t) qi * 2 + w) * dump.cap;
			for (int i = lane; i < nw && i < dump.cap; i += 32) {
				for (int d = 0; d < 8; ++d) dump.states[(base + i) * 8 + d] = Tw.t.v[(size_t) d * Tw.t.cap + i];
				for (int d = 0; d < 10; ++d) dump.actions[(base + i) * 10 + d] = Tw.t.act[(size_t) d * Tw.t.cap + i];
				dump.parent[base + i] = Tw.t.parent[i];
				dump.g[base + i] = Tw.t.g[i];
				dump.y[base + i] = Tw.t.y[i];
			}
		}
	}
	if (lane == 0) stats[qi] = st;
}

template <typename M, bool STAR, bool WIDE = false>
// Occupancy over registers: the kernel is 21 k SASS instructions and its warps sit at unrelated program counters, so
// at 255 registers (8 warps / SM) ncu shows 8.3 of the 12.2 cycles between two issues of a warp waiting for
// instruction fetch (profiles/r1b_planner_2ctas_ncu_summary.csv).  Capping the registers at 80 (24 warps / SM, ~2.7 KB of
// spills per thread, L1-resident) hides that latency: 9.2 k -> 15.8 k plans/s (4 CTAs / SM: 12.3 k, 8: 16.2 k).
#ifndef GBP_PLAN_MINBLOCKS
#define GBP_PLAN_MINBLOCKS 6
#endif
__global__ void __launch_bounds__(128, GBP_PLAN_MINBLOCKS) k_plan_batch(TerrainView Tv, int64_t nq, const double *__restrict__ starts,
													 const double *__restrict__ goals, uint64_t seed, uint64_t query0,
													 gbp_plan_params P, PlanArena A, int *__restrict__ counts,
													 gbp_plan_stats *__restrict__ stats, double *__restrict__ path_states,
													 double *__restrict__ path_actions, int path_cap, PlanTreeDump dump) {
	const int lane = threadIdx.x & 31;
	const int64_t slot = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
	unsigned long long *next_query = (unsigned long long *) (counts + 2 * (((int64_t) gridDim.x * blockDim.x) >> 5));
	volatile int *solved_count = (volatile int *) (next_query + 1);  // anytime use: queries solved so far in this launch
	const GroupMap gm = make_group_map(P.k_candidates, lane);
	while (true) {
		// queries differ widely in iterations: warps pull the next query from a device counter (no static round-robin tail)
		unsigned long long grabbed = 0;
		if (lane == 0) grabbed = atomicAdd(next_query, 1ull);
		const int64_t qi = (int64_t) __shfl_sync(FULL, grabbed, 0);
		if (qi >= nq) break;
		if (P.stop_after_solved > 0 && __shfl_sync(FULL, *solved_count, 0) >= P.stop_after_solved) {  // enough solved: skip the rest
			if (lane == 0) { gbp_plan_stats z = {}; stats[qi] = z; }
			continue;
		}
		PlanTree Ta = arena_tree(A, (int) slot, 0, counts + 2 * slot), Tb = arena_tree(A, (int) slot, 1, counts + 2 * slot + 1);
		double start[8], goal[8];
#pragma unroll
		for (int d = 0; d < 8; ++d) { start[d] = starts[8 * qi + d]; goal[d] = goals[8 * qi + d]; }
		if (lane == 0) { plan_tree_init(Ta, start); plan_tree_init(Tb, goal); }
		__syncwarp();
		int na = 1, nb = 1, it = 0;
		bool solved = false, full = false;
		double rs[8];          // this lane's random state of the current batch of 32 STATE cells
		unsigned rs_valid = 0;  // isValidState(STANCE) of the 32 states
		long long pair_checks = 0, nn_queries = 0;
		const uint64_t query = query0 + (uint64_t) qi;
		for (; it < P.max_iters && !solved && !full; ++it) {
			if (P.stop_after_solved > 0 && __shfl_sync(FULL, *solved_count, 0) >= P.stop_after_solved) break;
			for (int half = 0; half < 2 && !solved; ++half) {
				PlanTree &Tx = half == 0 ? Ta : Tb, &Ty = half == 0 ? Tb : Ta;
				int &nx = half == 0 ? na : nb, &ny = half == 0 ? nb : na;
				const int dir_ext = half == 0 ? GBP_FORWARD : GBP_REVERSE, dir_con = half == 0 ? GBP_REVERSE : GBP_FORWARD;
				if (nx >= A.cap || ny >= A.cap) { full = true; break; }
				const uint64_t cell = 2 * (uint64_t) it + (uint64_t) half;
				// Random states come from a counter-based stream (STATE cell = 2 * iter + half) and their validity does not
				// depend on the trees: the warp draws and checks the next 32 cells at once, lane L holding cell base + L,
				// instead of all 32 lanes redundantly producing one (that was ~75 % of the kernel's instructions).
				double s_rand[8];
				if (P.state_direction_sampling) {
					// randomState(terrain, flag, threshold, speed flag, s_from, s_to) between the start-side tree's newest vertex /
					// root and the goal-side tree's root / newest vertex (rrt_connect.cpp:246-251, :281-286): the sample depends on
					// the trees, so it is drawn when it is needed (every lane draws the same cell)
					double s_from[8], s_to[8];
					tree_get(Ta.t, half == 0 ? na - 1 : 0, s_from);
					tree_get(Tb.t, half == 0 ? 0 : nb - 1, s_to);
					sample_state<M>(Tv, seed, query, cell, true, P.state_direction_threshold, P.state_direction_speed != 0, s_from, s_to, s_rand);
					Counters c = {0, 0, 0, 0};
					if (!is_valid_state_auto<M>(Tv, pose6(s_rand), GBP_STANCE, c)) continue;  // rrt_connect.cpp:254
				} else {
					if ((cell & 31ull) == 0) {
						sample_state<M>(Tv, seed, query, cell + (uint64_t) lane, false, 0.0, false, nullptr, nullptr, rs);
						Counters c = {0, 0, 0, 0};
						rs_valid = __ballot_sync(FULL, is_valid_state_auto<M>(Tv, pose6(rs), GBP_STANCE, c));
					}
					const int src = (int) (cell & 31ull);
					if (!((rs_valid >> src) & 1u)) continue;  // rrt_connect.cpp:254
#pragma unroll
					for (int d = 0; d < 8; ++d) s_rand[d] = __shfl_sync(FULL, rs[d], src);
				}
				++nn_queries;
				const int r = STAR ? warp_extend_star<M>(Tv, Tx, nx, s_rand, dir_ext, seed, query, cell, P, gm, A, (int) slot, lane, pair_checks)
								   : warp_extend<M, WIDE>(Tv, Tx, nx, s_rand, dir_ext, seed, query, cell, P, gm, lane, pair_checks);
				if (r == GBP_TRAPPED) continue;
				double s_new[8];
				tree_get(Tx.t, nx - 1, s_new);
				++nn_queries;
				if (warp_connect<M>(Tv, Ty, ny, s_new, dir_con, P, lane, pair_checks) == GBP_REACHED) solved = true;
			}
		}
		plan_finish<M>(Tv, P, A, slot, Ta, Tb, na, nb, solved, it, pair_checks, nn_queries, qi, stats, path_states, path_actions, path_cap, dump, lane);
		if (lane == 0 && solved && P.stop_after_solved > 0) atomicAdd((int *) solved_count, 1);
		__syncwarp();
	}
}

// host-side launcher: sizes the grid to the SM count, allocates the tree arena for the resident warps
// `arena` / `arena_bytes`: grow-only device scratch owned by the caller's terrain handle
// the tree arena is stream-ordered scratch of the call (cudaMallocAsync / cudaFreeAsync on `st`): concurrent calls on one
// terrain handle from different streams or threads do not share it
inline int plan_batch_launch(const TerrainView &Tv, int64_t nq, const double *starts, const double *goals, uint64_t seed,
							 uint64_t query0, const gbp_plan_params &P, gbp_plan_stats *stats, double *path_states, double *path_actions,
							 int path_cap, cudaStream_t st, const PlanTreeDump &dump, std::string &err) {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int threads = 128, warps_per_block = threads / 32;
	int64_t slots = (int64_t) sms * 4 * GBP_PLAN_MINBLOCKS;  // resident warps (register-limited)
	if (slots > nq) slots = ((nq + warps_per_block - 1) / warps_per_block) * warps_per_block;
	const unsigned grid = (unsigned) (slots / warps_per_block);
	PlanArena A;
	A.cap = P.max_vertices;
	const size_t cap = (size_t) A.cap, per = (size_t) slots * 2 * cap;
	const size_t n_doubles = per * (8 + 10 + 1 + 1) + (size_t) slots * cap * 20 + (size_t) slots * 2 * cap * 18;
	const size_t n_ints = per * 3 + (size_t) slots * cap * 3 + (size_t) slots * 2 + 4;  // + the 8-byte work counter and the solved count
	cudaError_t e;
	const size_t need = n_doubles * sizeof(double) + n_ints * sizeof(int);
	void *mem = nullptr;
	if ((e = cudaMallocAsync(&mem, need, st)) != cudaSuccess) { err = std::string("plan arena: ") + cudaGetErrorString(e); return GBP_E_CUDA; }
	double *dp = (double *) mem;
	A.v = dp; dp += per * 8;
	A.act = dp; dp += per * 10;
	A.g = dp; dp += per;
	A.y = dp; dp += per;
	A.near_ain = dp; dp += (size_t) slots * cap * 10;
	A.near_aout = dp; dp += (size_t) slots * cap * 10;
	A.pstate = dp; dp += (size_t) slots * 2 * cap * 8;
	A.paction = dp; dp += (size_t) slots * 2 * cap * 10;
	int *ip = (int *) dp;
	A.parent = ip; ip += per;
	A.child = ip; ip += per;
	A.sibling = ip; ip += per;
	A.near_id = ip; ip += (size_t) slots * cap;
	A.near_in = ip; ip += (size_t) slots * cap;
	A.near_out = ip; ip += (size_t) slots * cap;
	int *counts = ip;
	{  // 8-byte aligned work counter right after the per-slot vertex counts
		unsigned long long *next_query = (unsigned long long *) (counts + 2 * slots);
		if ((e = cudaMemsetAsync(next_query, 0, sizeof(unsigned long long) + 2 * sizeof(int), st)) != cudaSuccess) { err = cudaGetErrorString(e); cudaFreeAsync(mem, st); return GBP_E_CUDA; }
	}
	const bool wide = !P.rrt_star && !P.adaptive && P.best_of_k && P.k_candidates > 32;  // the refill form of newConfig
#define GBP_PLAN_(M) do { if (P.rrt_star) k_plan_batch<M, true><<<grid, threads, 0, st>>>(Tv, nq, starts, goals, seed, query0, P, A, counts, stats, path_states, path_actions, path_cap, dump); \
						  else if (wide) k_plan_batch<M, false, true><<<grid, threads, 0, st>>>(Tv, nq, starts, goals, seed, query0, P, A, counts, stats, path_states, path_actions, path_cap, dump); \
						  else k_plan_batch<M, false><<<grid, threads, 0, st>>>(Tv, nq, starts, goals, seed, query0, P, A, counts, stats, path_states, path_actions, path_cap, dump); } while (0)
	if (Tv.cell_f32) { if (Tv.uniform) GBP_PLAN_(MapF32U); else GBP_PLAN_(MapF32N); }
	else { if (Tv.uniform) GBP_PLAN_(MapF64U); else GBP_PLAN_(MapF64N); }
#undef GBP_PLAN_
	e = cudaGetLastError();
	cudaFreeAsync(mem, st);  // stream-ordered: released after the kernel
	if (e != cudaSuccess) { err = std::string("k_plan_batch: ") + cudaGetErrorString(e); return GBP_E_CUDA; }
	return GBP_OK;
}

// ------------------------------------------------------------------ the stepped planner
// The same searches, split by role instead of run as one megakernel: ONE launch per half-iteration of runRRTConnect
// (rrt_connect.cpp:246-279 / :281-312) advances every query of the batch by that half — extend towards the query's random
// state and, when the tree grew, connect from the other tree — with the trees and a few words of per-query state in HBM
// between launches, and one launch per 16 iterations draws and validity-checks the next 32 STATE cells of every query.
// Why: k_plan_batch is a 21 k-instruction kernel whose warps sit at unrelated program counters (7.3 of 15 stall cycles per
// issue are instruction fetch, profiles/r1c_planner_ncu_summary.csv) and whose 80-register cap spills 2.7 KB per thread
// (23.5 GB of DRAM writes per 98 ms launch).  Here every resident warp of the chip runs the same ~3 k instructions at
// the same time and each kernel gets the registers it needs.  Arithmetic, Philox cells and the order of tree updates
// per query are the megakernel's, so trees, statistics and paths are bit-identical (tests/test_gpu_planner.py runs both).
// Used for large batches of plain RRT-Connect queries at the fixed step (the launcher decides); everything else — RRT*,
// adaptive step, directional state sampling, K > 32, anytime rounds — stays on the megakernel.
struct StepState {
	int *na, *nb;              // [Q] vertex counts (the trees' `n` words)
	int *status;               // [Q] 0 running, 1 solved, 2 a tree is full
	int *iters;                // [Q] started iterations when the query stopped (max_iters while running)
	long long *pair_checks, *nn_queries;  // [Q]
	double *rs;                // [Q][32][8] the 32 random states of the current batch of STATE cells
	unsigned *rs_valid;        // [Q] isValidState(STANCE) of those 32 states
};

template <typename M>
__global__ void __launch_bounds__(128) k_step_init(StepState S, PlanArena A, int64_t Q, const double *__restrict__ starts,
													const double *__restrict__ goals, int max_iters) {
	const int64_t q = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (q >= Q) return;
	PlanTree Ta = arena_tree(A, (int) q, 0, S.na + q), Tb = arena_tree(A, (int) q, 1, S.nb + q);
	double s[8], g[8];
#pragma unroll
	for (int d = 0; d < 8; ++d) { s[d] = starts[8 * q + d]; g[d] = goals[8 * q + d]; }
	plan_tree_init(Ta, s);
	plan_tree_init(Tb, g);
	S.status[q] = 0; S.iters[q] = max_iters; S.pair_checks[q] = 0; S.nn_queries[q] = 0; S.rs_valid[q] = 0;
}

// STATE cells base .. base + 31 of every running query: lane L draws cell base + L and checks it (their validity does not
// depend on the trees)
template <typename M>
__global__ void __launch_bounds__(128) k_step_sample(TerrainView Tv, StepState S, int64_t Q, uint64_t seed, uint64_t query0, uint64_t cell_base) {
	const int lane = threadIdx.x & 31;
	const int64_t q = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
	if (q >= Q || S.status[q] != 0) return;
	double rs[8];
	sample_state<M>(Tv, seed, query0 + (uint64_t) q, cell_base + (uint64_t) lane, false, 0.0, false, nullptr, nullptr, rs);
	Counters c = {0, 0, 0, 0};
	const unsigned valid = __ballot_sync(FULL, is_valid_state_auto<M>(Tv, pose6(rs), GBP_STANCE, c));
	double2 *o = reinterpret_cast<double2 *>(S.rs + ((size_t) q * 32 + lane) * 8);
#pragma unroll
	for (int d = 0; d < 4; ++d) o[d] = make_double2(rs[2 * d], rs[2 * d + 1]);
	if (lane == 0) S.rs_valid[q] = valid;
}

// one half-iteration of every running query: warp per query
#ifndef GBP_STEP_MINBLOCKS
#define GBP_STEP_MINBLOCKS 4
#endif
template <typename M>
__global__ void __launch_bounds__(128, GBP_STEP_MINBLOCKS) k_step_half(TerrainView Tv, StepState S, PlanArena A, int64_t Q, uint64_t seed, uint64_t query0,
													gbp_plan_params P, int it, int half) {
	const int lane = threadIdx.x & 31;
	const int64_t q = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
	if (q >= Q || S.status[q] != 0) return;
	int na = S.na[q], nb = S.nb[q];
	int &nx = half == 0 ? na : nb, &ny = half == 0 ? nb : na;
	if (nx >= A.cap || ny >= A.cap) {  // a tree is full at the start of a half: the query stops unsolved (as the megakernel)
		if (lane == 0) { S.status[q] = 2; S.iters[q] = it + 1; }
		return;
	}
	const uint64_t cell = 2 * (uint64_t) it + (uint64_t) half;
	const int src = (int) (cell & 31ull);
	if (!((S.rs_valid[q] >> src) & 1u)) return;  // rrt_connect.cpp:254
	double s_rand[8];
	{
		const double2 *r = reinterpret_cast<const double2 *>(S.rs + ((size_t) q * 32 + src) * 8);
#pragma unroll
		for (int d = 0; d < 4; ++d) { const double2 v = r[d]; s_rand[2 * d] = v.x; s_rand[2 * d + 1] = v.y; }
	}
	PlanTree Ta = arena_tree(A, (int) q, 0, S.na + q), Tb = arena_tree(A, (int) q, 1, S.nb + q);
	PlanTree &Tx = half == 0 ? Ta : Tb, &Ty = half == 0 ? Tb : Ta;
	const int dir_ext = half == 0 ? GBP_FORWARD : GBP_REVERSE, dir_con = half == 0 ? GBP_REVERSE : GBP_FORWARD;
	const GroupMap gm = make_group_map(P.k_candidates, lane);
	long long pair_checks = 0, nn_queries = 1;
	bool solved = false;
	if (warp_extend<M, false>(Tv, Tx, nx, s_rand, dir_ext, seed, query0 + (uint64_t) q, cell, P, gm, lane, pair_checks) != GBP_TRAPPED) {
		double s_new[8];
		tree_get(Tx.t, nx - 1, s_new);
		++nn_queries;
		solved = warp_connect<M>(Tv, Ty, ny, s_new, dir_con, P, lane, pair_checks) == GBP_REACHED;
	}
	if (lane == 0) {
		S.pair_checks[q] += pair_checks;
		S.nn_queries[q] += nn_queries;
		if (solved) { S.status[q] = 1; S.iters[q] = it + 1; }
	}
}

// statistics, stitched paths, postProcessPath: warps pull queries from a counter (path scratch per resident warp)
template <typename M>
__global__ void __launch_bounds__(128) k_step_finish(TerrainView Tv, StepState S, PlanArena A, PlanArena scratch, int64_t Q, gbp_plan_params P,
													  unsigned long long *__restrict__ next_query, gbp_plan_stats *__restrict__ stats,
													  double *__restrict__ path_states, double *__restrict__ path_actions, int path_cap, PlanTreeDump dump) {
	const int lane = threadIdx.x & 31;
	const int64_t slot = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
	while (true) {
		unsigned long long grabbed = 0;
		if (lane == 0) grabbed = atomicAdd(next_query, 1ull);
		const int64_t q = (int64_t) __shfl_sync(FULL, grabbed, 0);
		if (q >= Q) break;
		PlanTree Ta = arena_tree(A, (int) q, 0, S.na + q), Tb = arena_tree(A, (int) q, 1, S.nb + q);
		plan_finish<M>(Tv, P, scratch, slot, Ta, Tb, S.na[q], S.nb[q], S.status[q] == 1, S.iters[q], S.pair_checks[q], S.nn_queries[q], q, stats,
					   path_states, path_actions, path_cap, dump, lane);
		__syncwarp();
	}
}

// Measured on configs[4] (65,536 queries, B200): 1.18-1.78 s against the megakernel's 0.90 s at 80 / 168 / 255 registers —
// every launch streams ~60 KB of straight-line code through each warp ONCE, so the instruction-fetch stalls of the megakernel
// (no_instruction 6.9 cycles per issue in profiles/r2_step_half_ncu_summary.csv, against 7.3) stay and the HBM round trips of
// the per-query state are added.  Kept as an opt-in form (GBP_PLAN_MODE=step) because it is the one whose launches are short
// and uniform (e.g. for interleaving with other work on the stream); results are bit-identical either way.
inline bool plan_step_applies(const gbp_plan_params &P, int64_t nq) {
	const char *mode = getenv("GBP_PLAN_MODE");  // "step": plain fixed-step RRT-Connect batches take the stepped form
	(void) nq;
	if (P.rrt_star || P.adaptive || P.state_direction_sampling || P.stop_after_solved > 0 || P.k_candidates > 32) return false;
	return mode && !strcmp(mode, "step");
}

template <typename M>
inline int plan_step_launch_kind(const TerrainView &Tv, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
								 const gbp_plan_params &P, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap,
								 cudaStream_t st, const PlanTreeDump &dump, std::string &err) {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const size_t cap = (size_t) P.max_vertices, Q = (size_t) nq, per = Q * 2 * cap;
	const int64_t fin_slots = (int64_t) sms * 4 * 4;
	const size_t n_doubles = per * (8 + 10 + 1 + 1) + Q * 32 * 8 + (size_t) fin_slots * 2 * cap * 18;
	const size_t n_ll = Q * 2 + 1;
	const size_t n_ints = per * 3 + Q * 5;
	const size_t need = n_doubles * 8 + n_ll * 8 + n_ints * 4;
	void *mem = nullptr;
	cudaError_t e;
	if ((e = cudaMallocAsync(&mem, need, st)) != cudaSuccess) { err = std::string("stepped planner arena: ") + cudaGetErrorString(e); return GBP_E_CUDA; }
	PlanArena A = {}, Sc = {};
	StepState S;
	A.cap = Sc.cap = P.max_vertices;
	double *dp = (double *) mem;
	A.v = dp; dp += per * 8;
	A.act = dp; dp += per * 10;
	A.g = dp; dp += per;
	A.y = dp; dp += per;
	S.rs = dp; dp += Q * 32 * 8;
	Sc.pstate = dp; dp += (size_t) fin_slots * 2 * cap * 8;
	Sc.paction = dp; dp += (size_t) fin_slots * 2 * cap * 10;
	long long *lp = (long long *) dp;
	S.pair_checks = lp; lp += Q;
	S.nn_queries = lp; lp += Q;
	unsigned long long *next_query = (unsigned long long *) lp; lp += 1;
	int *ip = (int *) lp;
	A.parent = ip; ip += per;
	A.child = ip; ip += per;
	A.sibling = ip; ip += per;
	S.na = ip; ip += Q;
	S.nb = ip; ip += Q;
	S.status = ip; ip += Q;
	S.iters = ip; ip += Q;
	S.rs_valid = (unsigned *) ip; ip += Q;
	const unsigned warp_blocks = (unsigned) ((Q + 3) / 4);
	k_step_init<M><<<(unsigned) ((Q + 127) / 128), 128, 0, st>>>(S, A, nq, starts, goals, P.max_iters);
	for (int it = 0; it < P.max_iters; ++it) {
		if ((it & 15) == 0) k_step_sample<M><<<warp_blocks, 128, 0, st>>>(Tv, S, nq, seed, query0, 2 * (uint64_t) it);
		k_step_half<M><<<warp_blocks, 128, 0, st>>>(Tv, S, A, nq, seed, query0, P, it, 0);
		k_step_half<M><<<warp_blocks, 128, 0, st>>>(Tv, S, A, nq, seed, query0, P, it, 1);
	}
	cudaMemsetAsync(next_query, 0, sizeof(unsigned long long), st);
	const int64_t fin_warps = (int64_t) Q < fin_slots ? (int64_t) ((Q + 3) / 4 * 4) : fin_slots;
	k_step_finish<M><<<(unsigned) (fin_warps / 4), 128, 0, st>>>(Tv, S, A, Sc, nq, P, next_query, stats, path_states, path_actions, path_cap, dump);
	e = cudaGetLastError();
	cudaFreeAsync(mem, st);
	if (e != cudaSuccess) { err = std::string("stepped planner: ") + cudaGetErrorString(e); return GBP_E_CUDA; }
	return GBP_OK;
}
inline int plan_step_launch(const TerrainView &Tv, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
							const gbp_plan_params &P, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap,
							cudaStream_t st, const PlanTreeDump &dump, std::string &err) {
#define GBP_STEP_(M) return plan_step_launch_kind<M>(Tv, nq, starts, goals, seed, query0, P, stats, path_states, path_actions, path_cap, st, dump, err)
	if (Tv.cell_f32) { if (Tv.uniform) GBP_STEP_(MapF32U); else GBP_STEP_(MapF32N); }
	else { if (Tv.uniform) GBP_STEP_(MapF64U); else GBP_STEP_(MapF64N); }
#undef GBP_STEP_
}

// ------------------------------------------------------------------ extend through the host call (rrt.cpp:20-102)
// gbp_extend in ONE launch (measured before: k_nearest 4.3 us + k_extend_candidates 12-23 us + k_extend_select 9.7 us and
// three launch gaps, 56 us per host call).  Every CTA finds the nearest neighbour itself (the same arithmetic and
// (distance, id) argmin as k_nearest: trees are a few hundred vertices, 64 B each, out of L2), then one thread per
// candidate j = ACTION cell idx0 + j from s_near (S lanes per candidate on the fixed step); the last CTA to finish (device counter) does selection + acceptance +
// append + status and writes the result words to device memory and to the caller's mapped host buffer.  The target
// travels as a kernel argument.
template <typename M, bool GROUP>
__device__ __forceinline__ void extend_fused_body(const TerrainView &T, const TreeView &tree, const Target8 &tgt, int direction, int K, int best_of_k,
												  int S, uint64_t seed, uint64_t stream, uint64_t idx0, double dir_thresh, const ExtendScratch &S_,
												  unsigned *__restrict__ done, int *__restrict__ host_result) {
	__shared__ double sd[4];
	__shared__ int si[4];
	__shared__ int s_near_idx, s_last;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	double tg[8];
#pragma unroll
	for (int d = 0; d < 8; ++d) tg[d] = tgt.v[d];
	// ---- getNearestNeighbor (planner_class.cpp:185-200)
	{
		const int nv = *tree.n;
		double bd = INFINITY;
		int bi = 0x7fffffff;
		for (int j = threadIdx.x; j < nv; j += blockDim.x) argmin_combine(bd, bi, vertex_distance(tree, j, tg), j);
		warp_argmin(bd, bi);
		if (lane == 0) { sd[wib] = bd; si[wib] = bi; }
		__syncthreads();
		if (threadIdx.x < 32) {
			bd = threadIdx.x < 4 ? sd[threadIdx.x] : INFINITY;
			bi = threadIdx.x < 4 ? si[threadIdx.x] : 0x7fffffff;
			warp_argmin(bd, bi);
			if (threadIdx.x == 0) {
				s_near_idx = bi == 0x7fffffff ? 0 : bi;  // reference default index 0 (planner_class.cpp:186)
				if (blockIdx.x == 0) { *S_.near_idx = s_near_idx; *S_.near_dist = bd; }
			}
		}
		__syncthreads();
	}
	const int near = s_near_idx;
	double s_near[8], nn[3], R[9];
	tree_get(tree, near, s_near);
	unsigned fl = 0;
	surface_normal(T, tg[0], tg[1], nn, fl);  // rrt.cpp:25 — normal at the TARGET sample
	grf_rotation(nn, R);
	// directional action sampling (planning_utils.cpp:379-391): a negative threshold switches it off (flag && p <= thr)
	const bool dirs = dir_thresh >= 0.0;
	const double *a_from = direction == GBP_FORWARD ? s_near : tg, *a_to = direction == GBP_FORWARD ? tg : s_near;
	// ---- candidates (newConfig, rrt.cpp:20-70, generalised to K)
	if (GROUP) {
		// fixed step: S lanes per candidate speculate its sub-states S at a time (group_validate: only the verdict of a
		// candidate matters here, the exact end state of a valid one comes from finish_output) — ceil(19 / S) evaluator
		// passes of latency instead of up to 19
		const int G = 32 / S, g = lane / S, r = lane - g * S;
		const int j = (blockIdx.x * 4 + wib) * G + g;
		const bool has = g < G && j < K;
		const unsigned gmask = S == 32 ? FULL : ((1u << S) - 1u);
		double a[10];
		sample_action(seed, stream, idx0 + (uint64_t) (has ? j : 0), R, dirs, dir_thresh, a_from, a_to, a);
		const bool ok = group_validate<M>(T, s_near, a, direction, S, r, gmask, has ? g * S : 0, has);
		if (has && r == 0) {
			double sn[8];
			if (ok) finish_output(s_near, a, direction == GBP_FORWARD ? OUT_LAND : OUT_REV, 0.0, sn);
			S_.valid[j] = ok ? 1 : 0;
			S_.dist[j] = ok ? state_distance(sn, tg) : INFINITY;
			if (ok) store_state(S_.s_test + 8 * (size_t) j, sn);
		}
	} else {     // adaptive step: a thread per candidate walks its sub-states in sequence
		const int j = blockIdx.x * blockDim.x + threadIdx.x;
		if (j < K) {
			double a[10], sn[8], tn;
			sample_action(seed, stream, idx0 + (uint64_t) j, R, dirs, dir_thresh, a_from, a_to, a);
			Counters c = {0, 0, 0, 0};
			const bool ok = validate_pair_seq<M>(T, s_near, a, direction, true, sn, tn, c);
			S_.valid[j] = ok ? 1 : 0;
			S_.dist[j] = ok ? state_distance(sn, tg) : INFINITY;
			store_state(S_.s_test + 8 * (size_t) j, sn);
		}
	}
	// ---- the last CTA selects
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1 ? 1 : 0;
	__syncthreads();
	if (!s_last) return;
	__threadfence();
	// best_of_k: argmin of dist (ties: lowest j); first-valid: lowest valid j
	double bd = INFINITY;
	int bi = 0x7fffffff;
	for (int c = threadIdx.x; c < K; c += blockDim.x) {
		if (!__ldcg(S_.valid + c)) continue;
		if (best_of_k) argmin_combine(bd, bi, __ldcg(S_.dist + c), c);
		else if (c < bi) { bi = c; bd = __ldcg(S_.dist + c); }
	}
	if (!best_of_k) {  // reduce on index only
		for (int o = 16; o > 0; o >>= 1) {
			const double od = __shfl_xor_sync(FULL, bd, o);
			const int oi = __shfl_xor_sync(FULL, bi, o);
			if (oi < bi) { bi = oi; bd = od; }
		}
	} else {
		warp_argmin(bd, bi);
	}
	if (lane == 0) { sd[wib] = bd; si[wib] = bi; }
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < 4; ++w) {
			if (best_of_k) argmin_combine(bd, bi, sd[w], si[w]);
			else if (si[w] < bi) { bi = si[w]; bd = sd[w]; }
		}
		int status = GBP_TRAPPED, new_id = -1;
		const int checks = best_of_k ? K : (bi == 0x7fffffff ? K : bi + 1);
		const bool accept = bi != 0x7fffffff && bd < state_distance(s_near, tg);  // rrt.cpp:55-66
		if (accept && !(*tree.n < tree.cap)) status = -1;  // tree full: reported as GBP_E_CAPACITY, not as TRAPPED
		else if (accept) {
			double sn[8], a[10];
			for (int d = 0; d < 8; ++d) sn[d] = __ldcg(S_.s_test + 8 * (size_t) bi + d);
			sample_action(seed, stream, idx0 + (uint64_t) bi, R, dirs, dir_thresh, a_from, a_to, a);
			new_id = tree_push(tree, near, sn, a);
			status = state_distance(sn, tg) <= GOAL_BOUNDS ? GBP_REACHED : GBP_ADVANCED;  // rrt.cpp:96-99
		}
		S_.result[0] = status; S_.result[1] = new_id; S_.result[2] = checks;
		host_result[0] = status; host_result[1] = new_id; host_result[2] = checks;
		*done = 0;  // ready for the next call on this tree
	}
}
template <typename M>
__global__ void __launch_bounds__(128) k_extend_fused(TerrainView T, TreeView tree, Target8 tgt, int direction, int K, int best_of_k,
													   int lanes_per_candidate, uint64_t seed, uint64_t stream, uint64_t idx0, double dir_thresh, ExtendScratch S,
													   unsigned *__restrict__ done, int *__restrict__ host_result) {
	extend_fused_body<M, true>(T, tree, tgt, direction, K, best_of_k, lanes_per_candidate, seed, stream, idx0, dir_thresh, S, done, host_result);
}
template <typename M>
__global__ void __launch_bounds__(128) k_extend_fused_adaptive(TerrainView T, TreeView tree, Target8 tgt, int direction, int K, int best_of_k,
																int unused, uint64_t seed, uint64_t stream, uint64_t idx0, double dir_thresh, ExtendScratch S,
																unsigned *__restrict__ done, int *__restrict__ host_result) {
	extend_fused_body<M, false>(T, tree, tgt, direction, K, best_of_k, 1, seed, stream, idx0, dir_thresh, S, done, host_result);
}

}  // namespace gbp
