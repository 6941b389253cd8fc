// The device-wide form of the batch planner: ONE search at a time on the whole chip.
//
// runRRTConnect (rrt_connect.cpp:230-314) is a chain — extend n+1 needs the tree extend n left — so a single search cannot
// use more than the K candidates of one newConfig (rrt.cpp:34-50) at a time.  Through the host calls that chain costs a
// launch and a synchronisation per link (gbp_extend 40 us + gbp_connect 24 us); in the megakernel one warp walks a query's
// candidates 32 at a time (K = 4096: 128 batches per extend).  Here a cooperative grid of one CTA per SM runs the search:
//
//   * every CTA executes the whole control flow REDUNDANTLY on its own copy of the two trees (same Philox cells, same
//     arithmetic, same inputs -> same decisions), so the trees never travel between SMs and no CTA waits for a broadcast;
//   * only newConfig is split: candidate j goes to lanes [S*j, S*j + S) of the grid (S lanes speculate the sub-states of a
//     candidate, group_validate), each CTA reduces its candidates to one (distance, index, end state, action) record;
//   * ONE exchange per extend and no separate barrier: a CTA stores its record and then, with release, a tag carrying the
//     exchange number; thread i of every CTA polls CTA i's tag with acquire and reads that record — two L2 round trips after
//     the last CTA is through — then every CTA reduces the G records itself and appends the winner to its own tree copy;
//   * connect (rrt_connect.cpp:98-120: nearest neighbour, the connect primitive's pair check spread over the lanes of a warp,
//     append) runs redundantly in every warp: no exchange (one warp per SM with the result through shared memory measured
//     slower: 19.1 against 18.2 us per extend at K = 4096).
//
// Cells, arithmetic and update order are the megakernel's: statistics, paths and both trees are bit-identical
// (tests/test_gpu_planner.py).  Queries of a batch run one after the other.
#pragma once
#include "gbp_planner.cuh"

namespace gbp {

constexpr int WIDE_THREADS = 256, WIDE_WARPS = WIDE_THREADS / 32;

struct WideSlots {     // what a CTA hands to the others after its share of a newConfig; two parities
	double *d;         // [2][G] distance of the CTA's selected candidate to the target (INFINITY: none valid)
	double *sn;        // [2][G][8] its exact end state
	double *a;         // [2][G][10] its action
	unsigned long long *tag;  // [2][G] (exchange number << 32) | candidate index (0x7fffffff: none); stored last, with release
	int *abort;        // set when an exchange timed out (a CTA was not resident or left the common control flow)
	long long *trace;  // GBP_WIDE_TRACE builds: cycles per stage, accumulated by thread 0 of CTA 0
};
#ifdef GBP_WIDE_TRACE
#define WIDE_T(k) do { if (cta == 0 && threadIdx.x == 0) { const long long now_ = clock64(); W.trace[k] += now_ - tr_; tr_ = now_; } } while (0)
#else
#define WIDE_T(k) do { } while (0)
#endif

__device__ __forceinline__ unsigned long long wide_ld_acquire(const unsigned long long *p) {
	unsigned long long v;
	asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void wide_st_release(unsigned long long *p, unsigned long long v) {
	asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// (distance, index) selection: closest valid (ties: lowest index) or first valid in stream order
__device__ __forceinline__ bool wide_better(bool best_of_k, double d1, int j1, double d2, int j2) {
	return best_of_k ? (d1 < d2 || (d1 == d2 && j1 < j2)) : j1 < j2;
}
// block-wide selection over one (d, j) per thread; every thread returns the winner.  sd / si: WIDE_WARPS entries each.
__device__ __forceinline__ void wide_block_select(bool best_of_k, double &d, int &j, double *sd, int *si) {
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		const double od = __shfl_xor_sync(FULL, d, o);
		const int oj = __shfl_xor_sync(FULL, j, o);
		if (wide_better(best_of_k, od, oj, d, j)) { d = od; j = oj; }
	}
	if (lane == 0) { sd[wib] = d; si[wib] = j; }
	__syncthreads();
	d = sd[0]; j = si[0];
#pragma unroll
	for (int w = 1; w < WIDE_WARPS; ++w)
		if (wide_better(best_of_k, sd[w], si[w], d, j)) { d = sd[w]; j = si[w]; }
	__syncthreads();
}
// getNearestNeighbor (planner_class.cpp:185-200) by the whole CTA on its own tree copy; every thread returns the index
__device__ __forceinline__ int wide_nearest(const TreeView &T, int nv, const double q[8], double *sd, int *si) {
	double bd = INFINITY;
	int bi = 0x7fffffff;
	for (int j = threadIdx.x; j < nv; j += WIDE_THREADS) argmin_combine(bd, bi, vertex_distance(T, j, q), j);
	wide_block_select(true, bd, bi, sd, si);
	return bi == 0x7fffffff ? 0 : bi;
}

// addVertex + addEdge + addAction + the g value (rrt.cpp:87-92) with the yaw sum left for the end of the search: nothing in the
// plain RRT-Connect loop reads y, and its two atan2 calls (~3 k cycles) would sit on the critical path of every append.
__device__ __forceinline__ void wide_push(PlanTree &T, int parent, const double s[8], const double a[10]) {
	const int i = *T.t.n;
	*T.t.n = i + 1;
	double p[8];
	tree_get(T.t, parent, p);
#pragma unroll
	for (int d = 0; d < 8; ++d) T.t.v[(size_t) d * T.t.cap + i] = s[d];
#pragma unroll
	for (int d = 0; d < 10; ++d) T.t.act[(size_t) d * T.t.cap + i] = a[d];
	T.t.g[i] = T.t.g[parent] + pose_distance(p, s);
	T.t.y[i] = 0;
	T.child[i] = -1; T.sibling[i] = -1;
	plan_link(T, parent, i);
}
// y[i] = y[parent] + stateYawDistance(parent, i) for the whole tree (graph_class.cpp:36-42 accumulates it per append; ids
// ascend from parent to child, so one pass in id order after the per-edge terms gives the same sums): whole CTA
__device__ __forceinline__ void wide_fill_yaw(PlanTree &T, int n) {
	for (int i = 1 + threadIdx.x; i < n; i += WIDE_THREADS) {
		double a[8], b[8];
		tree_get(T.t, T.t.parent[i], a);
		tree_get(T.t, i, b);
		T.t.y[i] = yaw_distance(a, b);
	}
	__syncthreads();
	if (threadIdx.x == 0)
		for (int i = 1; i < n; ++i) T.t.y[i] = T.t.y[T.t.parent[i]] + T.t.y[i];
	__syncthreads();
}

// small trees (most of a search): every warp scans the tree itself — no shared memory, no CTA barrier
__device__ __forceinline__ int wide_nearest_auto(const TreeView &T, int nv, const double q[8], double *sd, int *si) {
	if (nv <= 64) return warp_nearest(T, nv, q, threadIdx.x & 31);
	return wide_nearest(T, nv, q, sd, si);
}

// group_validate (gbp_planner.cuh) for 4 ... 16 lanes per candidate, in two phases, by the whole CTA.  Phase 1: the S lanes
// of a group check the candidate's first S sub-states at once.  Phase 2: the candidates still running go to a list in shared
// memory and the warps of the CTA take one each, finishing it on all 32 lanes (32 sub-states per evaluator pass).  The
// slowest warp of the grid sets the pace of an extend: it needs 1 + ceil(survivors of its CTA / 8) passes instead of
// ceil(sub-states / S) of them.  Same sub-states, same order, same early exit: same verdicts.  All threads of the CTA call it
// (two CTA barriers).  Tried: phase 1 on S sub-states SPREAD over the path (only the verdict matters to newConfig, so the
// order is free) — lanes in six different phases made that pass 2x as long: 22.2 against 18.6 us per extend at K = 4096.
struct WideSurvivors {
	double a[WIDE_THREADS / 4][10];  // at most one candidate per 4 lanes
	double t[WIDE_THREADS / 4];      // base cursor after phase 1
	int ph[WIDE_THREADS / 4];
	int verdict[WIDE_THREADS / 4];   // 1 invalid, 2 valid
	int n;
};
template <typename M>
__device__ __forceinline__ bool wide_validate(const TerrainView &Tv, const double s_near[8], const double a[10], int dir, int S, int r,
											  unsigned gmask, int gshift, bool has_candidate, WideSurvivors &sv) {
	if (S == 32 || S < 4) return group_validate<M>(Tv, s_near, a, dir, S, r, gmask, gshift, has_candidate);
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	Cursor q;
#pragma unroll
	for (int i = 0; i < 8; ++i) q.s[i] = s_near[i];
#pragma unroll
	for (int i = 0; i < 10; ++i) q.a[i] = a[i];
	cursor_start(q, dir);
	int state = has_candidate ? 0 : 1;  // 0 running, 1 invalid / no candidate, 2 valid
	int slot = -1;
	if (threadIdx.x == 0) sv.n = 0;
	__syncthreads();
	{   // phase 1
		const double ts = q.a[6], tf = q.a[7];
		int ph = q.phase;
		double t = q.t;
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
		int lph = __shfl_sync(FULL, ph, gshift + S - 1);
		double lt = __shfl_sync(FULL, t, gshift + S - 1);
		if (state == 0) {
			if (bad) state = 1;
			else if (term) state = 2;
			else if (r == 0) {  // still running: the group's first lane files it
				walk_step(lph, lt, ts, tf);
				slot = atomicAdd(&sv.n, 1);
#pragma unroll
				for (int i = 0; i < 10; ++i) sv.a[slot][i] = a[i];
				sv.t[slot] = lt;
				sv.ph[slot] = lph;
			}
		}
	}
	__syncthreads();
	const int n = sv.n;
	for (int e = wib; e < n; e += WIDE_WARPS) {  // phase 2: warp-uniform
#pragma unroll
		for (int i = 0; i < 10; ++i) q.a[i] = sv.a[e][i];
		const double ts = q.a[6], tf = q.a[7];
		q.f.inv6ts = 1.0 / (6.0 * ts);
		q.f.inv2ts = 1.0 / (2.0 * ts);
		int cph = sv.ph[e], verdict = 0;
		double ct = sv.t[e];
		while (verdict == 0) {
			int ph = cph;
			double t = ct;
			for (int i = 0; i < lane && ph != PH_DONE; ++i) walk_step(ph, t, ts, tf);
			const bool active = ph != PH_DONE;
			bool valid = true;
			if (active) {
				q.phase = ph;
				q.t = t;
				valid = cursor_check<M>(Tv, q);
			}
			const unsigned bad = __ballot_sync(FULL, active && !valid);
			const unsigned term = __ballot_sync(FULL, active && (ph == PH_FWD_LAND || ph == PH_REV_START));
			if (bad) verdict = 1;
			else if (term) verdict = 2;
			else {
				cph = __shfl_sync(FULL, ph, 31);
				ct = __shfl_sync(FULL, t, 31);
				walk_step(cph, ct, ts, tf);
			}
		}
		if (lane == 0) sv.verdict[e] = verdict;
	}
	__syncthreads();
	slot = __shfl_sync(FULL, slot, gshift);  // the group's entry, if it filed one
	if (slot >= 0) state = sv.verdict[slot];
	return state == 2;
}

template <typename M>
__global__ void __launch_bounds__(WIDE_THREADS, 1) k_plan_wide(TerrainView Tv, int64_t nq, const double *__restrict__ starts,
															   const double *__restrict__ goals, uint64_t seed, uint64_t query0, gbp_plan_params P,
															   PlanArena A, int *__restrict__ counts, WideSlots W, int S,
															   gbp_plan_stats *__restrict__ stats, double *__restrict__ path_states,
															   double *__restrict__ path_actions, int path_cap, PlanTreeDump dump) {
	__shared__ double sd[WIDE_WARPS];
	__shared__ int si[WIDE_WARPS];
	__shared__ double s_win[18];
	__shared__ double s_rs[32][17];  // the current batch of 32 random states: state[8], GRF rotation[9]
	__shared__ unsigned s_rs_valid;  // isValidState(STANCE) of those states
	__shared__ WideSurvivors s_sv;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, G = gridDim.x, cta = blockIdx.x;
	// candidate groups: S lanes per candidate, 32 / S candidates per warp; warp w of CTA c is grid warp w * G + c, so that the
	// first candidates (all of them at K = 6) land on different SMs
	const int GW = 32 / S, g = lane / S, r = lane - g * S, gshift = g * S;
	const unsigned gmask = S == 32 ? FULL : ((1u << S) - 1u);
	const int64_t grid_warp = (int64_t) wib * G + cta, per_pass = (int64_t) WIDE_WARPS * G * GW;
	const int K = P.k_candidates;
	const bool two_phase = S >= 4 && S < 32;
	const bool best_of_k = P.best_of_k != 0, dirs = P.action_direction_sampling != 0;
	unsigned exchange = 0;
	int parity = 0, solved_so_far = 0;
	PlanTree Ta = arena_tree(A, cta, 0, counts + 2 * cta), Tb = arena_tree(A, cta, 1, counts + 2 * cta + 1);
	for (int64_t qi = 0; qi < nq; ++qi) {
		if (P.stop_after_solved > 0 && solved_so_far >= P.stop_after_solved) {  // enough solved: the rest report nothing
			if (cta == 0 && threadIdx.x == 0) { gbp_plan_stats z = {}; stats[qi] = z; }
			continue;
		}
		double start[8], goal[8];
#pragma unroll
		for (int d = 0; d < 8; ++d) { start[d] = starts[8 * qi + d]; goal[d] = goals[8 * qi + d]; }
		__syncthreads();
		if (threadIdx.x == 0) { plan_tree_init(Ta, start); plan_tree_init(Tb, goal); }
		__syncthreads();
		int na = 1, nb = 1, it = 0;
		bool solved = false, full = false;
		long long pair_checks = 0, nn_queries = 0;
		const uint64_t query = query0 + (uint64_t) qi;
		for (; it < P.max_iters && !solved && !full; ++it) {
			for (int half = 0; half < 2 && !solved; ++half) {
				PlanTree &Tx = half == 0 ? Ta : Tb, &Ty = half == 0 ? Tb : Ta;
				int &nx = half == 0 ? na : nb, &ny = half == 0 ? nb : na;
				const int dir_ext = half == 0 ? GBP_FORWARD : GBP_REVERSE, dir_con = half == 0 ? GBP_REVERSE : GBP_FORWARD;
				if (nx >= A.cap || ny >= A.cap) { full = true; break; }
				const uint64_t cell = 2 * (uint64_t) it + (uint64_t) half;
				double s_rand[8];
				if (P.state_direction_sampling) {  // rrt_connect.cpp:246-251, :281-286
					double s_from[8], s_to[8];
					tree_get(Ta.t, half == 0 ? na - 1 : 0, s_from);
					tree_get(Tb.t, half == 0 ? 0 : nb - 1, s_to);
					sample_state<M>(Tv, seed, query, cell, true, P.state_direction_threshold, P.state_direction_speed != 0, s_from, s_to, s_rand);
					Counters c = {0, 0, 0, 0};
					if (!is_valid_state_auto<M>(Tv, pose6(s_rand), GBP_STANCE, c)) continue;  // rrt_connect.cpp:254
				} else {
					if ((cell & 31ull) == 0) {
						// warp 0 draws and checks the next 32 STATE cells, lane L holding cell + L, and prepares what an extend
						// towards each needs before it can sample actions: the surface normal there and the GRF rotation (rrt.cpp:25)
						__syncthreads();
						if (wib == 0) {
							double rs[8], nrm[3], Rl[9];
							sample_state<M>(Tv, seed, query, cell + (uint64_t) lane, false, 0.0, false, nullptr, nullptr, rs);
							Counters c = {0, 0, 0, 0};
							const unsigned v = __ballot_sync(FULL, is_valid_state_auto<M>(Tv, pose6(rs), GBP_STANCE, c));
							unsigned fl0 = 0;
							surface_normal(Tv, rs[0], rs[1], nrm, fl0);
							grf_rotation(nrm, Rl);
#pragma unroll
							for (int d = 0; d < 8; ++d) s_rs[lane][d] = rs[d];
#pragma unroll
							for (int d = 0; d < 9; ++d) s_rs[lane][8 + d] = Rl[d];
							if (lane == 0) s_rs_valid = v;
						}
						__syncthreads();
					}
					const int src = (int) (cell & 31ull);
					if (!((s_rs_valid >> src) & 1u)) continue;  // rrt_connect.cpp:254
#pragma unroll
					for (int d = 0; d < 8; ++d) s_rand[d] = s_rs[src][d];
				}
				// ---- extend (rrt.cpp:77-102): nearest neighbour, newConfig split over the grid
#ifdef GBP_WIDE_TRACE
				long long tr_ = clock64();
				if (cta == 0 && threadIdx.x == 0) W.trace[15] += 1;
#endif
				++nn_queries;
				double s_near[8], R[9], a_first[10];
				if (P.state_direction_sampling) {
					double nn[3];
					unsigned fl = 0;
					surface_normal(Tv, s_rand[0], s_rand[1], nn, fl);  // rrt.cpp:25 — at the TARGET sample
					grf_rotation(nn, R);
				} else {
#pragma unroll
					for (int d = 0; d < 9; ++d) R[d] = s_rs[(int) (cell & 31ull)][8 + d];
				}
				// without directional sampling an action does not depend on the nearest vertex: this lane's first candidate is
				// drawn before the search instead of after it
				const bool pre = !dirs && grid_warp * GW < K;
				if (pre) sample_action(seed, query, cell * (uint64_t) K + (uint64_t) min((int64_t) K - 1, grid_warp * GW + g), R, false, 0.0, s_rand, s_rand, a_first);
				WIDE_T(1);
				const int near = wide_nearest_auto(Tx.t, nx, s_rand, sd, si);
				WIDE_T(0);
				tree_get(Tx.t, near, s_near);
				const double best0 = state_distance(s_near, s_rand);
				const double *a_from = dir_ext == GBP_FORWARD ? s_near : s_rand, *a_to = dir_ext == GBP_FORWARD ? s_rand : s_near;
				double my_d = INFINITY, my_sn[8], my_a[10];
				int my_j = 0x7fffffff;
				for (int64_t pass0 = 0; pass0 < K; pass0 += per_pass) {  // CTA-uniform: a pass holds CTA barriers at S = 4 ... 16
					const int64_t base = pass0 + grid_warp * GW;
					if (!two_phase && base >= K) continue;  // elsewhere a warp without candidates skips the pass
					const int64_t j = base + g;
					const bool has = j < K;
					double a[10];
					if (base >= K) {  // a warp without candidates only takes part in the CTA's phase 2
#pragma unroll
						for (int i = 0; i < 10; ++i) a[i] = 0.0;
					} else if (pre && pass0 == 0) {
#pragma unroll
						for (int i = 0; i < 10; ++i) a[i] = a_first[i];
					} else sample_action(seed, query, cell * (uint64_t) K + (uint64_t) (has ? j : 0), R, dirs, P.action_direction_threshold, a_from, a_to, a);
					WIDE_T(2);
					const bool ok = wide_validate<M>(Tv, s_near, a, dir_ext, S, r, gmask, has ? gshift : 0, has, s_sv);
					WIDE_T(3);
					if (ok && r == 0) {
						double sn[8];
						finish_output(s_near, a, dir_ext == GBP_FORWARD ? OUT_LAND : OUT_REV, 0.0, sn);
						const double d = state_distance(sn, s_rand);
						if (wide_better(best_of_k, d, (int) j, my_d, my_j)) {
							my_d = d; my_j = (int) j;
#pragma unroll
							for (int i = 0; i < 8; ++i) my_sn[i] = sn[i];
#pragma unroll
							for (int i = 0; i < 10; ++i) my_a[i] = a[i];
						}
					}
				}
				WIDE_T(4);
				double bd = my_d;
				int bj = my_j;
				wide_block_select(best_of_k, bd, bj, sd, si);
				// the thread that holds the CTA's candidate (thread 0 when there is none) publishes the record, tag last
				const size_t slot = (size_t) parity * G + cta;
				exchange += 1;
				if (bj != 0x7fffffff ? my_j == bj : threadIdx.x == 0) {
					W.d[slot] = bd;
					if (bj != 0x7fffffff) {
#pragma unroll
						for (int i = 0; i < 8; ++i) W.sn[slot * 8 + i] = my_sn[i];
#pragma unroll
						for (int i = 0; i < 10; ++i) W.a[slot * 10 + i] = my_a[i];
					}
					wide_st_release(W.tag + slot, ((unsigned long long) exchange << 32) | (unsigned) bj);
				}
				WIDE_T(5);
				// every CTA collects all CTAs' records (G <= WIDE_THREADS: one record per thread) and selects among them
				double cd = INFINITY, rec[18];
				int cj = 0x7fffffff;
				bool timed_out = false;
				if (threadIdx.x < G) {
					const size_t theirs = (size_t) parity * G + threadIdx.x;
					unsigned long long tag = wide_ld_acquire(W.tag + theirs);
					for (unsigned spins = 0; (unsigned) (tag >> 32) != exchange; tag = wide_ld_acquire(W.tag + theirs))
						if (++spins > (1u << 22)) { timed_out = true; break; }
					cj = (int) (unsigned) tag;
					cd = __ldcg(W.d + theirs);
					if (cj != 0x7fffffff) {
#pragma unroll
						for (int i = 0; i < 8; ++i) rec[i] = __ldcg(W.sn + theirs * 8 + i);
#pragma unroll
						for (int i = 0; i < 10; ++i) rec[8 + i] = __ldcg(W.a + theirs * 10 + i);
					}
				}
				if (__syncthreads_or(timed_out)) { if (threadIdx.x == 0) *W.abort = 1; return; }
				WIDE_T(6);
				const int mine = cj;
				wide_block_select(best_of_k, cd, cj, sd, si);
				pair_checks += best_of_k ? K : (cj == 0x7fffffff ? K : cj + 1);
				parity ^= 1;
				if (cj == 0x7fffffff || !(cd < best0)) { WIDE_T(10); continue; }  // TRAPPED (rrt.cpp:55-66)
				if (threadIdx.x < G && mine == cj) {  // candidate indices are unique: one thread holds the winner's record
#pragma unroll
					for (int i = 0; i < 18; ++i) s_win[i] = rec[i];
				}
				__syncthreads();
				double s_new[8], a_new[10];
#pragma unroll
				for (int i = 0; i < 8; ++i) s_new[i] = s_win[i];
#pragma unroll
				for (int i = 0; i < 10; ++i) a_new[i] = s_win[8 + i];
				if (threadIdx.x == 0) wide_push(Tx, near, s_new, a_new);
				nx += 1;
				__syncthreads();
				WIDE_T(7);
				// ---- connect (rrt_connect.cpp:98-120) from the other tree, redundantly in every warp of every CTA
				++nn_queries;
				const int near2 = wide_nearest_auto(Ty.t, ny, s_new, sd, si);
				WIDE_T(8);
				double s_near2[8], sn2[8], an2[10];
				tree_get(Ty.t, near2, s_near2);
				Counters c = {0, 0, 0, 0};
				unsigned checks = 0;
				const int rc = attempt_connect_warp<M>(Tv, s_near2, s_new, dir_con, sn2, an2, c, checks);
				pair_checks += checks;
				if (rc != GBP_TRAPPED) {
					if (threadIdx.x == 0) wide_push(Ty, near2, sn2, an2);
					ny += 1;
					__syncthreads();
				}
				WIDE_T(9);
				if (rc == GBP_REACHED) solved = true;
			}
		}
		if (solved) solved_so_far += 1;
		__syncthreads();
		if (cta == 0) { wide_fill_yaw(Ta, na); wide_fill_yaw(Tb, nb); }
		if (cta == 0 && wib == 0)
			plan_finish<M>(Tv, P, A, 0, Ta, Tb, na, nb, solved, it, pair_checks, nn_queries, qi, stats, path_states, path_actions, path_cap, dump, lane);
		__syncthreads();
	}
}

inline bool plan_wide_applies(const gbp_plan_params &P, int64_t nq) {
	if (P.rrt_star || P.adaptive || nq < 1) return false;
	const char *mode = getenv("GBP_PLAN_MODE");
	if (mode && !strcmp(mode, "wide")) return true;
	if (mode && (!strcmp(mode, "mega") || !strcmp(mode, "step") || !strcmp(mode, "pipe"))) return false;
	// one query after the other, ~10 us per extend + connect whatever K: ahead of the megakernel's warp per query (32 candidates
	// at a time) when a newConfig holds more candidates than the batch has queries to fill the chip with
	return (int64_t) P.k_candidates >= 256 * nq;
}

template <typename M>
inline int plan_wide_launch_kind(const TerrainView &Tv, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
								 const gbp_plan_params &P, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap,
								 cudaStream_t st, const PlanTreeDump &dump, std::string &err) {
	int dev = 0, sms = 148, coop = 0, per_sm = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_plan_wide<M>, WIDE_THREADS, 0);
	if (!coop || per_sm < 1) { err = "device-wide planner: cooperative launch not available"; return GBP_E_CUDA; }
	int G = sms < WIDE_THREADS ? sms : WIDE_THREADS;
	// lanes per candidate: the largest power of two that still gives every candidate of a newConfig its lanes in one pass
	int S = 32;
	while (S > 1 && (int64_t) G * WIDE_THREADS / S < (int64_t) P.k_candidates) S >>= 1;
	const size_t cap = (size_t) P.max_vertices, per = (size_t) G * 2 * cap;
	const size_t n_doubles = per * (8 + 10 + 1 + 1) + 2 * cap * 18 + (size_t) 2 * G * (1 + 8 + 10 + 1) + 16;
	const size_t n_ints = per * 3 + (size_t) 2 * G + 4;
	const size_t need = n_doubles * 8 + n_ints * 4;
	void *mem = nullptr;
	cudaError_t e;
	if ((e = cudaMallocAsync(&mem, need, st)) != cudaSuccess) { err = std::string("device-wide planner arena: ") + cudaGetErrorString(e); return GBP_E_CUDA; }
	PlanArena A = {};
	WideSlots W;
	A.cap = P.max_vertices;
	double *dp = (double *) mem;
	A.v = dp; dp += per * 8;
	A.act = dp; dp += per * 10;
	A.g = dp; dp += per;
	A.y = dp; dp += per;
	A.pstate = dp; dp += 2 * cap * 8;      // one scratch slot: CTA 0 stitches the path
	A.paction = dp; dp += 2 * cap * 10;
	W.d = dp; dp += 2 * G;
	W.sn = dp; dp += (size_t) 2 * G * 8;
	W.a = dp; dp += (size_t) 2 * G * 10;
	W.tag = (unsigned long long *) dp; dp += 2 * G;
	W.trace = (long long *) dp; dp += 16;
	int *ip = (int *) dp;
	A.parent = ip; ip += per;
	A.child = ip; ip += per;
	A.sibling = ip; ip += per;
	int *counts = ip; ip += 2 * G;
	W.abort = ip; ip += 4;
	if ((e = cudaMemsetAsync(W.tag, 0, (size_t) (2 * G + 16) * 8, st)) != cudaSuccess || (e = cudaMemsetAsync(W.abort, 0, 16, st)) != cudaSuccess) {
		err = cudaGetErrorString(e); cudaFreeAsync(mem, st); return GBP_E_CUDA;
	}
	TerrainView tv = Tv;
	gbp_plan_params p = P;
	PlanTreeDump dmp = dump;
	void *args[] = {&tv, &nq, &starts, &goals, &seed, &query0, &p, &A, &counts, &W, &S, &stats, &path_states, &path_actions, &path_cap, &dmp};
	e = cudaLaunchCooperativeKernel((const void *) k_plan_wide<M>, dim3((unsigned) G), dim3(WIDE_THREADS), args, 0, st);
	int aborted = 0;
	if (e == cudaSuccess) e = cudaMemcpyAsync(&aborted, W.abort, sizeof(int), cudaMemcpyDeviceToHost, st);
#ifdef GBP_WIDE_TRACE
	long long tr[16] = {};
	if (e == cudaSuccess) e = cudaMemcpyAsync(tr, W.trace, sizeof tr, cudaMemcpyDeviceToHost, st);
#endif
	if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the abort word is the call's error status
#ifdef GBP_WIDE_TRACE
	if (tr[15] > 0) {
		static const char *nm[11] = {"nearest", "normal+R", "sample", "validate", "finish", "select+publish", "collect", "global select+push", "nearest 2", "connect", "global select (trapped)"};
		fprintf(stderr, "wide trace, cycles per extend over %lld extends (S = %d):", tr[15], S);
		for (int k = 0; k < 11; ++k) fprintf(stderr, " %s %.0f;", nm[k], (double) tr[k] / (double) tr[15]);
		fprintf(stderr, "\n");
	}
#endif
	cudaFreeAsync(mem, st);
	if (e != cudaSuccess) { err = std::string("k_plan_wide: ") + cudaGetErrorString(e); return GBP_E_CUDA; }
	if (aborted) { err = "k_plan_wide: exchange between CTAs timed out"; return GBP_E_CUDA; }
	return GBP_OK;
}

}  // namespace gbp
