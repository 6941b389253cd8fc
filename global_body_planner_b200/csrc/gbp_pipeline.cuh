// The pipelined batch planner: runRRTConnect (rrt_connect.cpp:230-314) for a large batch of independent queries as a
// sequence of ROUNDS, each advancing every running query by up to B extend attempts (B = 1 .. 8 speculated half-iterations,
// see k_pipe_prep), with the three roles of an extend in three kernels that all queries go through together:
//
//   k_pipe_prep    4 lanes per segment: skips the query's invalid random states (rrt_connect.cpp:254; STATE cells are drawn and
//                  validity-checked 32 at a time, one batch ahead), checks the budget and the tree capacity, and for each of
//                  the query's next B half-iterations finds the nearest neighbour (rrt.cpp:78), the surface normal at the
//                  target (rrt.cpp:25) and its GRF rotation, and appends one SEGMENT {s_near, R, s_rand, direction, Philox
//                  cell} to the round's dense list
//   k_walk_seg     the candidates of all segments FLATTENED: candidate i = (segment i / K, action j = i % K).  This is
//                  k_walk_sv (gbp_sv.cuh) with per-segment parameters: a warp produces 32 candidates convergently (state
//                  row gathered with cp.async while the action is sampled), a lane walks one candidate's sub-states in the
//                  reference's order through the mixed-precision evaluator, a finished lane takes the next candidate.  The
//                  start state of a candidate is a tree vertex — valid by construction — so its own check (the reference's
//                  first sub-state) is not repeated (roots are checked once by k_pipe_init).  Out: one VALID bit and one
//                  UNDECIDED bit per candidate.
//   k_pipe_triage  thread per query: counts the leading TRAPPED segments (98 % of all), hands the first other one to select
//   k_pipe_select  warp per segment: resolves undecided candidates with the fp64 evaluator, takes the first valid action
//                  in stream order (or the closest valid one), rebuilds its exact end state (finish_output), applies
//                  newConfig's acceptance (rrt.cpp:55-66), appends the vertex (rrt.cpp:87-92) and runs connect from the
//                  other tree (rrt_connect.cpp:98-120, the pair check spread over the lanes); then moves the query to its
//                  next half-iteration or marks it solved.
//
// Why (profiles/r1c_planner_ncu_summary.csv, profiles/r2_step_half_ncu_summary.csv): in the megakernel a warp is one query,
// its 6 candidates x 5 speculative sub-states fill 30 lanes, and every lane of a candidate's group samples the same action
// — sampling alone is ~900 of the ~3800 warp instructions of an extend, the code of an extend is ~60 KB of straight line
// that every warp streams through once per iteration (instruction fetch is 7 of the 15 stall cycles per issue), and the
// 80-register cap spills 2.7 KB per thread.  Flattened, a lane samples ONE action for ONE candidate (28 warp instructions per
// candidate instead of 150), the walk is the microbenchmark's 1.2 k-instruction loop (no fetch stalls, 128 registers, no
// spills), and invalid random states cost nothing: ~2500 rounds instead of 4000 half-iterations at B = 1, ~780 at B = 4.
// Results are the megakernel's bit for bit (same Philox cells, same arithmetic, same order of tree updates per query):
// tests/test_gpu_planner.py runs both forms against the oracle and the reference-loop golden.
// Applies to plain RRT-Connect at the fixed step with K <= 32 on terrains with the mixed-precision evaluator; everything
// else (RRT*, adaptive step, directional STATE sampling, anytime rounds, K > 32, other map kinds) stays on k_plan_batch.
#pragma once
#include <chrono>
#include <type_traits>

#include "gbp_planner.cuh"
#include "gbp_sv.cuh"

namespace gbp {

constexpr int PIPE_BATCH = 64;  // STATE cells drawn per request (k_pipe_batch): what a query can consume per round
constexpr int PIPE_ROW = 26;  // doubles per segment row: s_near[8], R[9], s_rand[8], pad (rows stay 16-byte aligned)

struct PipeState {  // per query
	int *na, *nb;              // vertex counts (the trees' `n` words)
	int *status;               // 0 running, 1 solved, 2 stopped (budget used up or a tree full)
	int *it, *half;            // the half-iteration the query is at
	int *iters;                // started iterations when it stopped
	long long *pair_checks, *nn_queries;
	double *rs;                // [Q][2][64][8] two batches of 64 STATE cells: batch b (cells 64 b .. 64 b + 63) lives in slot b & 1
	unsigned long long *rs_valid;  // [Q][2] isValidState(STANCE) of a batch
	long long *rs_base;        // [Q][2] first cell of the batch a slot holds (-1: none yet)
	long long *rs_want;        // [Q] first cell of the batch k_pipe_batch is asked to draw
	unsigned char *root_valid; // [Q][2] isValidState(root, STANCE) of the start-side / goal-side tree
	int *busy_until;           // [Q] first round that may touch the query again (its connect runs on the second stream meanwhile)
};
// per-round counters (one block per round parity): the kernels of a round on the second stream read them while the first
// stream is already in the next round
// Rounds a query sits out after triage handed it to select / connect: those run on their own streams while the following
// rounds work on the other queries, and everything they use (heavy list, counter block, events) exists PIPE_DEPTH times.
// With 2, prep(r) waits 17-37 us per round for connect(r - 2) (a walk holds every SM until its candidates run out: the last
// connect warps get one in the walk's tail); with 3 that wait is 4 us but the queries that grew sit out one more round and the
// batch needs 13 % more rounds: 0.376 s against 0.358 s on configs[4].  Releasing each query on its own (release store at the
// end of its select / connect, acquire load in prep, no wait for the whole of connect(r - 2)) was measured too: no gain at
// 65,536 queries (0.283 against 0.277 s), slower at 16,384 (0.158 against 0.136 s) — without the wait the side kernels are
// pushed behind one more walk and their queries come back a round later.
constexpr int PIPE_DEPTH = 2;

enum : int { CNT_SEGS = 0, CNT_BUSY = 1, CNT_HEAVY = 2, CNT_CONNECTS = 3, CNT_BATCHES = 4, CNT_WORK = 5, CNT_RUNNING = 6, CNT_WORDS = 8 };
struct PipeHeavy {             // segments with a valid or an undecided candidate, copied out of the round's segment arrays
	double *rows;              // [Q][PIPE_ROW]
	int *q, *near;
	unsigned char *flags;
	unsigned long long *idx0;
	unsigned *vmask, *umask;
	int *connects;             // [Q] connect requests of the round: query * 2 + half
};
struct PipeSegs {  // the round's dense segment list: the segments of a query are consecutive, in cell order
	double *rows;              // [Q * B][PIPE_ROW]
	int *q, *near;             // query, id of s_near in the tree being extended
	unsigned char *flags;      // bit 0: direction, bit 1: s_near is known valid
	unsigned char *ord;        // position of the segment among its query's segments of this round | (their number - 1) << 4
	unsigned long long *idx0;  // ACTION cell of candidate 0: cell * K
	int *count;                // the round's counter block (CNT_*)
	unsigned *vbits, *ubits;   // one bit per candidate; zeroed at the start of every round
};

template <typename M>
__global__ void __launch_bounds__(128) k_pipe_init(TerrainView Tv, PipeState S, PlanArena A, int64_t Q, const double *__restrict__ starts,
													const double *__restrict__ goals, int max_iters) {
	const int64_t q = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (q >= Q) return;
	PlanTree Ta = arena_tree(A, (int) q, 0, S.na + q), Tb = arena_tree(A, (int) q, 1, S.nb + q);
	double s[8], g[8];
#pragma unroll
	for (int d = 0; d < 8; ++d) { s[d] = starts[8 * q + d]; g[d] = goals[8 * q + d]; }
	plan_tree_init(Ta, s);
	plan_tree_init(Tb, g);
	Counters c = {0, 0, 0, 0};
	S.root_valid[2 * q] = is_valid_state_auto<M>(Tv, pose6(s), GBP_STANCE, c) ? 1 : 0;
	S.root_valid[2 * q + 1] = is_valid_state_auto<M>(Tv, pose6(g), GBP_STANCE, c) ? 1 : 0;
	S.status[q] = 0; S.it[q] = 0; S.half[q] = 0; S.iters[q] = max_iters; S.pair_checks[q] = 0; S.nn_queries[q] = 0;
	S.rs_valid[2 * q] = S.rs_valid[2 * q + 1] = 0; S.rs_base[2 * q] = S.rs_base[2 * q + 1] = -1; S.rs_want[q] = 0; S.busy_until[q] = 0;
}

// STATE cells base .. base + 63 of the queries that ask for random states: warp per listed query, lane L draws cells
// base + L and base + 32 + L and checks them (their validity does not depend on the trees).  Runs on a side stream while the
// round's walk runs.
template <typename M>
__global__ void __launch_bounds__(128) k_pipe_batch(TerrainView Tv, PipeState S, const int *__restrict__ list, const int *__restrict__ cnt, uint64_t seed,
													 uint64_t query0) {
	const int lane = threadIdx.x & 31;
	const int n = cnt[CNT_BATCHES], warps = (gridDim.x * blockDim.x) >> 5;
	for (int e = (int) ((blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5); e < n; e += warps) {
		const int q = list[e];
		const long long base = S.rs_want[q];
		const size_t slot = 2 * (size_t) q + (size_t) ((base >> 6) & 1);
		unsigned long long valid = 0;
		for (int p = 0; p < 2; ++p) {
			double rs[8];
			sample_state<M>(Tv, seed, query0 + (uint64_t) q, (uint64_t) base + (uint64_t) (32 * p + lane), false, 0.0, false, nullptr, nullptr, rs);
			Counters c = {0, 0, 0, 0};
			valid |= (unsigned long long) __ballot_sync(FULL, is_valid_state_auto<M>(Tv, pose6(rs), GBP_STANCE, c)) << (32 * p);
			double2 *o = reinterpret_cast<double2 *>(S.rs + (slot * PIPE_BATCH + 32 * p + lane) * 8);
#pragma unroll
			for (int d = 0; d < 4; ++d) o[d] = make_double2(rs[2 * d], rs[2 * d + 1]);
		}
		if (lane == 0) { S.rs_valid[slot] = valid; S.rs_base[slot] = base; }
	}
}

// The control flow of a query's next half-iterations up to newConfig's candidates — budget and capacity checks, the next
// valid random states of the query (invalid ones are skipped, rrt_connect.cpp:254), and per half the nearest neighbour
// (rrt.cpp:78, a sequential scan of the query's tree: a few dozen vertices), the surface normal at the target (rrt.cpp:25), its
// GRF rotation and the segment row.
//
// SPECULATION over half-iterations: an extend that comes back TRAPPED leaves both trees as they were (rrt.cpp:77-102), the
// STATE cell of half h is 2 * it + half and its ACTION cells are h * K + j whatever happened before, and 98 % of the extends
// on these terrains are TRAPPED — so the next B halves of a query are independent unless one of them grows a tree.  A query
// emits up to B segments per round (its next B valid random states, each with the nearest neighbour in the tree that half
// extends); k_pipe_triage takes them in order, counts the leading TRAPPED ones, hands the first segment with a valid or an
// undecided candidate to k_pipe_select and drops the segments after it (they are emitted again, against the grown tree).
// Same Philox cells, same arithmetic, same order of tree updates: results do not depend on B.  A round then advances a query
// by up to B halves for one round's worth of launch gaps, waits and prep.
//
// 4 * B lanes per query (B = 1, 2, 4, 8, 16): every lane of a query runs the control flow, FOUR LANES per segment share the tree
// scan.  (A warp per query spent its time on dependent HBM round trips; several queries per warp issue them side by side.)
// The two batches of 64 STATE cells a query holds are drawn one ahead: a query only sits a round out for random states when
// it jumps over a whole batch of invalid ones.
template <typename M>
__global__ void __launch_bounds__(128, 6) k_pipe_prep(TerrainView Tv, PipeState S, PlanArena A, PipeSegs G, int *__restrict__ batch_list, int64_t Q,
													gbp_plan_params P, int round, int B, const int *solved_count) {
	__shared__ int s_cnt[32], s_off[32], s_run[32];
	const int lane = threadIdx.x & 31, sub = lane & 3;
	const int lpq = 4 * B;                                   // lanes per query
	const int qib = (int) threadIdx.x / lpq;                 // query within the block
	const int j = ((int) threadIdx.x - qib * lpq) >> 2;      // segment within the query
	const int64_t q = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) / lpq;
	const bool lead = sub == 0, qlead = (int) threadIdx.x == qib * lpq;
	int cnt = 0, na = 0, nb = 0, status = -1, it0 = 0, half0 = 0, busy_until = 0;
	long long cell = 0;
	unsigned long long m = 0;
	longlong2 rs_base = make_longlong2(-2, -2);
	ulonglong2 rs_valid = make_ulonglong2(0ull, 0ull);
	if (q < Q) {  // every lane of a query reads its state; its first lane writes what changes after the block's barrier below
		status = S.status[q]; it0 = S.it[q]; half0 = S.half[q]; na = S.na[q]; nb = S.nb[q]; busy_until = S.busy_until[q];
		rs_base = *reinterpret_cast<const longlong2 *>(S.rs_base + 2 * q);
		rs_valid = *reinterpret_cast<const ulonglong2 *>(S.rs_valid + 2 * q);
	}
	int stop_iters = -1;      // >= 0: the query stops here with this iteration count
	long long want = -1, jump = -1;  // batch of random states to draw / cell the query jumps to
	bool busy = false, sit_out = false;
	if (q < Q && status == 0) {
		const int it = it0, half = half0;
		if (round < busy_until) busy = true;  // its select / connect or its next batch of random states is in flight
		else if (P.stop_after_solved > 0 && *(const volatile int *) solved_count >= P.stop_after_solved) stop_iters = it;  // anytime use: enough queries have solved
		else if (it >= P.max_iters) stop_iters = P.max_iters;  // budget used up
		else if (na >= A.cap || nb >= A.cap) stop_iters = it + 1;  // a tree is full at the start of a half
		else {
			cell = 2ll * it + half;
			const long long base = cell & ~63ll, limit = 2ll * P.max_iters;  // cells below `limit` are within the budget
			const bool odd = ((cell >> 6) & 1) != 0;
			const long long base0 = odd ? rs_base.y : rs_base.x, base1 = odd ? rs_base.x : rs_base.y;  // the slot `cell` belongs to, the other one
			const unsigned long long valid0 = odd ? rs_valid.y : rs_valid.x, valid1 = odd ? rs_valid.x : rs_valid.y;
			if (base0 != base) { want = base; sit_out = true; }  // the batch holding `cell` is drawn by k_pipe_batch while this round's walk runs
			else {
				const int off = (int) (cell & 63ll);
				const bool have_next = base1 == base + PIPE_BATCH;
				// the next 64 cells from `cell` on, as far as their random states are known
				m = valid0 >> off;
				if (have_next && off) m |= valid1 << (64 - off);
				const long long end = have_next ? cell + 64 : base + PIPE_BATCH;  // first cell this round cannot look at
				if (limit - cell < 64) m &= (1ull << (int) (limit - cell)) - 1ull;
				if (m == 0) {  // no valid state among them (invalid random states are skipped, rrt_connect.cpp:254)
					if (end >= limit) stop_iters = P.max_iters;
					else {
						jump = end; sit_out = true;
						if (!have_next) want = end;  // (otherwise `end` lies in the batch already drawn)
					}
				} else {
					cnt = min(B, __popcll(m));
					if (!have_next && base + PIPE_BATCH < limit) want = base + PIPE_BATCH;  // drawn one batch ahead: ready for the next round
				}
			}
		}
	}
	// dense segment numbers, a query's segments consecutive: one atomic per block
	if (qlead) { s_cnt[qib] = cnt; s_run[qib] = (q < Q && status == 0 && stop_iters < 0) ? 1 : 0; }
	__syncthreads();
	if (threadIdx.x == 0) {
		const int nqb = (int) blockDim.x / lpq;
		int total = 0, running = 0;
		for (int i = 0; i < nqb; ++i) { s_off[i] = total; total += s_cnt[i]; running += s_run[i]; }
		const int first = total ? atomicAdd(G.count + CNT_SEGS, total) : 0;
		for (int i = 0; i < nqb; ++i) s_off[i] += first;
		if (running) atomicAdd(G.count + CNT_RUNNING, running);  // queries still searching (emitting, sitting a round out, or with select / connect in flight)
	}
	if (qlead) {
		if (stop_iters >= 0) { S.status[q] = 2; S.iters[q] = stop_iters; }
		if (jump >= 0) { S.it[q] = (int) (jump >> 1); S.half[q] = (int) (jump & 1); }
		if (want >= 0) {
			S.rs_want[q] = want;
			batch_list[atomicAdd(G.count + CNT_BATCHES, 1)] = (int) q;
		}
		if (sit_out) S.busy_until[q] = round + 1;
		if (busy || sit_out) atomicAdd(G.count + CNT_BUSY, 1);
	}
	__syncthreads();
	if (j >= cnt) return;  // the 4 lanes of a segment leave together
	const int seg = s_off[qib] + j;
	for (int i = 0; i < j; ++i) m &= m - 1ull;
	cell += __ffsll((long long) m) - 1;  // the j-th valid random state from `cell` on
	const int it = (int) (cell >> 1), half = (int) (cell & 1);
	const unsigned gmask = 0xfu << (lane & ~3);
	double s_rand[8];
	{
		const double2 *r = reinterpret_cast<const double2 *>(S.rs + ((2 * (size_t) q + (size_t) ((cell >> 6) & 1)) * PIPE_BATCH + (size_t) (cell & 63ll)) * 8);
#pragma unroll
		for (int d = 0; d < 4; ++d) { const double2 v = r[d]; s_rand[2 * d] = v.x; s_rand[2 * d + 1] = v.y; }
	}
	PlanTree Tx = arena_tree(A, (int) q, half, (half == 0 ? S.na : S.nb) + q);
	const int nx = half == 0 ? na : nb;
	// getNearestNeighbor (planner_class.cpp:185-200): (distance, id) argmin — the lowest id among equal distances
	double bd = INFINITY;
	int near = 0x7fffffff;
	for (int vi = sub; vi < nx; vi += 8) {  // two vertices per trip: the scan is a chain of HBM round trips (profiles/r2b_pipe_prep_lines.txt), 16 loads in flight halve it
		const int v2 = vi + 4;
		const bool two = v2 < nx;
		const int v2c = two ? v2 : vi;
		double v[8], w[8], sum = 0, sum2 = 0;
#pragma unroll
		for (int d = 0; d < 8; ++d) { v[d] = Tx.t.v[(size_t) d * Tx.t.cap + vi]; w[d] = Tx.t.v[(size_t) d * Tx.t.cap + v2c]; }
#pragma unroll
		for (int d = 0; d < 8; ++d) sum = sum + 1.0 * (v[d] - s_rand[d]) * (v[d] - s_rand[d]);
#pragma unroll
		for (int d = 0; d < 8; ++d) sum2 = sum2 + 1.0 * (w[d] - s_rand[d]) * (w[d] - s_rand[d]);
		const double dj = sqrt(sum), dj2 = sqrt(sum2);
		if (dj < bd) { bd = dj; near = vi; }  // ids ascend within a lane: a strict < keeps the lowest
		if (two && dj2 < bd) { bd = dj2; near = v2; }
	}
#pragma unroll
	for (int o = 1; o < 4; o <<= 1) {
		const double od = __shfl_xor_sync(gmask, bd, o);
		const int oj = __shfl_xor_sync(gmask, near, o);
		if (od < bd || (od == bd && oj < near)) { bd = od; near = oj; }
	}
	if (near == 0x7fffffff) near = 0;  // no finite distance: the reference's default index 0 (planner_class.cpp:186)
	double2 *row = reinterpret_cast<double2 *>(G.rows + (size_t) seg * PIPE_ROW);
	// s_near: each of the 4 lanes fetches and stores a quarter of it (the vertex was just scanned: L1 / L2 hits)
	row[sub] = make_double2(Tx.t.v[(size_t) (2 * sub) * Tx.t.cap + near], Tx.t.v[(size_t) (2 * sub + 1) * Tx.t.cap + near]);
	if (!lead) return;
	double R[9];
	if (Tv.nz3 || Tv.nz3d) {
		double nn[3];
		unsigned fl = 0;
		surface_normal(Tv, s_rand[0], s_rand[1], nn, fl);  // rrt.cpp:25 — at the TARGET sample
		grf_rotation(nn, R);
	} else {
		// a terrain without normal layers has dx = dy = 0 everywhere (fast_terrain_map.cpp:68-72): the interpolated normal is
		// (0, 0, ~1), its cross product with +z is exactly zero and rotate_grf takes the identity (planning_utils.cpp:213)
#pragma unroll
		for (int d = 0; d < 9; ++d) R[d] = (d == 0 || d == 4 || d == 8) ? 1.0 : 0.0;
	}
	row[4] = make_double2(R[0], R[1]); row[5] = make_double2(R[2], R[3]); row[6] = make_double2(R[4], R[5]); row[7] = make_double2(R[6], R[7]);
	row[8] = make_double2(R[8], s_rand[0]); row[9] = make_double2(s_rand[1], s_rand[2]); row[10] = make_double2(s_rand[3], s_rand[4]);
	row[11] = make_double2(s_rand[5], s_rand[6]); row[12] = make_double2(s_rand[7], 0.0);
	G.q[seg] = (int) q;
	G.near[seg] = near;
	// a tree vertex other than the root is the end state of a fully valid pair check: STANCE-valid by construction
	const bool known_valid = near != 0 || S.root_valid[2 * q + half] != 0;
	G.flags[seg] = (unsigned char) ((half == 0 ? GBP_FORWARD : GBP_REVERSE) | (known_valid ? 2 : 0));
	G.ord[seg] = (unsigned char) (j | ((cnt - 1) << 4));
	G.idx0[seg] = (unsigned long long) cell * (unsigned long long) P.k_candidates;
	(void) it;
}

// k_walk_sv with per-segment parameters; see the file header.  `count` = segments of this round (device).
template <bool TEX>
__global__ void __launch_bounds__(RF_WARPS * 32, GBP_WALK_CTAS) k_walk_seg(TerrainView T, PipeSegs G, int K, uint64_t seed, uint64_t query0, int dir_sampling,
																		 double dir_thresh) {
	__shared__ __align__(16) double ringS[RF_WARPS][SV_CAP][8];
	__shared__ __align__(16) double ringA[RF_WARPS][SV_CAP][10];
	__shared__ __align__(16) double stash[8][RF_WARPS * 32];
	__shared__ uint8_t ringD[RF_WARPS][SV_CAP];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	__shared__ int ringI[RF_WARPS][SV_CAP];
	const int n = G.count[CNT_SEGS] * K;
	double *const st = &stash[0][threadIdx.x];
	const uint64_t pol = l2_evict_first_policy();
	// warps claim batches of 32 candidates from a counter (a round is a few batches per warp: static ranges would leave the
	// warps with short candidates idle while the longest range finishes)
	int next = 0, filled = 0;  // candidates handed out / produced by this warp
	bool exhausted = false;
	WalkCursor q;
	q.phase = PH_IDLE;
	int mine = -1;
	while (true) {
		const unsigned need = __ballot_sync(FULL, q.phase == PH_IDLE);
		if (need && (next < filled || !exhausted)) {
			const int nidle = __popc(need);
			if (filled - next < nidle && !exhausted && filled - next <= SV_CAP - 32) {
				int base = 0;
				if (lane == 0) base = atomicAdd(G.count + CNT_WORK, 32);
				base = __shfl_sync(FULL, base, 0);
				const int cnt = min(32, n - base);
				if (cnt <= 0) exhausted = true;
				if (lane < cnt) {
					const int e = (filled + lane) % SV_CAP;
					const int i = base + lane;
					const int seg = i / K, j = i - seg * K;
					const double *row = G.rows + (size_t) seg * PIPE_ROW;
#pragma unroll
					for (int d = 0; d < 4; ++d) cp_async16_hint(&ringS[wib][e][2 * d], row + 2 * d, pol);
					const int f = (int) G.flags[seg];
					ringD[wib][e] = (uint8_t) f;
					ringI[wib][e] = i;
					double vx = 0, vy = 0, tvx = 0, tvy = 0;
					if (dir_sampling) { vx = row[3]; vy = row[4]; tvx = row[17 + 3]; tvy = row[17 + 4]; }
					sv_sample_to_ring(seed, query0 + (uint64_t) G.q[seg], G.idx0[seg] + (uint64_t) j, row + 8, dir_sampling, dir_thresh, tvx, tvy, f & 1, vx,
									  vy, &ringA[wib][e][0]);
				}
				cp_async_wait_all();
				__syncwarp();
				filled += max(cnt, 0);
			}
			if (q.phase == PH_IDLE) {
				const int rel = next + __popc(need & ((1u << lane) - 1));
				if (rel < filled) {
					const int e = rel % SV_CAP;
					const double2 *ps = reinterpret_cast<const double2 *>(&ringS[wib][e][0]);
					const double2 *pa = reinterpret_cast<const double2 *>(&ringA[wib][e][0]);
					double s[8], a[10];
#pragma unroll
					for (int d = 0; d < 4; ++d) { const double2 v = ps[d]; s[2 * d] = v.x; s[2 * d + 1] = v.y; }
#pragma unroll
					for (int d = 0; d < 5; ++d) { const double2 v = pa[d]; a[2 * d] = v.x; a[2 * d + 1] = v.y; }
					mine = ringI[wib][e];
					const int f = (int) ringD[wib][e];
					walk_start(q, s, a, f & 1, st);
					// the start state's own check (the reference's first sub-state) is not repeated for a vertex known valid
					if ((f & 2) && q.t == 0 && (q.phase == PH_FWD_ST || q.phase == PH_REV_FL)) {
						OutRecipe dummy;
						(void) walk_advance<false>(q, true, dummy, st);
					}
				}
			}
			next = min(filled, next + nidle);
			__syncwarp();
		}
		if (__ballot_sync(FULL, q.phase != PH_IDLE) == 0) {
			if (exhausted && next >= filled) break;
			continue;
		}
		bool valid = false, decided = true;
		if (q.phase != PH_IDLE) {
			const int ph = (q.phase == PH_FWD_FL || q.phase == PH_REV_FL) ? GBP_FLIGHT : GBP_STANCE;
			decided = is_valid_state_mixed<MapF32U, TEX>(T, walk_pose(q), ph, q.c_, valid);
		}
		if (!decided) {  // resolved by k_pipe_select with the fp64 evaluator
			atomicOr(G.ubits + (mine >> 5), 1u << (mine & 31));
			q.phase = PH_IDLE;
		}
		if (q.phase != PH_IDLE) {
			OutRecipe out;
			const int r = walk_advance<false>(q, valid, out, st);
			if (r) {
				if (r == 2) atomicOr(G.vbits + (mine >> 5), 1u << (mine & 31));
				q.phase = PH_IDLE;
			}
		}
	}
}

__device__ __forceinline__ unsigned pipe_bits(const unsigned *__restrict__ words, int first, int count) {
	const unsigned long long two = (unsigned long long) words[first >> 5] | ((unsigned long long) words[(first >> 5) + 1] << 32);
	return (unsigned) (two >> (first & 31)) & (count >= 32 ? 0xffffffffu : ((1u << count) - 1u));
}

// triage, thread per query that emitted segments this round (the thread of its first segment): its segments in cell order.
// A segment whose candidates are all invalid (98 % of them) is TRAPPED — counted, and the query moves on to the next one.  The
// first segment with a valid or an undecided candidate is copied to the heavy list for k_pipe_select, which runs on the second
// stream while the next round works on the other queries (its query sits that round out); the query is left AT that half,
// the speculated segments after it are dropped.
static __global__ void __launch_bounds__(256) k_pipe_triage(PipeState S, PipeSegs G, PipeHeavy H, int K, int round, int *diag) {  // H: this round's slot; diag: GBP_PIPE_TRACE counters or null
	const int seg0 = blockIdx.x * blockDim.x + threadIdx.x;
	if (seg0 >= G.count[CNT_SEGS]) return;
	const int ord = (int) G.ord[seg0];
	if (ord & 15) return;
	const int cnt = (ord >> 4) + 1, q = G.q[seg0];
	int trapped = 0;
	long long cell = 0;
	for (; trapped < cnt; ++trapped) {
		const int seg = seg0 + trapped, first = seg * K;
		const unsigned vmask = pipe_bits(G.vbits, first, K), umask = pipe_bits(G.ubits, first, K);
		cell = (long long) (G.idx0[seg] / (unsigned long long) K);
		if (vmask | umask) {
			const int h = atomicAdd(G.count + CNT_HEAVY, 1);
			const double2 *src = reinterpret_cast<const double2 *>(G.rows + (size_t) seg * PIPE_ROW);
			double2 *dst = reinterpret_cast<double2 *>(H.rows + (size_t) h * PIPE_ROW);
#pragma unroll
			for (int d = 0; d < PIPE_ROW / 2; ++d) dst[d] = src[d];
			H.q[h] = q; H.near[h] = G.near[seg]; H.flags[h] = G.flags[seg]; H.idx0[h] = G.idx0[seg]; H.vmask[h] = vmask; H.umask[h] = umask;
			S.busy_until[q] = round + PIPE_DEPTH;
			if (diag) { atomicAdd(diag + 2, 1); if (vmask) atomicAdd(diag + 3, 1); atomicAdd(diag + 1, cnt - trapped - 1); }
			break;
		}
	}
	if (trapped == cnt) cell += 1;  // all TRAPPED: on to the half after the last one; otherwise select / connect move the query on from the heavy half
	if (diag) atomicAdd(diag, trapped);
	S.pair_checks[q] += (long long) K * trapped;
	S.nn_queries[q] += trapped;
	S.it[q] = (int) (cell >> 1); S.half[q] = (int) (cell & 1);
}

template <typename M>
__global__ void __launch_bounds__(128, 4) k_pipe_select(TerrainView Tv, PipeState S, PlanArena A, PipeHeavy H, int *__restrict__ cnt, uint64_t seed,
													 uint64_t query0, gbp_plan_params P, int *diag) {
	const int lane = threadIdx.x & 31;
	const int nheavy = cnt[CNT_HEAVY], warps = (gridDim.x * blockDim.x) >> 5;
	for (int h = (int) ((blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5); h < nheavy; h += warps) {
	const int q = H.q[h], K = P.k_candidates;
	unsigned vmask = H.vmask[h];
	const unsigned umask = H.umask[h];
	const double *row = H.rows + (size_t) h * PIPE_ROW;
	double s_near[8], R[9], s_rand[8];
#pragma unroll
	for (int d = 0; d < 8; ++d) { s_near[d] = row[d]; s_rand[d] = row[17 + d]; }
#pragma unroll
	for (int d = 0; d < 9; ++d) R[d] = row[8 + d];
	const int dir = (int) H.flags[h] & 1, half = dir == GBP_FORWARD ? 0 : 1, near = H.near[h];
	const uint64_t stream = query0 + (uint64_t) q, idx0 = H.idx0[h];
	const bool dirs = P.action_direction_sampling != 0;
	const double *a_from = dir == GBP_FORWARD ? s_near : s_rand, *a_to = dir == GBP_FORWARD ? s_rand : s_near;
	// first-valid selection only needs the undecided candidates that come before the first valid one
	const unsigned um0 = (!P.best_of_k && vmask) ? (umask & ((vmask & (0u - vmask)) - 1u)) : umask;
	for (unsigned um = um0; um; um &= um - 1) {
		// a candidate the mixed-precision walk could not decide: the pair check again with the exact evaluators, its
		// sub-states spread over the lanes (validate_pair_warp: same verdict as the sequential walk)
		const int j = __ffs(um) - 1;
		double a[10], sn[8], tn;
		sample_action(seed, stream, idx0 + (uint64_t) j, R, dirs, P.action_direction_threshold, a_from, a_to, a);
		Counters c = {0, 0, 0, 0};
		if (validate_pair_warp<M>(Tv, s_near, a, dir, sn, tn, c)) vmask |= 1u << j;
	}
	long long pair_checks;
	const long long nn_queries = 1;
	bool found = false;
	double sn[8], a[10];
	const double best0 = state_distance(s_near, s_rand);
	if (!P.best_of_k) {
		const int j = vmask ? __ffs(vmask) - 1 : -1;  // the first valid action decides (rrt.cpp:44-47)
		pair_checks = j < 0 ? K : j + 1;
		if (j >= 0) {
			sample_action(seed, stream, idx0 + (uint64_t) j, R, dirs, P.action_direction_threshold, a_from, a_to, a);
			finish_output(s_near, a, dir == GBP_FORWARD ? OUT_LAND : OUT_REV, 0.0, sn);
			found = state_distance(sn, s_rand) < best0;  // rrt.cpp:55-66
		}
	} else {
		pair_checks = K;
		double my_d = INFINITY, my_sn[8], my_a[10];
		int my_j = 0x7fffffff;
		if (lane < K && ((vmask >> lane) & 1u)) {
			sample_action(seed, stream, idx0 + (uint64_t) lane, R, dirs, P.action_direction_threshold, a_from, a_to, my_a);
			finish_output(s_near, my_a, dir == GBP_FORWARD ? OUT_LAND : OUT_REV, 0.0, my_sn);
			my_d = state_distance(my_sn, s_rand);
			my_j = lane;
		}
		double bd = my_d;
		int bj = my_j;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const double od = __shfl_xor_sync(FULL, bd, o);
			const int oj = __shfl_xor_sync(FULL, bj, o);
			if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
		}
		if (bj != 0x7fffffff && bd < best0) {
			found = true;
#pragma unroll
			for (int d = 0; d < 8; ++d) sn[d] = __shfl_sync(FULL, my_sn[d], bj);
#pragma unroll
			for (int d = 0; d < 10; ++d) a[d] = __shfl_sync(FULL, my_a[d], bj);
		}
	}
	if (diag && found && lane == 0) atomicAdd(diag, 1);
	if (found && lane == 0) {
		PlanTree Tx = arena_tree(A, q, half, (half == 0 ? S.na : S.nb) + q);
		plan_push(Tx, near, sn, a);  // rrt.cpp:87-92
		H.connects[atomicAdd(cnt + CNT_CONNECTS, 1)] = q * 2 + half;  // connect from the other tree (rrt_connect.cpp:261, :296): k_pipe_connect, next on this stream
	}
	if (lane == 0) {
		S.pair_checks[q] += pair_checks;
		S.nn_queries[q] += nn_queries;
		if (!found) {  // TRAPPED: on to the next half (a query that extended is moved on by k_pipe_connect)
			if (half == 0) S.half[q] = 1;
			else { S.it[q] += 1; S.half[q] = 0; }
		}
	}
	__syncwarp();
	}
}

// connect (rrt_connect.cpp:98-120) for the queries whose tree grew in round `round`: warps pull requests from the list
template <typename M>
__global__ void __launch_bounds__(128, 4) k_pipe_connect(TerrainView Tv, PipeState S, PlanArena A, PipeHeavy H, const int *__restrict__ cnt, gbp_plan_params P,
														  int *solved_count) {
	const int lane = threadIdx.x & 31;
	const int n = cnt[CNT_CONNECTS];
	const int warps = (gridDim.x * blockDim.x) >> 5;
	for (int e = (int) ((blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5); e < n; e += warps) {
		const int code = H.connects[e], q = code >> 1, half = code & 1;
		PlanTree Ta = arena_tree(A, q, 0, S.na + q), Tb = arena_tree(A, q, 1, S.nb + q);
		PlanTree &Tx = half == 0 ? Ta : Tb, &Ty = half == 0 ? Tb : Ta;
		int na = S.na[q], nb = S.nb[q];
		const int nx = half == 0 ? na : nb;
		int &ny = half == 0 ? nb : na;
		double s_new[8];
		tree_get(Tx.t, nx - 1, s_new);
		long long pair_checks = 0;
		const bool solved = warp_connect<M>(Tv, Ty, ny, s_new, half == 0 ? GBP_REVERSE : GBP_FORWARD, P, lane, pair_checks) == GBP_REACHED;
		if (lane == 0) {
			S.pair_checks[q] += pair_checks;
			S.nn_queries[q] += 1;
			const int it = S.it[q];
			if (solved) { S.status[q] = 1; S.iters[q] = it + 1; if (P.stop_after_solved > 0) atomicAdd(solved_count, 1); }
			else if (half == 0) S.half[q] = 1;
			else { S.it[q] = it + 1; S.half[q] = 0; }
		}
		__syncwarp();
	}
}

// The tail of a batch: the queries still running when most are done are the ones whose trees keep growing, and in the
// pipeline every vertex costs them a round of their own plus the rounds they sit out (~0.3 ms), while a round costs ~0.1 ms
// however few queries are left — 1,100 of the 1,900 rounds of configs[4] ran for a few thousand such queries.  Once no more
// than `resume_at` queries are running, the side streams are drained and each of them continues from its (iteration, half,
// trees, counters) on a warp of its own with the megakernel's loop (k_plan_batch, gbp_planner.cuh: same cells, same
// arithmetic, same order of tree updates), where an extend that grows the tree costs what a TRAPPED one does.
template <typename M>
__global__ void __launch_bounds__(128, GBP_PLAN_MINBLOCKS) k_pipe_resume(TerrainView Tv, PipeState S, PlanArena A, int64_t Q, uint64_t seed, uint64_t query0,
																		  gbp_plan_params P, unsigned long long *__restrict__ next_query, int *solved_count) {
	const int lane = threadIdx.x & 31;
	const GroupMap gm = make_group_map(P.k_candidates, lane);
	while (true) {
		unsigned long long grabbed = 0;
		if (lane == 0) grabbed = atomicAdd(next_query, 1ull);
		const int64_t q = (int64_t) __shfl_sync(FULL, grabbed, 0);
		if (q >= Q) break;
		if (S.status[q] != 0) continue;
		PlanTree Ta = arena_tree(A, (int) q, 0, S.na + q), Tb = arena_tree(A, (int) q, 1, S.nb + q);
		int na = S.na[q], nb = S.nb[q], it = S.it[q], half0 = S.half[q];
		bool solved = false, full = false;
		double rs[8];             // this lane's random state of the current batch of 32 STATE cells
		unsigned rs_valid = 0;
		long long drawn = -1;     // first cell of that batch
		long long pair_checks = 0, nn_queries = 0;
		const uint64_t query = query0 + (uint64_t) q;
		for (; it < P.max_iters && !solved && !full; ++it) {
			if (P.stop_after_solved > 0 && __shfl_sync(FULL, *(volatile int *) solved_count, 0) >= P.stop_after_solved) break;
			for (int half = half0; half < 2 && !solved; ++half) {
				PlanTree &Tx = half == 0 ? Ta : Tb, &Ty = half == 0 ? Tb : Ta;
				int &nx = half == 0 ? na : nb, &ny = half == 0 ? nb : na;
				if (nx >= A.cap || ny >= A.cap) { full = true; break; }
				const uint64_t cell = 2 * (uint64_t) it + (uint64_t) half;
				const long long base = (long long) (cell & ~31ull);
				if (base != drawn) {
					sample_state<M>(Tv, seed, query, (uint64_t) base + (uint64_t) lane, false, 0.0, false, nullptr, nullptr, rs);
					Counters c = {0, 0, 0, 0};
					rs_valid = __ballot_sync(FULL, is_valid_state_auto<M>(Tv, pose6(rs), GBP_STANCE, c));
					drawn = base;
				}
				const int src = (int) (cell & 31ull);
				if (!((rs_valid >> src) & 1u)) continue;  // rrt_connect.cpp:254
				double s_rand[8];
#pragma unroll
				for (int d = 0; d < 8; ++d) s_rand[d] = __shfl_sync(FULL, rs[d], src);
				++nn_queries;
				if (warp_extend<M, false>(Tv, Tx, nx, s_rand, half == 0 ? GBP_FORWARD : GBP_REVERSE, seed, query, cell, P, gm, lane, pair_checks) == GBP_TRAPPED) continue;
				double s_new[8];
				tree_get(Tx.t, nx - 1, s_new);
				++nn_queries;
				if (warp_connect<M>(Tv, Ty, ny, s_new, half == 0 ? GBP_REVERSE : GBP_FORWARD, P, lane, pair_checks) == GBP_REACHED) solved = true;
			}
			half0 = 0;
		}
		if (lane == 0) {
			S.pair_checks[q] += pair_checks;
			S.nn_queries[q] += nn_queries;
			S.status[q] = solved ? 1 : 2;
			S.iters[q] = it;
			if (solved && P.stop_after_solved > 0) atomicAdd(solved_count, 1);
		}
		__syncwarp();
	}
}

template <typename M>
__global__ void __launch_bounds__(128) k_pipe_finish(TerrainView Tv, PipeState S, PlanArena A, PlanArena scratch, int64_t Q, gbp_plan_params P,
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

inline bool plan_pipe_applies(const TerrainView &Tv, const gbp_plan_params &P, int64_t nq) {
	const char *mode = getenv("GBP_PLAN_MODE");  // "mega" / "pipe": force one form (A/B measurements, tests); results are identical
	if (P.rrt_star || P.adaptive || P.state_direction_sampling || P.k_candidates > 32 || !Tv.mixed_ok || !Tv.uniform) return false;
	if (mode && !strcmp(mode, "pipe")) return true;
	if (mode && (!strcmp(mode, "mega") || !strcmp(mode, "step"))) return false;
	// a round costs ~0.1 ms whatever the batch size: below ~3 k queries the megakernel's independent warps win (configs[4] queries,
	// megakernel against this form with 8 speculated halves per round and the tail on k_pipe_resume: 2,048 queries 0.042 / 0.041 s,
	// 4,096: 0.063 / 0.045 s, 8,192: 0.114 / 0.056 s, 16,384: 0.218 / 0.079 s)
	if (P.stop_after_solved > 0) return nq >= 2048;  // anytime use (many attempts at one query): depth decides, and 16 speculated halves per round are a 4x shorter iteration than a megakernel warp's (tools/ttfs_sweep.py)
	return nq >= 4096;
}

// What one pipeline needs on the host besides its main stream.  Pooled (gbp_capi_pipeline.cu): a call takes one per group of
// queries it runs concurrently and gives it back.
struct PipeHostRes {
	int *h_count = nullptr;
	cudaStream_t main = nullptr, sb = nullptr, sc = nullptr, sd = nullptr;  // main: the stream of a group that does not run on the caller's
	cudaEvent_t ev_tri[PIPE_DEPTH] = {}, ev_con[PIPE_DEPTH] = {}, ev_prep[PIPE_DEPTH] = {}, ev_bat[PIPE_DEPTH] = {}, ev_sel[PIPE_DEPTH] = {}, done = nullptr;
	cudaError_t create() {
		cudaError_t e = cudaHostAlloc((void **) &h_count, CNT_WORDS * sizeof(int), cudaHostAllocDefault);
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&main, cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
		for (int k = 0; k < PIPE_DEPTH && e == cudaSuccess; ++k) {
			e = cudaEventCreateWithFlags(&ev_tri[k], cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_sel[k], cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_con[k], cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_prep[k], cudaEventDisableTiming);
			if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_bat[k], cudaEventDisableTiming);
		}
		return e;
	}
};

template <typename M>
inline int plan_pipe_launch_kind(const TerrainView &Tv, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
								 const gbp_plan_params &P, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap,
								 cudaStream_t st, const PlanTreeDump &dump, PipeHostRes &R, std::string &err) {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const size_t cap = (size_t) P.max_vertices, Q = (size_t) nq, per = Q * 2 * cap, K = (size_t) P.k_candidates;
	const int64_t fin_slots = (int64_t) sms * 4 * 4;
	// segments a query may emit per round (see k_pipe_prep): GBP_PIPE_SPEC = 1, 2, 4, 8 or 16
	// measured on configs[4] (4-8 m, 2000 iterations; profiles/r2b_spec_sweep.txt, r2b_resume_sweep.txt), with the tail on
	// k_pipe_resume, B = 2 / 4 / 8 / 16: 65,536 queries 0.260 / 0.2266 / 0.2226 / - s, 32,768 queries - / 0.1356 / 0.1257 / - s,
	// 16,384 queries - / 0.088 / 0.086 / 0.080 s, 4,096 queries - / - / 0.046 / 0.043 s
	int B = nq >= 32768 ? 8 : 16;
	if (const char *b = getenv("GBP_PIPE_SPEC")) B = atoi(b);
	B = B >= 16 ? 16 : B >= 8 ? 8 : B >= 4 ? 4 : B >= 2 ? 2 : 1;
	const size_t QB = Q * (size_t) B;
	const size_t bit_words = ((QB * K + 31) / 32 + 2 + 3) & ~(size_t) 3;  // a multiple of 16 bytes: vbits, ubits (and the counter blocks before them) stay adjacent in the arena
	PlanArena A = {}, Sc = {};
	PipeState S;
	PipeSegs G;
	PipeHeavy H[PIPE_DEPTH];
	A.cap = Sc.cap = P.max_vertices;
	unsigned long long *next_query = nullptr;
	int *batch_list = nullptr, *cnt = nullptr, *solved_count = nullptr;
	// one arena per call: laid out twice by the same code, first to size it, then over the allocation
	auto layout = [&](char *base) -> size_t {
		size_t off = 0;
		auto take = [&](auto *&ptr, size_t n) {
			typedef typename std::remove_reference<decltype(*ptr)>::type T;
			off = (off + 15) & ~(size_t) 15;
			ptr = reinterpret_cast<T *>(base + off);
			off += n * sizeof(T);
		};
		take(A.v, per * 8); take(A.act, per * 10); take(A.g, per); take(A.y, per);
		take(A.parent, per); take(A.child, per); take(A.sibling, per);
		take(Sc.pstate, (size_t) fin_slots * 2 * cap * 8); take(Sc.paction, (size_t) fin_slots * 2 * cap * 10);
		take(S.rs, Q * 2 * PIPE_BATCH * 8); take(S.rs_valid, Q * 2); take(S.rs_base, Q * 2); take(S.rs_want, Q);
		take(S.pair_checks, Q); take(S.nn_queries, Q);
		take(S.na, Q); take(S.nb, Q); take(S.status, Q); take(S.it, Q); take(S.half, Q); take(S.iters, Q); take(S.busy_until, Q);
		take(S.root_valid, 2 * Q);
		take(G.rows, QB * PIPE_ROW); take(G.idx0, QB); take(G.q, QB); take(G.near, QB); take(G.flags, QB); take(G.ord, QB);
		for (int k = 0; k < PIPE_DEPTH; ++k) {
			take(H[k].rows, Q * PIPE_ROW); take(H[k].idx0, Q); take(H[k].q, Q); take(H[k].near, Q); take(H[k].vmask, Q); take(H[k].umask, Q);
			take(H[k].connects, Q); take(H[k].flags, Q);
		}
		take(next_query, 2);
		take(solved_count, 12);  // [0] solved queries (stop_after_solved), [4..7] GBP_PIPE_TRACE: TRAPPED segments, dropped, heavy, heavy with a valid candidate
		take(batch_list, PIPE_DEPTH * Q);  // queries that need a batch of random states, one list per slot
		take(cnt, PIPE_DEPTH * CNT_WORDS);  // the counter blocks, then the two bit arrays (one memset)
		take(G.vbits, bit_words); take(G.ubits, bit_words);
		return off + 64;
	};
	const size_t need = layout(nullptr);
	void *mem = nullptr;
	cudaError_t e;
	const auto host_t0 = std::chrono::steady_clock::now();
	auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
	if ((e = cudaMallocAsync(&mem, need, st)) != cudaSuccess) { err = std::string("pipelined planner arena: ") + cudaGetErrorString(e); return GBP_E_CUDA; }
	const double host_alloc_ms = host_ms();
	layout((char *) mem);
	// host-side resources of this pipeline (a pinned word pair for the running count, the two side streams and their events)
	int *const h_count = R.h_count;
	const cudaStream_t sb = R.sb, sc = R.sc, sd = R.sd;
	cudaEvent_t *const ev_tri = R.ev_tri, *const ev_con = R.ev_con, *const ev_prep = R.ev_prep, *const ev_bat = R.ev_bat, *const ev_sel = R.ev_sel;
	int walk_ctas = GBP_WALK_CTAS;
	if (const char *w = getenv("GBP_PIPE_WALK_CTAS")) walk_ctas = atoi(w) > 0 ? atoi(w) : walk_ctas;
	const unsigned walk_grid = (unsigned) sms * walk_ctas;
	k_pipe_init<M><<<(unsigned) ((Q + 127) / 128), 128, 0, st>>>(Tv, S, A, nq, starts, goals, P.max_iters);
	cudaMemsetAsync(cnt, 0, (PIPE_DEPTH * CNT_WORDS + 2 * bit_words) * sizeof(int), st);
	cudaMemsetAsync(solved_count, 0, 12 * sizeof(int), st);
	// A round runs up to B extend attempts of every running query; a query needs at most 2 * max_iters of them (every random
	// state valid) plus the rounds it sits out.  The number of running queries is read back every few rounds to stop launching once
	// all are done.
	//   stream st: prep, walk, triage of round r.
	//   stream sd: the batches of random states requested by prep(r) (needed by prep(r + 1)).
	//   stream sb: select for the few queries of round r that have a valid or undecided candidate.
	//   stream sc: connect for the queries whose tree grew in select(r).
	//              Select and connect overlap the next rounds, which their queries sit out, and are over before
	//              prep(r + PIPE_DEPTH).
	const int max_rounds = 4 * P.max_iters + 4;
	const int check_mask = P.stop_after_solved > 0 ? 3 : B > 1 ? 7 : 31;  // rounds between two looks at the running count (a speculating round is several times longer)
	// queries left when the rest of the batch moves to k_pipe_resume: one wave of its warps at most, an eighth of the batch at most
	int resume_at = (int) std::min<int64_t>((int64_t) sms * 4 * GBP_PLAN_MINBLOCKS, nq / 8);
	if (const char *r = getenv("GBP_PIPE_RESUME")) resume_at = atoi(r);
	bool resume = false;
	// the side kernels are warp-per-item and latency-bound: as many resident warps as their registers allow (80 / 128 per thread)
	int side_ctas_batch = 6, side_ctas_connect = 4;
	if (const char *c = getenv("GBP_PIPE_SIDE_CTAS")) { side_ctas_batch = atoi(c) > 0 ? atoi(c) : 4; side_ctas_connect = atoi(c) > 0 ? (atoi(c) + 1) / 2 : 2; }
	int round = 0;
	// GBP_PIPE_TRACE=1: device timestamps of 32 consecutive rounds (from round GBP_PIPE_TRACE on, default 640 / 96 when speculating) on the first stream, printed to stderr
	const bool trace = getenv("GBP_PIPE_TRACE") != nullptr;
	const int tr0 = trace && atoi(getenv("GBP_PIPE_TRACE")) > 1 ? atoi(getenv("GBP_PIPE_TRACE")) : (B > 1 ? 96 : 640);
	cudaEvent_t tr[32][4];
	if (trace) for (auto &r4 : tr) for (auto &ev : r4) cudaEventCreate(&ev);
	for (; round < max_rounds; ++round) {
		const int par = round % PIPE_DEPTH, prev = (round + PIPE_DEPTH - 1) % PIPE_DEPTH;
		const bool tr_on = trace && round >= tr0 && round < tr0 + 32;
		if (tr_on) cudaEventRecord(tr[round - tr0][0], st);
		G.count = cnt + par * CNT_WORDS;
		if (round >= PIPE_DEPTH) cudaStreamWaitEvent(st, ev_con[par], 0);  // select + connect of round - PIPE_DEPTH: its queries, heavy buffers and counters are free again
		if (round >= 1) cudaStreamWaitEvent(st, ev_bat[prev], 0);  // the batches drawn for the previous round's requests
		if (round >= PIPE_DEPTH) cudaMemsetAsync(G.count, 0, CNT_WORDS * sizeof(int), st);
		if (round >= 1) cudaMemsetAsync(G.vbits, 0, 2 * bit_words * sizeof(int), st);  // the candidate bits of the previous round (vbits and ubits are adjacent)
		if (tr_on) cudaEventRecord(tr[round - tr0][1], st);
		k_pipe_prep<M><<<(unsigned) ((4 * QB + 127) / 128), 128, 0, st>>>(Tv, S, A, G, batch_list + (size_t) par * Q, nq, P, round, B, solved_count);
		cudaEventRecord(ev_prep[par], st);
		if (tr_on) cudaEventRecord(tr[round - tr0][2], st);
		// (queued before or after the walk makes no difference: its blocks get their SMs as walk CTAs retire)
		cudaStreamWaitEvent(sd, ev_prep[par], 0);
		k_pipe_batch<M><<<(unsigned) sms * side_ctas_batch, 128, 0, sd>>>(Tv, S, batch_list + (size_t) par * Q, G.count, seed, query0);
		cudaEventRecord(ev_bat[par], sd);
		if (Tv.ztex) k_walk_seg<true><<<walk_grid, RF_WARPS * 32, 0, st>>>(Tv, G, P.k_candidates, seed, query0, P.action_direction_sampling, P.action_direction_threshold);
		else k_walk_seg<false><<<walk_grid, RF_WARPS * 32, 0, st>>>(Tv, G, P.k_candidates, seed, query0, P.action_direction_sampling, P.action_direction_threshold);
		k_pipe_triage<<<(unsigned) ((QB + 255) / 256), 256, 0, st>>>(S, G, H[par], P.k_candidates, round, trace ? solved_count + 4 : nullptr);
		cudaEventRecord(ev_tri[par], st);
		if (tr_on) cudaEventRecord(tr[round - tr0][3], st);
		cudaStreamWaitEvent(sb, ev_tri[par], 0);
		k_pipe_select<M><<<(unsigned) sms * 4, 128, 0, sb>>>(Tv, S, A, H[par], G.count, seed, query0, P, trace ? solved_count + 8 : nullptr);
		cudaEventRecord(ev_sel[par], sb);
		cudaStreamWaitEvent(sc, ev_sel[par], 0);
		k_pipe_connect<M><<<(unsigned) sms * side_ctas_connect, 128, 0, sc>>>(Tv, S, A, H[par], G.count, P, solved_count);
		cudaEventRecord(ev_con[par], sc);
		if ((round & check_mask) == check_mask) {
			cudaMemcpyAsync(h_count, G.count, CNT_WORDS * sizeof(int), cudaMemcpyDeviceToHost, st);
			if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
			if (h_count[CNT_RUNNING] == 0) { ++round; break; }
			if (h_count[CNT_RUNNING] <= resume_at) { ++round; resume = true; break; }
		}
	}
	if (trace) {
		cudaStreamSynchronize(st);
		if (round >= tr0 + 32) {
			float wait = 0, prep = 0, rest = 0, gap = 0, t;
			for (int r = 0; r < 32; ++r) {
				cudaEventElapsedTime(&t, tr[r][0], tr[r][1]); wait += t;
				cudaEventElapsedTime(&t, tr[r][1], tr[r][2]); prep += t;
				cudaEventElapsedTime(&t, tr[r][2], tr[r][3]); rest += t;
				if (r + 1 < 32) { cudaEventElapsedTime(&t, tr[r][3], tr[r + 1][0]); gap += t; }
			}
			fprintf(stderr, "pipe trace (us per round): waits+memset %.1f, prep %.1f, walk+triage %.1f, round-to-round gap %.1f\n", wait / 32 * 1e3,
					prep / 32 * 1e3, rest / 32 * 1e3, gap / 31 * 1e3);
		}
		for (auto &r4 : tr) for (auto &ev : r4) cudaEventDestroy(ev);
		int d[5] = {};
		cudaMemcpy(d, solved_count + 4, sizeof d, cudaMemcpyDeviceToHost);
		fprintf(stderr, "pipe trace (host): arena of %.2f GB allocated in %.2f ms, rounds issued and drained after %.1f ms\n", need / 1e9, host_alloc_ms, host_ms());
		fprintf(stderr, "pipe trace: %d rounds, B = %d; segments TRAPPED at triage %d, heavy %d (%d with a valid candidate, the others undecided only), dropped behind a heavy one %d; heavy segments that grew a tree %d\n",
				round, B, d[0], d[2], d[3], d[1], d[4]);
	}
	for (int k = 0; k < PIPE_DEPTH && k < round; ++k) {  // the last batches / selects / connects precede the statistics
		cudaStreamWaitEvent(st, ev_con[k], 0);
		cudaStreamWaitEvent(st, ev_bat[k], 0);
	}
	if (resume && e == cudaSuccess) {
		cudaEvent_t r0 = nullptr, r1 = nullptr;
		if (trace) { cudaEventCreate(&r0); cudaEventCreate(&r1); cudaEventRecord(r0, st); }
		cudaMemsetAsync(next_query, 0, sizeof(unsigned long long), st);
		k_pipe_resume<M><<<(unsigned) sms * GBP_PLAN_MINBLOCKS, 128, 0, st>>>(Tv, S, A, nq, seed, query0, P, next_query, solved_count);
		if (trace) {
			cudaEventRecord(r1, st);
			cudaEventSynchronize(r1);
			float ms = 0;
			cudaEventElapsedTime(&ms, r0, r1);
			fprintf(stderr, "pipe trace: k_pipe_resume took over %d running queries after round %d: %.2f ms\n", h_count[CNT_RUNNING], round, ms);
			cudaEventDestroy(r0); cudaEventDestroy(r1);
		}
	}
	cudaMemsetAsync(next_query, 0, sizeof(unsigned long long), st);
	const int64_t fin_warps = (int64_t) Q < fin_slots ? (int64_t) ((Q + 3) / 4 * 4) : fin_slots;
	k_pipe_finish<M><<<(unsigned) (fin_warps / 4), 128, 0, st>>>(Tv, S, A, Sc, nq, P, next_query, stats, path_states, path_actions, path_cap, dump);
	e = cudaGetLastError();
	cudaFreeAsync(mem, st);
	if (e != cudaSuccess) { err = std::string("pipelined planner: ") + cudaGetErrorString(e); return GBP_E_CUDA; }
	return GBP_OK;
}

}  // namespace gbp
