// Sample + validate: the unit of work of RRTClass::newConfig (rrt.cpp:34-50) — getRandomAction, then
// isValidStateActionPair[Reverse] from a tree vertex — for n candidates at once, behind a NARROW wire format
// (the B_io = 77 B form):
//   in   per candidate: a 4-byte row number into a device-resident table of start states (the vertices candidates start
//        from) and a direction byte; per call: the Philox cell range (seed, stream, idx0) and the surface normal
//   out  one verdict BIT per candidate, and — for the valid candidates only, in ascending candidate order — the row
//        {candidate index, s_new, t_new, action}: everything newConfig reads afterwards (rrt.cpp:44-62 uses s_test of a
//        valid pair check only)
// against 145 B in / 74 B out per candidate of the full-fidelity call (gbp_validate_pairs).
//
// k_walk_sv is k_walk_mixed (gbp_walk.cuh: lane per candidate, warp-level refill, polynomial-segment cursor,
// mixed-precision evaluator, fp64 redo list) with a different front and back end:
//   * front: a warp PRODUCES its next 32 candidates convergently — lane j fetches the row number of candidate j, starts
//     the asynchronous gather of its 64-byte state row into the warp's shared-memory ring (cp.async, L2 only, evict-first:
//     the table is read once and must not displace the height grid) and samples ACTION cell idx0 + j while the gather is
//     in flight (the ~900 instructions of Philox + Box-Muller + force rotation hide the HBM round trip).  Sampling at
//     refill time instead would run the same code at ~8 active lanes.
//   * back: a finished lane writes nothing unless its candidate is valid (0.3-0.5 % are): one atomicOr on the verdict
//     word.  Exact s_new / t_new / action rows are produced afterwards for the valid candidates only (k_sv_outputs),
//     so neither the walk nor a second pass streams 64-byte rows for candidates nobody reads.
// Verdict bits, valid rows and the k / L work counters are identical to the dense path's (tests/test_gpu_sample_validate.py).
#pragma once
#include "gbp_walk.cuh"

namespace gbp {

struct SvParams {
	const double *table;     // [rows][8] start states, AoS, 16-byte aligned (device)
	const int *state_idx;    // [n] row numbers or null: candidate i starts from row row0 + i
	const uint8_t *dir;      // [n] directions or null: every candidate uses dir0 (or carries its own: dir_packed)
	int dir_packed;          // state_idx[i] = row | direction << 31: one 4-byte word per candidate on the wire (dir is null then)
	long long row0;
	long long rows;          // table size: a row number outside [0, rows) is counted in cnt[6] and read as row 0
	int dir0;
	uint64_t seed, stream, idx0;
	double normal[3];        // rrt.cpp:25: one surface normal per newConfig
	int states_valid;        // the caller's promise that every table row is a valid STANCE state (tree vertices are): see below
	int dir_sampling;        // getRandomAction(..., flag, threshold, s, s_near) (planning_utils.cpp:379-391)
	double dir_thresh;
	double target[8];        // `s` of newConfig (directional sampling only)
};

// row number and direction of candidate i as the call's wire format carries them.  STREAM = true: read-once loads (the
// walks); a row outside the table is reported through `bad` and read as row 0.
template <bool STREAM>
__device__ __forceinline__ void sv_row_dir(const SvParams &P, int64_t i, long long &row, int &dir, bool &bad) {
	dir = P.dir0;
	if (P.state_idx) {
		const int raw = STREAM ? __ldcs(P.state_idx + i) : P.state_idx[i];
		if (P.dir_packed) { row = (long long) (raw & 0x7fffffff); dir = (int) ((unsigned) raw >> 31); }
		else row = (long long) raw;
	} else row = P.row0 + i;
	if (P.dir) dir = (int) (STREAM ? __ldcs(P.dir + i) : P.dir[i]);
	bad = (unsigned long long) row >= (unsigned long long) P.rows;
	if (bad) row = 0;
}

constexpr int SV_CAP = 40;  // ring entries per warp: a batch of 32 is produced when at most SV_CAP - 32 are still buffered

__device__ __forceinline__ void cp_async16_hint(void *dst, const void *src, uint64_t policy) {
	asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ACTION cell `cell` with the directional option resolved for a candidate that starts from a state whose x / y
// velocities are (vx, vy): FORWARD samples towards the target from s_near, REVERSE the other way round (:385-388).
// (tvx, tvy) = the target's velocities.  Scalars, not the parameter block: a reference to a __global__ parameter would
// make the compiler copy the block to local memory.
__device__ __forceinline__ void sv_sample(uint64_t seed, uint64_t stream, uint64_t cell, const double *R, int dir_sampling, double dir_thresh,
										  double tvx, double tvy, int dir, double vx, double vy, double a[10]) {
	double sf[8], st[8];
	if (dir_sampling) {
#pragma unroll
		for (int d = 0; d < 8; ++d) { sf[d] = 0; st[d] = 0; }
		if (dir == GBP_FORWARD) { sf[3] = vx; sf[4] = vy; st[3] = tvx; st[4] = tvy; }
		else { st[3] = vx; st[4] = vy; sf[3] = tvx; sf[4] = tvy; }
	}
	sample_action(seed, stream, cell, R, dir_sampling != 0, dir_thresh, sf, st, a);
}
// out of line: called once per 32 candidates with all lanes active; keeps the sampler's registers out of the walk loop
static __device__ __noinline__ void sv_sample_to_ring(uint64_t seed, uint64_t stream, uint64_t cell, const double *R, int dir_sampling, double dir_thresh,
												double tvx, double tvy, int dir, double vx, double vy, double *dst) {
	double a[10];
	sv_sample(seed, stream, cell, R, dir_sampling, dir_thresh, tvx, tvy, dir, vx, vy, a);
	double2 *o = reinterpret_cast<double2 *>(dst);
#pragma unroll
	for (int d = 0; d < 5; ++d) o[d] = make_double2(a[2 * d], a[2 * d + 1]);
}

template <bool TEX, bool ADAPTIVE>
__global__ void __launch_bounds__(RF_WARPS * 32, GBP_WALK_CTAS) k_walk_sv(TerrainView T, SvParams P, int n, int per_warp, unsigned *__restrict__ bits,
																		uint8_t *__restrict__ flags, unsigned long long *__restrict__ cnt,
																		int *__restrict__ redo_idx, unsigned long long *__restrict__ redo_count) {
	__shared__ __align__(16) double ringS[RF_WARPS][SV_CAP][8];
	__shared__ __align__(16) double ringA[RF_WARPS][SV_CAP][10];
	__shared__ __align__(16) double stash[8][RF_WARPS * 32];
	__shared__ uint8_t ringD[RF_WARPS][SV_CAP];
	__shared__ double sR[9];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	if (threadIdx.x == 0) grf_rotation(P.normal, sR);
	__syncthreads();
	const int64_t warp = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
	const int64_t wbase64 = warp * (int64_t) per_warp;  // per_warp is a multiple of 32: a warp owns whole verdict words
	const int wbase = (int) min(wbase64, (int64_t) n);
	const int total = (int) min((int64_t) n - wbase, (int64_t) per_warp);
	double *const st = &stash[0][threadIdx.x];
	const uint64_t pol = l2_evict_first_policy();
	int next = 0, filled = 0;  // candidates handed out / produced so far, relative to wbase (warp-uniform)
	WalkCursor q;
	q.phase = PH_IDLE;
	int mine = -1;
	unsigned k = 0, L = 0, np = 0, nvalid = 0;
	while (true) {
		const unsigned need = __ballot_sync(FULL, q.phase == PH_IDLE);
		if (need && next < total) {
			const int nidle = __popc(need);
			if (filled - next < nidle && filled < total && filled - next <= SV_CAP - 32) {
				// produce the next batch of 32: start the state gather, sample while it is in flight
				const int r = filled + lane;
				if (r < total) {
					const int e = r % SV_CAP;
					const int i = wbase + r;
					long long row;
					int dv;
					bool bad;
					sv_row_dir<true>(P, i, row, dv, bad);
					if (bad) atomicAdd(cnt + 6, 1ull);
					const double *src = P.table + 8 * row;
#pragma unroll
					for (int d = 0; d < 4; ++d) cp_async16_hint(&ringS[wib][e][2 * d], src + 2 * d, pol);
					ringD[wib][e] = (uint8_t) dv;
					double vx = 0, vy = 0;
					if (P.dir_sampling) { vx = __ldg(src + 3); vy = __ldg(src + 4); }
					sv_sample_to_ring(P.seed, P.stream, P.idx0 + (uint64_t) i, sR, P.dir_sampling, P.dir_thresh, P.target[3], P.target[4], dv, vx, vy, &ringA[wib][e][0]);
				}
				cp_async_wait_all();
				__syncwarp();
				filled = min(total, filled + 32);
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
					mine = wbase + rel;
					walk_start(q, s, a, (int) ringD[wib][e], st);
					// The reference's first sub-state is the start state itself (FORWARD: the stance sample at t = 0, :718-730;
					// REVERSE: the flight sample at t = 0, :842-848, and a STANCE-valid state is FLIGHT-valid: the flight check only
					// drops the leg-reach test).  For a row promised valid it is counted as the reference counts it (1 sub-state,
					// 5 NaN probes, 9 lookups) and not evaluated again: every candidate of a tree vertex repeats that same check.
					if (P.states_valid && q.t == 0 && (q.phase == PH_FWD_ST || q.phase == PH_REV_FL)) {
						OutRecipe dummy;
						q.c_.substates = 1; q.c_.nanprobes = 5; q.c_.lookups = 9;
						(void) walk_advance<ADAPTIVE>(q, true, dummy, st);
					}
				}
			}
			next = min(filled, next + nidle);
			__syncwarp();  // ring entries are read before a later batch may overwrite them
		}
		if (__ballot_sync(FULL, q.phase != PH_IDLE) == 0) break;
		bool valid = false, decided = true;
		if (q.phase != PH_IDLE) {
			const int ph = (q.phase == PH_FWD_FL || q.phase == PH_REV_FL) ? GBP_FLIGHT : GBP_STANCE;
			decided = is_valid_state_mixed<MapF32U, TEX>(T, walk_pose(q), ph, q.c_, valid);
		}
		if (!decided) {  // hand the whole candidate to the fp64 pass
			redo_idx[atomicAdd(redo_count, 1ull)] = mine;
			q.phase = PH_IDLE;
		}
		if (q.phase != PH_IDLE) {
			OutRecipe out;
			const int r = walk_advance<ADAPTIVE>(q, valid, out, st);
			if (r) {
				if (r == 2) atomicOr(bits + (mine >> 5), 1u << (mine & 31));
				if (flags) __stcs(flags + mine, (uint8_t) (r == 2 ? GBP_FLAG_VALID : 0));  // no OOG / NEAR flag can arise on this path
				k += q.c_.substates; L += q.c_.lookups; np += q.c_.nanprobes; nvalid += r == 2 ? 1 : 0;
				q.phase = PH_IDLE;
			}
		}
	}
	flush_counters(cnt, k, L, np, 0, 0, nvalid);
}

// k_walk_sv for inputs that are still ARRIVING: the host-pointer call (gbp_sample_validate) launches this kernel ONCE over
// the whole call and copies row numbers and directions block by block on a second stream, each block followed by a copy
// that sets the block's word in `arrived`.  Warps claim batches of 32 candidates from a counter, in ascending order, and
// wait for the block a batch lies in — after the first block the copies run ahead of the walk (5 B against ~0.17 ns of
// walk per candidate), so nothing waits again.  One launch instead of one per chunk: a launch ends in a tail of one long
// candidate (~50 us), 8-9 of them were 0.5 ms of a 3.6 ms call.  Verdicts and counters do not depend on which warp walks a
// candidate.  A block that never arrives (failed copy) ends the wait after ~2^24 polls and is reported in cnt[7].
constexpr int SV_BLOCK_SHIFT = 19;  // 524,288 candidates per arrival word
__device__ __forceinline__ unsigned sv_ld_acquire(const unsigned *p) {
	unsigned v;
	asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
template <bool TEX, bool ADAPTIVE>
__global__ void __launch_bounds__(RF_WARPS * 32, GBP_WALK_CTAS) k_walk_sv_stream(TerrainView T, SvParams P, int n, const unsigned *__restrict__ arrived,
																				   unsigned long long *__restrict__ work, unsigned *__restrict__ bits,
																				   uint8_t *__restrict__ flags, unsigned long long *__restrict__ cnt,
																				   int *__restrict__ redo_idx, unsigned long long *__restrict__ redo_count) {
	__shared__ __align__(16) double ringS[RF_WARPS][SV_CAP][8];
	__shared__ __align__(16) double ringA[RF_WARPS][SV_CAP][10];
	__shared__ __align__(16) double stash[8][RF_WARPS * 32];
	__shared__ uint8_t ringD[RF_WARPS][SV_CAP];
	__shared__ int ringI[RF_WARPS][SV_CAP];
	__shared__ double sR[9];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	if (threadIdx.x == 0) grf_rotation(P.normal, sR);
	__syncthreads();
	double *const st = &stash[0][threadIdx.x];
	const uint64_t pol = l2_evict_first_policy();
	int next = 0, filled = 0;  // candidates handed out / produced by this warp
	bool exhausted = false;
	WalkCursor q;
	q.phase = PH_IDLE;
	int mine = -1;
	unsigned k = 0, L = 0, np = 0, nvalid = 0;
	while (true) {
		const unsigned need = __ballot_sync(FULL, q.phase == PH_IDLE);
		if (need && (next < filled || !exhausted)) {
			const int nidle = __popc(need);
			if (filled - next < nidle && !exhausted && filled - next <= SV_CAP - 32) {
				long long base = 0;
				if (lane == 0) {
					base = (long long) atomicAdd(work, 32ull);
					if (base < n) {  // wait for the block of row numbers / directions this batch lies in
						unsigned spins = 0;
						while (sv_ld_acquire(arrived + (base >> SV_BLOCK_SHIFT)) == 0u)
							if (++spins > (1u << 24)) { atomicAdd(cnt + 7, 1ull); base = n; break; }
					}
				}
				base = __shfl_sync(FULL, base, 0);
				const int c = (int) min(32ll, (long long) n - base);
				if (c <= 0) exhausted = true;
				if (lane < c) {
					const int e = (filled + lane) % SV_CAP;
					const int i = (int) base + lane;
					long long row;
					int dv;
					bool bad;
					sv_row_dir<true>(P, i, row, dv, bad);
					if (bad) atomicAdd(cnt + 6, 1ull);
					const double *src = P.table + 8 * row;
#pragma unroll
					for (int d = 0; d < 4; ++d) cp_async16_hint(&ringS[wib][e][2 * d], src + 2 * d, pol);
					ringD[wib][e] = (uint8_t) dv;
					ringI[wib][e] = i;
					double vx = 0, vy = 0;
					if (P.dir_sampling) { vx = __ldg(src + 3); vy = __ldg(src + 4); }
					sv_sample_to_ring(P.seed, P.stream, P.idx0 + (uint64_t) i, sR, P.dir_sampling, P.dir_thresh, P.target[3], P.target[4], dv, vx, vy, &ringA[wib][e][0]);
				}
				cp_async_wait_all();
				__syncwarp();
				filled += max(c, 0);
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
					walk_start(q, s, a, (int) ringD[wib][e], st);
					if (P.states_valid && q.t == 0 && (q.phase == PH_FWD_ST || q.phase == PH_REV_FL)) {  // see k_walk_sv
						OutRecipe dummy;
						q.c_.substates = 1; q.c_.nanprobes = 5; q.c_.lookups = 9;
						(void) walk_advance<ADAPTIVE>(q, true, dummy, st);
					}
				}
			}
			next = min(filled, next + nidle);
			__syncwarp();  // ring entries are read before a later batch may overwrite them
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
		if (!decided) {  // hand the whole candidate to the fp64 pass
			redo_idx[atomicAdd(redo_count, 1ull)] = mine;
			q.phase = PH_IDLE;
		}
		if (q.phase != PH_IDLE) {
			OutRecipe out;
			const int r = walk_advance<ADAPTIVE>(q, valid, out, st);
			if (r) {
				if (r == 2) atomicOr(bits + (mine >> 5), 1u << (mine & 31));
				if (flags) __stcs(flags + mine, (uint8_t) (r == 2 ? GBP_FLAG_VALID : 0));
				k += q.c_.substates; L += q.c_.lookups; np += q.c_.nanprobes; nvalid += r == 2 ? 1 : 0;
				q.phase = PH_IDLE;
			}
		}
	}
	flush_counters(cnt, k, L, np, 0, 0, nvalid);
}

// the candidate's inputs, rebuilt from its index (redo pass, general path, output pass)
__device__ __forceinline__ void sv_candidate(const SvParams &P, const double *R, int64_t i, double s[8], double a[10], int &dir) {
	long long row;
	bool bad;  // counted by the pass that decides the candidate
	sv_row_dir<false>(P, i, row, dir, bad);
	load_state(P.table + 8 * row, s);
	sv_sample(P.seed, P.stream, P.idx0 + (uint64_t) i, R, P.dir_sampling, P.dir_thresh, P.target[3], P.target[4], dir, s[3], s[4], a);
}
// the fp64 evaluator walked to the verdict (k_validate_redo's loop)
template <typename M>
__device__ __forceinline__ bool sv_walk_fp64(const TerrainView &T, const double s[8], const double a[10], int dir, bool adaptive, Counters &c) {
	Cursor q;
#pragma unroll
	for (int d = 0; d < 8; ++d) q.s[d] = s[d];
#pragma unroll
	for (int d = 0; d < 10; ++d) q.a[d] = a[d];
	cursor_start(q, dir);
	OutRecipe out;
	int r = 0;
	while (!r) {
		Pose6 p;
		double tmp[8];
		switch (q.phase) {
		case PH_FWD_ST: p = stance_fast(q.s, q.a, q.f, q.t); break;
		case PH_FWD_FL:
		case PH_FWD_LAND: stance_fast8(q.s, q.a, q.f, q.a[6], tmp); p = flight_fast(tmp, q.phase == PH_FWD_FL ? q.t : q.a[7]); break;
		case PH_REV_FL: p = flight_fast(q.s, -q.t); break;
		default: apply_flight(q.s, -q.a[7], tmp); p = stance_reverse_fast(tmp, q.a, q.f, q.phase == PH_REV_ST ? q.t : 0.0); break;
		}
		const int ph = (q.phase == PH_FWD_FL || q.phase == PH_REV_FL) ? GBP_FLIGHT : GBP_STANCE;
		r = cursor_advance(q, is_valid_state_fast<M>(T, p, ph, q.c), adaptive, out);
	}
	c = q.c;
	return r == 2;
}
// Candidates listed in redo_idx (the ones the mixed-precision walk could not decide), or — list == null — every candidate
// of [0, n): the general path for terrains without the mixed evaluator (fp64 cells beyond the rounding budget, NaN cells,
// non-uniform axes).  One thread per candidate, fp64 evaluator (1e-11 m guard).
template <typename M>
__global__ void __launch_bounds__(128) k_sv_fp64(TerrainView T, SvParams P, int64_t n, const int *__restrict__ list,
												  const unsigned long long *__restrict__ list_count, int adaptive, unsigned *__restrict__ bits,
												  uint8_t *__restrict__ flags, unsigned long long *__restrict__ cnt) {
	__shared__ double sR[9];
	if (threadIdx.x == 0) grf_rotation(P.normal, sR);
	__syncthreads();
	const unsigned long long m = list ? *list_count : (unsigned long long) n;
	unsigned long long k = 0, L = 0, np = 0, oog = 0, near = 0, nvalid = 0;
	for (unsigned long long j = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; j < m; j += (unsigned long long) gridDim.x * blockDim.x) {
		const int64_t i = list ? (int64_t) list[j] : (int64_t) j;
		double s[8], a[10];
		int dir;
		if (!list) {  // the walk has counted the listed ones
			long long row;
			int d0;
			bool bad;
			sv_row_dir<false>(P, i, row, d0, bad);
			if (bad) atomicAdd(cnt + 6, 1ull);
		}
		sv_candidate(P, sR, i, s, a, dir);
		Counters c = {0, 0, 0, 0};
		const bool ok = sv_walk_fp64<M>(T, s, a, dir, adaptive != 0, c);
		if (ok) atomicOr(bits + (i >> 5), 1u << (i & 31));
		if (flags) flags[i] = (uint8_t) (c.flags | (ok ? GBP_FLAG_VALID : 0));
		k += c.substates; L += c.lookups; np += c.nanprobes;
		oog += (c.flags & GBP_FLAG_OOG) ? 1 : 0; near += (c.flags & GBP_FLAG_NEAR) ? 1 : 0; nvalid += ok ? 1 : 0;
	}
	flush_counters(cnt, k, L, np, oog, near, nvalid);
}

// ---- compaction of the verdict bits into the ascending list of valid candidates
// pass 1: block b counts the set bits of its span of words
static __global__ void __launch_bounds__(256) k_sv_count(const unsigned *__restrict__ bits, int64_t nwords, int64_t span, unsigned long long *__restrict__ sums) {
	__shared__ unsigned long long part[8];
	const int64_t lo = blockIdx.x * span, hi = min(nwords, lo + span);
	unsigned long long c = 0;
	for (int64_t w = lo + threadIdx.x; w < hi; w += blockDim.x) c += __popc(bits[w]);
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(FULL, c, o);
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long s = 0;
		for (int w = 0; w < 8; ++w) s += part[w];
		sums[blockIdx.x] = s;
	}
}
// pass 2: block b starts at the sum of the blocks before it and lists its set bits in ascending order; block 0 also
// publishes the call's result words {n_valid, sub-states k, lookups L, NaN probes, OOG, NEAR, rows out of range, 0}
static __global__ void __launch_bounds__(256) k_sv_list(const unsigned *__restrict__ bits, int64_t nwords, int64_t span, const unsigned long long *__restrict__ sums,
												  int64_t cap, int *__restrict__ index, const unsigned long long *__restrict__ cnt, long long *__restrict__ result) {
	__shared__ unsigned long long part[8];
	__shared__ unsigned long long s_base;
	__shared__ unsigned wsum[8];
	unsigned long long c = 0;
	const int nb = gridDim.x;
	for (int b = threadIdx.x; b < nb; b += blockDim.x) c += (b < (int) blockIdx.x || blockIdx.x == 0) ? sums[b] : 0ull;  // block 0 sums everything (the total)
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(FULL, c, o);
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long s = 0;
		for (int w = 0; w < 8; ++w) s += part[w];
		if (blockIdx.x == 0) {
			result[0] = (long long) s;
			for (int j = 0; j < 5; ++j) result[1 + j] = (long long) cnt[j];
			result[6] = (long long) cnt[6]; result[7] = (long long) cnt[7];  // [7]: batches of a streamed call whose inputs never arrived
			s = 0;
		}
		s_base = s;
	}
	__syncthreads();
	if (!index) return;
	unsigned long long base = s_base;
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int64_t lo = blockIdx.x * span, hi = min(nwords, lo + span);
	for (int64_t w0 = lo; w0 < hi; w0 += blockDim.x) {
		const int64_t w = w0 + threadIdx.x;
		unsigned word = w < hi ? bits[w] : 0u;
		const unsigned pc = __popc(word);
		unsigned incl = pc;  // inclusive scan over the warp
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned v = __shfl_up_sync(FULL, incl, o);
			if (lane >= o) incl += v;
		}
		if (lane == 31) wsum[wib] = incl;
		__syncthreads();
		unsigned before = 0, all = 0;
#pragma unroll
		for (int j = 0; j < 8; ++j) { before += j < wib ? wsum[j] : 0u; all += wsum[j]; }
		unsigned long long pos = base + before + (incl - pc);
		while (word) {
			const int b = __ffs(word) - 1;
			word &= word - 1;
			if ((long long) pos < cap) index[pos] = (int) (w * 32 + b);
			++pos;
		}
		base += all;
		__syncthreads();
	}
}
// pass 3: exact rows of the valid candidates: s_new = the landing state (FORWARD, planning_utils.cpp:743-749) or the exact
// start state (REVERSE, :866-872), t_new = t_s + t_f or t_s, and the sampled action — what rrt.cpp:44-62 goes on to use
static __global__ void __launch_bounds__(128) k_sv_outputs(SvParams P, const long long *__restrict__ result, int64_t cap, const int *__restrict__ index,
													 double *__restrict__ s_new, double *__restrict__ t_new, double *__restrict__ action) {
	__shared__ double sR[9];
	if (threadIdx.x == 0) grf_rotation(P.normal, sR);
	__syncthreads();
	const int64_t m = min((int64_t) result[0], cap);
	for (int64_t j = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; j < m; j += (int64_t) gridDim.x * blockDim.x) {
		const int64_t i = index[j];
		double s[8], a[10], sn[8];
		int dir;
		sv_candidate(P, sR, i, s, a, dir);
		finish_output(s, a, dir == GBP_FORWARD ? OUT_LAND : OUT_REV, 0.0, sn);
		if (s_new) store_state(s_new + 8 * j, sn);
		if (t_new) t_new[j] = dir == GBP_FORWARD ? a[6] + a[7] : a[6];
		if (action) {
			double2 *o = reinterpret_cast<double2 *>(action + 10 * j);
#pragma unroll
			for (int d = 0; d < 5; ++d) o[d] = make_double2(a[2 * d], a[2 * d + 1]);
		}
	}
}

}  // namespace gbp
