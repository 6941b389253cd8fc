// k_walk_mixed — the pair-check walk for fp32 maps on uniform axes (the kernel bench.py's headline number runs).
// Same contract as k_validate_refill<M> (gbp_kernels.cuh: lane per action, warp-level refill, TMA-staged candidate
// ring) with the mixed-precision isValidState alone — a candidate that reaches a sub-state it cannot decide is handed to
// the fp64 redo pass (k_validate_redo) — and the per-lane cursor rebuilt around what the walk's ncu profile showed
// (profiles/r1b_step1_*): half of the 940 warp instructions per trip were
// divergent bookkeeping, led by the per-phase propagation switch (stance / flight-after-stance / reverse flight /
// reverse stance executed one after the other, 177 instructions at 5-14 active lanes).
//
//   * A primitive segment is a cubic in time for x, y, z and pitch (planning_utils.cpp:237-306, :324-367).  The cursor
//     keeps the 16 coefficients of the CURRENT segment instead of (state, action): every lane evaluates its pose with
//     the same 20 FMAs whatever its phase (velocities are the derivative), and the coefficients are rebuilt only at a
//     segment change (take-off in FORWARD, end of the backward flight in REVERSE) — at most once per candidate.
//     The reverse-stance segment is expressed as the forward stance from its own start state, which is algebraically
//     the reference's formula (:324-367) and differs from it by rounding only (~1e-14 m, far inside the evaluator's
//     1e-5 m guard band; the guard sends anything closer to the fp64 pass).
//   * REVERSE candidates park the 8 stance accelerations they need at that change in shared memory (8 x 128 doubles).
//   * 32-bit candidate indices, no adaptive-step state unless ADAPTIVE, t_new derived from the last valid stance time.
// Verdicts, outputs and the k / L work counters are identical to the other variants (tests/test_gpu_parity.py).
#pragma once
#include "gbp_kernels.cuh"

namespace gbp {

struct WalkCursor {
	double c[4][4];        // x, y, z, pitch: c0 + c1 tau + c2 tau^2 + c3 tau^3 on the current segment
	double t, ts, tf, t_ls;  // t_ls: time of the last valid stance sample
	double step, t_ok;     // adaptive step only
	int phase, have_ls;
	Counters c_;
};

__device__ __forceinline__ double cubic(const double c[4], double x) { return __fma_rn(__fma_rn(__fma_rn(c[3], x, c[2]), x, c[1]), x, c[0]); }
__device__ __forceinline__ double cubic_d(const double c[4], double x) { return __fma_rn(__fma_rn(3.0 * c[3], x, c[2] + c[2]), x, c[1]); }

__device__ __forceinline__ Pose6 walk_pose(const WalkCursor &q) {
	const double tau = q.phase == PH_REV_FL ? -q.t : q.t;  // LAND holds t = t_f, START holds t = 0
	Pose6 o;
	o.x = cubic(q.c[0], tau);
	o.y = cubic(q.c[1], tau);
	o.z = cubic(q.c[2], tau);
	o.pitch = cubic(q.c[3], tau);
	o.dx = cubic_d(q.c[0], tau);
	o.dy = cubic_d(q.c[1], tau);
	return o;
}
// stance segment from state s (applyStance, :237-277)
__device__ __forceinline__ void walk_set_stance(WalkCursor &q, const double s[8], const double a[10]) {
	const double inv6ts = __drcp_rn(6.0 * q.ts);
#pragma unroll
	for (int d = 0; d < 4; ++d) {
		const int ip = d < 3 ? d : 6, iv = d < 3 ? 3 + d : 7, itd = d < 3 ? d : 8, ito = d < 3 ? 3 + d : 9;
		q.c[d][0] = s[ip]; q.c[d][1] = s[iv]; q.c[d][2] = 0.5 * a[itd]; q.c[d][3] = (a[ito] - a[itd]) * inv6ts;
	}
}
// ballistic segment from state s (applyFlight, :282-306)
__device__ __forceinline__ void walk_set_flight(WalkCursor &q, const double s[8]) {
#pragma unroll
	for (int d = 0; d < 4; ++d) {
		const int ip = d < 3 ? d : 6, iv = d < 3 ? 3 + d : 7;
		q.c[d][0] = s[ip]; q.c[d][1] = s[iv]; q.c[d][2] = d == 2 ? -0.5 * 9.81 : 0.0; q.c[d][3] = 0.0;
	}
}
// FORWARD take-off: the stance segment evaluated at t_s becomes the base of the flight segment (:731-737)
__device__ __forceinline__ void walk_take_off(WalkCursor &q) {
#pragma unroll
	for (int d = 0; d < 4; ++d) {
		const double p = cubic(q.c[d], q.ts), v = cubic_d(q.c[d], q.ts);
		q.c[d][0] = p; q.c[d][1] = v; q.c[d][2] = d == 2 ? -0.5 * 9.81 : 0.0; q.c[d][3] = 0.0;
	}
}
// REVERSE: the backward flight evaluated at -t_f is the take-off state; the stance before it, written as the forward
// stance from its start state (applyStanceReverse at t = 0, :324-367).  st = this thread's column of the parked
// accelerations {a_td x,y,z,pitch, a_to x,y,z,pitch}, stride RF_WARPS * 32 doubles.
__device__ __forceinline__ void walk_reverse_stance(WalkCursor &q, const double *st) {
	const double ts = q.ts, mtf = -q.tf, inv6ts = __drcp_rn(6.0 * ts);
#pragma unroll
	for (int d = 0; d < 4; ++d) {
		// the segment being left is a flight: c3 = 0 and c2 is the constant -g/2 (z) or 0
		const double tp = d == 2 ? __fma_rn(__fma_rn(q.c[d][2], mtf, q.c[d][1]), mtf, q.c[d][0]) : __fma_rn(q.c[d][1], mtf, q.c[d][0]);
		const double tv = d == 2 ? __fma_rn(q.c[d][2] + q.c[d][2], mtf, q.c[d][1]) : q.c[d][1];
		const double atd = st[d * (RF_WARPS * 32)], j = st[(4 + d) * (RF_WARPS * 32)] - atd;
		const double c2 = 0.5 * atd, c3 = j * inv6ts;
		const double cc = __fma_rn(-ts, __fma_rn(0.5, j, atd), tv);                      // tv - a_td ts - j ts / 2
		q.c[d][0] = __fma_rn(-ts, __fma_rn(ts, __fma_rn(c3, ts, c2), cc), tp);           // tp - cc ts - a_td ts^2 / 2 - j ts^3 / (6 ts)
		q.c[d][1] = cc; q.c[d][2] = c2; q.c[d][3] = c3;
	}
}

// cursor_start (gbp_kernels.cuh) for the polynomial cursor
__device__ __forceinline__ void walk_start(WalkCursor &q, const double s[8], const double a[10], int dir, double *st) {
	q.ts = a[6]; q.tf = a[7];
	q.step = KINEMATICS_RES; q.t_ok = 0; q.t_ls = 0; q.have_ls = 0;
	q.c_ = {0, 0, 0, 0};
	const double ts = q.ts, tf = q.tf;
	if (dir == GBP_FORWARD) {
		walk_set_stance(q, s, a);
		q.t = 0; q.phase = PH_FWD_ST;
		if (!(0 <= ts)) {
			walk_take_off(q);
			if (0 < tf) q.phase = PH_FWD_FL; else { q.phase = PH_FWD_LAND; q.t = tf; }
		}
	} else {
		walk_set_flight(q, s);
#pragma unroll
		for (int d = 0; d < 4; ++d) {
			st[d * (RF_WARPS * 32)] = a[d < 3 ? d : 8];
			st[(4 + d) * (RF_WARPS * 32)] = a[d < 3 ? 3 + d : 9];
		}
		q.t = 0; q.phase = PH_REV_FL;
		if (!(0 < tf)) {
			walk_reverse_stance(q, st);
			if (ts >= 0) { q.t = ts; q.phase = PH_REV_ST; } else { q.t = 0; q.phase = PH_REV_START; }
		}
	}
}

// cursor_advance (gbp_kernels.cuh) for the polynomial cursor: 0 = continue, 1 = finished invalid, 2 = finished valid
template <bool ADAPTIVE>
__device__ __forceinline__ int walk_advance(WalkCursor &q, bool valid, OutRecipe &out, const double *st) {
	const double ts = q.ts, tf = q.tf;
	const int ph = q.phase;
	const bool fwd = ph <= PH_FWD_LAND;
	const bool stance_seg = ph == PH_FWD_ST || ph == PH_REV_ST;
	const bool terminal = ph == PH_FWD_LAND || ph == PH_REV_START;
	if (!valid) {
		if (stance_seg) {
			if (ADAPTIVE && !(KINEMATICS_RES - 0.01 <= q.step && q.step <= KINEMATICS_RES + 0.01)) {
				q.step = KINEMATICS_RES;  // rewind to the last success (:672-675, :812-815)
				q.t = fwd ? q.t_ok + q.step : q.t_ok - q.step;
				if (fwd ? !(q.t <= ts) : !(q.t >= 0)) {
					if (fwd) {
						walk_take_off(q);
						if (0 < tf) { q.t = 0; q.phase = PH_FWD_FL; } else { q.t = tf; q.phase = PH_FWD_LAND; }
					} else { q.t = 0; q.phase = PH_REV_START; }
				}
				return 0;
			}
			out.kind = OUT_STANCE;
			out.tau = fwd ? (1.0 - BACKUP_RATIO) * q.t : q.t + BACKUP_RATIO * (ts - q.t);
		} else {
			const bool keep = q.have_ls != 0 && ph != PH_REV_FL;
			out.kind = keep ? (ph == PH_REV_START ? OUT_REV : OUT_STANCE) : OUT_SAME;
			out.tau = q.t_ls;
		}
		return 1;
	}
	if (terminal) {
		out.kind = fwd ? OUT_LAND : OUT_REV;
		out.tau = 0;
		return 2;
	}
	if (stance_seg) {
		q.t_ls = q.t; q.have_ls = 1;
		if (ADAPTIVE) q.t_ok = q.t;
	}
	if (ADAPTIVE) q.step += KINEMATICS_RES;
	const double step = ADAPTIVE ? q.step : KINEMATICS_RES;
	q.t = ph == PH_REV_ST ? q.t - step : q.t + step;
	const bool done = ph == PH_FWD_ST ? !(q.t <= ts) : (ph == PH_REV_ST ? !(q.t >= 0) : !(q.t < tf));
	if (done) {
		if (ADAPTIVE) q.step = KINEMATICS_RES;
		if (ph == PH_FWD_ST) {
			walk_take_off(q);
			if (0 < tf) { q.t = 0; q.phase = PH_FWD_FL; } else { q.t = tf; q.phase = PH_FWD_LAND; }
		} else if (ph == PH_FWD_FL) {
			q.t = tf; q.phase = PH_FWD_LAND;
		} else if (ph == PH_REV_FL) {
			walk_reverse_stance(q, st);
			if (ts >= 0) { q.t = ts; q.phase = PH_REV_ST; } else { q.t = 0; q.phase = PH_REV_START; }
		} else {
			q.t = 0; q.phase = PH_REV_START;
		}
	}
	return 0;
}

#ifndef GBP_WALK_CTAS
#define GBP_WALK_CTAS 4
#endif
// The terrain gathers live in the L1 / texture cache, which shares 256 KB per SM with shared memory: every KB of ring
// is a KB of cache lost (measured: 46 KB per CTA at 4 CTAs / SM ran 4.5 ms, at 3 CTAs / SM 3.3 ms).  16-candidate
// chunks, 2 deep, are still ~8 trips of work per warp ahead of the TMA's latency.
constexpr int WK_CHUNK = 16, WK_NBUF = 2;
constexpr int WK_SLOT_BYTES = (WK_CHUNK * (64 + 80) + WK_CHUNK + 127) / 128 * 128;  // 2432
static_assert(RF_CHUNK % WK_CHUNK == 0, "per_warp (a multiple of RF_CHUNK) must be a multiple of WK_CHUNK");

template <bool TEX, bool ADAPTIVE>
__global__ void __launch_bounds__(RF_WARPS * 32, GBP_WALK_CTAS) k_walk_mixed(TerrainView T, int n, int per_warp, const double *__restrict__ states,
																			   const double *__restrict__ actions, const uint8_t *__restrict__ dir,
																			   uint8_t *__restrict__ verdict, uint8_t *__restrict__ flags,
																			   double *__restrict__ s_new, double *__restrict__ t_new,
																			   unsigned long long *__restrict__ cnt, int *__restrict__ redo_idx,
																			   unsigned long long *__restrict__ redo_count, double2 *__restrict__ recipe) {
	__shared__ __align__(128) unsigned char ring[RF_WARPS][WK_NBUF][WK_SLOT_BYTES];
	__shared__ __align__(16) double stash[8][RF_WARPS * 32];
	__shared__ __align__(8) uint64_t bars[RF_WARPS][WK_NBUF];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int64_t warp = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
	const int64_t wbase64 = warp * (int64_t) per_warp;  // per_warp is a multiple of RF_CHUNK
	const int wbase = (int) min(wbase64, (int64_t) n);
	const int total = (int) min((int64_t) n - wbase, (int64_t) per_warp);  // candidates of this warp (0 past the end)
	const int nchunks = (total + WK_CHUNK - 1) / WK_CHUNK;
	double *const st = &stash[0][threadIdx.x];
	int next = 0, issued = 0;  // relative to wbase
	if (lane == 0) {
		for (int k = 0; k < WK_NBUF; ++k) mbar_init(&bars[wib][k], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
	const uint64_t pol = l2_evict_first_policy();
	auto issue = [&](int c) {  // lane 0 only: chunk c -> ring slot c % WK_NBUF
		const int slot = c % WK_NBUF;
		const int64_t c0 = (int64_t) wbase + (int64_t) c * WK_CHUNK;
		const int m = min(WK_CHUNK, total - c * WK_CHUNK);
		unsigned char *dst = ring[wib][slot];
		const unsigned dbytes = (m == WK_CHUNK) ? WK_CHUNK : 0;  // a partial (tail) chunk reads its directions from global memory
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy reads of this slot precede the async write
		mbar_expect_tx(&bars[wib][slot], (unsigned) m * (64 + 80) + dbytes);
		tma_load_1d_hint(dst, states + 8 * c0, (unsigned) m * 64, &bars[wib][slot], pol);
		tma_load_1d_hint(dst + WK_CHUNK * 64, actions + 10 * c0, (unsigned) m * 80, &bars[wib][slot], pol);
		if (dbytes) tma_load_1d_hint(dst + WK_CHUNK * 144, dir + c0, dbytes, &bars[wib][slot], pol);
	};
	if (lane == 0) {
		for (; issued < nchunks && issued < WK_NBUF; ++issued) issue(issued);
	}
	issued = __shfl_sync(FULL, issued, 0);
	WalkCursor q;
	q.phase = PH_IDLE;
	int mine = -1;
	unsigned k = 0, L = 0, np = 0, nvalid = 0;
	while (true) {
		// refill idle lanes from the warp's range, in lane order
		const unsigned need = __ballot_sync(FULL, q.phase == PH_IDLE);
		if (need && next < total) {
			// only chunks already issued can be handed out (a round of 32 refills may span 3 chunks, the ring holds 2):
			// lanes beyond them stay idle for one trip, lane 0 issues the next chunks below
			const int limit = min(total, issued * WK_CHUNK);
			if (q.phase == PH_IDLE) {
				const int rel = next + __popc(need & ((1u << lane) - 1));
				if (rel < limit) {
					const int c = rel / WK_CHUNK, slot = c % WK_NBUF, j = rel % WK_CHUNK;
					mbar_wait(&bars[wib][slot], (unsigned) ((c / WK_NBUF) & 1));
					const unsigned char *src = ring[wib][slot];
					const double2 *ps = reinterpret_cast<const double2 *>(src + j * 64);
					const double2 *pa = reinterpret_cast<const double2 *>(src + WK_CHUNK * 64 + j * 80);
					double s[8], a[10];
#pragma unroll
					for (int d = 0; d < 4; ++d) { const double2 v = ps[d]; s[2 * d] = v.x; s[2 * d + 1] = v.y; }
#pragma unroll
					for (int d = 0; d < 5; ++d) { const double2 v = pa[d]; a[2 * d] = v.x; a[2 * d + 1] = v.y; }
					const bool full_chunk = (c + 1) * WK_CHUNK <= total;
					mine = wbase + rel;
					const int dv = full_chunk ? (int) src[WK_CHUNK * 144 + j] : (int) __ldg(dir + mine);
					walk_start(q, s, a, dv, st);
				}
			}
			next = min(limit, next + __popc(need));
			__syncwarp();
			// ring slots whose chunk is fully consumed are refilled two chunks ahead
			const int consumed = next >= total ? nchunks : next / WK_CHUNK;
			if (lane == 0) {
				for (; issued < nchunks && issued < consumed + WK_NBUF; ++issued) issue(issued);
			}
			issued = __shfl_sync(FULL, issued, 0);
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
			const bool fwd = q.phase <= PH_FWD_LAND;
			const int r = walk_advance<ADAPTIVE>(q, valid, out, st);
			if (r) {
				const bool ok = r == 2;
				__stcs(verdict + mine, (uint8_t) (ok ? 1 : 0));
				if (flags) __stcs(flags + mine, (uint8_t) (ok ? GBP_FLAG_VALID : 0));  // no OOG / NEAR flag can arise on this path
				// s_new is finished by k_pair_outputs (convergent, exact) from the recipe: kept in a compact side array when the
				// caller runs both kernels (a 16-byte write into a 64-byte s_new row costs k_pair_outputs a read of the whole row),
				// else in the row itself (variant 5 / gbp_pair_outputs_dev contract)
				if (s_new) __stcs(recipe ? recipe + mine : reinterpret_cast<double2 *>(s_new + 8 * (int64_t) mine), make_double2(out.tau, (double) out.kind));
				if (t_new) {
					const double tn = ok ? (fwd ? q.ts + q.tf : q.ts) : (q.have_ls ? (fwd ? q.t_ls : q.ts - q.t_ls) : 0.0);
					__stcs(t_new + mine, tn);
				}
				k += q.c_.substates; L += q.c_.lookups; np += q.c_.nanprobes; nvalid += ok ? 1 : 0;
				q.phase = PH_IDLE;
			}
		}
	}
	flush_counters(cnt, k, L, np, 0, 0, nvalid);
}

}  // namespace gbp
