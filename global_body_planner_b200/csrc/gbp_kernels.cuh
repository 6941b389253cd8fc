// Kernels of the extend path (sm_100a).  See DESIGN.md for the data layout and the roofline of each.
#pragma once
#include "gbp_device.cuh"

namespace gbp {

constexpr unsigned FULL = 0xffffffffu;

// device-resident GraphClass/PlannerClass store (graph_class.h:155-170) as SoA
struct TreeView {
	int cap;
	int *n;          // vertex count (device)
	double *v;       // [8][cap]
	double *act;     // [10][cap]
	int *parent;     // [cap]
	double *g, *y;   // [cap]
};

__device__ __forceinline__ void tree_get(const TreeView &T, int i, double s[8]) {
#pragma unroll
	for (int d = 0; d < 8; ++d) s[d] = T.v[(size_t) d * T.cap + i];
}
// addVertex + addEdge + addAction + updateGYValue (rrt.cpp:87-92, graph_class.cpp:36-42)
__device__ __forceinline__ int tree_push(const TreeView &T, int parent, const double s[8], const double a[10]) {
	int i = *T.n;
	*T.n = i + 1;
	double p[8];
	tree_get(T, parent, p);
#pragma unroll
	for (int d = 0; d < 8; ++d) T.v[(size_t) d * T.cap + i] = s[d];
#pragma unroll
	for (int d = 0; d < 10; ++d) T.act[(size_t) d * T.cap + i] = a[d];
	T.parent[i] = parent;
	T.g[i] = T.g[parent] + pose_distance(p, s);
	T.y[i] = T.y[parent] + yaw_distance(p, s);
	return i;
}

// ------------------------------------------------------------------ small batched queries
template <typename M>
__global__ void k_terrain_query(TerrainView T, int64_t n, const double *__restrict__ x, const double *__restrict__ y, int what,
								double *__restrict__ out, uint8_t *__restrict__ out8) {
	int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i >= n) return;
	unsigned fl = 0;
	if (what == 0) {
		out[i] = ground_height<M>(T, x[i], y[i], fl);
		if (out8) out8[i] = (uint8_t) fl;
	} else if (what == 1) {
		out8[i] = height_is_nan<M>(T, x[i], y[i], fl) ? 1 : 0;
	} else {
		double nn[3];
		surface_normal(T, x[i], y[i], nn, fl);
		out[3 * i] = nn[0]; out[3 * i + 1] = nn[1]; out[3 * i + 2] = nn[2];
	}
}

static __global__ void k_propagate(int kind, int64_t n, const double *__restrict__ s, const double *__restrict__ a,
							const double *__restrict__ t, double *__restrict__ out) {
	int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i >= n) return;
	double ss[8], aa[10], o[8];
#pragma unroll
	for (int d = 0; d < 8; ++d) ss[d] = s[8 * i + d];
	if (kind != 1) {
#pragma unroll
		for (int d = 0; d < 10; ++d) aa[d] = a[10 * i + d];
	}
	if (kind == 0) apply_stance(ss, aa, t[i], o);
	else if (kind == 1) apply_flight(ss, t[i], o);
	else apply_stance_reverse(ss, aa, t[i], o);
#pragma unroll
	for (int d = 0; d < 8; ++d) out[8 * i + d] = o[d];
}

// rotate_grf (planning_utils.cpp:198-231) and calculateCurvature (:884-899), one thread per element
static __global__ void k_rotate_grf(int64_t n, const double *__restrict__ normal, const double *__restrict__ force, double *__restrict__ out) {
	const int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double nn[3] = {normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]}, f[3] = {force[3 * i], force[3 * i + 1], force[3 * i + 2]};
	double R[9], o[3];
	grf_rotation(nn, R);
	mat3_apply(R, f, o);
	out[3 * i] = o[0]; out[3 * i + 1] = o[1]; out[3 * i + 2] = o[2];
}
__device__ __forceinline__ double curvature3(double x1, double y1, double x2, double y2, double x3, double y3) {
	if ((x1 == x2 && x2 == x3) || (y1 == y2 && y2 == y3)) return 0.0;
	const double dis12 = sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
	const double dis13 = sqrt((x1 - x3) * (x1 - x3) + (y1 - y3) * (y1 - y3));
	const double dis23 = sqrt((x2 - x3) * (x2 - x3) + (y2 - y3) * (y2 - y3));
	const double dis = dis12 * dis12 + dis23 * dis23 - dis13 * dis13;
	const double cosA = dis / (2 * dis12 * dis23);
	const double sinA = sqrt(1 - cosA * cosA);
	double c = 0.5 * dis13 / sinA;
	c = 1 / c;
	return c;
}
static __global__ void k_curvature(int64_t n, const double *__restrict__ p, double *__restrict__ out) {
	const int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i >= n) return;
	out[i] = curvature3(p[6 * i], p[6 * i + 1], p[6 * i + 2], p[6 * i + 3], p[6 * i + 4], p[6 * i + 5]);
}
static __global__ void k_valid_actions(int64_t n, const double *__restrict__ a, uint8_t *__restrict__ out) {
	int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i >= n) return;
	double aa[10];
#pragma unroll
	for (int d = 0; d < 10; ++d) aa[d] = a[10 * i + d];
	out[i] = is_valid_action(aa) ? 1 : 0;
}

template <typename M>
__global__ void k_valid_states(TerrainView T, int64_t n, const double *__restrict__ s, const uint8_t *__restrict__ phase,
							   uint8_t *__restrict__ out, uint8_t *__restrict__ flags) {
	int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i >= n) return;
	double ss[8];
#pragma unroll
	for (int d = 0; d < 8; ++d) ss[d] = s[8 * i + d];
	Counters c = {0, 0, 0, 0};
	bool ok = is_valid_state_auto<M>(T, pose6(ss), phase[i], c);
	out[i] = ok ? 1 : 0;
	if (flags) flags[i] = (uint8_t) (c.flags | (ok ? GBP_FLAG_VALID : 0));
}

static __global__ void k_distance(int kind, int64_t n, const double *__restrict__ q1, const double *__restrict__ q2,
						   double *__restrict__ out) {
	int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i >= n) return;
	double a[8], b[8];
#pragma unroll
	for (int d = 0; d < 8; ++d) { a[d] = q1[8 * i + d]; b[d] = q2[8 * i + d]; }
	out[i] = kind == 0 ? pose_distance(a, b) : kind == 1 ? state_distance(a, b) : yaw_distance(a, b);
}

// ------------------------------------------------------------------ counters
__device__ __forceinline__ void flush_counters(unsigned long long *cnt, unsigned long long k, unsigned long long L,
											   unsigned long long np, unsigned long long oog, unsigned long long near,
											   unsigned long long valid) {
	// warp-reduce then one atomic per warp per counter
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		k += __shfl_down_sync(FULL, k, o);
		L += __shfl_down_sync(FULL, L, o);
		np += __shfl_down_sync(FULL, np, o);
		oog += __shfl_down_sync(FULL, oog, o);
		near += __shfl_down_sync(FULL, near, o);
		valid += __shfl_down_sync(FULL, valid, o);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(cnt + 0, k); atomicAdd(cnt + 1, L); atomicAdd(cnt + 2, np);
		atomicAdd(cnt + 3, oog); atomicAdd(cnt + 4, near); atomicAdd(cnt + 5, valid);
	}
}

__device__ __forceinline__ void load_state(const double *__restrict__ p, double s[8]) {
	const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
	for (int d = 0; d < 4; ++d) { double2 v = __ldg(q + d); s[2 * d] = v.x; s[2 * d + 1] = v.y; }
}
__device__ __forceinline__ void load_action(const double *__restrict__ p, double a[10]) {
	const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
	for (int d = 0; d < 5; ++d) { double2 v = __ldg(q + d); a[2 * d] = v.x; a[2 * d + 1] = v.y; }
}
__device__ __forceinline__ void store_state(double *__restrict__ p, const double s[8]) {
	double2 *q = reinterpret_cast<double2 *>(p);
#pragma unroll
	for (int d = 0; d < 4; ++d) q[d] = make_double2(s[2 * d], s[2 * d + 1]);
}

// ------------------------------------------------------------------ the pair-check cursor machine
// Lane-per-action with warp-level refill: every lane walks the sub-states of its own candidate in
// the reference's order, ONE sub-state per loop trip, so the expensive isValidState is executed
// convergently by all 32 lanes on 32 different candidates; a lane whose candidate finished (early
// exit or fully valid) pulls the next candidate of the warp's contiguous range.  Work done equals
// the reference's early-exit work (k sub-states per candidate), unlike the warp-per-action form.
enum : int { PH_FWD_ST = 0, PH_FWD_FL = 1, PH_FWD_LAND = 2, PH_REV_FL = 3, PH_REV_ST = 4, PH_REV_START = 5, PH_IDLE = 6 };

struct Cursor {
	double s[8], a[10];
	double t, step, t_ok, t_new, t_ls;  // t_ls: time of the last valid stance sample (s_new on flight failure)
	int phase, have_ls;
	FastPrim f;
	Counters c;
};

__device__ __forceinline__ void cursor_start(Cursor &q, int dir) {
	q.step = KINEMATICS_RES;
	q.t_ok = 0;
	q.t_new = 0;
	q.t_ls = 0;
	q.have_ls = 0;
	q.c = {0, 0, 0, 0};
	const double ts = q.a[6], tf = q.a[7];
	q.f.inv6ts = 1.0 / (6.0 * ts);
	q.f.inv2ts = 1.0 / (2.0 * ts);
	if (dir == GBP_FORWARD) {
		q.t = 0;
		q.phase = (0 <= ts) ? PH_FWD_ST : ((0 < tf) ? PH_FWD_FL : PH_FWD_LAND);
	} else {
		q.t = 0;
		if (0 < tf) q.phase = PH_REV_FL;
		else { q.t = ts; q.phase = (ts >= 0) ? PH_REV_ST : PH_REV_START; }
	}
}
// isValidState of the sub-state at the cursor, through the fast validity path (guard-band accuracy;
// the exact propagation is only used for outputs, see cursor_advance)
template <typename M>
__device__ __forceinline__ bool cursor_check(const TerrainView &T, Cursor &q) {
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
	return is_valid_state_auto<M>(T, p, ph, q.c);
}
// How the reference's s_new output is obtained once a pair check has finished.  The (cheap) decision is
// taken where the walk ends; the (expensive, exact, division-heavy) evaluation is done once per
// candidate by finish_output(), convergently, instead of inside the divergent walk.
enum : int { OUT_SAME = 0, OUT_STANCE = 1, OUT_LAND = 2, OUT_REV = 3 };
struct OutRecipe {
	int kind;    // OUT_SAME: s_new = s (defined value where the reference leaves it unwritten)
	double tau;  // OUT_STANCE: applyStance(s, a, tau); OUT_LAND: applyFlight(applyStance(s, a), t_f);
};               // OUT_REV: applyStanceReverse(applyFlight(s, -t_f), a, tau)
__device__ __forceinline__ void finish_output(const double s[8], const double a[10], int kind, double tau, double s_new[8]) {
	double tmp[8];
	switch (kind) {
	case OUT_STANCE: apply_stance(s, a, tau, s_new); break;
	case OUT_LAND: apply_stance(s, a, a[6], tmp); apply_flight(tmp, a[7], s_new); break;  // exact s_land (:743-749)
	case OUT_REV: apply_flight(s, -a[7], tmp); apply_stance_reverse(tmp, a, tau, s_new); break;  // (:866-872)
	default:
#pragma unroll
		for (int i = 0; i < 8; ++i) s_new[i] = s[i];
		break;
	}
}

// Advance after the verdict of the current sub-state.  Returns 0 = continue, 1 = finished invalid,
// 2 = finished valid.  On finish, `out` says how to compute s_new and q.t_new holds t_new.
// Written as one predicated flow (not a switch over the six phases): the lanes of a warp are in different
// phases almost every trip, and a switch would execute all its arms one after the other.
__device__ __forceinline__ int cursor_advance(Cursor &q, bool valid, bool adaptive, OutRecipe &out) {
	const double ts = q.a[6], tf = q.a[7];
	const int ph = q.phase;
	const bool fwd = ph <= PH_FWD_LAND;
	const bool stance_seg = ph == PH_FWD_ST || ph == PH_REV_ST;   // the sampled stance loop (:718-730, :850-868)
	const bool terminal = ph == PH_FWD_LAND || ph == PH_REV_START;  // landing / exact start state
	if (!valid) {
		if (stance_seg) {
			if (adaptive && !(KINEMATICS_RES - 0.01 <= q.step && q.step <= KINEMATICS_RES + 0.01)) {
				q.step = KINEMATICS_RES;  // adaptive step: rewind to the last success (:672-675, :812-815)
				q.t = fwd ? q.t_ok + q.step : q.t_ok - q.step;
				if (fwd ? !(q.t <= ts) : !(q.t >= 0)) {
					q.t = 0; q.phase = fwd ? ((0 < tf) ? PH_FWD_FL : PH_FWD_LAND) : PH_REV_START;
				}
				return 0;
			}
			out.kind = OUT_STANCE;  // back up half of the failed step; REVERSE applies the FORWARD stance to the END state (sic, :862)
			out.tau = fwd ? (1.0 - BACKUP_RATIO) * q.t : q.t + BACKUP_RATIO * (ts - q.t);
		} else {
			// flight / landing failure keeps the last valid stance sample; REVERSE flight failure leaves s_new untouched
			const bool keep = q.have_ls != 0 && ph != PH_REV_FL;
			out.kind = keep ? (ph == PH_REV_START ? OUT_REV : OUT_STANCE) : OUT_SAME;
			out.tau = q.t_ls;
		}
		return 1;
	}
	if (terminal) {
		out.kind = fwd ? OUT_LAND : OUT_REV;
		out.tau = 0;
		q.t_new = fwd ? ts + tf : ts;
		return 2;
	}
	if (stance_seg) {
		q.t_ls = q.t; q.have_ls = 1;
		q.t_new = fwd ? q.t : ts - q.t;
		if (adaptive) q.t_ok = q.t;
	}
	if (adaptive) q.step += KINEMATICS_RES;
	q.t = ph == PH_REV_ST ? q.t - q.step : q.t + q.step;
	const bool done = ph == PH_FWD_ST ? !(q.t <= ts) : (ph == PH_REV_ST ? !(q.t >= 0) : !(q.t < tf));
	if (done) {
		q.step = KINEMATICS_RES;
		if (ph == PH_FWD_ST) { q.t = 0; q.phase = (0 < tf) ? PH_FWD_FL : PH_FWD_LAND; }
		else if (ph == PH_FWD_FL) q.phase = PH_FWD_LAND;
		else if (ph == PH_REV_FL) { q.t = ts; q.phase = (ts >= 0) ? PH_REV_ST : PH_REV_START; }
		else q.phase = PH_REV_START;
	}
	return 0;
}

// Sequential walk of one pair by one thread: the cursor machine below run to completion.  Same
// semantics as planning_utils.cpp:651-753 (forward) and :774-876 (reverse).
template <typename M>
__device__ __forceinline__ bool validate_pair_seq(const TerrainView &T, const double s[8], const double a[10], int direction,
												  bool adaptive, double s_new[8], double &t_new, Counters &c) {
	Cursor q;
#pragma unroll
	for (int i = 0; i < 8; ++i) q.s[i] = s[i];
#pragma unroll
	for (int i = 0; i < 10; ++i) q.a[i] = a[i];
	cursor_start(q, direction);
	int r = 0;
	OutRecipe out;
	while (true) {
		const bool valid = cursor_check<M>(T, q);
		r = cursor_advance(q, valid, adaptive, out);
		if (r) break;
	}
	finish_output(s, a, out.kind, out.tau, s_new);
	t_new = q.t_new;
	c.substates += q.c.substates; c.lookups += q.c.lookups; c.nanprobes += q.c.nanprobes; c.flags |= q.c.flags;
	return r == 2;
}

// attemptConnect (src/rrt_connect.cpp:20-91), recursion unrolled into a loop (see oracle/gbp_oracle.c)
template <typename M>
__device__ int attempt_connect(const TerrainView &T, const double s_existing[8], const double s_in[8], int direction,
							   bool adaptive, double s_new[8], double a_new[10], Counters &c, unsigned &pair_checks,
							   double ts0 = -1.0) {
	double target[8], ts = ts0 >= 0.0 ? ts0 : pose_distance(s_in, s_existing) / V_NOM;  // :89
#pragma unroll
	for (int i = 0; i < 8; ++i) target[i] = s_in[i];
	for (int depth = 0;; ++depth) {
		if (ts <= KINEMATICS_RES) return GBP_TRAPPED;
		if (direction == GBP_FORWARD) connect_action(s_existing, target, ts, a_new);
		else connect_action(target, s_existing, ts, a_new);
		if (!is_valid_action(a_new)) return GBP_TRAPPED;
		double out[8], tn;
		++pair_checks;
		bool ok = validate_pair_seq<M>(T, s_existing, a_new, direction, adaptive, out, tn, c);
#pragma unroll
		for (int i = 0; i < 8; ++i) { s_new[i] = out[i]; target[i] = out[i]; }
		if (ok) return depth == 0 ? GBP_REACHED : GBP_ADVANCED;
		ts = tn;
	}
}

// ------------------------------------------------------------------ validate_pairs, variant 1
// One thread per action, sub-states walked sequentially in the reference's order.
template <typename M>
__global__ void __launch_bounds__(128) k_validate_thread(TerrainView T, int64_t n, const double *__restrict__ states,
														  const double *__restrict__ actions, const uint8_t *__restrict__ dir,
														  int adaptive, uint8_t *__restrict__ verdict, uint8_t *__restrict__ flags,
														  double *__restrict__ s_new, double *__restrict__ t_new,
														  unsigned long long *__restrict__ cnt) {
	int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	Counters c = {0, 0, 0, 0};
	bool ok = false;
	if (i < n) {
		double s[8], a[10], sn[8], tn;
		load_state(states + 8 * i, s);
		load_action(actions + 10 * i, a);
		ok = validate_pair_seq<M>(T, s, a, dir[i], adaptive != 0, sn, tn, c);
		verdict[i] = ok ? 1 : 0;
		if (flags) flags[i] = (uint8_t) (c.flags | (ok ? GBP_FLAG_VALID : 0));
		if (s_new) store_state(s_new + 8 * i, sn);
		if (t_new) t_new[i] = tn;
	}
	flush_counters(cnt, c.substates, c.lookups, c.nanprobes, (c.flags & GBP_FLAG_OOG) ? 1 : 0, (c.flags & GBP_FLAG_NEAR) ? 1 : 0, ok ? 1 : 0);
}

// ------------------------------------------------------------------ validate_pairs, variant 2
// One warp per action, one lane per interpolated sub-state (fixed step only).  Lane l of a round
// evaluates sub-state 32*round + l of the reference's sequence (stance samples, flight samples,
// landing / exact start); the sample times are the reference's fp64-ACCUMULATED values, obtained by
// walking l steps from the round's base cursor.  __ballot_sync finds the first failing lane, which
// reproduces the sequential early-exit outputs; lanes past it are speculative work that the
// counters do not include.  Lowest latency per action: used for connect primitives (t_s = d/0.75
// can span hundreds of sub-states) and inside the per-query planner warp.
enum : int { PH_DONE = 7 };
__device__ __forceinline__ void walk_step(int &ph, double &t, double ts, double tf) {
	switch (ph) {
	case PH_FWD_ST: t += KINEMATICS_RES; if (!(t <= ts)) { t = 0; ph = (0 < tf) ? PH_FWD_FL : PH_FWD_LAND; } break;
	case PH_FWD_FL: t += KINEMATICS_RES; if (!(t < tf)) ph = PH_FWD_LAND; break;
	case PH_REV_FL: t += KINEMATICS_RES; if (!(t < tf)) { t = ts; ph = (ts >= 0) ? PH_REV_ST : PH_REV_START; } break;
	case PH_REV_ST: t -= KINEMATICS_RES; if (!(t >= 0)) ph = PH_REV_START; break;
	default: ph = PH_DONE; break;  // LAND / START are terminal
	}
}
// All 32 lanes call this with identical (warp-uniform) s, a, direction; outputs are uniform too.
template <typename M>
__device__ bool validate_pair_warp(const TerrainView &T, const double s[8], const double a[10], int direction, double s_new[8],
								   double &t_new, Counters &c) {
	const int lane = threadIdx.x & 31;
	const double ts = a[6], tf = a[7];
	Cursor q;
#pragma unroll
	for (int i = 0; i < 8; ++i) { q.s[i] = s[i]; s_new[i] = s[i]; }
#pragma unroll
	for (int i = 0; i < 10; ++i) q.a[i] = a[i];
	cursor_start(q, direction);
	int bph = q.phase;
	double bt = q.t, t_ls = 0;
	bool have_ls = false;
	t_new = 0;
	while (true) {
		int ph = bph;
		double t = bt;
		for (int i = 0; i < lane && ph != PH_DONE; ++i) walk_step(ph, t, ts, tf);
		const bool active = ph != PH_DONE;
		Counters lc = {0, 0, 0, 0};
		bool valid = true;
		if (active) {
			q.phase = ph;
			q.t = t;
			q.c = {0, 0, 0, 0};
			valid = cursor_check<M>(T, q);
			lc = q.c;
		}
		const unsigned bad = __ballot_sync(FULL, active && !valid);
		const int f = bad ? __ffs(bad) - 1 : 32;
		const bool counted = active && lane <= f;  // the sub-states the reference would have started
		c.substates += __reduce_add_sync(FULL, counted ? lc.substates : 0u);
		c.lookups += __reduce_add_sync(FULL, counted ? lc.lookups : 0u);
		c.nanprobes += __reduce_add_sync(FULL, counted ? lc.nanprobes : 0u);
		c.flags |= __reduce_or_sync(FULL, counted ? lc.flags : 0u);
		const unsigned st_ok = __ballot_sync(FULL, active && lane < f && (ph == PH_FWD_ST || ph == PH_REV_ST));
		if (st_ok) {
			t_ls = __shfl_sync(FULL, t, 31 - __clz(st_ok));
			have_ls = true;
			t_new = direction == GBP_FORWARD ? t_ls : ts - t_ls;
		}
		if (f < 32) {  // reference outputs at the first failing sub-state
			const int pf = __shfl_sync(FULL, ph, f);
			const double tfail = __shfl_sync(FULL, t, f);
			double tmp[8];
			switch (pf) {
			case PH_FWD_ST: apply_stance(s, a, (1.0 - BACKUP_RATIO) * tfail, s_new); break;
			case PH_FWD_FL:
			case PH_FWD_LAND: if (have_ls) apply_stance(s, a, t_ls, s_new); break;
			case PH_REV_FL: break;
			case PH_REV_ST: apply_stance(s, a, tfail + BACKUP_RATIO * (ts - tfail), s_new); break;
			default: if (have_ls) { apply_flight(s, -tf, tmp); apply_stance_reverse(tmp, a, t_ls, s_new); } break;
			}
			return false;
		}
		const unsigned term = __ballot_sync(FULL, active && (ph == PH_FWD_LAND || ph == PH_REV_START));
		if (term) {  // landing / exact start state reached and valid: exact outputs
			double tmp[8];
			if (direction == GBP_FORWARD) { apply_stance(s, a, ts, tmp); apply_flight(tmp, tf, s_new); t_new = ts + tf; }
			else { apply_flight(s, -tf, tmp); apply_stance_reverse(tmp, a, 0, s_new); t_new = ts; }
			return true;
		}
		bph = __shfl_sync(FULL, ph, 31);
		bt = __shfl_sync(FULL, t, 31);
		walk_step(bph, bt, ts, tf);
	}
}
template <typename M>
__global__ void __launch_bounds__(128) k_validate_warp(TerrainView T, int64_t n, const double *__restrict__ states,
														const double *__restrict__ actions, const uint8_t *__restrict__ dir,
														uint8_t *__restrict__ verdict, uint8_t *__restrict__ flags,
														double *__restrict__ s_new, double *__restrict__ t_new,
														unsigned long long *__restrict__ cnt) {
	const int64_t i = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
	if (i >= n) return;  // warp-uniform
	double s[8], a[10], sn[8], tn;
	load_state(states + 8 * i, s);
	load_action(actions + 10 * i, a);
	Counters c = {0, 0, 0, 0};
	const bool ok = validate_pair_warp<M>(T, s, a, dir[i], sn, tn, c);
	if ((threadIdx.x & 31) == 0) {
		verdict[i] = ok ? 1 : 0;
		if (flags) flags[i] = (uint8_t) (c.flags | (ok ? GBP_FLAG_VALID : 0));
		if (s_new) store_state(s_new + 8 * i, sn);
		if (t_new) t_new[i] = tn;
		atomicAdd(cnt + 0, (unsigned long long) c.substates); atomicAdd(cnt + 1, (unsigned long long) c.lookups);
		atomicAdd(cnt + 2, (unsigned long long) c.nanprobes); atomicAdd(cnt + 3, (c.flags & GBP_FLAG_OOG) ? 1ull : 0ull);
		atomicAdd(cnt + 4, (c.flags & GBP_FLAG_NEAR) ? 1ull : 0ull); atomicAdd(cnt + 5, ok ? 1ull : 0ull);
	}
}

// ------------------------------------------------------------------ validate_pairs, variant 3 (refill)
// Candidate inputs are streamed through shared memory by the TMA (1-D bulk copies, cp.async.bulk ->
// SASS UBLKCP): each warp owns a contiguous, 32-aligned range of candidates and keeps a 2-deep ring
// of 32-candidate chunks {states 2 KiB, actions 2.5 KiB, directions 32 B}; one elected lane issues
// the copies two chunks ahead and an mbarrier per ring slot signals arrival, so a lane that refills
// reads its next candidate from shared memory instead of stalling on an HBM round trip.
constexpr int RF_CHUNK = 32, RF_NBUF = 2, RF_WARPS = 4;
constexpr int RF_SLOT_BYTES = RF_CHUNK * (64 + 80) + RF_CHUNK;  // 4640

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_LOOP:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra WAIT_DONE;\n"
		"bra WAIT_LOOP;\n"
		"WAIT_DONE:\n"
		"}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
				 "l"(src), "r"(bytes), "r"(smem_u32(bar))
				 : "memory");
}
// candidate streams are read exactly once: evict-first in L2, so that they do not displace the height grid
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void *dst, const void *src, unsigned bytes, uint64_t *bar, uint64_t policy) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
					 smem_u32(dst)),
				 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
				 : "memory");
}


// General form (any map kind: fp64 cells, non-uniform axes, NaN cells, small maps), evaluator is_valid_state_auto.
// fp32 maps on uniform axes without NaN take k_walk_mixed (gbp_walk.cuh) instead.
template <typename M>
__global__ void __launch_bounds__(RF_WARPS * 32, 2) k_validate_refill(TerrainView T, int64_t n, int64_t per_warp,
																   const double *__restrict__ states, const double *__restrict__ actions,
																   const uint8_t *__restrict__ dir, int adaptive,
																   uint8_t *__restrict__ verdict, uint8_t *__restrict__ flags,
																   double *__restrict__ s_new, double *__restrict__ t_new,
																   unsigned long long *__restrict__ cnt) {
	__shared__ __align__(128) unsigned char ring[RF_WARPS][RF_NBUF][RF_SLOT_BYTES + 112];  // slots padded to 128 B multiples
	__shared__ __align__(8) uint64_t bars[RF_WARPS][RF_NBUF];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int64_t warp = (blockIdx.x * (int64_t) blockDim.x + threadIdx.x) >> 5;
	const int64_t wbase = warp * per_warp;  // per_warp is a multiple of RF_CHUNK
	const int64_t end = min(n, wbase + per_warp);
	const int nchunks = wbase < end ? (int) ((end - wbase + RF_CHUNK - 1) / RF_CHUNK) : 0;
	int64_t next = wbase;
	int issued = 0;
	if (lane == 0) {
		for (int k = 0; k < RF_NBUF; ++k) mbar_init(&bars[wib][k], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
#ifndef GBP_NO_L2_HINTS
	const uint64_t pol = l2_evict_first_policy();
#define GBP_TMA_LOAD(d, s, b, m) tma_load_1d_hint(d, s, b, m, pol)
#else
#define GBP_TMA_LOAD(d, s, b, m) tma_load_1d(d, s, b, m)
#endif
	auto issue = [&](int c) {  // lane 0 only: chunk c -> ring slot c % RF_NBUF
		const int slot = c % RF_NBUF;
		const int64_t c0 = wbase + (int64_t) c * RF_CHUNK;
		const int m = (int) min((int64_t) RF_CHUNK, end - c0);
		unsigned char *dst = ring[wib][slot];
		const unsigned dbytes = (m == RF_CHUNK) ? RF_CHUNK : 0;  // a partial (tail) chunk reads its directions from global memory
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy reads of this slot precede the async write
		mbar_expect_tx(&bars[wib][slot], (unsigned) m * (64 + 80) + dbytes);
		GBP_TMA_LOAD(dst, states + 8 * c0, (unsigned) m * 64, &bars[wib][slot]);
		GBP_TMA_LOAD(dst + RF_CHUNK * 64, actions + 10 * c0, (unsigned) m * 80, &bars[wib][slot]);
		if (dbytes) GBP_TMA_LOAD(dst + RF_CHUNK * 144, dir + c0, dbytes, &bars[wib][slot]);
	};
#undef GBP_TMA_LOAD
	if (lane == 0) {
		for (; issued < nchunks && issued < RF_NBUF; ++issued) issue(issued);
	}
	issued = __shfl_sync(FULL, issued, 0);
	Cursor q;
	q.phase = PH_IDLE;
	int64_t mine = -1;
	unsigned long long k = 0, L = 0, np = 0, oog = 0, near = 0, nvalid = 0;
	while (true) {
		// refill idle lanes from the warp's range, in lane order
		const unsigned need = __ballot_sync(FULL, q.phase == PH_IDLE);
		if (need && next < end) {
			if (q.phase == PH_IDLE) {
				const int64_t idx = next + __popc(need & ((1u << lane) - 1));
				if (idx < end) {
					const int c = (int) ((idx - wbase) / RF_CHUNK), slot = c % RF_NBUF, j = (int) ((idx - wbase) % RF_CHUNK);
					mbar_wait(&bars[wib][slot], (unsigned) ((c / RF_NBUF) & 1));
					const unsigned char *src = ring[wib][slot];
					const double2 *ps = reinterpret_cast<const double2 *>(src + j * 64);
					const double2 *pa = reinterpret_cast<const double2 *>(src + RF_CHUNK * 64 + j * 80);
#pragma unroll
					for (int d = 0; d < 4; ++d) { double2 v = ps[d]; q.s[2 * d] = v.x; q.s[2 * d + 1] = v.y; }
#pragma unroll
					for (int d = 0; d < 5; ++d) { double2 v = pa[d]; q.a[2 * d] = v.x; q.a[2 * d + 1] = v.y; }
					const bool full_chunk = (wbase + (int64_t) (c + 1) * RF_CHUNK) <= end;
					const int dv = full_chunk ? (int) src[RF_CHUNK * 144 + j] : (int) __ldg(dir + idx);
					mine = idx;
					cursor_start(q, dv);
				}
			}
			next = min(end, next + (int64_t) __popc(need));
			__syncwarp();
			// ring slots whose chunk is fully consumed are refilled two chunks ahead
			const int consumed = next >= end ? nchunks : (int) ((next - wbase) / RF_CHUNK);
			if (lane == 0) {
				for (; issued < nchunks && issued < consumed + RF_NBUF; ++issued) issue(issued);
			}
			issued = __shfl_sync(FULL, issued, 0);
		}
		if (__ballot_sync(FULL, q.phase != PH_IDLE) == 0) break;
		bool valid = true;
		if (q.phase != PH_IDLE) valid = cursor_check<M>(T, q);
		if (q.phase != PH_IDLE) {
			OutRecipe out;
			int r = cursor_advance(q, valid, adaptive != 0, out);
			if (r) {
				const bool ok = r == 2;
#ifndef GBP_NO_L2_HINTS
				__stcs(verdict + mine, (uint8_t) (ok ? 1 : 0));
				if (flags) __stcs(flags + mine, (uint8_t) (q.c.flags | (ok ? GBP_FLAG_VALID : 0)));
				// s_new is finished by k_pair_outputs (convergent, exact); its slot carries the recipe meanwhile
				if (s_new) __stcs(reinterpret_cast<double2 *>(s_new + 8 * mine), make_double2(out.tau, (double) out.kind));
#else
				verdict[mine] = ok ? 1 : 0;
				if (flags) flags[mine] = (uint8_t) (q.c.flags | (ok ? GBP_FLAG_VALID : 0));
				if (s_new) *reinterpret_cast<double2 *>(s_new + 8 * mine) = make_double2(out.tau, (double) out.kind);
#endif
				if (t_new) __stcs(t_new + mine, q.t_new);
				k += q.c.substates; L += q.c.lookups; np += q.c.nanprobes;
				oog += (q.c.flags & GBP_FLAG_OOG) ? 1 : 0; near += (q.c.flags & GBP_FLAG_NEAR) ? 1 : 0; nvalid += ok ? 1 : 0;
				q.phase = PH_IDLE;
			}
		}
	}
	flush_counters(cnt, k, L, np, oog, near, nvalid);
}

// fp64 pass over the candidates the mixed-precision walk could not decide (about 1 %)
template <typename M>
__global__ void __launch_bounds__(128) k_validate_redo(TerrainView T, const int *__restrict__ redo_idx,
														const unsigned long long *__restrict__ redo_count, const double *__restrict__ states,
														const double *__restrict__ actions, const uint8_t *__restrict__ dir, int adaptive,
														uint8_t *__restrict__ verdict, uint8_t *__restrict__ flags, double *__restrict__ s_new,
														double *__restrict__ t_new, unsigned long long *__restrict__ cnt,
														double2 *__restrict__ recipe) {
	const unsigned long long m = *redo_count;
	unsigned long long k = 0, L = 0, np = 0, oog = 0, near = 0, nvalid = 0;
	for (unsigned long long j = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; j < m; j += (unsigned long long) gridDim.x * blockDim.x) {
		const int64_t i = redo_idx[j];
		Cursor q;
		load_state(states + 8 * i, q.s);
		load_action(actions + 10 * i, q.a);
		cursor_start(q, dir[i]);
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
			r = cursor_advance(q, is_valid_state_fast<M>(T, p, ph, q.c), adaptive != 0, out);
		}
		const bool ok = r == 2;
		verdict[i] = ok ? 1 : 0;
		if (flags) flags[i] = (uint8_t) (q.c.flags | (ok ? GBP_FLAG_VALID : 0));
		if (s_new) *(recipe ? recipe + i : reinterpret_cast<double2 *>(s_new + 8 * i)) = make_double2(out.tau, (double) out.kind);
		if (t_new) t_new[i] = q.t_new;
		k += q.c.substates; L += q.c.lookups; np += q.c.nanprobes;
		oog += (q.c.flags & GBP_FLAG_OOG) ? 1 : 0; near += (q.c.flags & GBP_FLAG_NEAR) ? 1 : 0; nvalid += ok ? 1 : 0;
	}
	flush_counters(cnt, k, L, np, oog, near, nvalid);
}

// Second pass of the refill variant: s_new[i] = finish_output(recipe left in s_new[i][0..1]).  One thread per candidate,
// no divergence beyond the 4 recipe kinds; streaming loads / stores.  Latency-bound on the exact fp64 divisions of
// applyStance (8 per candidate).  Measured and rejected: staging the rows through shared memory with cp.async.bulk
// (per-CTA tiles 1.11 ms; persistent CTAs with a 2-stage ring and recipes prefetched one tile ahead 1.27 ms), 128-thread
// CTAs capped at 64 registers with streaming (ld.cs) input loads (1.35 ms), persistent warps each with a private 2-stage
// TMA ring and no block barrier at all (1.41 ms), against 1.05 ms for this form — the
// block-wide barriers per tile cost more than the row-strided accesses they remove; a non-persistent warp-private tile with plain
// coalesced 512-byte LDG / STG requests and __syncwarp only (13 full-line wavefronts per warp instead of ~230 single-sector
// ones): 0.927 against 0.937 ms, i.e. the access pattern is NOT what limits this pass (a warp lives ~11 k cycles for 281
// instructions, 44 % of them waiting on its one batch of loads; 24 warps / SM at 66 registers) — not kept; the same tile
// software-pipelined (persistent warps, cp.async double buffer: the loads of tile t+1 in flight while tile t is computed
// and stored, 92 registers, 2 / 3 / 4 CTAs of 4 warps per SM): 0.914 / 0.918 / 0.960 against 0.926 ms in the same run.
// Five structurally different forms land within 4 % of each other at ~4.1 TB/s of DRAM traffic (2.70 GB read in three
// streams + 1.06 GB written): the pass sits at what the memory system delivers for this mix, not at a kernel-side limit;
// and an in-kernel shared-memory output queue inside the walk
// (12.5-22.6 ms against 10.7 ms at the time: it shrinks the L1 the terrain gathers live on).
static __global__ void __launch_bounds__(256) k_pair_outputs(int64_t n, const double *__restrict__ states, const double *__restrict__ actions,
													   double *__restrict__ s_new, const double2 *__restrict__ recipe) {
	const int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double2 rc = recipe ? __ldcs(recipe + i) : *reinterpret_cast<const double2 *>(s_new + 8 * i);
	double s[8], a[10], sn[8];
	load_state(states + 8 * i, s);
	load_action(actions + 10 * i, a);
	finish_output(s, a, (int) rc.y, rc.x, sn);
	double2 *o = reinterpret_cast<double2 *>(s_new + 8 * i);
#pragma unroll
	for (int d = 0; d < 4; ++d) __stcs(o + d, make_double2(sn[2 * d], sn[2 * d + 1]));
}

// ------------------------------------------------------------------ plan output
// getInterpPath / interpStateActionPair (planning_utils.cpp:142-193).  The sample grid (which primitive, which phase,
// local time: the reference's fp64-accumulated t += dt loops) is laid out by the host; one thread per sample evaluates
// the state with the exact primitives.  kind 0: applyStance(s, a, t); 1: applyFlight(applyStance(s, a), t) (the
// landing sample is kind 1 at t = t_f); 2: the state itself (the closing sample of the path).
static __global__ void k_interp_samples(int64_t m, const int *__restrict__ prim, const uint8_t *__restrict__ kind, const double *__restrict__ tloc,
								 const double *__restrict__ states, const double *__restrict__ actions, double *__restrict__ out) {
	const int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i >= m) return;
	double s[8], a[10], o[8], tmp[8];
	const int p = prim[i], k = kind[i];
#pragma unroll
	for (int d = 0; d < 8; ++d) s[d] = states[8 * (size_t) p + d];
	if (k != 2) {
#pragma unroll
		for (int d = 0; d < 10; ++d) a[d] = actions[10 * (size_t) p + d];
	}
	if (k == 0) apply_stance(s, a, tloc[i], o);
	else if (k == 1) { apply_stance(s, a, a[6], tmp); apply_flight(tmp, tloc[i], o); }
	else {
#pragma unroll
		for (int d = 0; d < 8; ++d) o[d] = s[d];
	}
	store_state(out + 8 * i, o);
}
// calculateMaxCurvature (planning_utils.cpp:884-909): three-point curvature of consecutive plan states, maximum over the
// plan.  std::max(max, c) never takes a NaN c, curvatures are >= 0: the maximum is an atomicMax on the fp64 bit pattern.
static __global__ void k_max_curvature(int64_t n, const double *__restrict__ states, unsigned long long *__restrict__ max_bits) {
	const int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i + 2 >= n) return;
	const double x1 = states[8 * i], y1 = states[8 * i + 1], x2 = states[8 * (i + 1)], y2 = states[8 * (i + 1) + 1],
				 x3 = states[8 * (i + 2)], y3 = states[8 * (i + 2) + 1];
	const double c = curvature3(x1, y1, x2, y2, x3, y3);
	if (c > 0.0) atomicMax(max_bits, (unsigned long long) __double_as_longlong(c));
}

// ------------------------------------------------------------------ terrain generator of the publisher node
// TerrainMapPublisher::createOwnMap / changeOwnMapZDataRectangle[Random] (terrain_map_publisher.cpp:34-176): a flat
// 221 x 161 map at 5 cm on which a list of rectangles is filled, in order, with truncated-Gaussian heights.  The
// reference draws them from a time(0)-seeded engine; here every (rectangle, cell) owns a TERRAIN cell of the Philox
// stream (purpose 3, idx = iy * x_size + ix, stream = rectangle number), so a cell only evaluates the LAST rectangle that
// covers it.  One thread per cell; the float layer is written in grid_map index order (:88-93).
struct OwnMapRect { int x1, y1, x2, y2; double mu, delta; };
static __global__ void k_own_map(uint64_t seed, int x_size, int y_size, int n_rect, const OwnMapRect *__restrict__ rects,
						  float *__restrict__ elevation) {
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= x_size * y_size) return;
	const int iy = c / x_size, ix = c - iy * x_size;
	double val = 0.0;  // z_data starts as zeros (:63)
	for (int r = n_rect - 1; r >= 0; --r) {
		const OwnMapRect q = rects[r];
		if (ix < q.x1 || ix >= q.x2 || iy < q.y1 || iy >= q.y2) continue;
		val = q.mu;
		if (q.delta > 0.0 && q.mu == q.mu) {  // rejection loop of :171-173, two attempts per Box-Muller pair
			const double lo = q.mu - q.delta, hi = q.mu + q.delta;
			bool done = false;
			for (int b = 0; b < 16 && !done; ++b) {
				double ua, ub, z0, z1;
				uniform_pair(seed, (uint64_t) r, (uint64_t) c, 3, b, ua, ub);
				box_muller(ua, ub, z0, z1);
				const double v0 = z0 * q.delta + q.mu, v1 = z1 * q.delta + q.mu;
				if (!(v0 < lo || v0 > hi)) { val = v0; done = true; }
				else if (!(v1 < lo || v1 > hi)) { val = v1; done = true; }
			}
		}
		break;
	}
	elevation[(size_t) ((x_size - 1) - ix) * y_size + ((y_size - 1) - iy)] = (float) val;
}

// ------------------------------------------------------------------ samplers
static __global__ void k_sample_actions(uint64_t seed, uint64_t stream, uint64_t idx0, int64_t n, double n0, double n1, double n2,
								 int dir_flag, double dir_thresh, const double *__restrict__ s_from, const double *__restrict__ s_to,
								 double *__restrict__ out) {
	int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i >= n) return;
	double nn[3] = {n0, n1, n2}, R[9], a[10], sf[8], st[8];
	grf_rotation(nn, R);
	if (dir_flag) {
#pragma unroll
		for (int d = 0; d < 8; ++d) { sf[d] = s_from[d]; st[d] = s_to[d]; }
	}
	sample_action(seed, stream, idx0 + (uint64_t) i, R, dir_flag != 0, dir_thresh, sf, st, a);
	double2 *q = reinterpret_cast<double2 *>(out + 10 * i);
#pragma unroll
	for (int d = 0; d < 5; ++d) q[d] = make_double2(a[2 * d], a[2 * d + 1]);
}
template <typename M>
__global__ void k_sample_states(TerrainView T, uint64_t seed, uint64_t stream, uint64_t idx0, int64_t n, int dir_flag,
								double dir_thresh, int speed_dir, const double *__restrict__ s_from, const double *__restrict__ s_to,
								double *__restrict__ out) {
	int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i >= n) return;
	double q[8], sf[8], st[8];
	if (dir_flag) {
#pragma unroll
		for (int d = 0; d < 8; ++d) { sf[d] = s_from[d]; st[d] = s_to[d]; }
	}
	sample_state<M>(T, seed, stream, idx0 + (uint64_t) i, dir_flag != 0, dir_thresh, speed_dir != 0, sf, st, q);
	store_state(out + 8 * i, q);
}

// ------------------------------------------------------------------ tree queries
// getNearestNeighbor (planner_class.cpp:185-200): one CTA per query, fp64 distances in the
// reference's accumulation order, (distance, id) lexicographic argmin by warp shuffles.
__device__ __forceinline__ void argmin_combine(double &d, int &i, double od, int oi) {
	if (od < d || (od == d && oi < i)) { d = od; i = oi; }
}
__device__ __forceinline__ double vertex_distance(const TreeView &T, int j, const double q[8]) {
	double sum = 0;
#pragma unroll
	for (int d = 0; d < 8; ++d) {
		double vd = T.v[(size_t) d * T.cap + j];
		sum = sum + 1.0 * (vd - q[d]) * (vd - q[d]);
	}
	return sqrt(sum);
}
__device__ __forceinline__ void warp_argmin(double &d, int &i) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		double od = __shfl_xor_sync(FULL, d, o);
		int oi = __shfl_xor_sync(FULL, i, o);
		argmin_combine(d, i, od, oi);
	}
}
static __global__ void __launch_bounds__(256) k_nearest(TreeView T, int64_t m, const double *__restrict__ queries, int *__restrict__ idx,
												  double *__restrict__ dist) {
	__shared__ double sd[8];
	__shared__ int si[8];
	const int nv = *T.n;
	for (int64_t qi = blockIdx.x; qi < m; qi += gridDim.x) {
		double q[8];
#pragma unroll
		for (int d = 0; d < 8; ++d) q[d] = queries[8 * qi + d];
		double bd = INFINITY;
		int bi = 0x7fffffff;
		for (int j = threadIdx.x; j < nv; j += blockDim.x) argmin_combine(bd, bi, vertex_distance(T, j, q), j);
		warp_argmin(bd, bi);
		if ((threadIdx.x & 31) == 0) { sd[threadIdx.x >> 5] = bd; si[threadIdx.x >> 5] = bi; }
		__syncthreads();
		if (threadIdx.x < 32) {
			bd = threadIdx.x < (blockDim.x >> 5) ? sd[threadIdx.x] : INFINITY;
			bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : 0x7fffffff;
			warp_argmin(bd, bi);
			if (threadIdx.x == 0) {
				idx[qi] = bi == 0x7fffffff ? 0 : bi;  // reference default index 0 (planner_class.cpp:186)
				if (dist) dist[qi] = bd;
			}
		}
		__syncthreads();
	}
}
// Many queries against one tree: QT queries per CTA pass, so a vertex is loaded once for QT distance evaluations
// (the one-query form moves 64 B per 24 flop and sits at 41 % of the fp64 issue rate; profiles/r1b_measured_nn.json).
// Same arithmetic per (query, vertex) pair, same (distance, id) argmin: results are bit-identical to k_nearest.
template <int QT>
__global__ void __launch_bounds__(256) k_nearest_tiled(TreeView T, int64_t m, const double *__restrict__ queries, int *__restrict__ idx,
														double *__restrict__ dist) {
	__shared__ double sd[QT][8];
	__shared__ int si[QT][8];
	const int nv = *T.n;
	for (int64_t q0 = (int64_t) blockIdx.x * QT; q0 < m; q0 += (int64_t) gridDim.x * QT) {
		double q[QT][8], bd[QT];
		int bi[QT];
#pragma unroll
		for (int t = 0; t < QT; ++t) {
			const int64_t qi = min(q0 + t, m - 1);  // a partial tile repeats its last query (results of the repeats are not written)
#pragma unroll
			for (int d = 0; d < 8; ++d) q[t][d] = queries[8 * qi + d];
			bd[t] = INFINITY;
			bi[t] = 0x7fffffff;
		}
		for (int j = threadIdx.x; j < nv; j += blockDim.x) {
			double v[8];
#pragma unroll
			for (int d = 0; d < 8; ++d) v[d] = T.v[(size_t) d * T.cap + j];
#pragma unroll
			for (int t = 0; t < QT; ++t) {
				double sum = 0;
#pragma unroll
				for (int d = 0; d < 8; ++d) sum = sum + 1.0 * (v[d] - q[t][d]) * (v[d] - q[t][d]);
				argmin_combine(bd[t], bi[t], sqrt(sum), j);
			}
		}
#pragma unroll
		for (int t = 0; t < QT; ++t) {
			warp_argmin(bd[t], bi[t]);
			if ((threadIdx.x & 31) == 0) { sd[t][threadIdx.x >> 5] = bd[t]; si[t][threadIdx.x >> 5] = bi[t]; }
		}
		__syncthreads();
		if (threadIdx.x < 32) {
#pragma unroll
			for (int t = 0; t < QT; ++t) {
				double d = threadIdx.x < (blockDim.x >> 5) ? sd[t][threadIdx.x] : INFINITY;
				int i = threadIdx.x < (blockDim.x >> 5) ? si[t][threadIdx.x] : 0x7fffffff;
				warp_argmin(d, i);
				if (threadIdx.x == 0 && q0 + t < m) {
					idx[q0 + t] = i == 0x7fffffff ? 0 : i;  // reference default index 0 (planner_class.cpp:186)
					if (dist) dist[q0 + t] = d;
				}
			}
		}
		__syncthreads();
	}
}
// neighborhoodDist (planner_class.cpp:173-182): one warp, ballot/popc compaction keeps ascending ids
static __global__ void k_near(TreeView T, const double *__restrict__ query, double radius, int *__restrict__ ids, int cap,
					   int *__restrict__ count) {
	const int nv = *T.n, lane = threadIdx.x;
	double q[8];
#pragma unroll
	for (int d = 0; d < 8; ++d) q[d] = query[d];
	int base = 0;
	for (int j0 = 0; j0 < nv; j0 += 32) {
		int j = j0 + lane;
		bool in = false;
		if (j < nv) {
			double d = vertex_distance(T, j, q);
			in = (d <= radius) && (d > 0);
		}
		unsigned m = __ballot_sync(FULL, in);
		int pos = base + __popc(m & ((1u << lane) - 1));
		if (in && pos < cap) ids[pos] = j;
		base += __popc(m);
	}
	if (lane == 0) *count = base;
}

static __global__ void k_tree_init(TreeView T, const double *__restrict__ root) {
	if (threadIdx.x == 0) {
		*T.n = 1;
		for (int d = 0; d < 8; ++d) T.v[(size_t) d * T.cap] = root[d];
		for (int d = 0; d < 10; ++d) T.act[(size_t) d * T.cap] = 0.0;
		T.parent[0] = -1;
		T.g[0] = 0;
		T.y[0] = 0;
	}
}
static __global__ void k_tree_append(TreeView T, int parent, const double *__restrict__ s, const double *__restrict__ a, int *__restrict__ out) {
	if (threadIdx.x == 0) {
		double ss[8], aa[10];
		for (int d = 0; d < 8; ++d) ss[d] = s[d];
		for (int d = 0; d < 10; ++d) aa[d] = a[d];
		int n = *T.n;
		if (n >= T.cap || parent < 0 || parent >= n) { *out = -1; return; }
		*out = tree_push(T, parent, ss, aa);
	}
}
// bulk load: AoS host layout -> SoA, g / yaw rebuilt in id order (parents precede children)
static __global__ void k_tree_load(TreeView T, int n, const double *__restrict__ s, const double *__restrict__ a, const int *__restrict__ parent) {
	for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < n; i += blockDim.x * gridDim.x) {
		for (int d = 0; d < 8; ++d) T.v[(size_t) d * T.cap + i] = s[8 * (size_t) i + d];
		for (int d = 0; d < 10; ++d) T.act[(size_t) d * T.cap + i] = a ? a[10 * (size_t) i + d] : 0.0;
		T.parent[i] = i == 0 ? -1 : parent[i];
	}
}
static __global__ void k_tree_gy(TreeView T, int n) {
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		*T.n = n;
		T.g[0] = 0;
		T.y[0] = 0;
		for (int i = 1; i < n; ++i) {
			int p = T.parent[i];
			double a[8], b[8];
			tree_get(T, p, a);
			tree_get(T, i, b);
			T.g[i] = T.g[p] + pose_distance(a, b);
			T.y[i] = T.y[p] + yaw_distance(a, b);
		}
	}
}
static __global__ void k_tree_read(TreeView T, int first, int n, double *__restrict__ s, double *__restrict__ a, int *__restrict__ parent,
							double *__restrict__ g, double *__restrict__ y) {
	for (int k = threadIdx.x + blockIdx.x * blockDim.x; k < n; k += blockDim.x * gridDim.x) {
		int i = first + k;
		if (s) for (int d = 0; d < 8; ++d) s[8 * (size_t) k + d] = T.v[(size_t) d * T.cap + i];
		if (a) for (int d = 0; d < 10; ++d) a[10 * (size_t) k + d] = T.act[(size_t) d * T.cap + i];
		if (parent) parent[k] = T.parent[i];
		if (g) g[k] = T.g[i];
		if (y) y[k] = T.y[i];
	}
}

// ------------------------------------------------------------------ extend (rrt.cpp:20-102)
struct ExtendScratch {
	int *near_idx;       // [1] result of k_nearest
	double *near_dist;   // [1]
	uint8_t *valid;      // [K]
	double *dist;        // [K] stateDistance(s_test, target)
	double *s_test;      // [K][8]
	int *result;         // [4] status, new id, first-valid index / checks, pad
};
// (gbp_extend's kernel, k_extend_fused, lives in gbp_planner.cuh next to group_validate)

// ------------------------------------------------------------------ attemptConnect / connect
template <typename M>
__global__ void __launch_bounds__(128) k_attempt_connect(TerrainView T, int64_t n, const double *__restrict__ s_existing,
														  const double *__restrict__ s, const double *__restrict__ t_s,
														  const uint8_t *__restrict__ dir, int adaptive,
														  int *__restrict__ status, double *__restrict__ s_new,
														  double *__restrict__ a_new, uint8_t *__restrict__ flags) {
	int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
	if (i >= n) return;
	double se[8], sg[8], sn[8], an[10];
	load_state(s_existing + 8 * i, se);
	load_state(s + 8 * i, sg);
#pragma unroll
	for (int d = 0; d < 8; ++d) sn[d] = sg[d];
#pragma unroll
	for (int d = 0; d < 10; ++d) an[d] = 0;
	Counters c = {0, 0, 0, 0};
	unsigned checks = 0;
	status[i] = attempt_connect<M>(T, se, sg, dir[i], adaptive != 0, sn, an, c, checks, t_s ? t_s[i] : -1.0);
	store_state(s_new + 8 * i, sn);
#pragma unroll
	for (int d = 0; d < 10; ++d) a_new[10 * i + d] = an[d];
	if (flags) flags[i] = (uint8_t) c.flags;
}
// attemptConnect with the pair check spread over the warp (validate_pair_warp: a lane per sub-state, identical outputs);
// connect primitives are long (t_s = distance / 0.75: a 4 m connection is 107 sub-states).  Fixed step only; all lanes
// call with uniform arguments and get uniform results.
template <typename M>
__device__ int attempt_connect_warp(const TerrainView &T, const double s_existing[8], const double s_in[8], int direction, double s_new[8],
									double a_new[10], Counters &c, unsigned &pair_checks) {
	double target[8], ts = pose_distance(s_in, s_existing) / V_NOM;  // rrt_connect.cpp:29
#pragma unroll
	for (int i = 0; i < 8; ++i) target[i] = s_in[i];
	for (int depth = 0;; ++depth) {
		if (ts <= KINEMATICS_RES) return GBP_TRAPPED;
		if (direction == GBP_FORWARD) connect_action(s_existing, target, ts, a_new);
		else connect_action(target, s_existing, ts, a_new);
		if (!is_valid_action(a_new)) return GBP_TRAPPED;
		double out[8], tn;
		++pair_checks;
		const bool ok = validate_pair_warp<M>(T, s_existing, a_new, direction, out, tn, c);
#pragma unroll
		for (int i = 0; i < 8; ++i) { s_new[i] = out[i]; target[i] = out[i]; }
		if (ok) return depth == 0 ? GBP_REACHED : GBP_ADVANCED;
		ts = tn;
	}
}
// gbp_connect in one launch of one warp (rrt_connect.cpp:98-120): nearest neighbour by the warp (same arithmetic and
// (distance, id) argmin as k_nearest), attemptConnect, append; the target travels as a kernel argument and the result
// words also go to the caller's mapped host buffer.
struct Target8 { double v[8]; };
template <typename M>
__global__ void __launch_bounds__(32) k_connect(TerrainView T, TreeView tree, Target8 tgt, int direction, int adaptive, ExtendScratch S,
												  int *__restrict__ host_result) {
	const int lane = threadIdx.x;
	double tg[8], s_near[8], sn[8], an[10];
#pragma unroll
	for (int d = 0; d < 8; ++d) tg[d] = tgt.v[d];
	const int nv = *tree.n;
	double bd = INFINITY;
	int bi = 0x7fffffff;
	for (int j = lane; j < nv; j += 32) argmin_combine(bd, bi, vertex_distance(tree, j, tg), j);
	warp_argmin(bd, bi);
	const int near = bi == 0x7fffffff ? 0 : bi;  // reference default index 0 (planner_class.cpp:186)
	tree_get(tree, near, s_near);
	Counters c = {0, 0, 0, 0};
	unsigned checks = 0;
	int r = adaptive ? attempt_connect<M>(T, s_near, tg, direction, true, sn, an, c, checks)
					 : attempt_connect_warp<M>(T, s_near, tg, direction, sn, an, c, checks);
	if (lane != 0) return;
	*S.near_idx = near; *S.near_dist = bd;
	int new_id = -1;
	if (r != GBP_TRAPPED) {
		if (nv < tree.cap) new_id = tree_push(tree, near, sn, an);
		else r = -1;  // tree full: reported as GBP_E_CAPACITY, not as TRAPPED
	}
	S.result[0] = r; S.result[1] = new_id; S.result[2] = (int) checks;
	host_result[0] = r; host_result[1] = new_id; host_result[2] = (int) checks;
}

}  // namespace gbp
