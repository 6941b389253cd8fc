// TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
//
// C entry points around the UNMODIFIED reference core, compiled where it lies under
// /root/reference/src by oracle/Makefile into oracle/_ref/libgbp_ref.so.  Used by tests/ (to pin
// the oracle restatement and to mint tests/golden/), and by bench.py's reference arm / cpu_baseline
// leg (kind "reference").  Nothing here re-implements reference arithmetic: every function only
// marshals flat arrays into the reference's own State/Action/FastTerrainMap types and calls it.
//
// Layout conventions (shared with include/gbp_b200.h): State = 8 doubles, Action = 10 doubles,
// terrain layers x-major ([ix*ny + iy]), direction 0 = FORWARD, 1 = REVERSE, phase 0 = FLIGHT,
// 1 = STANCE (planning_utils.h:43-49).
#include <global_body_planner/rrt_star_connect.h>

#include <atomic>
#include <cstring>
#include <sstream>
#include <thread>
#include <vector>

using namespace planning_utils;

namespace {
State to_state(const double *p) { State s; for (int i = 0; i < 8; ++i) s[i] = p[i]; return s; }
Action to_action(const double *p) { Action a; for (int i = 0; i < 10; ++i) a[i] = p[i]; return a; }
void from_state(const State &s, double *p) { for (int i = 0; i < 8; ++i) p[i] = s[i]; }
void from_action(const Action &a, double *p) { for (int i = 0; i < 10; ++i) p[i] = a[i]; }

// Exposes the protected knobs/fields of the planner classes without touching reference code.
struct ConnectProbe : public RRTStarConnectClass {
	using RRTClass::goal_found;
	using RRTClass::path_length_;
	using RRTClass::path_cost_;
	using RRTClass::path_yaw_;
};

std::atomic<long long> g_pair_checks(0);   // 6-arg isValidStateActionPair[Reverse] calls (cross-TU)
std::atomic<long long> g_nn_queries(0);    // PlannerClass::getNearestNeighbor calls
}  // namespace

// ---- ld --wrap counters (see oracle/Makefile).  Cross-TU calls only; reference code is unmodified.
extern "C" {
bool __real__ZN14planning_utils22isValidStateActionPairESt5arrayIdLm8EES0_IdLm10EER14FastTerrainMapRS1_Rdb(
	State, Action, FastTerrainMap &, State &, double &, bool);
bool __wrap__ZN14planning_utils22isValidStateActionPairESt5arrayIdLm8EES0_IdLm10EER14FastTerrainMapRS1_Rdb(
	State s, Action a, FastTerrainMap &t, State &sn, double &tn, bool f) {
	g_pair_checks.fetch_add(1, std::memory_order_relaxed);
	return __real__ZN14planning_utils22isValidStateActionPairESt5arrayIdLm8EES0_IdLm10EER14FastTerrainMapRS1_Rdb(s, a, t, sn, tn, f);
}
bool __real__ZN14planning_utils29isValidStateActionPairReverseESt5arrayIdLm8EES0_IdLm10EER14FastTerrainMapRS1_Rdb(
	State, Action, FastTerrainMap &, State &, double &, bool);
bool __wrap__ZN14planning_utils29isValidStateActionPairReverseESt5arrayIdLm8EES0_IdLm10EER14FastTerrainMapRS1_Rdb(
	State s, Action a, FastTerrainMap &t, State &sn, double &tn, bool f) {
	g_pair_checks.fetch_add(1, std::memory_order_relaxed);
	return __real__ZN14planning_utils29isValidStateActionPairReverseESt5arrayIdLm8EES0_IdLm10EER14FastTerrainMapRS1_Rdb(s, a, t, sn, tn, f);
}
int __real__ZN12PlannerClass18getNearestNeighborESt5arrayIdLm8EE(PlannerClass *, State);
int __wrap__ZN12PlannerClass18getNearestNeighborESt5arrayIdLm8EE(PlannerClass *self, State q) {
	g_nn_queries.fetch_add(1, std::memory_order_relaxed);
	return __real__ZN12PlannerClass18getNearestNeighborESt5arrayIdLm8EE(self, q);
}
}

extern "C" {

// ---------------------------------------------------------------- terrain (fast_terrain_map.cpp)
void *ref_terrain_create(int nx, int ny, const double *x, const double *y, const double *z,
						 const double *dx, const double *dy, const double *dz) {
	std::vector<double> xv(x, x + nx), yv(y, y + ny);
	std::vector<std::vector<double> > zz(nx), dxx(nx), dyy(nx), dzz(nx);
	for (int i = 0; i < nx; ++i) {
		zz[i].assign(z + (size_t) i * ny, z + (size_t) (i + 1) * ny);
		dxx[i].assign(dx + (size_t) i * ny, dx + (size_t) (i + 1) * ny);
		dyy[i].assign(dy + (size_t) i * ny, dy + (size_t) (i + 1) * ny);
		dzz[i].assign(dz + (size_t) i * ny, dz + (size_t) (i + 1) * ny);
	}
	FastTerrainMap *t = new FastTerrainMap();
	t->loadData(nx, ny, xv, yv, zz, dxx, dyy, dzz);  // fast_terrain_map.cpp:10-28
	return t;
}

// Terrain through the GridMap ingest path (fast_terrain_map.cpp:31-91).  `elev` etc. are float
// layers in grid_map index order [i*ny + j] (index (0,0) = largest x,y); dxl may be NULL.
void *ref_terrain_create_gridmap(int nx, int ny, double res, double cx, double cy, const float *elev,
								 const float *dxl, const float *dyl, const float *dzl) {
	grid_map::GridMap m;
	m.setGeometry(nx, ny, res, cx, cy);
	m.add("elevation");
	if (dxl) { m.add("dx"); m.add("dy"); m.add("dz"); }
	for (int i = 0; i < nx; ++i)
		for (int j = 0; j < ny; ++j) {
			grid_map::Index idx(i, j);
			m.at("elevation", idx) = elev[(size_t) i * ny + j];
			if (dxl) {
				m.at("dx", idx) = dxl[(size_t) i * ny + j];
				m.at("dy", idx) = dyl[(size_t) i * ny + j];
				m.at("dz", idx) = dzl[(size_t) i * ny + j];
			}
		}
	FastTerrainMap *t = new FastTerrainMap();
	t->loadDataFromGridMap(m);
	return t;
}

void ref_terrain_destroy(void *h) { delete (FastTerrainMap *) h; }

int ref_terrain_axes(void *h, double *x, double *y) {  // returns nx | (ny << 16) when x == NULL
	FastTerrainMap *t = (FastTerrainMap *) h;
	std::vector<double> xv = t->getXData(), yv = t->getYData();
	if (x) std::memcpy(x, xv.data(), xv.size() * sizeof(double));
	if (y) std::memcpy(y, yv.data(), yv.size() * sizeof(double));
	return (int) xv.size() | ((int) yv.size() << 16);
}

void ref_ground_height(void *h, long long n, const double *x, const double *y, double *out) {
	FastTerrainMap *t = (FastTerrainMap *) h;
	for (long long i = 0; i < n; ++i) out[i] = t->getGroundHeight(x[i], y[i]);
}
void ref_height_is_nan(void *h, long long n, const double *x, const double *y, unsigned char *out) {
	FastTerrainMap *t = (FastTerrainMap *) h;
	for (long long i = 0; i < n; ++i) out[i] = t->heightIsNan(x[i], y[i]) ? 1 : 0;
}
void ref_surface_normal(void *h, long long n, const double *x, const double *y, double *out3) {
	FastTerrainMap *t = (FastTerrainMap *) h;
	for (long long i = 0; i < n; ++i) {
		std::array<double, 3> v = t->getSurfaceNormal(x[i], y[i]);
		out3[3 * i] = v[0]; out3[3 * i + 1] = v[1]; out3[3 * i + 2] = v[2];
	}
}

// ---------------------------------------------------------------- primitives (planning_utils.cpp)
void ref_apply_stance(long long n, const double *s, const double *a, const double *t, double *out) {
	for (long long i = 0; i < n; ++i) from_state(applyStance(to_state(s + 8 * i), to_action(a + 10 * i), t[i]), out + 8 * i);
}
void ref_apply_flight(long long n, const double *s, const double *t, double *out) {
	for (long long i = 0; i < n; ++i) from_state(applyFlight(to_state(s + 8 * i), t[i]), out + 8 * i);
}
void ref_apply_stance_reverse(long long n, const double *s, const double *a, const double *t, double *out) {
	for (long long i = 0; i < n; ++i) from_state(applyStanceReverse(to_state(s + 8 * i), to_action(a + 10 * i), t[i]), out + 8 * i);
}
void ref_apply_action(long long n, const double *s, const double *a, double *out) {
	for (long long i = 0; i < n; ++i) from_state(applyAction(to_state(s + 8 * i), to_action(a + 10 * i)), out + 8 * i);
}
void ref_rotate_grf(long long n, const double *normal3, const double *f3, double *out3) {
	for (long long i = 0; i < n; ++i) {
		std::array<double, 3> nn = {normal3[3 * i], normal3[3 * i + 1], normal3[3 * i + 2]};
		std::array<double, 3> ff = {f3[3 * i], f3[3 * i + 1], f3[3 * i + 2]};
		std::array<double, 3> r = rotate_grf(nn, ff);
		out3[3 * i] = r[0]; out3[3 * i + 1] = r[1]; out3[3 * i + 2] = r[2];
	}
}
void ref_is_valid_action(long long n, const double *a, unsigned char *out) {
	for (long long i = 0; i < n; ++i) out[i] = isValidAction(to_action(a + 10 * i)) ? 1 : 0;
}
void ref_is_valid_state(void *h, long long n, const double *s, const unsigned char *phase, unsigned char *out) {
	FastTerrainMap *t = (FastTerrainMap *) h;
	for (long long i = 0; i < n; ++i) out[i] = isValidState(to_state(s + 8 * i), *t, (int) phase[i]) ? 1 : 0;
}

// Distances (planning_utils.cpp:106-127, planning_utils.h:133-155). kind: 0 pose, 1 state, 2 yaw.
void ref_distance(long long n, const double *q1, const double *q2, int kind, double *out) {
	for (long long i = 0; i < n; ++i) {
		State a = to_state(q1 + 8 * i), b = to_state(q2 + 8 * i);
		out[i] = kind == 0 ? poseDistance(a, b) : kind == 1 ? stateDistance(a, b) : stateYawDistance(a, b);
	}
}

// isValidStateActionPair / ...Reverse (fixed step :713-753,:837-876; adaptive :651-712,:774-836),
// through the 6-arg dispatchers.  Outputs are pre-filled with NaN so that entries the reference
// leaves unwritten (t_new when the very first sub-state fails) are recognisable.
// `nthreads` > 1 splits [0,n) into disjoint slices (the functions only read the terrain).
void ref_validate_pairs(void *h, long long n, const double *s, const double *a, const unsigned char *dir,
						int adaptive, unsigned char *verdict, double *s_new, double *t_new, int nthreads) {
	FastTerrainMap *t = (FastTerrainMap *) h;
	auto work = [=](long long lo, long long hi) {
		for (long long i = lo; i < hi; ++i) {
			State sn; sn.fill(std::numeric_limits<double>::quiet_NaN());
			double tn = std::numeric_limits<double>::quiet_NaN();
			bool ok = dir[i] == FORWARD
				? isValidStateActionPair(to_state(s + 8 * i), to_action(a + 10 * i), *t, sn, tn, adaptive != 0)
				: isValidStateActionPairReverse(to_state(s + 8 * i), to_action(a + 10 * i), *t, sn, tn, adaptive != 0);
			verdict[i] = ok ? 1 : 0;
			if (s_new) from_state(sn, s_new + 8 * i);
			if (t_new) t_new[i] = tn;
		}
	};
	if (nthreads <= 1) { work(0, n); return; }
	std::vector<std::thread> th;
	for (int k = 0; k < nthreads; ++k) th.emplace_back(work, n * k / nthreads, n * (k + 1) / nthreads);
	for (auto &x : th) x.join();
}

// ---------------------------------------------------------------- tree queries (planner_class.cpp)
// Vertices are inserted with ids 0..nv-1 in order (as rrt.cpp:87 allocates them).
void ref_nearest(long long nv, const double *verts, long long nq, const double *q, int *idx, double *dist) {
	PlannerClass T;
	for (long long i = 0; i < nv; ++i) T.addVertex((int) i, to_state(verts + 8 * i));
	for (long long j = 0; j < nq; ++j) {
		State qs = to_state(q + 8 * j);
		int k = T.getNearestNeighbor(qs);  // planner_class.cpp:185-200
		idx[j] = k;
		if (dist) dist[j] = stateDistance(qs, T.getVertex(k));
	}
}
// neighborhoodDist (planner_class.cpp:173-182). Returns the count; ids written in map iteration order.
long long ref_near(long long nv, const double *verts, const double *q, double radius, int *ids, long long cap) {
	PlannerClass T;
	for (long long i = 0; i < nv; ++i) T.addVertex((int) i, to_state(verts + 8 * i));
	std::vector<int> r = T.neighborhoodDist(to_state(q), radius);
	for (size_t i = 0; i < r.size() && (long long) i < cap; ++i) ids[i] = r[i];
	return (long long) r.size();
}
// g / yaw bookkeeping of a chain built with addEdge+updateGYValue as extend does (rrt.cpp:87-92).
void ref_tree_gy(long long nv, const double *verts, const int *parent, double *g, double *yv) {
	PlannerClass T;
	T.init(to_state(verts), false, 1, 1);
	for (long long i = 1; i < nv; ++i) {
		State s = to_state(verts + 8 * i);
		State p = T.getVertex(parent[i]);
		T.addVertex((int) i, s);
		T.addEdge(parent[i], (int) i);
		T.updateGYValue((int) i, T.getGValue(parent[i]) + poseDistance(p, s), T.getYValue(parent[i]) + stateYawDistance(p, s));
	}
	for (long long i = 0; i < nv; ++i) { g[i] = T.getGValue((int) i); yv[i] = T.getYValue((int) i); }
}

// ---------------------------------------------------------------- planners (rrt*.cpp)
// attemptConnect without explicit t_s (rrt_connect.cpp:85-91 -> :20-84).
void ref_attempt_connect(void *h, long long n, const double *s_existing, const double *s, const unsigned char *dir,
						 int adaptive, int *status, double *s_new, double *a_new) {
	FastTerrainMap *t = (FastTerrainMap *) h;
	RRTConnectClass P;
	P.set_state_action_pair_check_adaptive_step_size_flag_(adaptive != 0);
	for (long long i = 0; i < n; ++i) {
		State sn; sn.fill(std::numeric_limits<double>::quiet_NaN());
		Action an; an.fill(std::numeric_limits<double>::quiet_NaN());
		status[i] = P.attemptConnect(to_state(s_existing + 8 * i), to_state(s + 8 * i), sn, an, *t, (int) dir[i]);
		from_state(sn, s_new + 8 * i);
		from_action(an, a_new + 10 * i);
	}
}

// postProcessPath (rrt_connect.cpp:139-227). In/out: ns states, ns-1 actions; returns new ns.
// stats3 = {path_length_, path_yaw_, path_cost_} as the reference leaves them.
int ref_post_process_path(void *h, int ns, double *states, double *actions, int cap, double *stats3) {
	FastTerrainMap *t = (FastTerrainMap *) h;
	ConnectProbe P;
	std::vector<State> ss; std::vector<Action> aa;
	for (int i = 0; i < ns; ++i) ss.push_back(to_state(states + 8 * i));
	for (int i = 0; i + 1 < ns; ++i) aa.push_back(to_action(actions + 10 * i));
	P.postProcessPath(ss, aa, *t);
	int m = (int) ss.size();
	for (int i = 0; i < m && i < cap; ++i) from_state(ss[i], states + 8 * i);
	for (int i = 0; i < (int) aa.size() && i + 1 < cap; ++i) from_action(aa[i], actions + 10 * i);
	if (stats3) { stats3[0] = P.path_length_; stats3[1] = P.path_yaw_; stats3[2] = P.path_cost_; }
	return m;
}

// getInterpPath (planning_utils.cpp:175-192) and calculateMaxCurvature (:900-909), unmodified
long long ref_interp_path(int n_actions, const double *states, const double *actions, double dt, long long cap, double *out_s,
						  double *out_t, int *out_phase) {
	std::vector<State> ss, path; std::vector<Action> aa; std::vector<double> tt; std::vector<int> ph;
	for (int i = 0; i <= n_actions; ++i) ss.push_back(to_state(states + 8 * i));
	for (int i = 0; i < n_actions; ++i) aa.push_back(to_action(actions + 10 * i));
	getInterpPath(ss, aa, dt, path, tt, ph);
	for (long long i = 0; i < (long long) path.size() && i < cap; ++i) { from_state(path[i], out_s + 8 * i); out_t[i] = tt[i]; }
	for (long long i = 0; i < (long long) ph.size() && i < cap; ++i) out_phase[i] = ph[i];
	return (long long) path.size();
}
double ref_max_curvature(long long n, const double *states) {
	std::vector<State> plan;
	for (long long i = 0; i < n; ++i) plan.push_back(to_state(states + 8 * i));
	return calculateMaxCurvature(plan);
}

// Unmodified buildRRTConnect / buildRRTStarConnect called the way callPlanner does
// (global_body_planner.cpp:113-124).  algorithm 0 = rrt-connect, 1 = rrt-star-connect.
// out: [plan_time, success, vertices, time_to_first, path_length(last cost), path_duration, n_states]
// pair_checks/nn_queries: wrapped counters accumulated during this call.
int ref_plan(void *h, int algorithm, const double *start, const double *goal, double max_time, int adaptive,
			 double *out7, long long *pair_checks, long long *nn_queries, double *states, double *actions, int cap) {
	FastTerrainMap *t = (FastTerrainMap *) h;
	std::streambuf *old = std::cout.rdbuf();
	std::ostringstream sink;
	std::cout.rdbuf(sink.rdbuf());
	long long c0 = g_pair_checks.load(), n0 = g_nn_queries.load();
	std::vector<State> ss; std::vector<Action> aa;
	double plan_time, ttf, dur; int succ, nv;
	std::vector<double> lv, yv, cv, cvt; std::vector<std::vector<double> > all;
	if (algorithm == 0) {
		RRTConnectClass P;
		P.set_state_action_pair_check_adaptive_step_size_flag_(adaptive != 0);
		P.buildRRTConnect(*t, to_state(start), to_state(goal), ss, aa, max_time);
		P.getStatistics(plan_time, succ, nv, ttf, lv, yv, cv, cvt, dur, all);
	} else {
		RRTStarConnectClass P;
		P.set_state_action_pair_check_adaptive_step_size_flag_(adaptive != 0);
		P.buildRRTStarConnect(*t, to_state(start), to_state(goal), ss, aa, max_time);
		P.getStatistics(plan_time, succ, nv, ttf, lv, yv, cv, cvt, dur, all);
	}
	std::cout.rdbuf(old);
	if (pair_checks) *pair_checks = g_pair_checks.load() - c0;
	if (nn_queries) *nn_queries = g_nn_queries.load() - n0;
	out7[0] = plan_time; out7[1] = succ; out7[2] = nv; out7[3] = ttf;
	out7[4] = cv.empty() ? -1.0 : cv.back(); out7[5] = dur; out7[6] = (double) ss.size();
	for (int i = 0; i < (int) ss.size() && i < cap; ++i) from_state(ss[i], states + 8 * i);
	for (int i = 0; i < (int) aa.size() && i < cap; ++i) from_action(aa[i], actions + 10 * i);
	return (int) ss.size();
}

// Bounded-time throughput probe of the unmodified extend/connect loop: runs runRRTConnect restarts
// (exactly buildRRTConnect's inner loop, rrt_connect.cpp:350-360) for `seconds` of wall clock and
// reports wrapped counters.  Used by bench.py --impl reference.
void ref_counters(long long *pair_checks, long long *nn_queries) {
	*pair_checks = g_pair_checks.load();
	*nn_queries = g_nn_queries.load();
}

}  // extern "C"
