// TEST INFRASTRUCTURE ONLY — never linked into, imported by or executed from the product path.
//
// Pins the Tier-2 oracle planner (oracle/gbp_oracle.c: orc_plan) to the UNMODIFIED reference's own loops:
// RRTClass::extend / newConfig (src/rrt.cpp:20-102), RRTConnectClass::connect / runRRTConnect
// (src/rrt_connect.cpp:98-120, :230-314) and RRTStarConnectClass::extend (src/rrt_star_connect.cpp:12-75), compiled where
// they lie under /root/reference/src by oracle/Makefile into oracle/_ref/libgbp_ref_pin.so.
//
// The reference's loops are not reproducible as shipped: they draw from rand() and from engines seeded with the wall
// clock (planning_utils.cpp:435-439, planner_class.cpp:51-60), and they read undefined memory for out-of-grid terrain
// probes.  This library leaves every reference translation unit untouched and interposes, with
// `ld --wrap` (cross-TU calls only), exactly the functions where reproducibility is lost:
//   * planning_utils::getRandomAction(surf_norm, direction, flag, threshold, s, s_near)   <- ACTION cells of the Philox stream
//   * PlannerClass::randomState(terrain, flag, threshold, speed flag, s_from, s_to), randomState(terrain)  <- STATE cells
//     (cell numbering as orc_plan: STATE cell = number of randomState calls so far, ACTION cell = cell * 6 + draws since)
//   * FastTerrainMap::getGroundHeight / heightIsNan: in-grid queries go to the REAL function; out-of-grid queries, where
//     the reference has undefined behaviour, get the defined semantics of the oracle (cell-0-anchored extrapolation)
//   * PlannerClass::neighborhoodDist: the REAL result, sorted into ascending id (the reference's order is that of
//     std::unordered_map and changes at every rehash, Appendix B-4; RRT* parent choice and rewiring depend on it)
// A randomState call beyond the cell budget throws, which ends runRRTConnect after a fixed number of iterations (its own
// exit is a wall-clock horizon, set out of reach here).  Everything else — nearest neighbour, newConfig's acceptance,
// attemptConnect, the pair checks, tree bookkeeping, g / yaw values, path extraction — is the reference's code.
#include <global_body_planner/rrt_star_connect.h>

#include <algorithm>
#include <cstring>
#include <sstream>
#include <vector>

extern "C" {
#include "gbp_oracle.h"
}

using namespace planning_utils;

namespace {
struct PinBudget {};  // thrown by the randomState wrapper when the cell budget is used up

struct PinState {
	orc_terrain terrain;
	bool have_terrain = false;
	uint64_t seed = 0, query = 0;
	long long cell = -1, max_cells = 0;  // STATE cell of the current half-iteration
	int draws = 0;                       // getRandomAction calls since the last randomState
	int k = 6;
	bool sort_near = true;
	long long oog_height = 0, oog_nan = 0, near_sets = 0, near_reordered = 0;
} g;

State to_state(const double *p) { State s; for (int i = 0; i < 8; ++i) s[i] = p[i]; return s; }
void from_state(const State &s, double *p) { for (int i = 0; i < 8; ++i) p[i] = s[i]; }
void from_action(const Action &a, double *p) { for (int i = 0; i < 10; ++i) p[i] = a[i]; }
bool in_grid(double x, double y) {
	const orc_terrain &t = g.terrain;
	return x >= t.x[0] && x < t.x[t.nx - 1] && y >= t.y[0] && y < t.y[t.ny - 1];
}

// exposes protected members without touching reference code
struct Probe : public RRTStarConnectClass {
	using RRTClass::goal_found;
	using RRTClass::path_length_;
	using RRTClass::path_yaw_;
	using RRTClass::path_cost_;
	using RRTConnectClass::anytime_horizon;
	// RRT-Connect with the plain extend: RRTStarConnectClass overrides the virtual extend, so the plain planner is a
	// different object type (below); both share this accessor set
};
struct ProbeConnect : public RRTConnectClass {
	using RRTClass::goal_found;
	using RRTClass::path_length_;
	using RRTClass::path_yaw_;
	using RRTClass::path_cost_;
	using RRTConnectClass::anytime_horizon;
};
}  // namespace

// --------------------------------------------------------------------------------------------- ld --wrap targets
extern "C" {
Action __wrap__ZN14planning_utils15getRandomActionESt5arrayIdLm3EEibdS0_IdLm8EES2_(std::array<double, 3> surf_norm, int direction, bool flag,
																					double threshold, State s, State s_near) {
	double n[3] = {surf_norm[0], surf_norm[1], surf_norm[2]}, a[10], sf[8], st[8];
	// planning_utils.cpp:385-388: FORWARD samples from s_near towards s, REVERSE from s towards s_near
	from_state(direction == FORWARD ? s_near : s, sf);
	from_state(direction == FORWARD ? s : s_near, st);
	orc_sample_action(g.seed, g.query, (uint64_t) g.cell * (uint64_t) g.k + (uint64_t) g.draws, n, flag ? 1 : 0, threshold, sf, st, a);
	++g.draws;
	Action out;
	for (int i = 0; i < 10; ++i) out[i] = a[i];
	return out;
}
State __wrap__ZN12PlannerClass11randomStateER14FastTerrainMapbdbSt5arrayIdLm8EES3_(PlannerClass *, FastTerrainMap &, bool flag, double threshold,
																				   bool speed_flag, State s_from, State s_to) {
	if (g.cell + 1 >= g.max_cells) throw PinBudget();
	++g.cell;
	g.draws = 0;
	double sf[8], st[8], q[8];
	from_state(s_from, sf);
	from_state(s_to, st);
	orc_sample_state(&g.terrain, g.seed, g.query, (uint64_t) g.cell, flag ? 1 : 0, threshold, speed_flag ? 1 : 0, sf, st, q);
	return to_state(q);
}
State __wrap__ZN12PlannerClass11randomStateER14FastTerrainMap(PlannerClass *, FastTerrainMap &) {
	if (g.cell + 1 >= g.max_cells) throw PinBudget();
	++g.cell;
	g.draws = 0;
	double q[8];
	orc_sample_state(&g.terrain, g.seed, g.query, (uint64_t) g.cell, 0, 0.0, 0, nullptr, nullptr, q);
	return to_state(q);
}
double __real__ZN14FastTerrainMap15getGroundHeightEdd(FastTerrainMap *, double, double);
double __wrap__ZN14FastTerrainMap15getGroundHeightEdd(FastTerrainMap *self, double x, double y) {
	if (!g.have_terrain || in_grid(x, y)) return __real__ZN14FastTerrainMap15getGroundHeightEdd(self, x, y);
	++g.oog_height;
	return orc_ground_height(&g.terrain, x, y, nullptr);
}
bool __real__ZN14FastTerrainMap11heightIsNanEdd(FastTerrainMap *, double, double);
bool __wrap__ZN14FastTerrainMap11heightIsNanEdd(FastTerrainMap *self, double x, double y) {
	if (!g.have_terrain || in_grid(x, y)) return __real__ZN14FastTerrainMap11heightIsNanEdd(self, x, y);
	++g.oog_nan;
	return orc_height_is_nan(&g.terrain, x, y, nullptr) != 0;
}
std::vector<int> __real__ZN12PlannerClass16neighborhoodDistESt5arrayIdLm8EEd(PlannerClass *, State, double);
std::vector<int> __wrap__ZN12PlannerClass16neighborhoodDistESt5arrayIdLm8EEd(PlannerClass *self, State q, double dist) {
	std::vector<int> r = __real__ZN12PlannerClass16neighborhoodDistESt5arrayIdLm8EEd(self, q, dist);
	++g.near_sets;
	if (!std::is_sorted(r.begin(), r.end())) ++g.near_reordered;
	if (g.sort_near) std::sort(r.begin(), r.end());
	return r;
}
}

namespace {
void dump_tree(PlannerClass &T, int cap, double *states, double *actions, int *parent, double *gv, double *yv) {
	const int n = T.getNumVertices();
	for (int i = 0; i < n && i < cap; ++i) {
		from_state(T.getVertex(i), states + 8 * i);
		if (i == 0) std::memset(actions, 0, 80); else from_action(T.getAction(i), actions + 10 * i);
		parent[i] = i == 0 ? -1 : T.getPredecessor(i);
		gv[i] = T.getGValue(i);
		yv[i] = T.getYValue(i);
	}
}
}  // namespace

extern "C" {

// terrain layers x-major as everywhere else; the arrays must stay alive until pin_terrain_clear
void pin_terrain_set(int nx, int ny, const double *x, const double *y, const double *z, const double *dx, const double *dy, const double *dz) {
	std::memset(&g.terrain, 0, sizeof g.terrain);
	g.terrain.nx = nx; g.terrain.ny = ny;
	g.terrain.x = x; g.terrain.y = y; g.terrain.z = z; g.terrain.dx = dx; g.terrain.dy = dy; g.terrain.dz = dz;
	g.have_terrain = true;
}
void pin_terrain_clear(void) { g.have_terrain = false; }

typedef struct {
	int star;           // 0: RRTConnectClass (plain extend), 1: RRTStarConnectClass (virtual extend override)
	int max_iters;      // loop iterations of runRRTConnect to allow (2 STATE cells each)
	int adaptive;
	int sort_near;      // 1: neighborhoodDist results in ascending id (the defined order)
	int state_direction_sampling, state_direction_speed, action_direction_sampling, cost_add_yaw;
	double state_direction_threshold, action_direction_threshold, cost_length_weight, cost_yaw_weight;
} pin_params;

typedef struct {
	int solved, cells_used, nv_a, nv_b, path_states, budget_hit;
	double path_length, path_yaw, path_cost;
	long long oog_height, oog_nan, near_sets, near_reordered;
} pin_result;

// One runRRTConnect of the unmodified reference from fresh trees (rrt_connect.cpp:352-359) on the Philox stream
// (seed, query), stopped after max_iters iterations.  Trees are copied out (cap vertices each); when solved, also the
// stitched path exactly as buildRRTConnect extracts it (rrt_connect.cpp:381-401: pathFromStart, getStateSequence,
// getActionSequence / getActionSequenceReverse), before postProcessPath.
int pin_run(void *terrain_handle, const double *start, const double *goal, uint64_t seed, uint64_t query, const pin_params *p, pin_result *out, int cap,
			double *a_states, double *a_actions, int *a_parent, double *a_g, double *a_y, double *b_states, double *b_actions, int *b_parent,
			double *b_g, double *b_y, int path_cap, double *path_states, double *path_actions) {
	FastTerrainMap *t = (FastTerrainMap *) terrain_handle;
	std::streambuf *old = std::cout.rdbuf();
	std::ostringstream sink;
	std::cout.rdbuf(sink.rdbuf());
	g.seed = seed; g.query = query; g.cell = -1; g.draws = 0; g.k = NUM_GEN_STATES; g.max_cells = 2ll * p->max_iters;
	g.sort_near = p->sort_near != 0;
	g.oog_height = g.oog_nan = g.near_sets = g.near_reordered = 0;
	std::memset(out, 0, sizeof *out);
	PlannerClass Ta, Tb;
	Ta.init(to_state(start), p->cost_add_yaw != 0, p->cost_length_weight, p->cost_yaw_weight);
	Tb.init(to_state(goal), p->cost_add_yaw != 0, p->cost_length_weight, p->cost_yaw_weight);
	Probe star;
	ProbeConnect plain;
	RRTConnectClass *P = p->star ? (RRTConnectClass *) &star : (RRTConnectClass *) &plain;
	P->set_state_action_pair_check_adaptive_step_size_flag_(p->adaptive != 0);
	P->set_cost_add_yaw(p->cost_add_yaw != 0, p->cost_length_weight, p->cost_yaw_weight);
	P->set_state_direction_sampling(p->state_direction_sampling != 0, p->state_direction_threshold, p->state_direction_speed != 0);
	P->set_action_direction_sampling(p->action_direction_sampling != 0, p->action_direction_threshold);
	if (p->star) { star.goal_found = false; star.anytime_horizon = 1e18; } else { plain.goal_found = false; plain.anytime_horizon = 1e18; }
	try {
		P->runRRTConnect(Ta, Tb, *t);  // rrt_connect.cpp:230-314; extend is virtual (RRT*: rrt_star_connect.cpp:12-75)
	} catch (const PinBudget &) {
		out->budget_hit = 1;
	}
	const bool found = p->star ? star.goal_found : plain.goal_found;
	out->solved = found ? 1 : 0;
	out->cells_used = (int) (g.cell + 1);
	out->nv_a = Ta.getNumVertices();
	out->nv_b = Tb.getNumVertices();
	out->path_length = p->star ? star.path_length_ : plain.path_length_;
	out->path_yaw = p->star ? star.path_yaw_ : plain.path_yaw_;
	out->path_cost = p->star ? star.path_cost_ : plain.path_cost_;
	dump_tree(Ta, cap, a_states, a_actions, a_parent, a_g, a_y);
	dump_tree(Tb, cap, b_states, b_actions, b_parent, b_g, b_y);
	if (found) {  // rrt_connect.cpp:381-401
		std::vector<int> path_a = P->pathFromStart(Ta, Ta.getNumVertices() - 1);
		std::vector<int> path_b = P->pathFromStart(Tb, Tb.getNumVertices() - 1);
		std::reverse(path_b.begin(), path_b.end());
		std::vector<Action> action_sequence_b = P->getActionSequenceReverse(Tb, path_b);
		path_b.erase(path_b.begin());
		std::vector<State> ss = P->getStateSequence(Ta, path_a);
		std::vector<State> sb = P->getStateSequence(Tb, path_b);
		ss.insert(ss.end(), sb.begin(), sb.end());
		std::vector<Action> aa = P->getActionSequence(Ta, path_a);
		aa.insert(aa.end(), action_sequence_b.begin(), action_sequence_b.end());
		out->path_states = (int) ss.size();
		for (int i = 0; i < (int) ss.size() && i < path_cap; ++i) from_state(ss[i], path_states + 8 * i);
		for (int i = 0; i < (int) aa.size() && i < path_cap; ++i) from_action(aa[i], path_actions + 10 * i);
	}
	out->oog_height = g.oog_height; out->oog_nan = g.oog_nan; out->near_sets = g.near_sets; out->near_reordered = g.near_reordered;
	std::cout.rdbuf(old);
	return out->solved;
}

// loadData as the other harness does (fast_terrain_map.cpp:10-28)
void *pin_terrain_create(int nx, int ny, const double *x, const double *y, const double *z, const double *dx, const double *dy, const double *dz) {
	std::vector<double> xv(x, x + nx), yv(y, y + ny);
	std::vector<std::vector<double> > zz(nx), dxx(nx), dyy(nx), dzz(nx);
	for (int i = 0; i < nx; ++i) {
		zz[i].assign(z + (size_t) i * ny, z + (size_t) (i + 1) * ny);
		dxx[i].assign(dx + (size_t) i * ny, dx + (size_t) (i + 1) * ny);
		dyy[i].assign(dy + (size_t) i * ny, dy + (size_t) (i + 1) * ny);
		dzz[i].assign(dz + (size_t) i * ny, dz + (size_t) (i + 1) * ny);
	}
	FastTerrainMap *t = new FastTerrainMap();
	t->loadData(nx, ny, xv, yv, zz, dxx, dyy, dzz);
	return t;
}
void pin_terrain_destroy(void *h) { delete (FastTerrainMap *) h; }

}  // extern "C"
