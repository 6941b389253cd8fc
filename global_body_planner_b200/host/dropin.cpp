// Host side of the drop-in C++ API (include/global_body_planner/*.h): the reference's class and
// function names over the C ABI of libgbp_b200.so.  This file holds marshalling and the planners'
// control flow only — every distance, propagation, validity verdict, sample and neighbour query is
// computed on the GPU through include/gbp_b200.h.  Reference citations: file:line under the
// reference root.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <stdexcept>
#include <string>

#include "../../include/global_body_planner/rrt_star_connect.h"

namespace gbp_dropin {
void raise(const char *what) { throw std::runtime_error(std::string(what) + ": " + gbp_last_error()); }
}  // namespace gbp_dropin
using gbp_dropin::check;

// =============================================================================== FastTerrainMap
FastTerrainMap::FastTerrainMap() {}

void FastTerrainMap::loadData(int x_size, int y_size, std::vector<double> x_data, std::vector<double> y_data,
							  std::vector<std::vector<double>> z_data, std::vector<std::vector<double>> dx_data,
							  std::vector<std::vector<double>> dy_data, std::vector<std::vector<double>> dz_data) {
	auto flat = [&](const std::vector<std::vector<double>> &L) {
		std::vector<double> out((size_t) x_size * y_size);
		for (int i = 0; i < x_size; ++i) std::copy(L[i].begin(), L[i].begin() + y_size, out.begin() + (size_t) i * y_size);
		return out;
	};
	std::vector<double> z = flat(z_data), a = flat(dx_data), b = flat(dy_data), c = flat(dz_data);
	gbp_terrain *t = nullptr;
	check(gbp_terrain_create(x_size, y_size, x_data.data(), y_data.data(), z.data(), a.data(), b.data(), c.data(), &t),
		  "FastTerrainMap::loadData");
	dev_.reset(t, gbp_terrain_destroy);
	x_data_.assign(x_data.begin(), x_data.begin() + x_size);
	y_data_.assign(y_data.begin(), y_data.begin() + y_size);
}
void FastTerrainMap::loadDataFromCSV(const std::string &directory, bool via_gridmap) {
	gbp_terrain *t = nullptr;
	check(gbp_terrain_create_csv(directory.c_str(), via_gridmap ? 1 : 0, &t), "FastTerrainMap::loadDataFromCSV");
	adopt(t, "FastTerrainMap::loadDataFromCSV");
}
void FastTerrainMap::adopt(gbp_terrain *t, const char *what) {
	dev_.reset(t, gbp_terrain_destroy);
	int nx = 0, ny = 0;
	check(gbp_terrain_dims(t, &nx, &ny, nullptr), what);
	x_data_.resize(nx);
	y_data_.resize(ny);
	check(gbp_terrain_axes(t, x_data_.data(), y_data_.data()), what);
}
void FastTerrainMap::createOwnMap(uint64_t seed) {
	gbp_terrain *t = nullptr;
	check(gbp_terrain_create_own_map(seed, 221, 161, -0.5, -4.0, 0.05, 0, nullptr, &t), "FastTerrainMap::createOwnMap");  // :36-38
	adopt(t, "FastTerrainMap::createOwnMap");
}
void FastTerrainMap::createMap() {
	gbp_terrain *t = nullptr;
	check(gbp_terrain_create_default_map(&t), "FastTerrainMap::createMap");
	adopt(t, "FastTerrainMap::createMap");
}
const gbp_terrain *FastTerrainMap::handle() const {
	if (!dev_) throw std::runtime_error("FastTerrainMap: no terrain loaded");
	return dev_.get();
}
double FastTerrainMap::getGroundHeight(const double x, const double y) {
	double h = 0;
	check(gbp_ground_height(handle(), 1, &x, &y, &h, nullptr), "getGroundHeight");
	return h;
}
std::vector<double> FastTerrainMap::getGroundHeight(const std::vector<double> &x, const std::vector<double> &y) {
	std::vector<double> h(x.size());
	check(gbp_ground_height(handle(), (int64_t) x.size(), x.data(), y.data(), h.data(), nullptr), "getGroundHeight");
	return h;
}
bool FastTerrainMap::heightIsNan(const double x, const double y) {
	uint8_t r = 0;
	check(gbp_height_is_nan(handle(), 1, &x, &y, &r), "heightIsNan");
	return r != 0;
}
std::array<double, 3> FastTerrainMap::getSurfaceNormal(const double x, const double y) {
	std::array<double, 3> n;
	check(gbp_surface_normal(handle(), 1, &x, &y, n.data()), "getSurfaceNormal");
	return n;
}
std::vector<double> FastTerrainMap::getXData() { return x_data_; }
std::vector<double> FastTerrainMap::getYData() { return y_data_; }

// =============================================================================== planning_utils
namespace planning_utils {

namespace {
std::atomic<std::uint64_t> g_cell(0);
std::uint64_t g_seed = 1, g_stream = 0;
State propagate1(int kind, const State &s, const Action *a, double t) {
	State o;
	check(gbp_propagate(kind, 1, s.data(), a ? a->data() : nullptr, &t, o.data()), "propagate");
	return o;
}
double dist1(int kind, const State &a, const State &b) {
	double d = 0;
	check(gbp_distance(kind, 1, a.data(), b.data(), &d), "distance");
	return d;
}
bool pair1(const State &s, const Action &a, FastTerrainMap &terrain, State &s_new, double &t_new, int direction, bool adaptive) {
	uint8_t v = 0, d = (uint8_t) direction;
	check(gbp_validate_pairs(terrain.handle(), 1, s.data(), a.data(), &d, adaptive ? 1 : 0, 1, &v, nullptr, s_new.data(), &t_new),
		  "isValidStateActionPair");
	return v != 0;
}
}  // namespace

void set_random_stream(std::uint64_t seed, std::uint64_t stream) { g_seed = seed; g_stream = stream; g_cell = 0; }
std::uint64_t next_random_cell() { return g_cell.fetch_add(1); }
std::uint64_t random_seed() { return g_seed; }
std::uint64_t random_stream() { return g_stream; }

double poseDistance(const State &q1, const State &q2) { return dist1(0, q1, q2); }
double stateDistance(const State &q1, const State &q2) { return dist1(1, q1, q2); }
double stateYawDistance(const State &q1, const State &q2) { return dist1(2, q1, q2); }
double stateDistance(const State &q1, const State &q2, bool flag, double lw, double yw) {
	return flag ? poseDistance(q1, q2) * lw + stateYawDistance(q1, q2) * yw : stateDistance(q1, q2);  // header :146-155
}
bool isWithinBounds(State s1, State s2) { return stateDistance(s1, s2) <= GOAL_BOUNDS; }

State applyStance(State s, Action a, double t) { return propagate1(0, s, &a, t); }
State applyStance(State s, Action a) { return propagate1(0, s, &a, a[6]); }
State applyFlight(State s, double t_f) { return propagate1(1, s, nullptr, t_f); }
State applyAction(State s, Action a) { return applyFlight(applyStance(s, a), a[7]); }
State applyStanceReverse(State s, Action a, double t) { return propagate1(2, s, &a, t); }
State applyStanceReverse(State s, Action a) { return propagate1(2, s, &a, 0.0); }

Action getRandomAction(std::array<double, 3> n) {
	Action a;
	check(gbp_sample_actions(g_seed, g_stream, next_random_cell(), 1, n.data(), nullptr, nullptr, 0.0, a.data()), "getRandomAction");
	return a;
}
Action getRandomActionDirection(std::array<double, 3> n, State s_from, State s_to) {
	Action a;  // threshold 2 > any uniform: always directional (planning_utils.cpp:443-515)
	check(gbp_sample_actions(g_seed, g_stream, next_random_cell(), 1, n.data(), s_from.data(), s_to.data(), 2.0, a.data()),
		  "getRandomActionDirection");
	return a;
}
Action getRandomAction(std::array<double, 3> n, int direction, bool flag, double threshold, State s, State s_near) {
	if (!flag) return getRandomAction(n);
	const State &from = direction == FORWARD ? s_near : s, &to = direction == FORWARD ? s : s_near;  // :384-387
	Action a;
	check(gbp_sample_actions(g_seed, g_stream, next_random_cell(), 1, n.data(), from.data(), to.data(), threshold, a.data()),
		  "getRandomAction");
	return a;
}

bool isValidAction(Action a) {
	uint8_t v = 0;
	check(gbp_valid_actions(1, a.data(), &v), "isValidAction");
	return v != 0;
}
bool isValidState(State s, FastTerrainMap &terrain, int phase) {
	uint8_t v = 0, p = (uint8_t) phase;
	check(gbp_valid_states(terrain.handle(), 1, s.data(), &p, &v, nullptr), "isValidState");
	return v != 0;
}
std::vector<unsigned char> isValidState(const std::vector<State> &s, FastTerrainMap &terrain, int phase) {
	std::vector<unsigned char> v(s.size()), p(s.size(), (unsigned char) phase);
	check(gbp_valid_states(terrain.handle(), (int64_t) s.size(), s.empty() ? nullptr : s[0].data(), p.data(), v.data(), nullptr), "isValidState");
	return v;
}
bool isValidStateActionPair(State s, Action a, FastTerrainMap &terrain, State &s_new, double &t_new, bool adaptive) {
	return pair1(s, a, terrain, s_new, t_new, FORWARD, adaptive);
}
bool isValidStateActionPairAdaptiveStepSize(State s, Action a, FastTerrainMap &terrain, State &s_new, double &t_new) {
	return pair1(s, a, terrain, s_new, t_new, FORWARD, true);
}
bool isValidStateActionPair(State s, Action a, FastTerrainMap &terrain, State &s_new, double &t_new) {
	return pair1(s, a, terrain, s_new, t_new, FORWARD, false);
}
bool isValidStateActionPair(State s, Action a, FastTerrainMap &terrain) {
	State d; double t;
	return pair1(s, a, terrain, d, t, FORWARD, false);
}
bool isValidStateActionPairReverse(State s, Action a, FastTerrainMap &terrain, State &s_new, double &t_new, bool adaptive) {
	return pair1(s, a, terrain, s_new, t_new, REVERSE, adaptive);
}
bool isValidStateActionPairReverseAdaptiveStepSize(State s, Action a, FastTerrainMap &terrain, State &s_new, double &t_new) {
	return pair1(s, a, terrain, s_new, t_new, REVERSE, true);
}
bool isValidStateActionPairReverse(State s, Action a, FastTerrainMap &terrain, State &s_new, double &t_new) {
	return pair1(s, a, terrain, s_new, t_new, REVERSE, false);
}
bool isValidStateActionPairReverse(State s, Action a, FastTerrainMap &terrain) {
	State d; double t;
	return pair1(s, a, terrain, d, t, REVERSE, false);
}
std::vector<unsigned char> isValidStateActionPair(const std::vector<State> &s, const std::vector<Action> &a,
												  const std::vector<unsigned char> &direction, FastTerrainMap &terrain,
												  std::vector<State> &s_new, std::vector<double> &t_new, bool adaptive) {
	const size_t n = s.size();
	std::vector<unsigned char> v(n);
	s_new.resize(n);
	t_new.resize(n);
	if (n)
		check(gbp_validate_pairs(terrain.handle(), (int64_t) n, s[0].data(), a[0].data(), direction.data(), adaptive ? 1 : 0, 0, v.data(),
								 nullptr, s_new[0].data(), t_new.data()), "isValidStateActionPair (batched)");
	return v;
}

// Output interpolation (planning_utils.cpp:142-193): the sample grid is laid out exactly as the reference's loops do
// (t += dt) and every state on it is evaluated in one launch (gbp_interp_path).
namespace {
void interp_append(const State *states, const Action *actions, int n_actions, bool closing_state, double t0, double dt,
				   std::vector<State> &interp_path, std::vector<double> &interp_t, std::vector<int> &interp_phase) {
	std::int64_t count = 0;
	std::vector<State> out(1);
	std::vector<double> tt(1);
	std::vector<int> ph(1);
	check(gbp_interp_path(n_actions, states[0].data(), n_actions ? actions[0].data() : nullptr, dt, 0, nullptr, nullptr, nullptr, &count), "getInterpPath");
	out.resize((size_t) count); tt.resize((size_t) count); ph.resize((size_t) count);
	check(gbp_interp_path(n_actions, states[0].data(), n_actions ? actions[0].data() : nullptr, dt, count, out[0].data(), tt.data(), ph.data(), &count),
		  "getInterpPath");
	const size_t keep = closing_state ? (size_t) count : (size_t) count - 1;  // a single pair has no closing sample
	for (size_t i = 0; i < keep; ++i) { interp_path.push_back(out[i]); interp_t.push_back(tt[i] + t0); }
	for (size_t i = 0; i + 1 < (size_t) count; ++i) interp_phase.push_back(ph[i]);
}
}  // namespace
void interpStateActionPair(State s, Action a, double t0, double dt, std::vector<State> &interp_path, std::vector<double> &interp_t,
						   std::vector<int> &interp_phase) {
	State two[2] = {s, s};
	// t0 is added on the host AFTER the device call, which would round differently from the reference's `t + t0`;
	// with a non-zero t0 the times are therefore rebuilt with the reference's own expressions below
	const size_t first = interp_t.size();
	interp_append(two, &a, 1, false, 0.0, dt, interp_path, interp_t, interp_phase);
	const double t_s = a[6], t_f = a[7];
	size_t k = first;
	for (double t = 0; t < t_s; t += dt) interp_t[k++] = t + t0;
	for (double t = 0; t < t_f; t += dt) interp_t[k++] = t_s + t + t0;
	if (t_f > 0) interp_t[k++] = t0 + t_s + t_f;
}
void getInterpPath(std::vector<State> state_sequence, std::vector<Action> action_sequence, double dt, std::vector<State> &interp_path,
				   std::vector<double> &interp_t, std::vector<int> &interp_phase) {
	if (state_sequence.empty()) return;
	if (state_sequence.size() != action_sequence.size() + 1) throw std::runtime_error("getInterpPath: need one more state than actions");
	interp_append(state_sequence.data(), action_sequence.data(), (int) action_sequence.size(), true, 0.0, dt, interp_path, interp_t, interp_phase);
}

double calculateMaxCurvature(std::vector<State> &body_plan) {  // :900-909
	double c = 0;
	check(gbp_max_curvature((std::int64_t) body_plan.size(), body_plan.empty() ? nullptr : body_plan[0].data(), &c), "calculateMaxCurvature");
	return c;
}

// ---- printing: same text as the reference emits (:16-104)
void printState(State vec) {
	std::cout << "{";
	for (size_t i = 0; i < vec.size(); i++) std::cout << vec[i] << ", ";
	std::cout << "\b\b}";
}
void printStateNewline(State vec) { printState(vec); std::cout << std::endl; }
void printAction(Action a) {
	std::cout << "{";
	for (size_t i = 0; i < a.size(); i++) std::cout << a[i] << ", ";
	std::cout << "\b\b}";
}
void printActionNewline(Action a) { printAction(a); std::cout << std::endl; }
void printStateSequence(std::vector<State> state_sequence) { for (const State &s : state_sequence) printStateNewline(s); }
void printActionSequence(std::vector<Action> action_sequence) { for (const Action &a : action_sequence) printActionNewline(a); }
void vectorToArray(State vec, double *new_array) {  // :5-8
	for (size_t i = 0; i < vec.size(); i++) new_array[i] = vec.at(i);
}
void printVectorInt(std::vector<int> vec) {
	std::cout << "{";
	for (size_t i = 0; i < vec.size(); i++) std::cout << vec[i] << ", ";
	std::cout << "\b\b}";
}
void printVectorIntNewline(std::vector<int> vec) { printVectorInt(vec); std::cout << std::endl; }
void printInterpStateSequence(std::vector<State> state_sequence, std::vector<double> interp_t) {
	for (size_t i = 0; i < state_sequence.size(); i++) { std::cout << interp_t[i] << "\t"; printStateNewline(state_sequence[i]); }
}
State interp(State q1, State q2, double x) {  // :97-103
	State q_out;
	for (size_t dim = 0; dim < q1.size(); dim++) q_out[dim] = (q2[dim] - q1[dim]) * x + q1[dim];
	return q_out;
}
std::array<double, 3> rotate_grf(std::array<double, 3> surface_norm, std::array<double, 3> grf) {  // :198-231
	std::array<double, 3> out;
	check(gbp_rotate_grf(1, surface_norm.data(), grf.data(), out.data()), "rotate_grf");
	return out;
}
double calculateCurvature(double x1, double y1, double x2, double y2, double x3, double y3) {  // :884-899
	const double p[6] = {x1, y1, x2, y2, x3, y3};
	double c = 0;
	check(gbp_curvature(1, p, &c), "calculateCurvature");
	return c;
}
void printStateXYZPYaw(const State &s) {
	std::cout << "x: " << std::setw(7) << std::setprecision(3) << s[0] << " | y: " << std::setw(7) << std::setprecision(3) << s[1]
			  << " | z: " << std::setw(7) << std::setprecision(3) << s[2] << " | p: " << std::setw(7) << std::setprecision(4) << s[6]
			  << " | yaw: " << std::setw(6) << std::setprecision(3) << std::atan2(s[4], s[3]) << " |" << std::endl;
}
void printStateSequenceXYZPYaw(const std::vector<State> &state_sequence) {
	std::cout << "---------- discrete state sequence ----------" << std::endl;
	if (state_sequence.empty()) return;
	double path_length = 0, path_yaw = 0;
	for (size_t i = 0; i + 1 < state_sequence.size(); ++i) {
		printStateXYZPYaw(state_sequence[i]);
		const double dl = poseDistance(state_sequence[i], state_sequence[i + 1]), dy = stateYawDistance(state_sequence[i], state_sequence[i + 1]);
		std::cout << "           |            |            |            |            "
				  << " | length: " << std::setw(6) << std::setprecision(3) << dl << " | yaw: " << std::setw(6) << std::setprecision(3) << dy << std::endl;
		path_length += dl;
		path_yaw += dy;
	}
	printStateXYZPYaw(state_sequence.back());
	std::cout << "path length: " << std::setw(6) << std::setprecision(3) << path_length << std::endl;
	std::cout << "path yaw:    " << std::setw(6) << std::setprecision(3) << path_yaw << std::endl;
}

}  // namespace planning_utils

// =============================================================================== GraphClass
struct GraphClass::Mirror {
	gbp_tree *tree = nullptr;
	int cap = 0;
	size_t synced = 0;
	bool dirty = false;
	~Mirror() { if (tree) gbp_tree_destroy(tree); }
};

GraphClass::GraphClass() : mirror_(new Mirror()) {}
GraphClass::~GraphClass() {}
GraphClass::GraphClass(const GraphClass &o)
	: nodes_(o.nodes_), order_(o.order_), cost_add_yaw_flag_(o.cost_add_yaw_flag_),
	  cost_add_yaw_length_weight_(o.cost_add_yaw_length_weight_), cost_add_yaw_yaw_weight_(o.cost_add_yaw_yaw_weight_), mirror_(new Mirror()) {}
GraphClass &GraphClass::operator=(const GraphClass &o) {
	if (this != &o) {
		nodes_ = o.nodes_; order_ = o.order_;
		cost_add_yaw_flag_ = o.cost_add_yaw_flag_;
		cost_add_yaw_length_weight_ = o.cost_add_yaw_length_weight_;
		cost_add_yaw_yaw_weight_ = o.cost_add_yaw_yaw_weight_;
		mirror_.reset(new Mirror());  // the device mirror is rebuilt on the next neighbour query
	}
	return *this;
}
GraphClass::Node &GraphClass::node(int idx) {
	auto it = nodes_.find(idx);
	if (it == nodes_.end()) {  // the reference's operator[] inserts on reads of unknown ids
		it = nodes_.emplace(idx, Node()).first;
		it->second.state.fill(0.0);
		it->second.action.fill(0.0);
		order_.push_back(idx);
	}
	return it->second;
}
gbp_tree *GraphClass::device_store() {
	Mirror &m = *mirror_;
	const int n = (int) order_.size();
	if (n == 0) throw std::runtime_error("GraphClass: empty graph");
	if (n > m.cap) {
		if (m.tree) gbp_tree_destroy(m.tree);
		m.tree = nullptr;
		m.cap = std::max(1024, 2 * n);
		check(gbp_tree_create(m.cap, &m.tree), "gbp_tree_create");
		m.synced = 0;
	}
	if (m.dirty || m.synced == 0 || n - (int) m.synced > 8) {  // bulk (re)load in slot order
		std::vector<double> s((size_t) 8 * n);
		std::vector<int> parent(n);
		for (int i = 0; i < n; ++i) {
			std::memcpy(&s[8 * (size_t) i], nodes_[order_[i]].state.data(), 64);
			parent[i] = i - 1;
		}
		check(gbp_tree_load(m.tree, n, s.data(), nullptr, parent.data()), "gbp_tree_load");
	} else {
		Action zero; zero.fill(0.0);
		for (int i = (int) m.synced; i < n; ++i) check(gbp_tree_append(m.tree, 0, nodes_[order_[i]].state.data(), zero.data(), nullptr), "gbp_tree_append");
	}
	m.synced = n;
	m.dirty = false;
	return m.tree;
}
State GraphClass::getVertex(int idx) { return node(idx).state; }
int GraphClass::getNumVertices() { return (int) nodes_.size(); }
void GraphClass::addVertex(int idx, State q) {
	const bool known = nodes_.count(idx) != 0;
	node(idx).state = q;
	if (known) mirror_->dirty = true;
}
void GraphClass::addEdge(int idx1, int idx2) {  // graph_class.cpp:36-42
	Node &p = node(idx1), &c = node(idx2);
	c.parents.push_back(idx1);
	p.children.push_back(idx2);
	c.g = p.g + poseDistance(p.state, c.state);
	c.y = p.y + stateYawDistance(p.state, c.state);
}
void GraphClass::removeEdge(int idx1, int idx2) {
	auto drop = [](std::vector<int> &v, int x) { auto it = std::find(v.begin(), v.end(), x); if (it != v.end()) v.erase(it); };
	drop(node(idx2).parents, idx1);
	drop(node(idx1).children, idx2);
}
int GraphClass::getPredecessor(int idx) {
	Node &n = node(idx);
	if (n.parents.size() > 1) { std::cout << "More than one predecessor, fix this!" << std::endl; throw("Error"); }  // :63-66
	if (n.parents.empty()) throw std::out_of_range("GraphClass::getPredecessor: vertex has no parent");
	return n.parents.front();
}
std::vector<int> GraphClass::getSuccessors(int idx) { return node(idx).children; }
void GraphClass::addAction(int idx, Action a) { Node &n = node(idx); n.action = a; n.has_action = true; }
Action GraphClass::getAction(int idx) { return node(idx).action; }
double GraphClass::getGValue(int idx) { return node(idx).g; }
double GraphClass::getYValue(int idx) { return node(idx).y; }
void GraphClass::updateGYValue(int idx, double g_val, double y_val) {  // :131-138, iterative instead of recursive
	std::vector<int> stack(1, idx);
	node(idx).g = g_val;
	node(idx).y = y_val;
	while (!stack.empty()) {
		const int i = stack.back();
		stack.pop_back();
		Node &p = node(i);
		for (int c : p.children) {
			Node &ch = node(c);
			ch.g = p.g + poseDistance(p.state, ch.state);
			ch.y = p.y + stateYawDistance(p.state, ch.state);
			stack.push_back(c);
		}
	}
}
void GraphClass::printVertex(State v) {
	std::cout << "{";
	for (size_t i = 0; i < v.size(); ++i) std::cout << v[i] << (i + 1 < v.size() ? ", " : "}");
}
void GraphClass::printVertices() {
	std::cout << "\nAll Vertices : \n";
	for (int id : order_) { std::cout << id << " "; printVertex(nodes_[id].state); std::cout << std::endl; }
}
void GraphClass::printIncomingEdges(int in_vertex) {
	for (int out_vertex : node(in_vertex).parents) std::cout << "{" << out_vertex << " -> " << in_vertex << "}" << std::endl;
}
void GraphClass::printEdges() {
	std::cout << "All Edges : \n";
	for (int id : order_) printIncomingEdges(id);
}
void GraphClass::init(State s, bool flag, double lw, double yw) {
	nodes_.clear();
	order_.clear();
	mirror_.reset(new Mirror());
	addVertex(0, s);
	node(0).g = 0;
	node(0).y = 0;
	cost_add_yaw_flag_ = flag;
	cost_add_yaw_length_weight_ = lw;
	cost_add_yaw_yaw_weight_ = yw;
}

// =============================================================================== PlannerClass
PlannerClass::PlannerClass() {}
PlannerClass::~PlannerClass() {}

State PlannerClass::randomState(FastTerrainMap &terrain) {
	State q;
	check(gbp_sample_states(terrain.handle(), planning_utils::random_seed(), planning_utils::random_stream(), planning_utils::next_random_cell(), 1, nullptr, nullptr,
							0.0, 0, q.data()), "randomState");
	return q;
}
State PlannerClass::randomStateDirection(FastTerrainMap &terrain, State s_from, State s_to, bool speed_direction_flag) {
	State q;
	check(gbp_sample_states(terrain.handle(), planning_utils::random_seed(), planning_utils::random_stream(), planning_utils::next_random_cell(), 1, s_from.data(),
							s_to.data(), 2.0, speed_direction_flag ? 1 : 0, q.data()), "randomStateDirection");
	return q;
}
State PlannerClass::randomState(FastTerrainMap &terrain, bool flag, double threshold, bool speed_direction_flag, State s_from, State s_to) {
	if (!flag) return randomState(terrain);
	State q;
	check(gbp_sample_states(terrain.handle(), planning_utils::random_seed(), planning_utils::random_stream(), planning_utils::next_random_cell(), 1, s_from.data(),
							s_to.data(), threshold, speed_direction_flag ? 1 : 0, q.data()), "randomState");
	return q;
}
int PlannerClass::getNearestNeighbor(State q) {  // planner_class.cpp:185-200 (ties: lowest insertion slot)
	int slot = 0;
	check(gbp_nearest(device_store(), 1, q.data(), &slot, nullptr), "getNearestNeighbor");
	return order_[slot];
}
std::vector<int> PlannerClass::neighborhoodDist(State q, double dist) {  // :173-182, insertion order
	gbp_tree *t = device_store();
	std::vector<int> slots(order_.size());
	int count = 0;
	check(gbp_near(t, q.data(), dist, slots.data(), (int) slots.size(), &count), "neighborhoodDist");
	std::vector<int> ids(count);
	for (int i = 0; i < count; ++i) ids[i] = order_[slots[i]];
	return ids;
}
std::vector<int> PlannerClass::neighborhoodN(State s, int N) {  // :151-171 (unused by the planners)
	const size_t n = order_.size();
	std::vector<State> a(n, s), b(n);
	for (size_t i = 0; i < n; ++i) b[i] = nodes_[order_[i]].state;
	std::vector<double> d(n), y(n);
	if (n) check(gbp_distance(cost_add_yaw_flag_ ? 0 : 1, (int64_t) n, a[0].data(), b[0].data(), d.data()), "neighborhoodN");
	if (n && cost_add_yaw_flag_) {
		check(gbp_distance(2, (int64_t) n, a[0].data(), b[0].data(), y.data()), "neighborhoodN");
		for (size_t i = 0; i < n; ++i) d[i] = d[i] * cost_add_yaw_length_weight_ + y[i] * cost_add_yaw_yaw_weight_;
	}
	std::vector<int> idx(n);
	for (size_t i = 0; i < n; ++i) idx[i] = (int) i;
	std::stable_sort(idx.begin(), idx.end(), [&](int i, int j) { return d[i] < d[j]; });
	std::vector<int> out;
	for (int i = 0; i < N && i < (int) n; ++i) out.push_back(order_[idx[i]]);
	return out;
}

