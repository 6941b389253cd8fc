// Planner classes of the drop-in C++ API (RRTClass, RRTConnectClass, RRTStarConnectClass): the
// reference's control flow over the C ABI.  buildRRTConnect is the B200-native entry: whole searches
// run resident on the device (gbp_plan_batch), many independent attempts per round.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <iostream>
#include <numeric>

#include "../../include/global_body_planner/rrt_star_connect.h"

using gbp_dropin::check;
typedef std::chrono::high_resolution_clock Clock;
static double seconds_since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

// =============================================================================== RRTClass
RRTClass::RRTClass() : elapsed_total(0), elapsed_to_first(0) {}
RRTClass::~RRTClass() {}

void RRTClass::set_candidates_per_extend(int k, bool best_of_k) { k_candidates_ = std::max(1, k); best_of_k_ = best_of_k; }
void RRTClass::set_random_stream(std::uint64_t seed, std::uint64_t stream) { seed_ = seed; stream_ = stream; cell_ = 0; }

// rrt.cpp:20-70: the K candidates are sampled (with the directional option when set: planning_utils.cpp:379-391),
// validated and selected in one device pass
bool RRTClass::newConfig(State s, State s_near, State &s_new, Action &a_new, FastTerrainMap &terrain, int direction) {
	int found = 0;
	const std::uint64_t idx0 = (cell_++) * (std::uint64_t) k_candidates_;
	const double dir_thresh = action_direction_sampling_flag_ ? action_direction_sampling_probability_threshold_ : -1.0;
	check(gbp_new_config(terrain.handle(), s.data(), s_near.data(), direction, k_candidates_, best_of_k_ ? 1 : 0,
						 state_action_pair_check_adaptive_step_size_flag_ ? 1 : 0, dir_thresh, seed_, stream_, idx0, &found, s_new.data(), a_new.data(),
						 nullptr),
		  "newConfig");
	return found != 0;
}

int RRTClass::extend(PlannerClass &T, State s, FastTerrainMap &terrain, int direction) {  // rrt.cpp:77-102
	const int near = T.getNearestNeighbor(s);
	const State s_near = T.getVertex(near);
	State s_new;
	Action a_new;
	if (!newConfig(s, s_near, s_new, a_new, terrain, direction)) return TRAPPED;
	const int id = T.getNumVertices();
	T.addVertex(id, s_new);
	T.addEdge(near, id);
	T.addAction(id, a_new);
	T.updateGYValue(id, T.getGValue(near) + poseDistance(s_near, s_new), T.getYValue(near) + stateYawDistance(s_near, s_new));
	return isWithinBounds(s_new, s) ? REACHED : ADVANCED;
}

std::vector<int> RRTClass::pathFromStart(PlannerClass &T, int s) {
	std::vector<int> path(1, s);
	while (s != 0) { s = T.getPredecessor(s); path.push_back(s); }
	std::reverse(path.begin(), path.end());
	return path;
}
std::vector<State> RRTClass::getStateSequence(PlannerClass &T, std::vector<int> path) {
	std::vector<State> out;
	for (int i : path) out.push_back(T.getVertex(i));
	return out;
}
std::vector<Action> RRTClass::getActionSequence(PlannerClass &T, std::vector<int> path) {
	std::vector<Action> out;  // the action stored at a vertex leads INTO it (rrt.cpp:127-135)
	for (size_t i = 1; i < path.size(); ++i) out.push_back(T.getAction(path[i]));
	return out;
}
void RRTClass::printPath(PlannerClass &T, std::vector<int> path) {
	std::cout << "Printing path:";
	for (int i : path) { std::cout << std::endl << i << " or "; T.printVertex(T.getVertex(i)); std::cout << " ->"; }
	std::cout << std::endl;
}
void RRTClass::getStatistics(double &plan_time, int &success_var, int &vertices_generated, double &time_to_first_solve,
							 std::vector<double> &length_vector, std::vector<double> &yaw_vector, std::vector<double> &cost_vector,
							 std::vector<double> &cost_vector_times, double &path_duration, std::vector<std::vector<double>> &allStatePosition) {
	plan_time = elapsed_total.count();
	success_var = success_;
	vertices_generated = num_vertices;
	time_to_first_solve = elapsed_to_first.count();
	length_vector = length_vector_;
	yaw_vector = yaw_vector_;
	cost_vector = cost_vector_;
	cost_vector_times = cost_vector_times_;
	path_duration = path_duration_;
	allStatePosition = allStatePosition_;
}
void RRTClass::saveStateSequence(PlannerClass &T) {
	for (int i = 0; i < T.getNumVertices(); ++i) {
		State s = T.getVertex(i);
		allStatePosition_.push_back({s[0], s[1], s[2]});
	}
}
void RRTClass::set_action_direction_sampling(bool flag, double threshold) {
	action_direction_sampling_flag_ = flag;
	action_direction_sampling_probability_threshold_ = threshold;
}
void RRTClass::set_state_direction_sampling(bool flag, double threshold, bool speed_direction_flag) {
	state_direction_sampling_flag_ = flag;
	state_direction_sampling_probability_threshold_ = threshold;
	state_direction_sampling_speed_direction_flag_ = speed_direction_flag;
}
void RRTClass::set_state_action_pair_check_adaptive_step_size_flag_(bool f) { state_action_pair_check_adaptive_step_size_flag_ = f; }
void RRTClass::set_cost_add_yaw(bool flag, double length_weight, double yaw_weight) {
	cost_add_yaw_flag_ = flag;
	cost_add_yaw_length_weight_ = length_weight;
	cost_add_yaw_yaw_weight_ = yaw_weight;
}
void RRTClass::print_setting_parameters() {
	std::cout << "adaptive step size: " << state_action_pair_check_adaptive_step_size_flag_ << "\ncost_add_yaw: " << cost_add_yaw_flag_ << " ("
			  << cost_add_yaw_length_weight_ << ", " << cost_add_yaw_yaw_weight_ << ")\nstate direction sampling: "
			  << state_direction_sampling_flag_ << " (" << state_direction_sampling_probability_threshold_ << ", "
			  << state_direction_sampling_speed_direction_flag_ << ")\naction direction sampling: " << action_direction_sampling_flag_ << " ("
			  << action_direction_sampling_probability_threshold_ << ")\ncandidates per extend: " << k_candidates_
			  << (best_of_k_ ? " (closest valid)" : " (first valid)") << std::endl;
}
// vanilla RRT (rrt.cpp:171-249): goal-biased single tree, 30 s cut-off
void RRTClass::buildRRT(FastTerrainMap &terrain, State s_start, State s_goal, std::vector<State> &state_sequence,
						std::vector<Action> &action_sequence) {
	const auto t0 = Clock::now();
	success_ = 0;
	length_vector_.clear(); yaw_vector_.clear(); cost_vector_.clear(); cost_vector_times_.clear();
	PlannerClass T;
	T.init(s_start, cost_add_yaw_flag_, cost_add_yaw_length_weight_, cost_add_yaw_yaw_weight_);
	goal_found = false;
	double u[1];
	while (seconds_since(t0) < 30) {
		State probe = T.randomState(terrain);
		(void) u;
		// goal bias: reuse the pitch draw of a second state sample as the uniform (pitch = 2u - 1)
		State r2 = T.randomState(terrain);
		const double prob_goal = 0.5 * (r2[6] / P_MAX + 1.0);
		State s = prob_goal <= prob_goal_thresh ? s_goal : probe;
		if (!isValidState(s, terrain, STANCE)) continue;
		const int result = extend(T, s, terrain, FORWARD);
		if (result == REACHED && isWithinBounds(s, s_goal)) { goal_found = true; elapsed_to_first = Clock::now() - t0; break; }
	}
	num_vertices = T.getNumVertices();
	elapsed_total = Clock::now() - t0;
	if (!goal_found) { std::cout << "Path not found" << std::endl; return; }
	const int goal_idx = T.getNumVertices() - 1;
	std::vector<int> path = pathFromStart(T, goal_idx);
	state_sequence = getStateSequence(T, path);
	action_sequence = getActionSequence(T, path);
	if (elapsed_total.count() <= 5.0) success_ = 1;
	path_duration_ = 0.0;
	for (const Action &a : action_sequence) path_duration_ += a[6] + a[7];
	path_length_ = T.getGValue(goal_idx);
	path_yaw_ = T.getYValue(goal_idx);
	path_cost_ = path_length_;
	length_vector_.push_back(path_length_); yaw_vector_.push_back(path_yaw_); cost_vector_.push_back(path_cost_);
	cost_vector_times_.push_back(elapsed_total.count());
}

// =============================================================================== RRTConnectClass
RRTConnectClass::RRTConnectClass() {}
RRTConnectClass::~RRTConnectClass() {}
void RRTConnectClass::set_max_time_solve(double seconds) { max_time_solve_ = seconds; }
void RRTConnectClass::set_parallel_attempts(int attempts, int iterations, int vertices) {
	parallel_attempts_ = std::max(1, attempts);
	iterations_per_attempt_ = std::max(1, iterations);
	vertices_per_tree_ = std::max(2, vertices);
	attempts_set_ = true;
}

int RRTConnectClass::attemptConnect(State s_existing, State s, double t_s, State &s_new, Action &a_new, FastTerrainMap &terrain, int direction) {
	int status = TRAPPED;
	const uint8_t d = (uint8_t) direction;
	State sn = s_new;
	Action an = a_new;
	check(gbp_attempt_connect_ts(terrain.handle(), 1, s_existing.data(), s.data(), &t_s, &d,
								 state_action_pair_check_adaptive_step_size_flag_ ? 1 : 0, &status, sn.data(), an.data(), nullptr), "attemptConnect");
	s_new = sn;
	a_new = an;
	return status;
}
int RRTConnectClass::attemptConnect(State s_existing, State s, State &s_new, Action &a_new, FastTerrainMap &terrain, int direction) {
	int status = TRAPPED;
	const uint8_t d = (uint8_t) direction;
	check(gbp_attempt_connect(terrain.handle(), 1, s_existing.data(), s.data(), &d, state_action_pair_check_adaptive_step_size_flag_ ? 1 : 0,
							  &status, s_new.data(), a_new.data(), nullptr), "attemptConnect");
	return status;
}
int RRTConnectClass::connect(PlannerClass &T, State s, FastTerrainMap &terrain, int direction) {  // rrt_connect.cpp:98-120
	const int near = T.getNearestNeighbor(s);
	const State s_near = T.getVertex(near);
	State s_new;
	Action a_new;
	const int result = attemptConnect(s_near, s, s_new, a_new, terrain, direction);
	if (result != TRAPPED) {
		const int id = T.getNumVertices();
		T.addVertex(id, s_new);
		T.addEdge(near, id);
		T.addAction(id, a_new);
		T.updateGYValue(id, T.getGValue(near) + poseDistance(s_near, s_new), T.getYValue(near) + stateYawDistance(s_near, s_new));
	}
	return result;
}
std::vector<Action> RRTConnectClass::getActionSequenceReverse(PlannerClass &T, std::vector<int> path) {
	std::vector<Action> out;  // in the goal tree the action stored at a vertex is executed AT it (:125-133)
	for (size_t i = 0; i + 1 < path.size(); ++i) out.push_back(T.getAction(path[i]));
	return out;
}

// rrt_connect.cpp:139-227.  The reference probes attemptConnect(s, s_next) from the LAST state backwards and
// takes the first REACHED; here all later states are probed in one batched launch and the farthest REACHED
// one is taken — the same choice.
void RRTConnectClass::postProcessPath(std::vector<State> &state_sequence, std::vector<Action> &action_sequence, FastTerrainMap &terrain) {
	if (state_sequence.empty()) return;
	const int n = (int) state_sequence.size();
	std::vector<State> new_states(1, state_sequence.front());
	std::vector<Action> new_actions;
	path_length_ = 0; path_yaw_ = 0; path_cost_ = 0;
	int cur = 0;
	while (cur < n - 1) {
		const int m = n - 1 - cur;  // candidates: states cur+1 .. n-1
		std::vector<State> from(m, state_sequence[cur]), sn(m);
		std::vector<Action> an(m);
		std::vector<uint8_t> dir(m, (uint8_t) FORWARD);
		std::vector<int> status(m);
		check(gbp_attempt_connect(terrain.handle(), m, from[0].data(), state_sequence[cur + 1].data(), dir.data(),
								  state_action_pair_check_adaptive_step_size_flag_ ? 1 : 0, status.data(), sn[0].data(), an[0].data(), nullptr),
			  "postProcessPath");
		int pick = -1;
		for (int j = m - 1; j >= 0; --j) if (status[j] == REACHED) { pick = j; break; }
		const State &s = state_sequence[cur];
		if (pick >= 0) {
			const State &s_next = state_sequence[cur + 1 + pick];
			new_states.push_back(s_next);
			new_actions.push_back(an[pick]);
			const double dl = poseDistance(s, s_next), dy = stateYawDistance(s, s_next);
			path_length_ += dl;
			path_yaw_ += dy;
			path_cost_ += cost_add_yaw_flag_ ? dl * cost_add_yaw_length_weight_ + dy * cost_add_yaw_yaw_weight_ : dl;
			cur = cur + 1 + pick;
		} else {  // keep the original next state and action (:206-220; only path_cost_ is updated there)
			const State &s_next = state_sequence[cur + 1];
			new_states.push_back(s_next);
			new_actions.push_back(action_sequence[cur]);
			const double dl = poseDistance(s, s_next), dy = stateYawDistance(s, s_next);
			path_cost_ += cost_add_yaw_flag_ ? dl * cost_add_yaw_length_weight_ + dy * cost_add_yaw_yaw_weight_ : dl;
			cur = cur + 1;
		}
	}
	state_sequence = new_states;
	action_sequence = new_actions;
}

// rrt_connect.cpp:230-314, host-driven (one extend / connect per call): the compatibility path.
void RRTConnectClass::runRRTConnect(PlannerClass &Ta, PlannerClass &Tb, FastTerrainMap &terrain) {
	const auto t0 = Clock::now();
	while (true) {
		if (seconds_since(t0) >= anytime_horizon) { anytime_horizon *= horizon_expansion_factor; return; }
		for (int half = 0; half < 2; ++half) {
			PlannerClass &Tx = half == 0 ? Ta : Tb, &Ty = half == 0 ? Tb : Ta;
			State s_from = half == 0 ? Ta.getVertex(Ta.getNumVertices() - 1) : Ta.getVertex(0);
			State s_to = half == 0 ? Tb.getVertex(0) : Tb.getVertex(Tb.getNumVertices() - 1);
			State s_rand = Tx.randomState(terrain, state_direction_sampling_flag_, state_direction_sampling_probability_threshold_,
										  state_direction_sampling_speed_direction_flag_, s_from, s_to);
			if (!isValidState(s_rand, terrain, STANCE)) continue;
			if (extend(Tx, s_rand, terrain, half == 0 ? FORWARD : REVERSE) == TRAPPED) continue;
			State s_new = Tx.getVertex(Tx.getNumVertices() - 1);
			if (connect(Ty, s_new, terrain, half == 0 ? REVERSE : FORWARD) == REACHED) {
				goal_found = true;
				elapsed_to_first = Clock::now() - t0;
				path_length_ = Ta.getGValue(Ta.getNumVertices() - 1) + Tb.getGValue(Tb.getNumVertices() - 1);
				path_yaw_ = Ta.getYValue(Ta.getNumVertices() - 1) + Tb.getYValue(Tb.getNumVertices() - 1);
				path_cost_ = cost_add_yaw_flag_ ? path_length_ * cost_add_yaw_length_weight_ + path_yaw_ * cost_add_yaw_yaw_weight_ : path_length_;
				return;
			}
		}
	}
}

// rrt_connect.cpp:323-467 re-designed for the GPU: instead of one search restarted under a growing anytime
// horizon, every round runs `parallel_attempts_` independent searches on the device and the shortest
// post-processed path so far is kept.  Same termination rule and statistics as the reference.
void RRTConnectClass::buildRRTConnect(FastTerrainMap &terrain, State s_start, State s_goal, std::vector<State> &state_sequence,
									  std::vector<Action> &action_sequence, double max_time_opt) {
	buildAnytime(terrain, s_start, s_goal, state_sequence, action_sequence, max_time_opt, false);
}
// star = true: every attempt is an RRT*-Connect search (choose parent + near-set rewiring on the device,
// rrt_star_connect.cpp:12-75); otherwise RRT-Connect.
void RRTConnectClass::buildAnytime(FastTerrainMap &terrain, State s_start, State s_goal, std::vector<State> &state_sequence,
								   std::vector<Action> &action_sequence, double max_time_opt, bool star) {
	const auto t0 = Clock::now();
	success_ = 0;
	length_vector_.clear(); yaw_vector_.clear(); cost_vector_.clear(); cost_vector_times_.clear();
	goal_found = false;
	num_vertices = 0;
	anytime_horizon = poseDistance(s_start, s_goal) / planning_rate_estimate;
	double cost_so_far = INFTY;
	const int R = star && !attempts_set_ ? 3552 : parallel_attempts_, cap = 256;
	// anytime use of the batch planner: all attempts work on the same query, the round ends once 8 of them have solved
	// (the 8 shortest raw paths are shortcut below)
	// the fork's options travel with the batch: directional state / action sampling and the yaw-aware cost run inside the
	// device planner exactly as the setters configured them (rrt_connect.cpp:246-251, rrt.cpp:34, rrt_connect.cpp:270-274)
	gbp_plan_params P = {k_candidates_, best_of_k_ ? 1 : 0, star && !attempts_set_ ? 8000 : iterations_per_attempt_, vertices_per_tree_,
						 state_action_pair_check_adaptive_step_size_flag_ ? 1 : 0, star ? 1 : 0, 0, 8,
						 state_direction_sampling_flag_ ? 1 : 0, state_direction_sampling_speed_direction_flag_ ? 1 : 0,
						 action_direction_sampling_flag_ ? 1 : 0, cost_add_yaw_flag_ ? 1 : 0, state_direction_sampling_probability_threshold_,
						 action_direction_sampling_probability_threshold_, cost_add_yaw_length_weight_, cost_add_yaw_yaw_weight_};
	std::vector<State> starts(R, s_start), goals(R, s_goal);
	std::vector<gbp_plan_stats> stats(R);
	std::vector<double> ps((size_t) R * cap * 8), pa((size_t) R * cap * 10);
	std::vector<State> best_states;
	std::vector<Action> best_actions;
	bool first = true;
	for (std::uint64_t round = 0;; ++round) {
		P.stop_after_solved = goal_found ? 8 : 1;  // until a first solution exists the round ends with the first attempt that solves
		check(gbp_plan_batch(terrain.handle(), R, starts[0].data(), goals[0].data(), seed_, (stream_ << 20) + round * (std::uint64_t) R, &P,
							 stats.data(), ps.data(), pa.data(), cap), star ? "buildRRTStarConnect" : "buildRRTConnect");
		std::vector<int> solved;
		for (int i = 0; i < R; ++i) {
			num_vertices += stats[i].nv_a + stats[i].nv_b;
			if (stats[i].solved && stats[i].path_states <= cap) solved.push_back(i);
		}
		if (!solved.empty() && first) { elapsed_to_first = Clock::now() - t0; first = false; }
		// shortcut the most promising raw paths (postProcessPath is itself batched on the device)
		std::sort(solved.begin(), solved.end(), [&](int a, int b) { return stats[a].path_length < stats[b].path_length; });
		for (size_t k = 0; k < solved.size() && k < 8; ++k) {
			const int i = solved[k], n = stats[i].path_states;
			std::vector<State> ss(n);
			std::vector<Action> aa(n - 1);
			std::memcpy(ss[0].data(), &ps[(size_t) i * cap * 8], sizeof(State) * n);
			if (n > 1) std::memcpy(aa[0].data(), &pa[(size_t) i * cap * 10], sizeof(Action) * (n - 1));
			postProcessPath(ss, aa, terrain);
			goal_found = true;
			if (path_cost_ < cost_so_far) {
				cost_so_far = path_cost_;
				best_states = ss;
				best_actions = aa;
				length_vector_.push_back(path_length_);
				yaw_vector_.push_back(path_yaw_);
				cost_vector_.push_back(cost_so_far);
				cost_vector_times_.push_back(seconds_since(t0));
			}
		}
		const double elapsed = seconds_since(t0);
		if (elapsed >= max_time_solve_) {
			std::cout << "Failed, exiting" << std::endl;
			elapsed_total = Clock::now() - t0;
			elapsed_to_first = elapsed_total;
			success_ = 0;
			return;
		}
		if (goal_found && elapsed >= max_time_opt) break;  // :423
	}
	state_sequence = best_states;
	action_sequence = best_actions;
	postProcessPath(state_sequence, action_sequence, terrain);  // :455 (idempotent on an already shortcut path)
	elapsed_total = Clock::now() - t0;
	if (elapsed_total.count() <= 5.0) success_ = 1;
	path_duration_ = 0.0;
	for (const Action &a : action_sequence) path_duration_ += a[6] + a[7];
}

// =============================================================================== RRTStarConnectClass
RRTStarConnectClass::RRTStarConnectClass() {}
RRTStarConnectClass::~RRTStarConnectClass() {}

// rrt_star_connect.cpp:12-75.  The attemptConnect probes of the near set are geometry only, so they are
// evaluated in two batched launches; parent choice and rewiring then replay the reference's order.
int RRTStarConnectClass::extend(PlannerClass &T, State s, FastTerrainMap &terrain, int direction) {
	const int nearest = T.getNearestNeighbor(s);
	const State s_nearest = T.getVertex(nearest);
	State s_new;
	Action a_new;
	if (!newConfig(s, s_nearest, s_new, a_new, terrain, direction)) return TRAPPED;
	const int id = T.getNumVertices();
	T.addVertex(id, s_new);
	std::vector<int> nb = T.neighborhoodDist(s_new, delta);
	const int m = (int) nb.size();
	const int adaptive = state_action_pair_check_adaptive_step_size_flag_ ? 1 : 0;
	std::vector<State> near_states(m), rep_new(m, s_new), sn(m);
	std::vector<Action> an_in(m), an_out(m);
	std::vector<uint8_t> dir(m, (uint8_t) direction);
	std::vector<int> st_in(m), st_out(m);
	for (int i = 0; i < m; ++i) near_states[i] = T.getVertex(nb[i]);
	if (m) {
		check(gbp_attempt_connect(terrain.handle(), m, near_states[0].data(), rep_new[0].data(), dir.data(), adaptive, st_in.data(),
								  sn[0].data(), an_in[0].data(), nullptr), "RRT* choose parent");
		check(gbp_attempt_connect(terrain.handle(), m, rep_new[0].data(), near_states[0].data(), dir.data(), adaptive, st_out.data(),
								  sn[0].data(), an_out[0].data(), nullptr), "RRT* rewire");
	}
	int parent = nearest;
	double g_new = T.getGValue(nearest) + poseDistance(s_new, s_nearest), y_new = T.getYValue(nearest) + stateYawDistance(s_new, s_nearest);
	for (int i = 0; i < m; ++i) {
		if (st_in[i] != REACHED) continue;
		const double g = T.getGValue(nb[i]) + poseDistance(near_states[i], s_new);
		if (g < g_new) {
			a_new = an_in[i];
			parent = nb[i];
			g_new = g;
			y_new = T.getYValue(nb[i]) + stateYawDistance(near_states[i], s_new);
		}
	}
	T.addEdge(parent, id);
	T.updateGYValue(id, g_new, y_new);
	T.addAction(id, a_new);
	for (int i = 0; i < m; ++i) {
		const int k = nb[i];
		if (k == parent || st_out[i] != REACHED) continue;
		const double through_new = T.getGValue(id) + poseDistance(near_states[i], s_new);
		if (T.getGValue(k) > through_new) {
			T.removeEdge(T.getPredecessor(k), k);
			T.addEdge(id, k);
			T.updateGYValue(k, through_new, T.getYValue(id) + stateYawDistance(near_states[i], s_new));
			T.addAction(k, an_out[i]);
		}
	}
	return isWithinBounds(s_new, s) ? REACHED : ADVANCED;
}

void RRTStarConnectClass::getStateAndActionSequences(PlannerClass &Ta, PlannerClass &Tb, int shared_a_idx, int shared_b_idx,
													 std::vector<State> &state_sequence, std::vector<Action> &action_sequence) {
	std::vector<int> path_a = pathFromStart(Ta, shared_a_idx), path_b = pathFromStart(Tb, shared_b_idx);
	std::reverse(path_b.begin(), path_b.end());
	std::vector<Action> actions_b = getActionSequenceReverse(Tb, path_b);
	path_b.erase(path_b.begin());
	state_sequence = getStateSequence(Ta, path_a);
	std::vector<State> states_b = getStateSequence(Tb, path_b);
	state_sequence.insert(state_sequence.end(), states_b.begin(), states_b.end());
	action_sequence = getActionSequence(Ta, path_a);
	action_sequence.insert(action_sequence.end(), actions_b.begin(), actions_b.end());
}

// rrt_star_connect.cpp:100-226.  Like buildRRTConnect, re-designed for the GPU: every round runs `parallel_attempts_`
// independent RRT*-Connect searches resident on the device (gbp_plan_batch with rrt_star = 1: choose parent and
// near-set rewiring inside the kernel) and the cheapest post-processed path so far is kept; the reference's loop never
// exits before a goal is found, here max_time_solve bounds it.  (The single-tree loop driven call by call from the
// host — extend() above is still available to callers — manages ~1 k iterations/s and does not solve the shipped maps.)
void RRTStarConnectClass::buildRRTStarConnect(FastTerrainMap &terrain, State s_start, State s_goal, std::vector<State> &state_sequence,
											  std::vector<Action> &action_sequence, double max_time) {
	buildAnytime(terrain, s_start, s_goal, state_sequence, action_sequence, max_time, true);
	if (!goal_found) std::cout << "Path not found" << std::endl;
}
