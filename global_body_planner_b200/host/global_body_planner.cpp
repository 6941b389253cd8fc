// GlobalBodyPlanner without ROS: the planning driver of the reference's src/global_body_planner.cpp:60-270 over the
// B200 planner classes.  Control flow, statistics and console output follow callPlanner(); every number is produced by
// the GPU path (planners, interpolation, curvature, terrain lookups) through the C ABI.
#include <cmath>
#include <iostream>
#include <stdexcept>

#include "../../include/global_body_planner/global_body_planner.h"

#include <fstream>
#include <map>

namespace {
// the YAML subset of the reference's config/params.yaml, flattened to rosparam names ("global_body_planner/cost_add_yaw/flag")
std::map<std::string, std::string> flatten_yaml(const std::string &path) {
	std::ifstream f(path);
	if (!f) throw std::runtime_error("loadParamsYaml: cannot open " + path);
	std::map<std::string, std::string> out;
	std::vector<std::pair<int, std::string>> stack;  // (indent, key) of the open mappings
	std::string line;
	while (std::getline(f, line)) {
		bool quoted = false;
		size_t cut = std::string::npos;
		for (size_t i = 0; i < line.size(); ++i) {  // a comment starts at a # outside quotes that begins the line or follows a blank
			if (line[i] == '"' || line[i] == '\'') quoted = !quoted;
			if (line[i] == '#' && !quoted && (i == 0 || line[i - 1] == ' ' || line[i - 1] == '\t')) { cut = i; break; }
		}
		if (cut != std::string::npos) line.erase(cut);
		while (!line.empty() && (line.back() == ' ' || line.back() == '\t' || line.back() == '\r')) line.pop_back();
		const size_t indent = line.find_first_not_of(' ');
		if (indent == std::string::npos || line[indent] == '-') continue;  // blank lines and sequence items (RViz section)
		const size_t colon = line.find(':', indent);
		if (colon == std::string::npos) continue;
		std::string key = line.substr(indent, colon - indent), value = colon + 1 < line.size() ? line.substr(colon + 1) : "";
		const size_t v0 = value.find_first_not_of(" \t");
		value = v0 == std::string::npos ? "" : value.substr(v0);
		if (value.size() >= 2 && (value.front() == '"' || value.front() == '\'') && value.back() == value.front()) value = value.substr(1, value.size() - 2);
		while (!stack.empty() && stack.back().first >= (int) indent) stack.pop_back();
		std::string name;
		for (const auto &s : stack) name += s.second + "/";
		name += key;
		if (value.empty()) stack.emplace_back((int) indent, key);
		else out[name] = value;
	}
	return out;
}
bool yaml_bool(const std::string &name, const std::string &v) {
	if (v == "true" || v == "True" || v == "TRUE" || v == "yes" || v == "on") return true;
	if (v == "false" || v == "False" || v == "FALSE" || v == "no" || v == "off") return false;
	throw std::runtime_error("loadParamsYaml: " + name + ": '" + v + "' is not a bool");
}
double yaml_double(const std::string &name, const std::string &v) {
	size_t used = 0;
	double d = 0;
	try { d = std::stod(v, &used); } catch (const std::exception &) { used = 0; }
	if (used != v.size() || v.empty()) throw std::runtime_error("loadParamsYaml: " + name + ": '" + v + "' is not a number");
	return d;
}
}  // namespace

std::vector<std::string> loadParamsYaml(const std::string &path, GlobalBodyPlannerParams &p) {
	const std::map<std::string, std::string> y = flatten_yaml(path);
	std::vector<std::string> applied;
	auto has = [&](const char *name) { const bool h = y.count(name) != 0; if (h) applied.push_back(name); return h; };
	auto B = [&](const char *name, bool &dst) { if (has(name)) dst = yaml_bool(name, y.at(name)); };
	auto D = [&](const char *name, double &dst) { if (has(name)) dst = yaml_double(name, y.at(name)); };
	if (has("global_body_planner/num_calls")) p.num_calls = (int) yaml_double("global_body_planner/num_calls", y.at("global_body_planner/num_calls"));
	D("global_body_planner/replan_time_limit", p.replan_time_limit);
	if (has("global_body_planner/algorithm")) {
		p.algorithm = y.at("global_body_planner/algorithm");
		if (p.algorithm != "rrt-connect" && p.algorithm != "rrt-star-connect")
			throw std::runtime_error("loadParamsYaml: global_body_planner/algorithm: '" + p.algorithm + "' (rrt-connect | rrt-star-connect)");
	}
	B("global_body_planner/state_action_pair_check_adaptive_step_size_flag", p.state_action_pair_check_adaptive_step_size_flag);
	B("global_body_planner/cost_add_yaw/flag", p.cost_add_yaw_flag);
	D("global_body_planner/cost_add_yaw/length_weight", p.cost_add_yaw_length_weight);
	D("global_body_planner/cost_add_yaw/yaw_weight", p.cost_add_yaw_yaw_weight);
	B("global_body_planner/action_direction_sampling/flag", p.action_direction_sampling_flag);
	D("global_body_planner/action_direction_sampling/probability_threshold", p.action_direction_sampling_probability_threshold);
	B("global_body_planner/state_direction_sampling/flag", p.state_direction_sampling_flag);
	D("global_body_planner/state_direction_sampling/probability_threshold", p.state_direction_sampling_probability_threshold);
	B("global_body_planner/state_direction_sampling/speed_direction_flag", p.state_direction_sampling_speed_direction_flag);
	D("state_publisher/start_position_x", p.start_position_x);
	D("state_publisher/start_position_y", p.start_position_y);
	D("state_publisher/start_yaw", p.start_yaw);
	D("state_publisher/goal_position_x", p.goal_position_x);
	D("state_publisher/goal_position_y", p.goal_position_y);
	D("state_publisher/goal_yaw", p.goal_yaw);
	return applied;
}

GlobalBodyPlanner::GlobalBodyPlanner(const GlobalBodyPlannerParams &params) : p_(params) {
	robot_start_.fill(0);
	robot_goal_.fill(0);
}

void GlobalBodyPlanner::setTerrain(const FastTerrainMap &terrain) { terrain_ = terrain; }
void GlobalBodyPlanner::loadTerrainFromCSV(const std::string &directory, bool via_gridmap) { terrain_.loadDataFromCSV(directory, via_gridmap); }

void GlobalBodyPlanner::clearPlan() {  // :53-63
	body_plan_.clear();
	t_plan_.clear();
	interp_phase_.clear();
	solve_time_info_.clear();
	vertices_generated_info_.clear();
	length_vectors_.clear();
	yaw_vectors_.clear();
	cost_vectors_.clear();
	cost_vectors_times_.clear();
}

void GlobalBodyPlanner::setPlannerParameter(RRTClass &rrt_obj) {  // :171-200
	rrt_obj.set_state_action_pair_check_adaptive_step_size_flag_(p_.state_action_pair_check_adaptive_step_size_flag);
	rrt_obj.set_cost_add_yaw(p_.cost_add_yaw_flag, p_.cost_add_yaw_length_weight, p_.cost_add_yaw_yaw_weight);
	rrt_obj.set_action_direction_sampling(p_.action_direction_sampling_flag, p_.action_direction_sampling_probability_threshold);
	rrt_obj.set_state_direction_sampling(p_.state_direction_sampling_flag, p_.state_direction_sampling_probability_threshold,
										 p_.state_direction_sampling_speed_direction_flag);
}

// heading (dx, dy) with atan2(dy, dx) = yaw, as :226-252 builds it (unit x component, tan for the y component,
// +-pi/2 within 2 degrees handled separately)
static void heading(double yaw, double &dx, double &dy) {
	const double delta = 0.0349066;
	if (yaw >= M_PI_2 - delta && yaw <= M_PI_2 + delta) { dx = 0; dy = 1; }
	else if (yaw >= -M_PI_2 - delta && yaw <= -M_PI_2 + delta) { dx = 0; dy = -1; }
	else if (yaw > M_PI_2) { dx = -1; dy = -std::tan(yaw - M_PI); }
	else if (yaw < -M_PI_2) { dx = -1; dy = -std::tan(yaw + M_PI); }
	else { dx = 1; dy = std::tan(yaw); }
}

void GlobalBodyPlanner::setStartAndGoalStates() {  // :203-257
	double sdx, sdy, gdx, gdy;
	heading(p_.start_yaw, sdx, sdy);
	heading(p_.goal_yaw, gdx, gdy);
	robot_start_ = {p_.start_position_x, p_.start_position_y, p_.body_height, sdx, sdy, 0, 0, 0};
	robot_goal_ = {p_.goal_position_x, p_.goal_position_y, p_.body_height, gdx, gdy, 0, 0, 0};
	if (p_.verbose) {
		std::cout << "start: x:" << robot_start_[0] << ", y:" << robot_start_[1] << ", z:" << robot_start_[2]
				  << ", yaw:" << std::atan2(robot_start_[4], robot_start_[3]) << std::endl;
		std::cout << "goal: x:" << robot_goal_[0] << ", y:" << robot_goal_[1] << ", z:" << robot_goal_[2]
				  << ", yaw:" << std::atan2(robot_goal_[4], robot_goal_[3]) << std::endl;
	}
	robot_start_[2] += terrain_.getGroundHeight(robot_start_[0], robot_start_[1]);
	robot_goal_[2] += terrain_.getGroundHeight(robot_goal_[0], robot_goal_[1]);
}

void GlobalBodyPlanner::callPlanner() {  // :60-168
	setStartAndGoalStates();
	clearPlan();
	double plan_time = 0, time_to_first_solve = 0, path_duration = 0, max_curvature = 0;
	int success = 0, vertices_generated = 0;
	double total_solve_time = 0, total_vertices_generated = 0, total_path_length = 0, total_path_yaw = 0, total_path_cost = 0,
		   total_path_duration = 0, total_max_curvature = 0;
	int successes = 0;
	RRTConnectClass rrt_connect_obj;
	RRTStarConnectClass rrt_star_connect_obj;
	setPlannerParameter(rrt_connect_obj);
	setPlannerParameter(rrt_star_connect_obj);
	rrt_connect_obj.set_parallel_attempts(p_.parallel_attempts, p_.iterations_per_attempt, p_.vertices_per_tree);
	rrt_connect_obj.set_max_time_solve(p_.max_time_solve);
	rrt_star_connect_obj.set_max_time_solve(p_.max_time_solve);
	if (p_.algorithm != "rrt-connect" && p_.algorithm != "rrt-star-connect") throw std::runtime_error("Invalid algorithm specified");  // :126
	for (int i = 0; i < p_.num_calls; ++i) {
		if (p_.verbose) std::cout << "----- plan times: " << (i + 1) << " / " << p_.num_calls << " -----" << std::endl;
		state_sequence_.clear();
		action_sequence_.clear();
		body_plan_.clear();
		t_plan_.clear();
		interp_phase_.clear();
		std::vector<double> length_vector, yaw_vector, cost_vector, cost_vector_times;
		// every call draws from its own Philox stream (the reference's rand() simply keeps running)
		if (p_.algorithm == "rrt-connect") {
			rrt_connect_obj.set_random_stream(p_.seed, (std::uint64_t) i);
			rrt_connect_obj.buildRRTConnect(terrain_, robot_start_, robot_goal_, state_sequence_, action_sequence_, p_.replan_time_limit);
			rrt_connect_obj.getStatistics(plan_time, success, vertices_generated, time_to_first_solve, length_vector, yaw_vector, cost_vector,
										  cost_vector_times, path_duration, allStatePosition);
		} else {
			rrt_star_connect_obj.set_random_stream(p_.seed, (std::uint64_t) i);
			planning_utils::set_random_stream(p_.seed, (std::uint64_t) i);
			rrt_star_connect_obj.buildRRTStarConnect(terrain_, robot_start_, robot_goal_, state_sequence_, action_sequence_, p_.replan_time_limit);
			rrt_star_connect_obj.getStatistics(plan_time, success, vertices_generated, time_to_first_solve, length_vector, yaw_vector,
											   cost_vector, cost_vector_times, path_duration, allStatePosition);
		}
		solve_time_info_.push_back(plan_time);
		vertices_generated_info_.push_back(vertices_generated);
		total_solve_time += plan_time;
		total_vertices_generated += vertices_generated;
		if (state_sequence_.empty() || length_vector.empty()) {  // the reference dereferences .back() of empty vectors here (UB)
			if (p_.verbose) std::cout << "No plan found" << std::endl;
			continue;
		}
		successes += success;
		const double dt = 0.05;  // :131
		getInterpPath(state_sequence_, action_sequence_, dt, body_plan_, t_plan_, interp_phase_);
		length_vectors_.push_back(length_vector);
		yaw_vectors_.push_back(yaw_vector);
		cost_vectors_.push_back(cost_vector);
		cost_vectors_times_.push_back(cost_vector_times);
		max_curvature = calculateMaxCurvature(body_plan_);
		total_path_length += length_vector.back();
		total_path_yaw += yaw_vector.back();
		total_path_cost += cost_vector.back();
		total_path_duration += path_duration;
		total_max_curvature += max_curvature;
		if (p_.verbose) {
			std::cout << "Vertices generated: " << vertices_generated << std::endl;
			std::cout << "Solve time: " << plan_time << std::endl;
			std::cout << "Time to first solve: " << time_to_first_solve << std::endl;
			std::cout << "Path length: " << length_vector.back() << std::endl;
			std::cout << "Path yaw: " << yaw_vector.back() << std::endl;
			std::cout << "Path cost: " << cost_vector.back() << std::endl;
			std::cout << "Path duration: " << path_duration << std::endl;
			std::cout << "Path max curvature: " << max_curvature << std::endl;
		}
	}
	if (p_.verbose) rrt_connect_obj.print_setting_parameters();
	const double n = p_.num_calls > 0 ? (double) p_.num_calls : 1.0;
	averages_ = {total_vertices_generated / n, total_solve_time / n, total_path_length / n, total_path_yaw / n, total_path_cost / n,
				 total_path_duration / n, total_max_curvature / n, p_.num_calls, successes};
	if (!p_.verbose) return;
	if (p_.num_calls > 1) {
		std::cout << "---------- Average ----------" << std::endl;
		std::cout << "Average vertices generated: " << averages_.vertices_generated << std::endl;
		std::cout << "Average solve time: " << averages_.solve_time << std::endl;
		std::cout << "Average path length: " << averages_.path_length << std::endl;
		std::cout << "Average path yaw: " << averages_.path_yaw << std::endl;
		std::cout << "Average path cost: " << averages_.path_cost << std::endl;
		std::cout << "Average path duration: " << averages_.path_duration << std::endl;
		std::cout << "Average path max curvature: " << averages_.max_curvature << std::endl;
	} else if (!state_sequence_.empty()) {
		printStateSequenceXYZPYaw(state_sequence_);
	}
}
