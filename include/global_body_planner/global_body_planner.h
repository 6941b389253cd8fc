// Drop-in for the reference's include/global_body_planner/global_body_planner.h (class GlobalBodyPlanner) WITHOUT ROS:
// the planning driver of src/global_body_planner.cpp:60-270 — parameter plumbing (setPlannerParameter), start / goal
// construction (setStartAndGoalStates), the num_calls loop with per-call and average statistics (callPlanner), plan
// interpolation and curvature — over the B200 planners.  What the reference reads from the ROS parameter server
// arrives in a plain struct with the same names and defaults; what it publishes (BodyPlan) is exposed through
// accessors.  Topics, publishers, the spin loop and RViz output are out of scope.
#ifndef GBP_DROPIN_GLOBAL_BODY_PLANNER_H
#define GBP_DROPIN_GLOBAL_BODY_PLANNER_H

#include <string>

#include "rrt_star_connect.h"

struct GlobalBodyPlannerParams {  // rosparam names: global_body_planner/* and state_publisher/* (config/params.yaml)
	int num_calls = 1;                       // global_body_planner.cpp:16
	double replan_time_limit = 0.0;          // :17
	std::string algorithm = "rrt-connect";   // :18  ("rrt-connect" | "rrt-star-connect")
	bool state_action_pair_check_adaptive_step_size_flag = false;  // :175-176
	bool cost_add_yaw_flag = false;          // :181-183
	double cost_add_yaw_length_weight = 1.0, cost_add_yaw_yaw_weight = 1.0;
	bool action_direction_sampling_flag = false;  // :189-190
	double action_direction_sampling_probability_threshold = 0.1;   // config/params.yaml:27 (the node's fallback without a rosparam is 0.15, :201)
	bool state_direction_sampling_flag = false;   // :196-198
	double state_direction_sampling_probability_threshold = 0.05;   // config/params.yaml:23 (fallback 0.15, :203)
	bool state_direction_sampling_speed_direction_flag = false;
	double start_position_x = 0, start_position_y = 0, start_yaw = 0;  // :214-218
	double goal_position_x = 0, goal_position_y = 0, goal_yaw = 0;     // :219-223
	double body_height = 0.375;              // hard-coded start_z / goal_z of the fork (:214, :219)
	// B200 additions (no reference counterpart): Philox seed, device searches per anytime round, quiet mode
	unsigned long long seed = 1;
	int parallel_attempts = 2048, iterations_per_attempt = 32000, vertices_per_tree = 2048;  // see rrt_connect.h
	double max_time_solve = 4000;            // rrt_connect.h:119-120
	bool verbose = true;                     // the reference prints every statistic to stdout
};

// What `rosparam load config/params.yaml` + the launch files give the reference's node: reads the YAML subset that file uses
// (nested block mappings, scalar values, # comments) and overwrites the fields of `p` whose rosparam names
// (global_body_planner/*, state_publisher/{start,goal}_*; global_body_planner.cpp:15-28, :181-204, :220-228) appear in it;
// names it does not know (topics, update rates, the publisher's and RViz's sections) are ignored as the planner node ignores
// them.  Returns the rosparam names it applied; throws std::runtime_error on an unreadable file or a malformed value.
std::vector<std::string> loadParamsYaml(const std::string &path, GlobalBodyPlannerParams &p);

class GlobalBodyPlanner {
public:
	explicit GlobalBodyPlanner(const GlobalBodyPlannerParams &params);

	// terrainMapCallback (:43-50) hands a GridMap to FastTerrainMap; here the caller hands over the FastTerrainMap,
	// or the reference's CSV directory (data/<terrain_type>, terrain_map_publisher.cpp:330-370)
	void setTerrain(const FastTerrainMap &terrain);
	void loadTerrainFromCSV(const std::string &directory, bool via_gridmap = false);

	void callPlanner();  // :60-168

	// what publishPlan() would put on the wire (BodyPlan: interpolated states, times) and the per-call statistics
	const std::vector<State> &bodyPlan() const { return body_plan_; }
	const std::vector<double> &planTimes() const { return t_plan_; }
	const std::vector<int> &planPhases() const { return interp_phase_; }
	const std::vector<State> &stateSequence() const { return state_sequence_; }
	const std::vector<Action> &actionSequence() const { return action_sequence_; }
	const std::vector<double> &solveTimeInfo() const { return solve_time_info_; }
	const std::vector<int> &verticesGeneratedInfo() const { return vertices_generated_info_; }
	State robotStart() const { return robot_start_; }
	State robotGoal() const { return robot_goal_; }
	struct Averages { double vertices_generated, solve_time, path_length, path_yaw, path_cost, path_duration, max_curvature; int calls, successes; };
	Averages averages() const { return averages_; }

private:
	void setPlannerParameter(RRTClass &rrt_obj);  // :171-200
	void setStartAndGoalStates();                 // :203-257
	void clearPlan();                             // :53-63

	GlobalBodyPlannerParams p_;
	FastTerrainMap terrain_;
	std::vector<State> body_plan_, state_sequence_;
	std::vector<double> t_plan_;
	std::vector<int> interp_phase_;
	std::vector<Action> action_sequence_;
	State robot_start_, robot_goal_;
	std::vector<double> solve_time_info_;
	std::vector<int> vertices_generated_info_;
	std::vector<std::vector<double>> length_vectors_, yaw_vectors_, cost_vectors_, cost_vectors_times_;
	std::vector<std::vector<double>> allStatePosition;
	Averages averages_ = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

#endif
