// Drop-in for the reference's include/global_body_planner/rrt.h (class RRTClass, :17-208).
#ifndef GBP_DROPIN_RRT_H
#define GBP_DROPIN_RRT_H

#include <chrono>

#include "planner_class.h"

#define TRAPPED 0
#define ADVANCED 1
#define REACHED 2

using namespace planning_utils;

class RRTClass {
public:
	RRTClass();
	virtual ~RRTClass();

	virtual int extend(PlannerClass &T, State s, FastTerrainMap &terrain, int direction);
	std::vector<int> pathFromStart(PlannerClass &T, int idx);
	void printPath(PlannerClass &T, std::vector<int> path);
	void buildRRT(FastTerrainMap &terrain, State s_start, State s_goal, std::vector<State> &state_sequence,
				  std::vector<Action> &action_sequence);
	void getStatistics(double &plan_time, int &success_var, int &vertices_generated, double &time_to_first_solve,
					   std::vector<double> &length_vector, std::vector<double> &yaw_vector,
					   std::vector<double> &cost_vector, std::vector<double> &cost_vector_times, double &path_duration,
					   std::vector<std::vector<double>> &allStatePosition);
	bool newConfig(State s, State s_near, State &s_new, Action &a_new, FastTerrainMap &terrain, int direction);
	std::vector<State> getStateSequence(PlannerClass &T, std::vector<int> path);
	std::vector<Action> getActionSequence(PlannerClass &T, std::vector<int> path);
	void saveStateSequence(PlannerClass &T);
	void set_action_direction_sampling(bool flag, double threshold);
	void set_state_direction_sampling(bool flag, double threshold, bool speed_direction_flag);
	void set_state_action_pair_check_adaptive_step_size_flag_(bool state_action_pair_check_adaptive_step_size_flag);
	void set_cost_add_yaw(bool flag, double length_weight, double yaw_weight);
	void print_setting_parameters();

	// B200 additions: candidates per extend (default NUM_GEN_STATES, first valid decides = the reference) and
	// the Philox (seed, stream) the planner draws from.
	void set_candidates_per_extend(int k, bool best_of_k);
	void set_random_stream(std::uint64_t seed, std::uint64_t stream);

protected:
	const double prob_goal_thresh = 0.05;
	bool goal_found = false;
	std::chrono::duration<double> elapsed_total;
	std::chrono::duration<double> elapsed_to_first;
	int success_ = 0;
	int num_vertices = 0;
	double path_length_ = 0, path_yaw_ = 0, path_cost_ = 0;
	std::vector<double> length_vector_, yaw_vector_, cost_vector_, cost_vector_times_;
	double path_duration_ = 0;
	std::vector<std::vector<double>> allStatePosition_;
	bool action_direction_sampling_flag_ = false;
	double action_direction_sampling_probability_threshold_ = 0.15;
	bool state_direction_sampling_flag_ = false;
	double state_direction_sampling_probability_threshold_ = 0.05;
	bool state_direction_sampling_speed_direction_flag_ = false;
	bool state_action_pair_check_adaptive_step_size_flag_ = false;
	bool cost_add_yaw_flag_ = false;
	double cost_add_yaw_length_weight_ = 1, cost_add_yaw_yaw_weight_ = 1;
	int k_candidates_ = NUM_GEN_STATES;
	bool best_of_k_ = false;
	std::uint64_t seed_ = 1, stream_ = 0, cell_ = 0;
};

#endif
