// Drop-in for the reference's include/global_body_planner/rrt_connect.h (class RRTConnectClass, :19-121).
#ifndef GBP_DROPIN_RRT_CONNECT_H
#define GBP_DROPIN_RRT_CONNECT_H

#include "rrt.h"

using namespace planning_utils;

class RRTConnectClass : public RRTClass {
public:
	RRTConnectClass();
	~RRTConnectClass();

	int attemptConnect(State s_existing, State s, double t_s, State &s_new, Action &a_new, FastTerrainMap &terrain,
					   int direction);
	int attemptConnect(State s_existing, State s, State &s_new, Action &a_new, FastTerrainMap &terrain, int direction);
	int connect(PlannerClass &T, State s, FastTerrainMap &terrain, int direction);
	std::vector<Action> getActionSequenceReverse(PlannerClass &T, std::vector<int> path);
	void postProcessPath(std::vector<State> &state_sequence, std::vector<Action> &action_sequence, FastTerrainMap &terrain);
	void runRRTConnect(PlannerClass &Ta, PlannerClass &Tb, FastTerrainMap &terrain);
	// Anytime planning.  B200-native: every round launches `parallel_attempts` independent bidirectional
	// searches (distinct Philox streams) resident on the device and keeps the shortest post-processed path;
	// terminates like the reference (:423): a solution exists and max_time has elapsed.
	void buildRRTConnect(FastTerrainMap &terrain, State s_start, State s_goal, std::vector<State> &state_sequence,
						 std::vector<Action> &action_sequence, double max_time);
	void set_parallel_attempts(int attempts, int iterations_per_attempt, int vertices_per_tree);
	void set_max_time_solve(double seconds);  // hard stop (the reference's constant max_time_solve, :119-120)

protected:
	// shared by buildRRTConnect and RRTStarConnectClass::buildRRTStarConnect
	void buildAnytime(FastTerrainMap &terrain, State s_start, State s_goal, std::vector<State> &state_sequence,
					  std::vector<Action> &action_sequence, double max_time, bool star);
	double anytime_horizon = 0;
	const double planning_rate_estimate = 16.0;
	double anytime_horizon_init = 0;
	double horizon_expansion_factor = 1.2;
	const int max_time_solve = 4000;
	// RRT-Connect rounds run on the pipelined device planner with 16 speculated half-iterations per round: deep attempts find the
	// first solution soonest (tools/ttfs_sweep.py on data/slope: 2048 x 32000 mean 0.13 s, 4096 x 8000 0.16 s; megakernel
	// 3552 x 8000 0.29 s).  RRT*-Connect rounds run on the megakernel: one wave of warps on 148 SMs (24 per SM).  Trees as large
	// as the reference's solves (1-2 k vertices).
	int parallel_attempts_ = 2048, iterations_per_attempt_ = 32000, vertices_per_tree_ = 2048;
	bool attempts_set_ = false;  // set_parallel_attempts was called: RRT*-Connect rounds use its values too, otherwise 3552 x 8000
	double max_time_solve_ = 4000;
};

#endif
