// Drop-in for the reference's include/global_body_planner/rrt_star_connect.h (class RRTStarConnectClass, :14-60).
#ifndef GBP_DROPIN_RRT_STAR_CONNECT_H
#define GBP_DROPIN_RRT_STAR_CONNECT_H

#include "rrt_connect.h"

using namespace planning_utils;

class RRTStarConnectClass : public RRTConnectClass {
public:
	RRTStarConnectClass();
	~RRTStarConnectClass();

	int extend(PlannerClass &T, State s, FastTerrainMap &terrain, int direction);
	void getStateAndActionSequences(PlannerClass &Ta, PlannerClass &Tb, int shared_a_idx, int shared_b_idx,
									std::vector<State> &state_sequence, std::vector<Action> &action_sequence);
	void buildRRTStarConnect(FastTerrainMap &terrain, State s_start, State s_goal, std::vector<State> &state_sequence,
							 std::vector<Action> &action_sequence, double max_time);

protected:
	const double delta = 3.0;
};

#endif
