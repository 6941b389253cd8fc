// Drop-in for the reference's include/global_body_planner/planner_class.h (class PlannerClass, :17-83).
#ifndef GBP_DROPIN_PLANNER_CLASS_H
#define GBP_DROPIN_PLANNER_CLASS_H

#include "graph_class.h"

using namespace planning_utils;

class PlannerClass : public GraphClass {
public:
	PlannerClass();
	~PlannerClass();

	State randomState(FastTerrainMap &terrain, bool state_direction_sampling_flag,
					  double state_direction_sampling_probability_threshold, bool speed_direction_flag, State s_from,
					  State s_to);
	State randomState(FastTerrainMap &terrain);
	State randomStateDirection(FastTerrainMap &terrain, State s_from, State s_to, bool speed_direction_flag);
	std::vector<int> neighborhoodN(State s, int N);
	std::vector<int> neighborhoodDist(State q, double dist);
	int getNearestNeighbor(State q);
};

#endif
