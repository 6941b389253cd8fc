// Drop-in for the reference's include/global_body_planner/planning_utils.h: same namespace, constants,
// State/Action types and free-function signatures (:18-156).  Every function forwards to the C ABI
// (include/gbp_b200.h) and therefore runs on the GPU; the vector overloads batch n elements per launch.
#ifndef GBP_DROPIN_PLANNING_UTILS_H
#define GBP_DROPIN_PLANNING_UTILS_H

#include <array>
#include <cstdint>
#include <limits>
#include <utility>
#include <vector>

#include "fast_terrain_map.h"

namespace planning_utils {

// kinematic / dynamic / planning constants (reference :21-54; the FORK's values)
const double H_MAX = 0.4, H_MIN = 0.075, V_MAX = 2.0, V_NOM = 0.75, P_MAX = 1.0, DP_MAX = 3.0, ANG_ACC_MAX = 7.0;
const double ROBOT_L = 0.3, ROBOT_W = 0.3, ROBOT_H = 0.05;
const double M_CONST = 13, G_CONST = 9.81, F_MAX = 637, MU = 1.0;
const double T_S_MIN = 0.3, T_S_MAX = 0.3, T_F_MIN = 0.0, T_F_MAX = 0.5;
const double KINEMATICS_RES = 0.05, BACKUP_TIME = 0.2, BACKUP_RATIO = 0.5, GOAL_BOUNDS = 0.5;
const int NUM_GEN_STATES = 6;
const int FLIGHT = 0, STANCE = 1, CONNECT_STANCE = 2, FORWARD = 0, REVERSE = 1;
const int POSEDIM = 3, STATEDIM = 8, ACTIONDIM = 10;
typedef std::array<double, STATEDIM> State;    // x,y,z,dx,dy,dz,pitch,dpitch
typedef std::array<double, ACTIONDIM> Action;  // a_td(xyz), a_to(xyz), t_stance, t_flight, apitch_td, apitch_to
typedef std::pair<State, Action> StateActionPair;
const double INFTY = std::numeric_limits<double>::max();
const double MY_PI = 3.14159;

// ---- interpolation of a plan for output (:142-193)
void interpStateActionPair(State s, Action a, double t0, double dt, std::vector<State> &interp_path,
						   std::vector<double> &interp_t, std::vector<int> &interp_phase);
void getInterpPath(std::vector<State> state_sequence, std::vector<Action> action_sequence, double dt,
				   std::vector<State> &interp_path, std::vector<double> &interp_t, std::vector<int> &interp_phase);

State interp(State q1, State q2, double x);                   // :97-103
double calculateCurvature(double x1, double y1, double x2, double y2, double x3, double y3);  // :884-899 (three-point curvature)
double calculateMaxCurvature(std::vector<State> &body_plan);  // :900-909 (maximum over the plan)

// ---- printing and conversion helpers (:5-94)
void vectorToArray(State vec, double *new_array);
void printVectorInt(std::vector<int> vec);
void printVectorIntNewline(std::vector<int> vec);
void printInterpStateSequence(std::vector<State> state_sequence, std::vector<double> interp_t);
void printStateXYZPYaw(const State &s);
void printState(State vec);
void printStateNewline(State vec);
void printAction(Action a);
void printActionNewline(Action a);
void printStateSequence(std::vector<State> state_sequence);
void printActionSequence(std::vector<Action> action_sequence);
void printStateSequenceXYZPYaw(const std::vector<State> &state_sequence);

// ---- distances (:106-132, header :133-155)
double poseDistance(const State &q1, const State &q2);
double stateDistance(const State &q1, const State &q2);
double stateYawDistance(const State &q1, const State &q2);
double stateDistance(const State &q1, const State &q2, bool cost_add_yaw_flag, double cost_add_yaw_length_weight,
					 double cost_add_yaw_yaw_weight);
bool isWithinBounds(State s1, State s2);
// rotates a ground reaction force by the Rodrigues rotation taking +z to the surface normal (:198-231)
std::array<double, 3> rotate_grf(std::array<double, 3> surface_norm, std::array<double, 3> grf);

// ---- primitives (:237-370)
State applyStance(State s, Action a, double t);
State applyStance(State s, Action a);
State applyFlight(State s, double t_f);
State applyAction(State s, Action a);
State applyStanceReverse(State s, Action a, double t);
State applyStanceReverse(State s, Action a);

// ---- samplers (:379-515) on the Philox stream; set_random_stream() picks (seed, stream)
Action getRandomAction(std::array<double, 3> surf_norm, int direction, bool action_direction_sampling_flag,
					   double action_direction_sampling_probability_threshold, State s, State s_near);
Action getRandomAction(std::array<double, 3> surf_norm);
Action getRandomActionDirection(std::array<double, 3> surf_norm, State s_from, State s_to);
void set_random_stream(std::uint64_t seed, std::uint64_t stream);
std::uint64_t next_random_cell();  // the process-wide Philox cell counter shared by the samplers
std::uint64_t random_seed();
std::uint64_t random_stream();

// ---- validity (:519-876)
bool isValidAction(Action a);
bool isValidState(State s, FastTerrainMap &terrain, int phase);
bool isValidStateActionPair(State s, Action a, FastTerrainMap &terrain, State &s_new, double &t_new,
							bool state_action_pair_check_adaptive_step_size_flag);
bool isValidStateActionPairAdaptiveStepSize(State s, Action a, FastTerrainMap &terrain, State &s_new, double &t_new);
bool isValidStateActionPair(State s, Action a, FastTerrainMap &terrain, State &s_new, double &t_new);
bool isValidStateActionPair(State s, Action a, FastTerrainMap &terrain);
bool isValidStateActionPairReverse(State s, Action a, FastTerrainMap &terrain, State &s_new, double &t_new,
								   bool state_action_pair_check_adaptive_step_size_flag);
bool isValidStateActionPairReverseAdaptiveStepSize(State s, Action a, FastTerrainMap &terrain, State &s_new, double &t_new);
bool isValidStateActionPairReverse(State s, Action a, FastTerrainMap &terrain, State &s_new, double &t_new);
bool isValidStateActionPairReverse(State s, Action a, FastTerrainMap &terrain);

// ---- batched forms (no reference counterpart): n candidates per launch
std::vector<unsigned char> isValidState(const std::vector<State> &s, FastTerrainMap &terrain, int phase);
std::vector<unsigned char> isValidStateActionPair(const std::vector<State> &s, const std::vector<Action> &a,
												  const std::vector<unsigned char> &direction, FastTerrainMap &terrain,
												  std::vector<State> &s_new, std::vector<double> &t_new, bool adaptive = false);

}  // namespace planning_utils

#endif
