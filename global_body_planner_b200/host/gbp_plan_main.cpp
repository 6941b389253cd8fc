// gbp_plan — command-line front end of the ROS-free GlobalBodyPlanner driver: the reference's benchmarking protocol
// ("num_calls planner calls, print every statistic, then the averages", global_body_planner.cpp:60-168) on a terrain in
// the reference's CSV format.
//   gbp_plan <csv-dir | create:SEED | default> [--algorithm rrt-connect|rrt-star-connect] [--num-calls N] [--replan-time-limit S]
//            [--start X Y YAW] [--goal X Y YAW] [--height H] [--seed S] [--gridmap] [--attempts A ITERS VERTS]
//            [--max-time-solve S] [--adaptive] [--plan-out FILE] [--discrete-out FILE] [--quiet]
//            [--params FILE]                       the reference's config/params.yaml (rosparam names; later options override it)
//            [--cost-add-yaw LENGTH_W YAW_W] [--state-direction-sampling P] [--speed-direction] [--action-direction-sampling P]
//            [--print-params]                      print the resolved parameters and exit (no device needed)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "../../include/global_body_planner/global_body_planner.h"

int main(int argc, char **argv) {
	if (argc < 2) {
		std::fprintf(stderr, "usage: %s <csv-dir> [options]  (see the header of gbp_plan_main.cpp)\n", argv[0]);
		return 2;
	}
	GlobalBodyPlannerParams p;
	p.goal_position_x = 8.0;  // launch/example.launch: (0, 0) -> (8, 0)
	bool gridmap = false;
	bool print_params = false;
	const char *plan_out = nullptr, *discrete_out = nullptr;
	for (int i = 2; i < argc; ++i) {
		auto need = [&](int k) { if (i + k >= argc) { std::fprintf(stderr, "%s needs %d value(s)\n", argv[i], k); std::exit(2); } };
		if (!std::strcmp(argv[i], "--algorithm")) { need(1); p.algorithm = argv[++i]; }
		else if (!std::strcmp(argv[i], "--num-calls")) { need(1); p.num_calls = std::atoi(argv[++i]); }
		else if (!std::strcmp(argv[i], "--replan-time-limit")) { need(1); p.replan_time_limit = std::atof(argv[++i]); }
		else if (!std::strcmp(argv[i], "--start")) { need(3); p.start_position_x = std::atof(argv[++i]); p.start_position_y = std::atof(argv[++i]); p.start_yaw = std::atof(argv[++i]); }
		else if (!std::strcmp(argv[i], "--goal")) { need(3); p.goal_position_x = std::atof(argv[++i]); p.goal_position_y = std::atof(argv[++i]); p.goal_yaw = std::atof(argv[++i]); }
		else if (!std::strcmp(argv[i], "--height")) { need(1); p.body_height = std::atof(argv[++i]); }
		else if (!std::strcmp(argv[i], "--seed")) { need(1); p.seed = std::strtoull(argv[++i], nullptr, 10); }
		else if (!std::strcmp(argv[i], "--gridmap")) gridmap = true;
		else if (!std::strcmp(argv[i], "--adaptive")) p.state_action_pair_check_adaptive_step_size_flag = true;
		else if (!std::strcmp(argv[i], "--attempts")) { need(3); p.parallel_attempts = std::atoi(argv[++i]); p.iterations_per_attempt = std::atoi(argv[++i]); p.vertices_per_tree = std::atoi(argv[++i]); }
		else if (!std::strcmp(argv[i], "--max-time-solve")) { need(1); p.max_time_solve = std::atof(argv[++i]); }
		else if (!std::strcmp(argv[i], "--plan-out")) { need(1); plan_out = argv[++i]; }
		else if (!std::strcmp(argv[i], "--discrete-out")) { need(1); discrete_out = argv[++i]; }
		else if (!std::strcmp(argv[i], "--quiet")) p.verbose = false;
		else if (!std::strcmp(argv[i], "--params")) {
			need(1);
			try { loadParamsYaml(argv[++i], p); } catch (const std::exception &e) { std::fprintf(stderr, "gbp_plan: %s\n", e.what()); return 2; }
		}
		else if (!std::strcmp(argv[i], "--cost-add-yaw")) { need(2); p.cost_add_yaw_flag = true; p.cost_add_yaw_length_weight = std::atof(argv[++i]); p.cost_add_yaw_yaw_weight = std::atof(argv[++i]); }
		else if (!std::strcmp(argv[i], "--state-direction-sampling")) { need(1); p.state_direction_sampling_flag = true; p.state_direction_sampling_probability_threshold = std::atof(argv[++i]); }
		else if (!std::strcmp(argv[i], "--speed-direction")) p.state_direction_sampling_speed_direction_flag = true;
		else if (!std::strcmp(argv[i], "--action-direction-sampling")) { need(1); p.action_direction_sampling_flag = true; p.action_direction_sampling_probability_threshold = std::atof(argv[++i]); }
		else if (!std::strcmp(argv[i], "--print-params")) print_params = true;
		else { std::fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
	}
	if (print_params) {  // rosparam names, one per line
		std::cout.precision(17);
		std::cout << "global_body_planner/num_calls " << p.num_calls << "\nglobal_body_planner/replan_time_limit " << p.replan_time_limit
				  << "\nglobal_body_planner/algorithm " << p.algorithm
				  << "\nglobal_body_planner/state_action_pair_check_adaptive_step_size_flag " << p.state_action_pair_check_adaptive_step_size_flag
				  << "\nglobal_body_planner/cost_add_yaw/flag " << p.cost_add_yaw_flag << "\nglobal_body_planner/cost_add_yaw/length_weight "
				  << p.cost_add_yaw_length_weight << "\nglobal_body_planner/cost_add_yaw/yaw_weight " << p.cost_add_yaw_yaw_weight
				  << "\nglobal_body_planner/action_direction_sampling/flag " << p.action_direction_sampling_flag
				  << "\nglobal_body_planner/action_direction_sampling/probability_threshold " << p.action_direction_sampling_probability_threshold
				  << "\nglobal_body_planner/state_direction_sampling/flag " << p.state_direction_sampling_flag
				  << "\nglobal_body_planner/state_direction_sampling/probability_threshold " << p.state_direction_sampling_probability_threshold
				  << "\nglobal_body_planner/state_direction_sampling/speed_direction_flag " << p.state_direction_sampling_speed_direction_flag
				  << "\nstate_publisher/start_position_x " << p.start_position_x << "\nstate_publisher/start_position_y " << p.start_position_y
				  << "\nstate_publisher/start_yaw " << p.start_yaw << "\nstate_publisher/goal_position_x " << p.goal_position_x
				  << "\nstate_publisher/goal_position_y " << p.goal_position_y << "\nstate_publisher/goal_yaw " << p.goal_yaw
				  << "\nbody_height " << p.body_height << "\nseed " << p.seed << "\nparallel_attempts " << p.parallel_attempts
				  << "\niterations_per_attempt " << p.iterations_per_attempt << "\nvertices_per_tree " << p.vertices_per_tree
				  << "\nmax_time_solve " << p.max_time_solve << std::endl;
		return 0;
	}
	try {
		GlobalBodyPlanner planner(p);
		if (!std::strncmp(argv[1], "create:", 7)) {  // the publisher's "create" source on Philox seed N
			FastTerrainMap m;
			m.createOwnMap(std::strtoull(argv[1] + 7, nullptr, 10));
			planner.setTerrain(m);
		} else if (!std::strcmp(argv[1], "default")) {  // the publisher's default source (createMap)
			FastTerrainMap m;
			m.createMap();
			planner.setTerrain(m);
		} else planner.loadTerrainFromCSV(argv[1], gridmap);
		planner.callPlanner();
		if (plan_out) {  // the BodyPlan wire content: t, 8 state components, phase (the closing state carries -1)
			std::ofstream f(plan_out);
			f.precision(17);
			const auto &bp = planner.bodyPlan();
			for (size_t i = 0; i < bp.size(); ++i) {
				f << planner.planTimes()[i];
				for (double v : bp[i]) f << "," << v;
				f << "," << (i < planner.planPhases().size() ? planner.planPhases()[i] : -1) << "\n";
			}
		}
		if (discrete_out) {  // the discrete plan of the last call: 8 state components, then the 10 action components leaving that state
			std::ofstream f(discrete_out);
			f.precision(17);
			const auto &ss = planner.stateSequence();
			const auto &aa = planner.actionSequence();
			for (size_t i = 0; i < ss.size(); ++i) {
				for (size_t d = 0; d < 8; ++d) f << (d ? "," : "") << ss[i][d];
				for (size_t d = 0; d < 10; ++d) f << "," << (i < aa.size() ? aa[i][d] : 0.0);
				f << "\n";
			}
		}
		const GlobalBodyPlanner::Averages a = planner.averages();
		std::cout << "SUMMARY calls " << a.calls << " successes " << a.successes << " avg_solve_time " << a.solve_time << " avg_path_length "
				  << a.path_length << " avg_vertices " << a.vertices_generated << std::endl;
	} catch (const std::exception &e) {
		std::fprintf(stderr, "gbp_plan: %s\n", e.what());
		return 1;
	}
	return 0;
}
