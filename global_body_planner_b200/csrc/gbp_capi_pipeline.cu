// Part 4 of libgbp_b200.so: the pipelined form of the batch planner (gbp_pipeline.cuh), in its own translation unit so
// that it compiles next to the megakernel instead of after it.  No C entry points: gbp_plan_batch* (gbp_capi_plan.cu)
// chooses between the two forms.
#include "gbp_host.h"
#include "gbp_pipeline.cuh"

bool gbp_plan_pipe_applies(const TerrainView &Tv, const gbp_plan_params &P, int64_t nq) { return plan_pipe_applies(Tv, P, nq); }

int gbp_plan_pipe_launch(const TerrainView &Tv, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
						 const gbp_plan_params &P, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap, cudaStream_t st,
						 const PlanTreeDump &dump, std::string &err) {
	if (Tv.cell_f32) return plan_pipe_launch_kind<MapF32U>(Tv, nq, starts, goals, seed, query0, P, stats, path_states, path_actions, path_cap, st, dump, err);
	return plan_pipe_launch_kind<MapF64U>(Tv, nq, starts, goals, seed, query0, P, stats, path_states, path_actions, path_cap, st, dump, err);
}
