// Part 5 of libgbp_b200.so: the device-wide form of the batch planner (gbp_wide.cuh: one search at a time on a cooperative
// grid), in its own translation unit.  No C entry points: gbp_plan_batch* (gbp_capi_plan.cu) chooses the form.
#include "gbp_host.h"
#include "gbp_wide.cuh"

bool gbp_plan_wide_applies(const gbp_plan_params &P, int64_t nq) { return plan_wide_applies(P, nq); }

int gbp_plan_wide_launch(const TerrainView &Tv, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
						 const gbp_plan_params &P, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap, cudaStream_t st,
						 const PlanTreeDump &dump, std::string &err) {
#define GBP_WIDE_(M) return plan_wide_launch_kind<M>(Tv, nq, starts, goals, seed, query0, P, stats, path_states, path_actions, path_cap, st, dump, err)
	if (Tv.cell_f32) { if (Tv.uniform) GBP_WIDE_(MapF32U); else GBP_WIDE_(MapF32N); }
	else { if (Tv.uniform) GBP_WIDE_(MapF64U); else GBP_WIDE_(MapF64N); }
#undef GBP_WIDE_
}
