// Part 4 of libgbp_b200.so: the pipelined form of the batch planner (gbp_pipeline.cuh), in its own translation unit so
// that it compiles next to the megakernel instead of after it.  No C entry points: gbp_plan_batch* (gbp_capi_plan.cu)
// chooses between the forms.
#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include "gbp_host.h"
#include "gbp_pipeline.cuh"

bool gbp_plan_pipe_applies(const TerrainView &Tv, const gbp_plan_params &P, int64_t nq) { return plan_pipe_applies(Tv, P, nq); }

// pool of host-side pipeline resources (streams, events, a pinned word pair): per device, grown on demand, never shrunk
namespace {
std::mutex pool_mutex;
std::vector<std::pair<int, PipeHostRes *>> pool;
PipeHostRes *res_acquire(int dev, std::string &err) {
	{
		std::lock_guard<std::mutex> lock(pool_mutex);
		for (size_t i = 0; i < pool.size(); ++i)
			if (pool[i].first == dev) { PipeHostRes *r = pool[i].second; pool.erase(pool.begin() + i); return r; }
	}
	PipeHostRes *r = new PipeHostRes();
	const cudaError_t e = r->create();
	if (e != cudaSuccess) { err = std::string("pipelined planner: ") + cudaGetErrorString(e); return nullptr; }  // (partially created: left to the process)
	return r;
}
void res_release(int dev, PipeHostRes *r) {
	std::lock_guard<std::mutex> lock(pool_mutex);
	pool.emplace_back(dev, r);
}
int run_kind(const TerrainView &Tv, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0, const gbp_plan_params &P,
			 gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap, cudaStream_t st, const PlanTreeDump &dump,
			 PipeHostRes &R, std::string &err) {
	if (Tv.cell_f32) return plan_pipe_launch_kind<MapF32U>(Tv, nq, starts, goals, seed, query0, P, stats, path_states, path_actions, path_cap, st, dump, R, err);
	return plan_pipe_launch_kind<MapF64U>(Tv, nq, starts, goals, seed, query0, P, stats, path_states, path_actions, path_cap, st, dump, R, err);
}
}  // namespace

// A round of the pipeline is a chain of dependent kernels, none of which fills the chip on its own (prep and select are
// latency-bound, the walk ends in a tail of long candidates).  GBP_PIPE_GROUPS=n splits a batch into n GROUPS of queries,
// each an independent pipeline driven by its own host thread on its own streams, so that the kernels of one group could
// fill the gaps of the others (queries are independent: results do not depend on the split; tests run 3 groups).  Measured
// on configs[4] (65,536 queries): 1 group 0.344 s, 2 groups 0.420 s, 3 groups 0.559 s, 4 groups 0.577 s — a group's
// latency-critical small kernels queue behind the other groups' walks, which occupy every SM.  The default is ONE group.
int gbp_plan_pipe_launch(const TerrainView &Tv, int64_t nq, const double *starts, const double *goals, uint64_t seed, uint64_t query0,
						 const gbp_plan_params &P, gbp_plan_stats *stats, double *path_states, double *path_actions, int path_cap, cudaStream_t st,
						 const PlanTreeDump &dump, std::string &err) {
	int dev = 0;
	cudaGetDevice(&dev);
	int groups = 1;
	if (const char *g = getenv("GBP_PIPE_GROUPS")) groups = atoi(g);
	if (groups > 8) groups = 8;
	if (groups > nq) groups = (int) nq;
	if (groups < 1) groups = 1;
	std::vector<PipeHostRes *> R(groups, nullptr);
	for (int g = 0; g < groups; ++g)
		if (!(R[g] = res_acquire(dev, err))) {
			for (int k = 0; k < g; ++k) res_release(dev, R[k]);
			return GBP_E_CUDA;
		}
	int rc = GBP_OK;
	if (groups == 1) {
		// the arena holds both trees of every query at full capacity (plus the round's segment rows): a batch larger than the
		// device can hold runs as consecutive sub-batches (queries are independent: same results)
		// (the device's total memory is read once: cudaMemGetInfo inside a call cost the timed region a sporadic 40 ms)
		static std::mutex total_mutex;
		static std::vector<std::pair<int, size_t>> totals;
		size_t total_b = 0;
		{
			std::lock_guard<std::mutex> lock(total_mutex);
			for (const auto &t : totals) if (t.first == dev) total_b = t.second;
			if (!total_b) {
				size_t free_b = 0;
				if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || !total_b) total_b = (size_t) 64 << 30;
				totals.emplace_back(dev, total_b);
			}
		}
		const size_t per_query = (size_t) 2 * (size_t) P.max_vertices * (20 * 8 + 3 * 4) + 16 * (PIPE_ROW * 8 + 32) + PIPE_DEPTH * (PIPE_ROW * 8 + 48) +
								 2 * PIPE_BATCH * 64 + 512;
		int64_t chunk = std::max<int64_t>(4096, (int64_t) (total_b / 10 * 4 / per_query));  // sub-batches of at most 40 % of the device
		if (const char *c = getenv("GBP_PIPE_CHUNK")) chunk = std::max(1, atoi(c));  // tests: queries per sub-batch
		for (int64_t lo = 0; lo < nq && rc == GBP_OK; lo += chunk) {
			const int64_t hi = std::min(nq, lo + chunk);
			PlanTreeDump d = dump;
			if (d.states) {
				const size_t rows = (size_t) lo * 2 * d.cap;
				d.states += rows * 8; d.actions += rows * 10; d.parent += rows; d.g += rows; d.y += rows;
			}
			rc = run_kind(Tv, hi - lo, starts + 8 * lo, goals + 8 * lo, seed, query0 + (uint64_t) lo, P, stats + lo,
						  path_states ? path_states + (size_t) lo * path_cap * 8 : nullptr, path_actions ? path_actions + (size_t) lo * path_cap * 10 : nullptr,
						  path_cap, st, d, *R[0], err);
		}
	} else {
		// group g runs queries [lo, hi) on its own main stream, after everything the caller queued on st; st then waits for all
		cudaEvent_t start = R[0]->done;  // borrowed as the fork event: group 0 runs on st itself and does not need its `done`
		cudaEventRecord(start, st);
		std::vector<int> rcs(groups, GBP_OK);
		std::vector<std::string> errs(groups);
		std::vector<std::thread> threads;
		auto body = [&](int g) {
			cudaSetDevice(dev);
			const int64_t lo = nq * g / groups, hi = nq * (g + 1) / groups;
			const cudaStream_t sg = g == 0 ? st : R[g]->main;
			if (g) cudaStreamWaitEvent(sg, start, 0);
			PlanTreeDump d = dump;
			if (d.states) {
				const size_t rows = (size_t) lo * 2 * d.cap;
				d.states += rows * 8; d.actions += rows * 10; d.parent += rows; d.g += rows; d.y += rows;
			}
			rcs[g] = run_kind(Tv, hi - lo, starts + 8 * lo, goals + 8 * lo, seed, query0 + (uint64_t) lo, P, stats + lo,
							  path_states ? path_states + (size_t) lo * path_cap * 8 : nullptr, path_actions ? path_actions + (size_t) lo * path_cap * 10 : nullptr,
							  path_cap, sg, d, *R[g], errs[g]);
			if (g) cudaEventRecord(R[g]->done, sg);
		};
		for (int g = 1; g < groups; ++g) threads.emplace_back(body, g);
		body(0);
		for (auto &t : threads) t.join();
		for (int g = 1; g < groups; ++g) cudaStreamWaitEvent(st, R[g]->done, 0);
		for (int g = 0; g < groups; ++g)
			if (rcs[g] != GBP_OK && rc == GBP_OK) { rc = rcs[g]; err = errs[g]; }
	}
	for (int g = 0; g < groups; ++g) res_release(dev, R[g]);
	return rc;
}
