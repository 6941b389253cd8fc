// extern "C" layer, part 2: isValidStateActionPair[Reverse] for batches (gbp_validate_pairs*, gbp_sample_validate*).
#include <chrono>
#include <vector>

#include "gbp_host.h"
#include "gbp_walk.cuh"
#include "gbp_sv.cuh"

extern "C" {

// ------------------------------------------------------------------------------ validate_pairs
// `scratch` / `scratch_cap`: redo list + recipe array of the mixed-precision walk ([cap] ints, an 8-byte counter at a 16-byte
// offset, [cap] double2) owned by the caller (the host pipeline keeps one per buffer set), or null: allocated for this call
// on its stream and released in stream order — concurrent calls on one terrain handle do not share scratch.
static int validate_dev_impl(const gbp_terrain *t, int64_t n, const double *states, const double *actions, const uint8_t *direction,
							 int adaptive, int variant, uint8_t *verdict, uint8_t *flags, double *s_new, double *t_new, void *stream,
							 bool zero_counters, int *scratch, size_t scratch_cap);
int gbp_validate_pairs_dev(const gbp_terrain *t, int64_t n, const double *states, const double *actions, const uint8_t *direction,
						   int adaptive, int variant, uint8_t *verdict, uint8_t *flags, double *s_new, double *t_new, void *stream) {
	return validate_dev_impl(t, n, states, actions, direction, adaptive, variant, verdict, flags, s_new, t_new, stream, true, nullptr, 0);
}
static int validate_dev_impl(const gbp_terrain *t, int64_t n, const double *states, const double *actions, const uint8_t *direction,
							 int adaptive, int variant, uint8_t *verdict, uint8_t *flags, double *s_new, double *t_new, void *stream,
							 bool zero_counters, int *scratch, size_t scratch_cap) {
	if (!t || n < 0 || (n && (!states || !actions || !direction || !verdict))) return fail(GBP_E_INVALID, "bad arguments");
	if (variant < 0 || variant > 5 || variant == 4) return fail(GBP_E_INVALID, "variant must be 0..3 or 5");
	const bool walk_only = variant == 5;
	if (walk_only) variant = 3;
	if (variant == 2 && adaptive) return fail(GBP_E_INVALID, "variant 2 (warp per action) supports the fixed step only");
	cudaStream_t st = (cudaStream_t) stream;
	Dev own(st);  // per-call scratch: released (in stream order) when the function returns, i.e. after k_pair_outputs is enqueued
	if (zero_counters) CU(cudaMemsetAsync(t->d_cnt, 0, 6 * sizeof(unsigned long long), st));
	if (n == 0) return GBP_OK;
	if (variant == 0) variant = 3;
	if (variant == 1) {
		GBP_DISPATCH(t->view, k_validate_thread, (blocks_for(n, 128), 128), st, t->view, n, states, actions, direction, adaptive, verdict, flags, s_new, t_new, t->d_cnt);
	} else if (variant == 2) {
		const int64_t warps_per_block = 4;
		unsigned grid = (unsigned) ((n + warps_per_block - 1) / warps_per_block);
		GBP_DISPATCH(t->view, k_validate_warp, (grid, 128), st, t->view, n, states, actions, direction, verdict, flags, s_new, t_new, t->d_cnt);
	} else {
		// persistent-style geometry: a multiple of the SM count; each warp owns a contiguous range
		const int threads = RF_WARPS * 32, warps_per_block = RF_WARPS;
		int64_t max_warps = (int64_t) sm_count() * RF_WARPS * (t->view.mixed_ok ? GBP_WALK_CTAS : 2);  // one wave of resident warps
		int64_t per_warp = (n + max_warps - 1) / max_warps;
		if (per_warp < 64) per_warp = 64;
		per_warp = (per_warp + RF_CHUNK - 1) / RF_CHUNK * RF_CHUNK;  // chunks of the TMA ring are 32-aligned
		if ((((uintptr_t) states) | ((uintptr_t) actions) | ((uintptr_t) direction) | ((uintptr_t) s_new)) & 15)
			return fail(GBP_E_INVALID, "states/actions/direction/s_new must be 16-byte aligned (TMA bulk copies)");
		int64_t warps = (n + per_warp - 1) / per_warp;
		unsigned grid = (unsigned) ((warps + warps_per_block - 1) / warps_per_block);
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
		cudaLaunchAttribute attr[1];
		cfg.attrs = attr; cfg.numAttrs = 0;
		if (t->l2_hit_ratio > 0.f && !getenv("GBP_NO_L2_WINDOW")) {
			attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
			attr[0].val.accessPolicyWindow.base_ptr = t->d_z;
			size_t win = t->z_bytes;
			int dev = 0, max_window = 0;
			cudaGetDevice(&dev); cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
			if (win > (size_t) max_window) win = (size_t) max_window;
			attr[0].val.accessPolicyWindow.num_bytes = win;
			attr[0].val.accessPolicyWindow.hitRatio = t->l2_hit_ratio;
			attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
			attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
			cfg.numAttrs = 1;
		}
		double2 *recipe = nullptr;
		if (t->view.mixed_ok) {
			// mixed-precision walk (k_walk_mixed, gbp_walk.cuh) + fp64 redo pass over the candidates it could not decide
			if (n > 0x7fffffff) return fail(GBP_E_INVALID, "at most 2^31-1 candidates per call");
			size_t cap = scratch_cap;
			int *redo = scratch;
			if (!redo || cap < (size_t) n) {
				cap = ((size_t) n + 1023) / 1024 * 1024;
				CU(own.alloc(cap * sizeof(int) + 16 + cap * sizeof(double2)));  // freed in stream order when `own` goes out of scope
				redo = own.as<int>();
			}
			unsigned long long *redo_count = (unsigned long long *) (redo + cap);
			// {tau, kind} per candidate between the walk and k_pair_outputs: compact side array, except when the caller
			// finishes the outputs itself (variant 5: the recipes stay in s_new[i][0..1])
			recipe = (s_new && !walk_only) ? (double2 *) ((char *) redo + cap * sizeof(int) + 16) : nullptr;
			CU(cudaMemsetAsync(redo_count, 0, sizeof(unsigned long long), st));
#define GBP_WALK_(TEX, AD) CU(cudaLaunchKernelEx(&cfg, k_walk_mixed<TEX, AD>, t->view, (int) n, (int) per_warp, states, actions, direction, verdict, flags, s_new, t_new, t->d_cnt, redo, redo_count, recipe))
			if (t->view.ztex) { if (adaptive) GBP_WALK_(true, true); else GBP_WALK_(true, false); }
			else { if (adaptive) GBP_WALK_(false, true); else GBP_WALK_(false, false); }
#undef GBP_WALK_
			if (t->view.cell_f32)
				k_validate_redo<MapF32U><<<sm_count() * 4, 128, 0, st>>>(t->view, redo, redo_count, states, actions, direction, adaptive, verdict,
																	   flags, s_new, t_new, t->d_cnt, recipe);
			else
				k_validate_redo<MapF64U><<<sm_count() * 4, 128, 0, st>>>(t->view, redo, redo_count, states, actions, direction, adaptive, verdict,
																	   flags, s_new, t_new, t->d_cnt, recipe);
		} else {
#define GBP_WALK_(M) CU(cudaLaunchKernelEx(&cfg, k_validate_refill<M>, t->view, n, per_warp, states, actions, direction, adaptive, verdict, flags, s_new, t_new, t->d_cnt))
			if (t->view.cell_f32) { if (t->view.uniform) GBP_WALK_(MapF32U); else GBP_WALK_(MapF32N); }
			else { if (t->view.uniform) GBP_WALK_(MapF64U); else GBP_WALK_(MapF64N); }
#undef GBP_WALK_
		}
		if (s_new && !walk_only) k_pair_outputs<<<blocks_for(n, 256), 256, 0, st>>>(n, states, actions, s_new, recipe);
	}
	CU(cudaGetLastError());
	return GBP_OK;
}

int gbp_pair_outputs_dev(int64_t n, const double *states, const double *actions, double *s_new, void *stream) {
	if (n < 0 || (n && (!states || !actions || !s_new))) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	if ((((uintptr_t) states) | ((uintptr_t) actions) | ((uintptr_t) s_new)) & 15)
		return fail(GBP_E_INVALID, "states/actions/s_new must be 16-byte aligned (TMA bulk copies)");
	k_pair_outputs<<<blocks_for(n, 256), 256, 0, (cudaStream_t) stream>>>(n, states, actions, s_new, (const double2 *) nullptr);
	CU(cudaGetLastError());
	return GBP_OK;
}
int gbp_validate_counters(const gbp_terrain *t, int64_t counters6[6]) {
	if (!t || !counters6) return fail(GBP_E_INVALID, "bad arguments");
	CU(cudaDeviceSynchronize());
	unsigned long long h[6];
	CU(cudaMemcpy(h, t->d_cnt, sizeof h, cudaMemcpyDeviceToHost));
	for (int i = 0; i < 6; ++i) counters6[i] = (int64_t) h[i];
	return GBP_OK;
}

// HOST buffers: chunks of candidates flow through the handle's ring of device buffer sets, each on its own
// stream, so the H2D copy of chunk c+1, the kernels of chunk c and the D2H copy of chunk c-1 overlap (PCIe is
// full duplex).  Nothing in the loop blocks the host; work counters accumulate on the device.
int gbp_validate_pairs(const gbp_terrain *t, int64_t n, const double *states, const double *actions, const uint8_t *direction,
					   int adaptive, int variant, uint8_t *verdict, uint8_t *flags, double *s_new, double *t_new) {
	if (!t || n < 0 || (n && (!states || !actions || !direction || !verdict))) return fail(GBP_E_INVALID, "bad arguments");
	if (n == 0) return GBP_OK;
	constexpr int NBUF = HostPipe::NBUF;
	gbp_terrain *tm = const_cast<gbp_terrain *>(t);
	std::lock_guard<std::mutex> lock(tm->pipe_mutex);  // the staging ring and the work counters belong to the handle: one host-pointer call at a time
	HostPipe &P = tm->pipe;
	const int64_t want = n < (1 << 19) ? (n + 1023) / 1024 * 1024 : (1 << 19);
	if (P.chunk < want) {
		for (int k = 0; k < NBUF; ++k) {
			if (P.st[k]) CU(cudaStreamSynchronize(P.st[k]));
			cudaFree(P.in[k]); cudaFree(P.out[k]); cudaFree(P.redo[k]);
			P.in[k] = P.out[k] = nullptr; P.redo[k] = nullptr;
		}
		P.chunk = 0;
		for (int k = 0; k < NBUF; ++k) {
			if (!P.st[k]) CU(cudaStreamCreateWithFlags(&P.st[k], cudaStreamNonBlocking));
			CU(cudaMalloc(&P.in[k], (size_t) want * (64 + 80 + 1)));
			CU(cudaMalloc(&P.out[k], (size_t) want * (64 + 8 + 1 + 1)));
			CU(cudaMalloc((void **) &P.redo[k], (size_t) want * sizeof(int) + 16 + (size_t) want * sizeof(double2)));
		}
		if (!P.ready) CU(cudaEventCreateWithFlags(&P.ready, cudaEventDisableTiming));
		P.chunk = want;
	}
	const int64_t chunk = P.chunk;
	// the counters are zeroed on the first pipeline stream; the other streams (non-blocking: nothing else orders them) wait for it
	CU(cudaMemsetAsync(t->d_cnt, 0, 6 * sizeof(unsigned long long), P.st[0]));
	CU(cudaEventRecord(P.ready, P.st[0]));
	for (int k = 1; k < NBUF; ++k) CU(cudaStreamWaitEvent(P.st[k], P.ready, 0));
	int64_t ci = 0;
	int rc = GBP_OK;
	cudaError_t e = cudaSuccess;
	for (int64_t off = 0; off < n && rc == GBP_OK && e == cudaSuccess; off += chunk, ++ci) {
		const int k = (int) (ci % NBUF);
		cudaStream_t st = P.st[k];
		const int64_t m = n - off < chunk ? n - off : chunk;
		double *d_s = (double *) P.in[k], *d_a = (double *) (P.in[k] + (size_t) chunk * 64);
		uint8_t *d_d = (uint8_t *) (P.in[k] + (size_t) chunk * 144);
		double *d_sn = (double *) P.out[k], *d_tn = (double *) (P.out[k] + (size_t) chunk * 64);
		uint8_t *d_v = (uint8_t *) (P.out[k] + (size_t) chunk * 72), *d_f = d_v + chunk;
		e = cudaMemcpyAsync(d_s, states + 8 * off, (size_t) m * 64, cudaMemcpyHostToDevice, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(d_a, actions + 10 * off, (size_t) m * 80, cudaMemcpyHostToDevice, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(d_d, direction + off, (size_t) m, cudaMemcpyHostToDevice, st);
		if (e != cudaSuccess) break;
		rc = validate_dev_impl(t, m, d_s, d_a, d_d, adaptive, variant, d_v, flags ? d_f : nullptr, s_new ? d_sn : nullptr, t_new ? d_tn : nullptr, st,
							   false, P.redo[k], (size_t) chunk);
		if (rc) break;
		e = cudaMemcpyAsync(verdict + off, d_v, (size_t) m, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess && flags) e = cudaMemcpyAsync(flags + off, d_f, (size_t) m, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess && s_new) e = cudaMemcpyAsync(s_new + 8 * off, d_sn, (size_t) m * 64, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess && t_new) e = cudaMemcpyAsync(t_new + off, d_tn, (size_t) m * 8, cudaMemcpyDeviceToHost, st);
	}
	// common epilogue, also after a failure: no copy into the caller's buffers stays in flight
	for (int k = 0; k < NBUF; ++k) {
		const cudaError_t es = cudaStreamSynchronize(P.st[k]);
		if (es != cudaSuccess && e == cudaSuccess) e = es;
	}
	if (rc == GBP_OK && e != cudaSuccess) rc = fail(GBP_E_CUDA, std::string("validate pipeline: ") + cudaGetErrorString(e));
	return rc;
}

// ------------------------------------------------------------------- sample + validate (narrow wire)
struct gbp_states {
	double *d = nullptr;
	int64_t rows = 0;
};

int gbp_states_create(int64_t rows, const double *states, gbp_states **out) {
	if (!out) return fail(GBP_E_INVALID, "out is NULL");
	*out = nullptr;
	if (rows < 1 || !states) return fail(GBP_E_INVALID, "a state table needs at least one row");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(GBP_E_CUDA, "no CUDA device (this library has no CPU fallback)");
	gbp_states *S = new gbp_states();
	cudaError_t e = cudaMalloc(&S->d, (size_t) rows * 64);
	if (e == cudaSuccess) e = cudaMemcpy(S->d, states, (size_t) rows * 64, cudaMemcpyHostToDevice);
	if (e != cudaSuccess) { cudaFree(S->d); delete S; return fail(GBP_E_CUDA, std::string("state table: ") + cudaGetErrorString(e)); }
	S->rows = rows;
	*out = S;
	return GBP_OK;
}
void gbp_states_destroy(gbp_states *s) {
	if (!s) return;
	cudaFree(s->d);
	delete s;
}
int gbp_states_rows(const gbp_states *s, int64_t *rows) {
	if (!s || !rows) return fail(GBP_E_INVALID, "bad arguments");
	*rows = s->rows;
	return GBP_OK;
}

static SvParams sv_device_params(const gbp_sv_params &p, const double *table, int64_t rows, const int *idx, const uint8_t *dir, int64_t off) {
	SvParams P;
	P.table = table;
	P.rows = (long long) rows;
	P.state_idx = idx ? idx + off : nullptr;
	P.dir = dir ? dir + off : nullptr;
	P.dir_packed = p.direction_in_row ? 1 : 0;
	P.row0 = (long long) (p.row0 + off);
	P.dir0 = p.direction0;
	P.seed = p.seed; P.stream = p.stream; P.idx0 = p.idx0 + (uint64_t) off;
	for (int k = 0; k < 3; ++k) P.normal[k] = p.normal[k];
	P.states_valid = p.start_states_valid;
	P.dir_sampling = p.action_direction_sampling;
	P.dir_thresh = p.action_direction_threshold;
	for (int k = 0; k < 8; ++k) P.target[k] = p.target[k];
	return P;
}
// the walk (+ fp64 redo pass) or, on terrains without the mixed-precision evaluator, the general fp64 kernel, over
// candidates [off, off + m) of a call; `bits`, `flags` are the call's arrays (off is a multiple of 32), `redo` = m ints
static int sv_launch_range(const gbp_terrain *t, const gbp_sv_params &p, const double *table, int64_t rows, const int *idx, const uint8_t *dir, int64_t off,
						   int64_t m, unsigned *bits, uint8_t *flags, unsigned long long *cnt, int *redo, unsigned long long *redo_count,
						   cudaStream_t st, const unsigned *arrived = nullptr, unsigned long long *work = nullptr) {
	const SvParams P = sv_device_params(p, table, rows, idx, dir, off);
	unsigned *b = bits + off / 32;
	uint8_t *f = flags ? flags + off : nullptr;
	if (t->view.mixed_ok) {
		const int64_t max_warps = (int64_t) sm_count() * RF_WARPS * GBP_WALK_CTAS;  // one wave of resident warps
		int64_t per_warp = (m + max_warps - 1) / max_warps;
		if (per_warp < 64) per_warp = 64;
		per_warp = (per_warp + 31) / 32 * 32;
		const int64_t warps = (m + per_warp - 1) / per_warp;
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3(arrived ? (unsigned) (sm_count() * GBP_WALK_CTAS) : (unsigned) ((warps + RF_WARPS - 1) / RF_WARPS));
		cfg.blockDim = dim3(RF_WARPS * 32); cfg.dynamicSmemBytes = 0; cfg.stream = st;
		cudaLaunchAttribute attr[1];
		cfg.attrs = attr; cfg.numAttrs = 0;
		if (t->l2_hit_ratio > 0.f && !getenv("GBP_NO_L2_WINDOW")) {
			attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
			attr[0].val.accessPolicyWindow.base_ptr = t->d_z;
			size_t win = t->z_bytes;
			int dev = 0, max_window = 0;
			cudaGetDevice(&dev); cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
			if (win > (size_t) max_window) win = (size_t) max_window;
			attr[0].val.accessPolicyWindow.num_bytes = win;
			attr[0].val.accessPolicyWindow.hitRatio = t->l2_hit_ratio;
			attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
			attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
			cfg.numAttrs = 1;
		}
#define GBP_SV_(TEX, AD) do { if (arrived) CU(cudaLaunchKernelEx(&cfg, k_walk_sv_stream<TEX, AD>, t->view, P, (int) m, arrived, work, b, f, cnt, redo, redo_count)); \
							  else CU(cudaLaunchKernelEx(&cfg, k_walk_sv<TEX, AD>, t->view, P, (int) m, (int) per_warp, b, f, cnt, redo, redo_count)); } while (0)
		if (t->view.ztex) { if (p.adaptive) GBP_SV_(true, true); else GBP_SV_(true, false); }
		else { if (p.adaptive) GBP_SV_(false, true); else GBP_SV_(false, false); }
#undef GBP_SV_
		if (t->view.cell_f32) k_sv_fp64<MapF32U><<<sm_count() * 4, 128, 0, st>>>(t->view, P, m, redo, redo_count, p.adaptive, b, f, cnt);
		else k_sv_fp64<MapF64U><<<sm_count() * 4, 128, 0, st>>>(t->view, P, m, redo, redo_count, p.adaptive, b, f, cnt);
	} else {
		const unsigned grid = (unsigned) std::min<int64_t>((m + 127) / 128, (int64_t) sm_count() * 64);
		GBP_DISPATCH(t->view, k_sv_fp64, (grid, 128), st, t->view, P, m, (const int *) nullptr, (const unsigned long long *) nullptr, p.adaptive, b, f, cnt);
	}
	CU(cudaGetLastError());
	return GBP_OK;
}
// verdict bits -> ascending list of valid candidates + result words
static int sv_compact(int64_t n, const unsigned *bits, unsigned long long *sums, int64_t cap, int *index, const unsigned long long *cnt,
					  long long *result, cudaStream_t st) {
	const int64_t nwords = (n + 31) / 32;
	int64_t nb = (nwords + 255) / 256;
	if (nb > 1024) nb = 1024;
	if (nb < 1) nb = 1;
	const int64_t span = (nwords + nb - 1) / nb;
	k_sv_count<<<(unsigned) nb, 256, 0, st>>>(bits, nwords, span, sums);
	k_sv_list<<<(unsigned) nb, 256, 0, st>>>(bits, nwords, span, sums, cap, index, cnt, result);
	CU(cudaGetLastError());
	return GBP_OK;
}
static int sv_check(const gbp_terrain *t, int64_t n, const gbp_sv_params *p, const void *bits, int64_t valid_cap, const void *idx, const void *dir) {
	if (!t || !p || n < 0 || (n && !bits) || valid_cap < 0) return fail(GBP_E_INVALID, "bad arguments");
	if (n > 0x7fffffff) return fail(GBP_E_INVALID, "at most 2^31-1 candidates per call");
	if (p->direction0 != GBP_FORWARD && p->direction0 != GBP_REVERSE) return fail(GBP_E_INVALID, "direction0 must be GBP_FORWARD or GBP_REVERSE");
	if (p->direction_in_row && (!idx || dir)) return fail(GBP_E_INVALID, "direction_in_row needs state_idx and a NULL direction array");
	return GBP_OK;
}

int gbp_sample_validate_dev(const gbp_terrain *t, const double *states_dev, int64_t table_rows, int64_t n, const int32_t *state_idx_dev, const uint8_t *direction_dev,
							const gbp_sv_params *p, uint32_t *bits, uint8_t *flags, int64_t valid_cap, int32_t *valid_index, double *valid_s_new,
							double *valid_t_new, double *valid_action, int64_t *result_dev, void *stream) {
	int rc;
	if ((rc = sv_check(t, n, p, bits, valid_cap, state_idx_dev, direction_dev))) return rc;
	if (!states_dev || !result_dev || table_rows < 1) return fail(GBP_E_INVALID, "states_dev (table_rows >= 1) and result_dev are required");
	if (((uintptr_t) states_dev) & 15) return fail(GBP_E_INVALID, "the state table must be 16-byte aligned");
	if ((valid_s_new || valid_t_new || valid_action) && !valid_index) return fail(GBP_E_INVALID, "valid rows need valid_index");
	cudaStream_t st = (cudaStream_t) stream;
	const int64_t nwords = (n + 31) / 32;
	Dev scratch(st);  // [cnt 8 | redo_count 1 | sums 1024] u64, then the redo list
	CU(scratch.alloc((9 + 1024) * sizeof(unsigned long long) + (size_t) (n ? n : 1) * sizeof(int)));
	unsigned long long *cnt = scratch.as<unsigned long long>(), *redo_count = cnt + 8, *sums = cnt + 9;
	int *redo = (int *) (sums + 1024);
	CU(cudaMemsetAsync(cnt, 0, 9 * sizeof(unsigned long long), st));
	if (nwords) CU(cudaMemsetAsync(bits, 0, (size_t) nwords * 4, st));
	if (n && (rc = sv_launch_range(t, *p, states_dev, table_rows, state_idx_dev, direction_dev, 0, n, bits, flags, cnt, redo, redo_count, st))) return rc;
	if ((rc = sv_compact(n, bits, sums, valid_index ? valid_cap : 0, valid_index, cnt, (long long *) result_dev, st))) return rc;
	if (valid_index && valid_cap > 0 && (valid_s_new || valid_t_new || valid_action)) {
		const SvParams P = sv_device_params(*p, states_dev, table_rows, state_idx_dev, direction_dev, 0);
		k_sv_outputs<<<sm_count() * 2, 128, 0, st>>>(P, (const long long *) result_dev, valid_cap, valid_index, valid_s_new, valid_t_new, valid_action);
		CU(cudaGetLastError());
	}
	return GBP_OK;
}

// the walk alone (k_walk_sv + the fp64 redo pass, or the general fp64 kernel): verdict bits and the 8 counter words
// {k, L, NaN probes, OOG, NEAR, valid, rows out of range, 0}.  For measurements: bench.py times the dominant kernel this way.
int gbp_sample_validate_walk_dev(const gbp_terrain *t, const double *states_dev, int64_t table_rows, int64_t n, const int32_t *state_idx_dev,
								 const uint8_t *direction_dev, const gbp_sv_params *p, uint32_t *bits, int64_t *counters8_dev, void *stream) {
	int rc;
	if ((rc = sv_check(t, n, p, bits, 0, state_idx_dev, direction_dev))) return rc;
	if (!states_dev || !counters8_dev || table_rows < 1) return fail(GBP_E_INVALID, "states_dev (table_rows >= 1) and counters8_dev are required");
	if (((uintptr_t) states_dev) & 15) return fail(GBP_E_INVALID, "the state table must be 16-byte aligned");
	cudaStream_t st = (cudaStream_t) stream;
	Dev scratch(st);
	CU(scratch.alloc(sizeof(unsigned long long) + (size_t) (n ? n : 1) * sizeof(int)));
	unsigned long long *redo_count = scratch.as<unsigned long long>();
	CU(cudaMemsetAsync(redo_count, 0, sizeof(unsigned long long), st));
	CU(cudaMemsetAsync(counters8_dev, 0, 8 * sizeof(long long), st));
	if (n) CU(cudaMemsetAsync(bits, 0, (size_t) ((n + 31) / 32) * 4, st));
	if (n && (rc = sv_launch_range(t, *p, states_dev, table_rows, state_idx_dev, direction_dev, 0, n, bits, nullptr, (unsigned long long *) counters8_dev,
								   (int *) (redo_count + 1), redo_count, st)))
		return rc;
	return GBP_OK;
}

// HOST buffers.  The per-candidate inputs (4-byte rows, direction bytes) are copied in chunks on a copy stream while the
// compute stream walks the chunks already there; the verdict bits come back as one copy; the rows of the valid
// candidates are produced and copied once their number is known.
int gbp_sample_validate(const gbp_terrain *t, const gbp_states *table, int64_t n, const int32_t *state_idx, const uint8_t *direction,
						const gbp_sv_params *p, uint32_t *verdict_bits, uint8_t *flags, int64_t valid_cap, int32_t *valid_index,
						double *valid_s_new, double *valid_t_new, double *valid_action, gbp_sv_result *result) {
	int rc;
	if ((rc = sv_check(t, n, p, verdict_bits, valid_cap, state_idx, direction))) return rc;
	if (!table || !result) return fail(GBP_E_INVALID, "table and result are required");
	if ((valid_s_new || valid_t_new || valid_action) && !valid_index) return fail(GBP_E_INVALID, "valid rows need valid_index");
	memset(result, 0, sizeof *result);
	if (n == 0) return GBP_OK;
	if (!state_idx && (p->row0 < 0 || p->row0 + n > table->rows)) return fail(GBP_E_INVALID, "row0 + n exceeds the state table");
	cudaStream_t st = lib_stream();
	const bool sv_trace = getenv("GBP_SV_TRACE") != nullptr;  // host timestamps of the call's phases, to stderr
	const auto sv_t0 = std::chrono::steady_clock::now();
	auto sv_us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - sv_t0).count(); };
	double sv_t_alloc = 0, sv_t_issued = 0, sv_t_walks = 0;
	static thread_local cudaStream_t cs = nullptr;
	static thread_local cudaEvent_t ev = nullptr;
	static thread_local long long *h_result = nullptr;  // pinned
	if (!cs) CU(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
	if (!ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
	if (!h_result) CU(cudaHostAlloc((void **) &h_result, 8 * sizeof(long long), cudaHostAllocDefault));
	// Inputs travel in chunks on a copy stream while the compute stream walks what is already there.  On terrains with the
	// mixed-precision walk the whole call is ONE launch (k_walk_sv_stream) whose warps wait for the block of inputs they
	// need: each chunk's copies are followed by a copy that sets its arrival words.  Elsewhere: one launch per chunk.
	// (not under CUDA_LAUNCH_BLOCKING: the launch would return only after the kernel, which waits for copies issued after it)
	const bool streamed = t->view.mixed_ok && (state_idx || direction) && n >= (2ll << SV_BLOCK_SHIFT) && !getenv("GBP_SV_NO_STREAM") &&
						  !getenv("CUDA_LAUNCH_BLOCKING");
	const int64_t BL = 1ll << SV_BLOCK_SHIFT, nblocks = (n + BL - 1) / BL;
	std::vector<int64_t> chunk_off;  // streamed: 1, 1, 2, 4, 4 ... arrival blocks per chunk (the first copy is the only exposed one)
	int64_t max_chunk_blocks = 4;
	if (const char *c = getenv("GBP_SV_CHUNK_BLOCKS")) max_chunk_blocks = std::min(64, std::max(1, atoi(c)));  // A/B runs: larger copies when many ranks share the host
	if (streamed) for (int64_t b = 0, w = 1, k = 0; b < nblocks; b += w, ++k, w = std::min<int64_t>(k < 2 ? 1 : 2 * w, max_chunk_blocks)) chunk_off.push_back(b * BL);
	else for (int64_t off = 0; off < n; off += 1 << 21) chunk_off.push_back(off);
	chunk_off.push_back(n);
	const int64_t nwords = (n + 31) / 32, nch = (int64_t) chunk_off.size() - 1;
	const int64_t cap = !valid_index ? 0 : (valid_cap < n ? valid_cap : n);
	static thread_local unsigned *h_ones = nullptr;  // pinned source of the arrival words
	if (!h_ones) {
		CU(cudaHostAlloc((void **) &h_ones, 64 * sizeof(unsigned), cudaHostAllocDefault));
		for (int k = 0; k < 64; ++k) h_ones[k] = 1u;
	}
	Dev scratch(st), d_idx(st), d_dir(st), d_bits(st), d_flags(st), d_index(st), d_res(st), d_sn(st), d_tn(st), d_act(st);
	const size_t n_words64 = 8 + 1024 + (size_t) nch + 1 + (size_t) (nblocks + 1) / 2;  // counters, block sums, redo counts, work counter, arrival words
	CU(scratch.alloc(n_words64 * sizeof(unsigned long long) + (size_t) n * sizeof(int)));
	unsigned long long *cnt = scratch.as<unsigned long long>(), *sums = cnt + 8, *redo_counts = sums + 1024, *work = redo_counts + nch;
	unsigned *arrived = (unsigned *) (work + 1);
	int *redo = (int *) (cnt + n_words64);
	if (state_idx) CU(d_idx.alloc((size_t) n * 4));
	if (direction) CU(d_dir.alloc((size_t) n));
	CU(d_bits.alloc((size_t) nwords * 4));
	if (flags) CU(d_flags.alloc((size_t) n));
	if (cap) CU(d_index.alloc((size_t) cap * 4));
	CU(d_res.alloc(8 * sizeof(long long)));
	CU(cudaMemsetAsync(cnt, 0, n_words64 * sizeof(unsigned long long), st));
	CU(cudaMemsetAsync(d_bits.p, 0, (size_t) nwords * 4, st));
	CU(cudaEventRecord(ev, st));
	CU(cudaStreamWaitEvent(cs, ev, 0));  // the copy stream may touch the buffers once they exist in stream order
	rc = GBP_OK;
	sv_t_alloc = sv_us();
	if (streamed)
		rc = sv_launch_range(t, *p, table->d, table->rows, state_idx ? d_idx.as<int>() : nullptr, direction ? d_dir.as<uint8_t>() : nullptr, 0, n,
							 d_bits.as<unsigned>(), flags ? d_flags.as<uint8_t>() : nullptr, cnt, redo, redo_counts, st, arrived, work);
	cudaError_t ce = cudaSuccess;
	for (int64_t c = 0; c < nch && rc == GBP_OK; ++c) {
		const int64_t off = chunk_off[c], m = chunk_off[c + 1] - off;
		cudaError_t e = cudaSuccess;
		if (state_idx) e = cudaMemcpyAsync(d_idx.as<int>() + off, state_idx + off, (size_t) m * 4, cudaMemcpyHostToDevice, cs);
		if (e == cudaSuccess && direction) e = cudaMemcpyAsync(d_dir.as<uint8_t>() + off, direction + off, (size_t) m, cudaMemcpyHostToDevice, cs);
		if (streamed) {
			// a failed copy must not leave the running kernel waiting: its blocks are released all the same, the call fails below
			const cudaError_t e1 = cudaMemcpyAsync(arrived + off / BL, h_ones, (size_t) ((m + BL - 1) / BL) * sizeof(unsigned), cudaMemcpyHostToDevice, cs);
			if (e == cudaSuccess) e = e1;
			if (e != cudaSuccess) {
				cudaMemsetAsync(arrived, 0xff, (size_t) nblocks * sizeof(unsigned), cs);
				ce = e;
				break;
			}
			continue;
		}
		if (e == cudaSuccess) e = cudaEventRecord(ev, cs);
		if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev, 0);
		if (e != cudaSuccess) { rc = fail(GBP_E_CUDA, std::string("sample_validate copy: ") + cudaGetErrorString(e)); break; }
		rc = sv_launch_range(t, *p, table->d, table->rows, state_idx ? d_idx.as<int>() : nullptr, direction ? d_dir.as<uint8_t>() : nullptr, off, m,
							 d_bits.as<unsigned>(), flags ? d_flags.as<uint8_t>() : nullptr, cnt, redo + off, redo_counts + c, st);
	}
	if (ce != cudaSuccess) rc = fail(GBP_E_CUDA, std::string("sample_validate copy: ") + cudaGetErrorString(ce));
	if (rc == GBP_OK) rc = sv_compact(n, d_bits.as<unsigned>(), sums, cap, cap ? d_index.as<int>() : nullptr, cnt, d_res.as<long long>(), st);
	cudaError_t e = cudaSuccess;
	if (rc == GBP_OK) {
		e = cudaMemcpyAsync(h_result, d_res.p, 8 * sizeof(long long), cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaMemcpyAsync(verdict_bits, d_bits.p, (size_t) nwords * 4, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess && flags) e = cudaMemcpyAsync(flags, d_flags.p, (size_t) n, cudaMemcpyDeviceToHost, st);
	}
	sv_t_issued = sv_us();
	cudaError_t e2 = cudaStreamSynchronize(cs), e3 = cudaStreamSynchronize(st);  // also on failure: no copy of the caller's buffers stays in flight
	sv_t_walks = sv_us();
	if (rc != GBP_OK) return rc;
	if (e != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
		return fail(GBP_E_CUDA, std::string("sample_validate: ") + cudaGetErrorString(e != cudaSuccess ? e : (e2 != cudaSuccess ? e2 : e3)));
	memcpy(result, h_result, sizeof *result);
	// row numbers are checked where they are read (a host-side scan of 16 M indices costs more than the whole call): a row
	// outside the table was read as row 0 and counted; the call's results are void then
	if (result->reserved[1] > 0) return fail(GBP_E_CUDA, "sample_validate: input blocks did not reach the running kernel (copy stream stalled)");
	if (result->reserved[0] > 0) return fail(GBP_E_INVALID, "state_idx entry outside the state table");
	const int64_t rows = result->n_valid < cap ? result->n_valid : cap;
	if (rows > 0) {
		CU(cudaMemcpyAsync(valid_index, d_index.p, (size_t) rows * 4, cudaMemcpyDeviceToHost, st));
		if (valid_s_new || valid_t_new || valid_action) {
			if (valid_s_new) CU(d_sn.alloc((size_t) rows * 64));
			if (valid_t_new) CU(d_tn.alloc((size_t) rows * 8));
			if (valid_action) CU(d_act.alloc((size_t) rows * 80));
			const SvParams P = sv_device_params(*p, table->d, table->rows, state_idx ? d_idx.as<int>() : nullptr, direction ? d_dir.as<uint8_t>() : nullptr, 0);
			k_sv_outputs<<<(unsigned) std::min<int64_t>((rows + 127) / 128, (int64_t) sm_count() * 2), 128, 0, st>>>(
				P, d_res.as<long long>(), rows, d_index.as<int>(), d_sn.as<double>(), d_tn.as<double>(), d_act.as<double>());
			CU(cudaGetLastError());
			if (valid_s_new) CU(cudaMemcpyAsync(valid_s_new, d_sn.p, (size_t) rows * 64, cudaMemcpyDeviceToHost, st));
			if (valid_t_new) CU(cudaMemcpyAsync(valid_t_new, d_tn.p, (size_t) rows * 8, cudaMemcpyDeviceToHost, st));
			if (valid_action) CU(cudaMemcpyAsync(valid_action, d_act.p, (size_t) rows * 80, cudaMemcpyDeviceToHost, st));
		}
		CU(cudaStreamSynchronize(st));
	}
	if (sv_trace)
		fprintf(stderr, "gbp_sample_validate trace (us): buffers %.0f, %d chunks issued %.0f, walks + compaction + verdict bits done %.0f, valid rows done %.0f\n",
				sv_t_alloc, (int) nch, sv_t_issued, sv_t_walks, sv_us());
	return GBP_OK;
}

}  // extern "C"
