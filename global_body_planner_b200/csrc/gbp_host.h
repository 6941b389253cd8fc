// Host-side plumbing shared by the translation units of libgbp_b200.so (gbp_capi.cu: handles, terrain, primitives, tree
// store; gbp_capi_validate.cu: pair checks; gbp_capi_plan.cu: extend / connect / batch planner).  Not part of the C ABI.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "gbp_kernels.cuh"

using namespace gbp;


std::string &gbp_err();  // message of the last failure on the calling thread (defined in gbp_capi.cu)
inline int fail(int code, const std::string &msg) { gbp_err() = msg; return code; }

#define CU(call)                                                                                          \
	do {                                                                                                  \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess) return fail(GBP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
	} while (0)

inline int sm_count() {
	static int sms = 0;
	if (!sms) {
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
			sms = 148;
	}
	return sms;
}

// The default memory pool of a device gives freed blocks back to the driver at the next synchronisation unless told
// otherwise; every call would then pay a driver allocation (measured: 5-45 ms spikes on the first call after a sync).
inline void keep_pool(void) {
	static thread_local int done_for = -1;
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
	cudaMemPool_t pool;
	if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
		unsigned long long keep = ~0ull;
		cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
	}
	done_for = dev;
}

// stream-ordered device scratch; the pool keeps freed blocks, so repeated calls do not hit the driver
struct Dev {
	void *p = nullptr;
	cudaStream_t st;
	explicit Dev(cudaStream_t s) : st(s) {}
	cudaError_t alloc(size_t bytes) { keep_pool(); return cudaMallocAsync(&p, bytes ? bytes : 1, st); }
	~Dev() { if (p) cudaFreeAsync(p, st); }
	template <class T> T *as() { return (T *) p; }
};

inline cudaStream_t lib_stream() {
	static thread_local cudaStream_t s = nullptr;
	if (!s) {
		if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) s = nullptr;
		keep_pool();
	}
	return s;
}

template <class T>
inline int upload(Dev &d, const T *host, size_t n, cudaStream_t st) {
	CU(d.alloc(n * sizeof(T)));
	if (n) CU(cudaMemcpyAsync(d.p, host, n * sizeof(T), cudaMemcpyHostToDevice, st));
	return GBP_OK;
}

// kernel instantiation by map kind: fp32 / fp64 cells x uniform / general axes
#define GBP_LAUNCH_(K, M, CFG, ST, ...) K<M><<<GBP_UNPACK CFG, 0, ST>>>(__VA_ARGS__)
#define GBP_UNPACK(...) __VA_ARGS__
#define GBP_DISPATCH(VIEW, K, CFG, ST, ...)                                                  \
	do {                                                                                     \
		if ((VIEW).cell_f32) {                                                               \
			if ((VIEW).uniform) GBP_LAUNCH_(K, MapF32U, CFG, ST, __VA_ARGS__);               \
			else GBP_LAUNCH_(K, MapF32N, CFG, ST, __VA_ARGS__);                              \
		} else {                                                                             \
			if ((VIEW).uniform) GBP_LAUNCH_(K, MapF64U, CFG, ST, __VA_ARGS__);               \
			else GBP_LAUNCH_(K, MapF64N, CFG, ST, __VA_ARGS__);                              \
		}                                                                                    \
	} while (0)

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned) ((n + threads - 1) / threads); }

// Staging pipeline of the HOST-pointer pair check: ring of device buffer sets, one stream each, kept by the
// terrain handle between calls (grow-only) so that a call costs copies + kernels, not allocations.
struct HostPipe {
	static constexpr int NBUF = 3;
	cudaStream_t st[NBUF] = {nullptr, nullptr, nullptr};
	char *in[NBUF] = {nullptr, nullptr, nullptr}, *out[NBUF] = {nullptr, nullptr, nullptr};
	int *redo[NBUF] = {nullptr, nullptr, nullptr};
	int64_t chunk = 0;  // candidates per buffer set
	cudaEvent_t ready = nullptr;
};

struct gbp_terrain {
	TerrainView view;
	HostPipe pipe;
	std::mutex pipe_mutex;                // serialises the host-pointer pair checks on this handle (they share `pipe` and d_cnt)
	double *d_x = nullptr, *d_y = nullptr;
	void *d_z = nullptr;
	void *d_n = nullptr;
	cudaArray_t z_arr = nullptr;          // block-linear copy of the fp32 height grid behind view.ztex (texture gathers of the walk)
	cudaTextureObject_t z_tex = 0;
	unsigned long long *d_cnt = nullptr;  // 6 work counters of the last validate launch
	size_t z_bytes = 0;                   // height grid bytes
	float l2_hit_ratio = 0.f;             // share of the grid that fits the persisting L2 carve-out (0 = no window)
	std::vector<double> hx, hy;
	int cell_bytes = 8;
};

struct gbp_tree {
	TreeView view;
	int *d_n = nullptr;
	double *d_v = nullptr, *d_act = nullptr, *d_g = nullptr, *d_y = nullptr;
	int *d_parent = nullptr;
	// scratch of extend/connect
	ExtendScratch S;
	int k_cap = 0;
	double *d_target = nullptr;
	unsigned *d_done = nullptr;  // CTA counter of k_extend_fused
	int *h_result = nullptr;     // pinned, mapped: the result words of an extend without a copy
};

