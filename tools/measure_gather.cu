// Terrain-probe microbenchmark: how fast can one SM fetch the 4 cells of a bilinear quad for scattered probes?
// Pattern of the validity walk: every thread owns a random centre on a 4096 x 4096 fp32 map, evaluates 9 probes within
// +-6 cells of it per "sub-state", then drifts by a cell.  Compared fetch paths:
//   ldg4     4 scalar __ldg per quad from the row-major map (what k_validate_refill does)
//   gather   1 tex2Dgather per quad from a block-linear cudaArray (texture units fetch the 2x2 footprint)
//   tiled4   4 scalar __ldg per quad from a linear map re-tiled into 4 x 8-cell (128-byte) tiles
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/measure_gather tools/measure_gather.cu && tools/measure_gather
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int N = 4096;

__device__ __forceinline__ unsigned lcg(unsigned &x) { x = x * 1664525u + 1013904223u; return x >> 8; }

template <int MODE>
__global__ void __launch_bounds__(128) k_probe(const float *__restrict__ map, cudaTextureObject_t tex, int iters, float *out, int check) {
	unsigned rng = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	int cx = 16 + lcg(rng) % (N - 32), cy = 16 + lcg(rng) % (N - 32);
	float acc = 0;
	for (int it = 0; it < iters; ++it) {
		int px[9], py[9];
#pragma unroll
		for (int p = 0; p < 9; ++p) {
			const unsigned r = lcg(rng);
			px[p] = cx + (int) (r % 13) - 6;
			py[p] = cy + (int) ((r >> 8) % 13) - 6;
		}
		float f[9][4];
#pragma unroll
		for (int p = 0; p < 9; ++p) {
			if (MODE == 0) {
				const float *q = map + px[p] * N + py[p];
				f[p][0] = __ldg(q); f[p][1] = __ldg(q + 1); f[p][2] = __ldg(q + N); f[p][3] = __ldg(q + N + 1);
			} else if (MODE == 1) {
				const float4 g = tex2Dgather<float4>(tex, (float) py[p] + 1.0f, (float) px[p] + 1.0f, 0);
				f[p][0] = g.w; f[p][1] = g.z; f[p][2] = g.x; f[p][3] = g.y;  // (x0,y0) (x1,y0) (x0,y1) (x1,y1), x = iy
			} else {
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					const int ix = px[p] + (c >> 1), iy = py[p] + (c & 1);
					const int tile = (ix >> 2) * (N >> 3) + (iy >> 3);
					f[p][c] = __ldg(map + tile * 32 + (ix & 3) * 8 + (iy & 7));
				}
			}
		}
#pragma unroll
		for (int p = 0; p < 9; ++p) acc += (f[p][0] - f[p][1]) + (f[p][2] - f[p][3]);
		if (check && it == 0 && blockIdx.x == 0 && threadIdx.x == 0)
			printf("mode %d probe (ix=%d, iy=%d): f11 %.0f f12 %.0f f21 %.0f f22 %.0f (expect %d %d %d %d)\n", MODE, px[0], py[0], f[0][0], f[0][1],
				   f[0][2], f[0][3], px[0] * N + py[0], px[0] * N + py[0] + 1, (px[0] + 1) * N + py[0], (px[0] + 1) * N + py[0] + 1);
		cx = min(max(cx + (int) (lcg(rng) % 3) - 1, 16), N - 17);
		cy = min(max(cy + (int) (lcg(rng) % 3) - 1, 16), N - 17);
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	std::vector<float> h((size_t) N * N), ht((size_t) N * N);
	for (int ix = 0; ix < N; ++ix)
		for (int iy = 0; iy < N; ++iy) {
			h[(size_t) ix * N + iy] = (float) (ix * N + iy);
			ht[(size_t) ((ix >> 2) * (N >> 3) + (iy >> 3)) * 32 + (ix & 3) * 8 + (iy & 7)] = (float) (ix * N + iy);
		}
	float *d_map, *d_tiled, *d_out;
	cudaMalloc(&d_map, h.size() * 4); cudaMalloc(&d_tiled, h.size() * 4);
	cudaMemcpy(d_map, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
	cudaMemcpy(d_tiled, ht.data(), h.size() * 4, cudaMemcpyHostToDevice);
	cudaArray_t arr;
	cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
	if (cudaMallocArray(&arr, &fd, N, N, cudaArrayTextureGather) != cudaSuccess) { printf("cudaMallocArray failed\n"); return 1; }
	cudaMemcpy2DToArray(arr, 0, 0, h.data(), N * 4, N * 4, N, cudaMemcpyHostToDevice);
	cudaResourceDesc rd = {};
	rd.resType = cudaResourceTypeArray;
	rd.res.array.array = arr;
	cudaTextureDesc td = {};
	td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
	td.filterMode = cudaFilterModePoint;
	td.readMode = cudaReadModeElementType;
	td.normalizedCoords = 0;
	cudaTextureObject_t tex;
	if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) != cudaSuccess) { printf("cudaCreateTextureObject failed\n"); return 1; }
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int threads = 128, iters = 512;
	printf("{\"gpu\": \"%s\", \"pattern\": \"9 probes within +-6 cells of a drifting per-thread centre, 4096^2 fp32 map, 128-thread CTAs\"", prop.name);
	for (int per_sm = 3; per_sm <= 12; per_sm *= 2) {
		const int blocks = prop.multiProcessorCount * per_sm;
		cudaMalloc(&d_out, sizeof(float) * blocks * threads);
		float ms[3];
		for (int mode = 0; mode < 3; ++mode) {
			const float *m = mode == 2 ? d_tiled : d_map;
			auto launch = [&](int it, int chk) {
				if (mode == 0) k_probe<0><<<blocks, threads>>>(m, tex, it, d_out, chk);
				else if (mode == 1) k_probe<1><<<blocks, threads>>>(m, tex, it, d_out, chk);
				else k_probe<2><<<blocks, threads>>>(m, tex, it, d_out, chk);
			};
			launch(8, per_sm == 3);
			cudaDeviceSynchronize();
			cudaEventRecord(e0); launch(iters, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
			cudaEventElapsedTime(&ms[mode], e0, e1);
		}
		const double quads = (double) blocks * threads * iters * 9;
		printf(",\n \"ctas_per_sm_%d\": {\"ldg4_quads_per_s\": %.4g, \"tex_gather_quads_per_s\": %.4g, \"tiled_ldg4_quads_per_s\": %.4g}", per_sm,
			   quads / (ms[0] * 1e-3), quads / (ms[1] * 1e-3), quads / (ms[2] * 1e-3));
		cudaFree(d_out);
	}
	printf("}\n");
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
	return 0;
}
