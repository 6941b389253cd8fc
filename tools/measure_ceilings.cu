// Measures the ceilings SURVEY §8(d) asks for and MEASURED_PEAKS.json does not hold: fp64 / fp32 FMA issue rate,
// random 4-byte gather bandwidth out of L2 and out of L1 (sector granularity), on the GPU it runs on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/measure_ceilings tools/measure_ceilings.cu && tools/measure_ceilings
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <typename T>
__global__ void k_fma(T *out, int iters) {
	T a0 = threadIdx.x * (T) 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const T b = (T) 1.0000001, c = (T) 1e-7;
	for (int i = 0; i < iters; ++i) {
		a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
		a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// each thread issues `iters` dependent-address-free random loads (LCG over `mask`+1 floats), 4 in flight
__global__ void k_gather(const float *__restrict__ src, unsigned mask, int iters, float *out) {
	unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	float acc = 0;
	for (int i = 0; i < iters; i += 4) {
		unsigned i0 = x & mask; x = x * 1664525u + 1013904223u;
		unsigned i1 = x & mask; x = x * 1664525u + 1013904223u;
		unsigned i2 = x & mask; x = x * 1664525u + 1013904223u;
		unsigned i3 = x & mask; x = x * 1664525u + 1013904223u;
		acc += __ldg(src + i0) + __ldg(src + i1) + __ldg(src + i2) + __ldg(src + i3);
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int blocks = p.multiProcessorCount * 8, threads = 256;
	double *d64; float *d32;
	cudaMalloc(&d64, sizeof(double) * blocks * threads);
	cudaMalloc(&d32, sizeof(float) * blocks * threads * 2);
	const int iters = 20000;
	k_fma<double><<<blocks, threads>>>(d64, 100);
	cudaEventRecord(e0); k_fma<double><<<blocks, threads>>>(d64, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
	const double fp64_tf = 2.0 * 8 * iters * (double) blocks * threads / (time_ms(e0, e1) * 1e-3) / 1e12;
	k_fma<float><<<blocks, threads>>>(d32, 100);
	cudaEventRecord(e0); k_fma<float><<<blocks, threads>>>(d32, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
	const double fp32_tf = 2.0 * 8 * iters * (double) blocks * threads / (time_ms(e0, e1) * 1e-3) / 1e12;
	// gathers: 64 MiB (L2-resident on a 126 MB L2) and 64 KiB (L1-resident)
	float *src;
	const size_t n_l2 = 16u << 20;
	cudaMalloc(&src, n_l2 * sizeof(float));
	cudaMemset(src, 0, n_l2 * sizeof(float));
	const int gblocks = p.multiProcessorCount * 16, giters = 4096;
	double res[2];
	const unsigned masks[2] = {(unsigned) n_l2 - 1, (16u << 10) - 1};
	for (int k = 0; k < 2; ++k) {
		k_gather<<<gblocks, threads>>>(src, masks[k], 256, d32);
		cudaEventRecord(e0); k_gather<<<gblocks, threads>>>(src, masks[k], giters, d32); cudaEventRecord(e1); cudaEventSynchronize(e1);
		res[k] = (double) gblocks * threads * giters / (time_ms(e0, e1) * 1e-3);  // loads / s
	}
	printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.2f, \"fp32_fma_tflops\": %.2f, "
		   "\"l2_random_4B_loads_per_s\": %.4g, \"l2_random_sector_GBps\": %.1f, \"l1_random_4B_loads_per_s\": %.4g, "
		   "\"how\": \"8 independent FMA chains per thread, 8 CTAs x 256 threads per SM; random 4-byte __ldg over 64 MiB (L2) / 64 KiB (L1), 4 in flight per thread, 16 CTAs x 256 threads per SM; sector GB/s = loads x 32 B\"}\n",
		   p.name, p.multiProcessorCount, fp64_tf, fp32_tf, res[0], res[0] * 32 / 1e9, res[1]);
	return 0;
}
