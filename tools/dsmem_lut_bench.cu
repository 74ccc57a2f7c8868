// dsmem_lut_bench.cu -- probe: can a thread-block cluster hold a RAW 65536-entry float gamma table (256 KiB: more than one SM's
// shared memory) spread over the shared memories of its CTAs, and serve random lookups with ld.shared::cluster fast enough to
// replace the exact one-byte decode of pb_lut.cuh (about 10 issue slots + 2 MUFU per lookup)?  DESIGN.md 4.3.
//
//   CL = 1: 64 Ki floats do not fit; a 32 Ki-entry local table as the LDS.32 reference point
//   CL = 2: 32 Ki floats (128 KiB) per CTA, half of the lookups remote
//   CL = 4: 16 Ki floats (64 KiB) per CTA, three quarters remote
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dsmem_lut_bench tools/dsmem_lut_bench.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                        \
	do {                                                                             \
		cudaError_t e = (x);                                                         \
		if (e != cudaSuccess) {                                                      \
			printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
			exit(1);                                                                 \
		}                                                                            \
	} while (0)

constexpr int kThreads = 640;
constexpr int kIter = 2048;

template <int CL, bool kRandom>
__global__ void __launch_bounds__(kThreads, 1) k_probe(float *out, unsigned long long *check) {
	extern __shared__ __align__(16) unsigned char sm[];
	constexpr uint32_t kPer = (CL == 1 ? 32768u : 65536u / CL);   // entries per CTA
	float *part = reinterpret_cast<float *>(sm);
	uint32_t rank = 0;
	if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
	for (uint32_t i = threadIdx.x; i < kPer; i += kThreads) part[i] = (float)(rank * kPer + i);   // table[i] = i
	const uint32_t local = (uint32_t)__cvta_generic_to_shared(part);
	uint32_t base0 = local, stride = 0;
	if (CL > 1) {
		uint32_t b1;
		asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(base0) : "r"(local));
		asm volatile("mapa.shared::cluster.u32 %0, %1, 1;" : "=r"(b1) : "r"(local));
		stride = b1 - base0;
		asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
	} else {
		__syncthreads();
	}
	uint32_t s = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 12345u;
	const uint32_t coh = (blockIdx.x * 977u + (threadIdx.x >> 5) * 131u) & 0xFFFFu;
	float acc = 0.f;
	unsigned long long want = 0;
#pragma unroll 8
	for (int it = 0; it < kIter; ++it) {
		uint32_t idx;
		if (kRandom) {
			s = s * 1664525u + 1013904223u;
			idx = s >> 16;
		} else {
			idx = (coh + it * 37u + (threadIdx.x & 31u) * 3u) & 0xFFFFu;
		}
		if (CL == 1) idx &= kPer - 1;
		float v;
		if (CL == 1) {
			asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(local + idx * 4u));
		} else {
			const uint32_t a = base0 + (idx / kPer) * stride + (idx % kPer) * 4u;
			asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a));
		}
		acc += v;
		want += idx;
	}
	out[blockIdx.x * kThreads + threadIdx.x] = acc;
	if ((unsigned long long)acc != want && want < (1ull << 24)) atomicAdd(check, 1ull);   // (float sums are exact below 2^24 only: informative)
	if (CL > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // nobody leaves while peers still read
}

template <int CL, bool R>
static void run(const char *name, float *out, unsigned long long *check, int sms) {
	const size_t smem = (CL == 1 ? 32768 : 65536 / CL) * 4;
	const size_t reserve = smem > 120 * 1024 ? smem : 120 * 1024;   // one CTA per SM, as in the real kernel
	CK(cudaFuncSetAttribute(k_probe<CL, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)reserve));
	if (CL > 1) CK(cudaFuncSetAttribute(k_probe<CL, R>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
	int grid = sms / CL * CL;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(grid);
	cfg.blockDim = dim3(kThreads);
	cfg.dynamicSmemBytes = reserve;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeClusterDimension;
	at[0].val.clusterDim.x = CL;
	at[0].val.clusterDim.y = 1;
	at[0].val.clusterDim.z = 1;
	cfg.attrs = at;
	cfg.numAttrs = CL > 1 ? 1 : 0;
	int max_clusters = -1;
	if (CL > 1) cudaOccupancyMaxActiveClusters(&max_clusters, k_probe<CL, R>, &cfg);
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	CK(cudaMemset(check, 0, 8));
	for (int w = 0; w < 2; ++w) CK(cudaLaunchKernelEx(&cfg, k_probe<CL, R>, out, check));
	CK(cudaEventRecord(e0));
	const int reps = 5;
	for (int w = 0; w < reps; ++w) CK(cudaLaunchKernelEx(&cfg, k_probe<CL, R>, out, check));
	CK(cudaEventRecord(e1));
	CK(cudaEventSynchronize(e1));
	float ms;
	CK(cudaEventElapsedTime(&ms, e0, e1));
	unsigned long long bad = 0;
	CK(cudaMemcpy(&bad, check, 8, cudaMemcpyDeviceToHost));
	const double lookups = (double)grid * kThreads * kIter * reps;
	printf("%-44s %s  grid %3d x %d, cluster %d (max active clusters %d)  %8.1f Glookup/s  %6.2f lookups/clk/SM @1.965GHz\n", name, R ? "random  " : "coherent", grid,
	       kThreads, CL, max_clusters, lookups / (ms * 1e-3) / 1e9, lookups / (ms * 1e-3) / 1.965e9 / grid);
}

int main() {
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	const int sms = prop.multiProcessorCount;
	printf("device %s, %d SMs\n", prop.name, sms);
	float *out;
	unsigned long long *check;
	CK(cudaMalloc(&out, (size_t)sms * kThreads * 4));
	CK(cudaMalloc(&check, 8));
	run<1, true>("local LDS.32, 32 Ki-entry table", out, check, sms);
	run<1, false>("local LDS.32, 32 Ki-entry table", out, check, sms);
	run<2, true>("cluster of 2, raw 64 Ki floats over DSMEM", out, check, sms);
	run<2, false>("cluster of 2, raw 64 Ki floats over DSMEM", out, check, sms);
	run<4, true>("cluster of 4, raw 64 Ki floats over DSMEM", out, check, sms);
	run<4, false>("cluster of 4, raw 64 Ki floats over DSMEM", out, check, sms);
	CK(cudaDeviceSynchronize());
	return 0;
}
