// lut_bench.cu -- design probe for the exact gamma-LUT problem.
//
// The reference's kernels gather from 65536-entry float tables (colourMaths.ts:130-169).  On B200 a
// random 4-byte gather from a 256 KiB table is an L1/L2 transaction per lane; this probe measures
// lossless in-shared-memory encodings of the same tables:
//   V0  global __ldg gather from the raw table                         (what the reference does)
//   V1  delta8 + MUFU base:  bits = bits(s*ex2(G*lg2(i*p+q))+o) + d8[i]  (64 KiB / table)
//   V2  16-entry segments {anchor, slope} + d8[i]                      (96 KiB / table)
//   V3  raw LDS.32 gather from a 16 Ki-entry dummy table               (smem gather ceiling)
//   V4  V1's arithmetic without the gather                             (ALU/MUFU ceiling)
//   V5  index generation only                                          (to subtract)
// and checks that every encoding reproduces the table bit for bit for all 65536 indices.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lut_bench tools/lut_bench.cu phaneron_b200/csrc/pb_colour.cpp
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/phaneron_b200.h"

#define CK(x)                                                                        \
	do {                                                                             \
		cudaError_t e = (x);                                                         \
		if (e != cudaSuccess) {                                                      \
			printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
			exit(1);                                                                 \
		}                                                                            \
	} while (0)

struct BaseParams {
	float p, q, G, s, o, kt;
	int J;   // indices below J use the linear toe i*kt
};

__device__ __forceinline__ float ex2a(float x) {
	float r;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
__device__ __forceinline__ float lg2a(float x) {
	float r;
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

__device__ __forceinline__ float base_of(uint32_t idx, const BaseParams &bp) {
	const float fi = __uint_as_float(0x4B000000u | idx);   // 2^23 + idx
	const float x = __fmaf_rn(fi, bp.p, bp.q);             // q already has -2^23*p folded in
	const float pw = __fmaf_rn(ex2a(__fmul_rn(bp.G, lg2a(x))), bp.s, bp.o);
	const float toe = __fmaf_rn(fi, bp.kt, -8388608.0f * bp.kt);
	return idx < (uint32_t)bp.J ? toe : pw;
}

__global__ void k_fit(const float *table, BaseParams bp, int8_t *d8, int *minmax) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 65536) return;
	const int d = (int)__float_as_uint(table[i]) - (int)__float_as_uint(base_of(i, bp));
	atomicMin(&minmax[0], d);
	atomicMax(&minmax[1], d);
	d8[i] = (int8_t)max(-128, min(127, d));
}

__global__ void k_verify_v1(const float *table, BaseParams bp, const int8_t *d8, int *bad) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 65536) return;
	const uint32_t b = __float_as_uint(base_of(i, bp)) + (int)d8[i];
	if (b != __float_as_uint(table[i])) atomicAdd(bad, 1);
}

// Option X: segments of 16
__global__ void k_fit_x(const float *table, float2 *seg, int8_t *d8, int *minmax) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 65536) return;
	const uint32_t k = i >> 4, j = i & 15;
	const float a = table[k * 16], b = table[k * 16 + 15];
	const float slope = (b - a) / 15.0f;
	if (j == 0) seg[k] = make_float2(a, slope);
	const float base = __fmaf_rn((float)j, slope, a);
	const int d = (int)__float_as_uint(table[i]) - (int)__float_as_uint(base);
	atomicMin(&minmax[0], d);
	atomicMax(&minmax[1], d);
	d8[i] = (int8_t)max(-128, min(127, d));
}

constexpr int kThreads = 512;
constexpr int kIter = 2048;

template <int V, bool kRandom>
__global__ void __launch_bounds__(kThreads) k_tp(const float *table, BaseParams bp, const int8_t *d8g, const float2 *segg, float *out) {
	extern __shared__ __align__(16) unsigned char sm[];
	int8_t *d8 = reinterpret_cast<int8_t *>(sm);                    // 64 KiB
	float2 *seg = reinterpret_cast<float2 *>(sm + 65536);           // 32 KiB (V2)
	float *raw = reinterpret_cast<float *>(sm);                     // V3: 16 Ki floats
	if (V == 1 || V == 2 || V == 4) {
		for (int i = threadIdx.x; i < 65536 / 16; i += kThreads) reinterpret_cast<uint4 *>(d8)[i] = reinterpret_cast<const uint4 *>(d8g)[i];
	}
	if (V == 2) {
		for (int i = threadIdx.x; i < 4096; i += kThreads) seg[i] = segg[i];
	}
	if (V == 3) {
		for (int i = threadIdx.x; i < 16384; i += kThreads) raw[i] = table[i * 4];
	}
	__syncthreads();
	uint32_t s = (blockIdx.x * kThreads + threadIdx.x) * 2654435761u + 12345u;
	const uint32_t coh = (blockIdx.x * 977u + (threadIdx.x >> 5) * 131u) & 0xFFFFu;
	uint32_t acc = 0;
#pragma unroll 8
	for (int it = 0; it < kIter; ++it) {
		uint32_t idx;
		if (kRandom) {
			s = s * 1664525u + 1013904223u;
			idx = s >> 16;
		} else {   // coherent: neighbouring lanes a few entries apart, drifting
			idx = (coh + it * 37u + (threadIdx.x & 31u) * 3u) & 0xFFFFu;
		}
		uint32_t v;
		if (V == 0) {
			v = __float_as_uint(__ldg(table + idx));
		} else if (V == 1) {
			v = __float_as_uint(base_of(idx, bp)) + (int)d8[idx];
		} else if (V == 2) {
			const float2 as = seg[idx >> 4];
			const float fj = __uint_as_float(0x4B000000u | (idx & 15u));
			v = __float_as_uint(__fmaf_rn(fj, as.y, as.x)) + (int)d8[idx];   // as.x has -2^23*slope folded in (not here: probe only)
		} else if (V == 3) {
			v = __float_as_uint(raw[idx & 16383u]);
		} else if (V == 4) {
			v = __float_as_uint(base_of(idx, bp));
		} else {
			v = idx;
		}
		acc ^= v;
	}
	out[blockIdx.x * kThreads + threadIdx.x] = __uint_as_float(acc);
}

template <int V, bool R>
static void run_tp(const char *name, const float *table, BaseParams bp, const int8_t *d8, const float2 *seg, float *out, int sms) {
	size_t smem = (V == 2) ? 65536 + 32768 : 65536;
	// reserve what the real kernel would: two tables resident
	size_t reserve = (V == 2) ? 2 * (65536 + 32768) : 2 * 65536;
	if (V == 0 || V == 5) reserve = smem = 0;
	if (V == 3) reserve = smem = 65536;
	CK(cudaFuncSetAttribute(k_tp<V, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(reserve > 48 * 1024 ? reserve : 48 * 1024)));
	int blocks_per_sm = 0;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_tp<V, R>, kThreads, reserve));
	const int grid = sms * (blocks_per_sm > 0 ? blocks_per_sm : 1) * 2;
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	for (int w = 0; w < 2; ++w) k_tp<V, R><<<grid, kThreads, reserve>>>(table, bp, d8, seg, out);
	CK(cudaEventRecord(e0));
	const int reps = 5;
	for (int w = 0; w < reps; ++w) k_tp<V, R><<<grid, kThreads, reserve>>>(table, bp, d8, seg, out);
	CK(cudaEventRecord(e1));
	CK(cudaEventSynchronize(e1));
	float ms;
	CK(cudaEventElapsedTime(&ms, e0, e1));
	const double lookups = (double)grid * kThreads * kIter * reps;
	printf("%-34s %s  grid %4d x %d (occ %d/SM)  %8.1f Glookup/s   %6.2f lookups/clk/SM @1.9GHz\n", name, R ? "random  " : "coherent", grid, kThreads,
	       blocks_per_sm, lookups / (ms * 1e-3) / 1e9, lookups / (ms * 1e-3) / 1.9e9 / sms);
}

static BaseParams params_g2l(double alpha, double beta, double gamma, double delta) {
	BaseParams bp;
	const double p = 1.0 / (65535.0 * alpha), q = (alpha - 1) / alpha;
	bp.p = (float)p;
	bp.q = (float)(q - 8388608.0 * (double)bp.p);
	bp.G = (float)(1.0 / gamma);
	bp.s = 1.0f;
	bp.o = 0.0f;
	bp.kt = (float)(1.0 / (65535.0 * delta));
	bp.J = (int)std::ceil(beta * delta * 65535.0);
	return bp;
}
static BaseParams params_l2g(double alpha, double beta, double gamma, double delta) {
	BaseParams bp;
	bp.p = (float)(1.0 / 65535.0);
	bp.q = (float)(0.0 - 8388608.0 * (double)bp.p);
	bp.G = (float)gamma;
	bp.s = (float)alpha;
	bp.o = (float)(-(alpha - 1));
	bp.kt = (float)(delta / 65535.0);
	bp.J = (int)std::ceil(beta * 65535.0);
	return bp;
}

int main() {
	int dev = 0;
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, dev));
	const int sms = prop.multiProcessorCount;
	printf("device %s, %d SMs\n", prop.name, sms);
	std::vector<float> h(65536);
	float *table;
	int8_t *d8;
	float2 *seg;
	int *mm;
	float *out;
	CK(cudaMalloc(&table, 65536 * 4));
	CK(cudaMalloc(&d8, 65536));
	CK(cudaMalloc(&seg, 4096 * 8));
	CK(cudaMalloc(&mm, 16));
	CK(cudaMalloc(&out, (size_t)sms * 8 * kThreads * 4));

	struct Case {
		const char *name;
		bool g2l;
		const char *spec;
		double alpha, beta, gamma, delta;
	} cases[] = {
		{"g2l 709", true, "709", 1.099, 0.018, 0.45, 4.5},
		{"l2g 709", false, "709", 1.099, 0.018, 0.45, 4.5},
		{"g2l sRGB", true, "sRGB", 1.055, 0.0031308, 1.0 / 2.4, 12.92},
		{"l2g sRGB", false, "sRGB", 1.055, 0.0031308, 1.0 / 2.4, 12.92},
	};
	BaseParams bp0{};
	for (const auto &c : cases) {
		if (c.g2l) pb_gamma2linear_lut(c.spec, h.data());
		else pb_linear2gamma_lut(c.spec, h.data());
		CK(cudaMemcpy(table, h.data(), 65536 * 4, cudaMemcpyHostToDevice));
		const BaseParams bp = c.g2l ? params_g2l(c.alpha, c.beta, c.gamma, c.delta) : params_l2g(c.alpha, c.beta, c.gamma, c.delta);
		int init[4] = {1 << 30, -(1 << 30), 0, 0};
		CK(cudaMemcpy(mm, init, 16, cudaMemcpyHostToDevice));
		k_fit<<<256, 256>>>(table, bp, d8, mm);
		k_verify_v1<<<256, 256>>>(table, bp, d8, mm + 2);
		int res[4];
		CK(cudaMemcpy(res, mm, 16, cudaMemcpyDeviceToHost));
		printf("%-9s V1 (MUFU base, J=%d): delta range [%d, %d], mismatches after decode: %d\n", c.name, bp.J, res[0], res[1], res[2]);
		CK(cudaMemcpy(mm, init, 16, cudaMemcpyHostToDevice));
		k_fit_x<<<256, 256>>>(table, seg, d8, mm);
		CK(cudaMemcpy(res, mm, 16, cudaMemcpyDeviceToHost));
		printf("%-9s V2 (16-entry linear segments): delta range [%d, %d]\n", c.name, res[0], res[1]);
		if (&c == &cases[0]) bp0 = bp;
	}
	// throughput on the first case's tables
	pb_gamma2linear_lut("709", h.data());
	CK(cudaMemcpy(table, h.data(), 65536 * 4, cudaMemcpyHostToDevice));
	int init[4] = {1 << 30, -(1 << 30), 0, 0};
	CK(cudaMemcpy(mm, init, 16, cudaMemcpyHostToDevice));
	k_fit_x<<<256, 256>>>(table, seg, d8, mm);
	CK(cudaDeviceSynchronize());
	run_tp<2, true>("V2 segments+delta8 (smem)", table, bp0, d8, seg, out, sms);
	run_tp<2, false>("V2 segments+delta8 (smem)", table, bp0, d8, seg, out, sms);
	k_fit<<<256, 256>>>(table, bp0, d8, mm);
	CK(cudaDeviceSynchronize());
	run_tp<0, true>("V0 global __ldg raw table", table, bp0, d8, seg, out, sms);
	run_tp<0, false>("V0 global __ldg raw table", table, bp0, d8, seg, out, sms);
	run_tp<1, true>("V1 MUFU base + delta8 (smem)", table, bp0, d8, seg, out, sms);
	run_tp<1, false>("V1 MUFU base + delta8 (smem)", table, bp0, d8, seg, out, sms);
	run_tp<3, true>("V3 raw LDS.32 16Ki table", table, bp0, d8, seg, out, sms);
	run_tp<3, false>("V3 raw LDS.32 16Ki table", table, bp0, d8, seg, out, sms);
	run_tp<4, true>("V4 MUFU base only (no gather)", table, bp0, d8, seg, out, sms);
	run_tp<5, true>("V5 index generation only", table, bp0, d8, seg, out, sms);
	CK(cudaDeviceSynchronize());
	return 0;
}
