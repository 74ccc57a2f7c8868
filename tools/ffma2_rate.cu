// issue-rate probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, and each mixed 1:1 with LOP3
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, int iters, float m, float c) {
	float2 a[8];
	unsigned u[8];
	for (int i = 0; i < 8; ++i) { a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f); u[i] = threadIdx.x + i; }
	const float2 mm = make_float2(m, m), cc = make_float2(c, c);
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			if (MODE == 0 || MODE == 2) { a[i].x = __fmaf_rn(a[i].x, m, c); a[i].y = __fmaf_rn(a[i].y, m, c); }
			if (MODE == 1 || MODE == 3) a[i] = __ffma2_rn(a[i], mm, cc);
			if (MODE >= 2) { u[i] = (u[i] & 0x3ff) | (u[(i + 1) & 7] ^ 0x4b000000u); u[i] ^= u[(i + 3) & 7] >> 3; }
		}
	}
	float s = 0; unsigned t = 0;
	for (int i = 0; i < 8; ++i) { s += a[i].x + a[i].y; t ^= u[i]; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)t;
}
template <int MODE> void run(const char *name, float *out) {
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int iters = 4096, grid = 148 * 4;
	k<MODE><<<grid, 512>>>(out, 64, 0.999f, 0.001f);
	cudaEventRecord(e0);
	k<MODE><<<grid, 512>>>(out, iters, 0.999f, 0.001f);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	double fma = (double)grid * 512 * iters * 16;
	printf("%-28s %7.3f ms  %7.2f T lane-FMA/s  = %6.1f lane-FMA/clk/SM @1.9GHz\n", name, ms, fma / ms / 1e9, fma / (ms * 1e-3) / 1.9e9 / 148);
}
int main() {
	float *out; cudaMalloc(&out, 148 * 4 * 512 * 4);
	run<0>("FFMA scalar", out); run<1>("FFMA2 packed", out); run<2>("FFMA scalar + 2 LOP3 per FMA-pair", out); run<3>("FFMA2 packed + 2 LOP3", out);
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
