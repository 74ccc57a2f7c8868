// does fma.rn.f32x2 / mul.rn.f32x2 / add.rn.f32x2 equal the scalar rn ops bit for bit?
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ uint32_t rng(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__global__ void k(unsigned long long *bad, int iters) {
	uint32_t s = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
	unsigned long long b0 = 0, b1 = 0, b2 = 0, b3 = 0;
	for (int i = 0; i < iters; ++i) {
		float a0 = (rng(s) & 1023), a1 = (rng(s) & 1023);
		float m0 = __uint_as_float(0x3A000000u | (rng(s) & 0x7FFFFF)), m1 = -__uint_as_float(0x39000000u | (rng(s) & 0x7FFFFF));
		float c0 = __uint_as_float(0x3F000000u | (rng(s) & 0x7FFFFF)) - 1.0f, c1 = __uint_as_float(0x3E800000u | (rng(s) & 0x7FFFFF));
		float2 p = __ffma2_rn(make_float2(a0, a1), make_float2(m0, m1), make_float2(c0, c1));
		float s0 = __fmaf_rn(a0, m0, c0), s1 = __fmaf_rn(a1, m1, c1);
		if (p.x != s0 || p.y != s1) b0++;
		float2 q = __fmul2_rn(make_float2(c0, c1), make_float2(65535.0f, 65535.0f));
		if (q.x != __fmul_rn(c0, 65535.0f) || q.y != __fmul_rn(c1, 65535.0f)) b1++;
		float2 r = __fadd2_rn(q, make_float2(8388608.0f, 8388608.0f));
		if (r.x != __fadd_rn(q.x, 8388608.0f) || r.y != __fadd_rn(q.y, 8388608.0f)) b2++;
		float t = c0 * 3.0f - 1.0f;
		float sat = __saturatef(t), cl = fminf(fmaxf(t, 0.0f), 1.0f);
		if (sat != cl) b3++;
	}
	atomicAdd(bad + 0, b0); atomicAdd(bad + 1, b1); atomicAdd(bad + 2, b2); atomicAdd(bad + 3, b3);
}
int main() {
	unsigned long long *d, h[4] = {0, 0, 0, 0};
	cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
	k<<<256, 256>>>(d, 2000);
	cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
	printf("mismatches of %d: ffma2=%llu fmul2=%llu fadd2=%llu saturate=%llu (%s)\n", 256 * 256 * 2000, h[0], h[1], h[2], h[3], cudaGetErrorString(cudaGetLastError()));
	return 0;
}
