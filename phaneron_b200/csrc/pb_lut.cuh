// pb_lut.cuh -- lossless shared-memory form of phaneron's 65536-entry gamma tables.
//
// The reference gathers gammaLut[convert_ushort_sat_rte(v * 65535)] from a 256 KiB float table
// (v210.ts:68-70, 148-150; tables from colourMaths.ts:130-169).  Two such tables do not fit the
// 227 KiB of shared memory of an SM, and random 4-byte gathers from L1/L2 are what bounds the
// reference's algorithm on a B200 (tools/lut_bench.cu: 0.8 Tlookup/s from L2 vs 2.0 from this
// encoding on incoherent indices).  So the fused kernel keeps, per table, ONE BYTE per entry:
//
//     table[i] == as_float( as_int( base(i) ) + d8[i] - 128 )     for all 65536 i
//     base(i)   = i < J ? i * kt : s * ex2.approx(lg2.approx(i * p + q) * G) + o
//
// base() is the analytic transfer function evaluated with the SFU approximations (a handful of
// ulps off); d8 is the integer distance from it to the exact table value.  d8 is computed on the
// device by the very same lut_base() the kernels decode with, so the decode is exact by
// construction; lut_fit_kernel also reports the min/max distance so the host can reject a table
// the model does not describe (it then stays a global-memory gather).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pb_desc.h"

namespace pb {

__device__ __forceinline__ float ex2_approx(float x) {
	float r;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
	float r;
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

// fi = the table index as an exact float (0 .. 65535).  The march kernel evaluates the same
// expression two indices at a time with packed f32x2 instructions (pb_march.cu lut2); every
// operation is IEEE round-to-nearest or an SFU approximation, so both forms agree bit for bit.
__device__ __forceinline__ float lut_base(float fi, const LutParams &lp) {
	const float x = __fmaf_rn(fi, lp.p, lp.q);
	float pw;
	if (lp.affine == 2) {   // MUFU-free model: Horner, highest coefficient first (the march kernel runs the same chain as FFMA2)
		pw = lp.c[kLutPolyDeg];
#pragma unroll
		for (int k = kLutPolyDeg - 1; k >= 0; --k) pw = __fmaf_rn(pw, x, lp.c[k]);
	} else {
		pw = ex2_approx(__fmul_rn(lg2_approx(x), lp.G));
		if (lp.affine) pw = __fmaf_rn(pw, lp.s, lp.o);
	}
	const float toe = __fmul_rn(fi, lp.kt);
	const float h = __saturatef(__fadd_rn(fi, lp.cJ));   // 0 below the knee, 1 from it on
	// below the knee exactly the toe; above it RN(RN(pw - toe) + toe), a few ulp from pw -- any deterministic
	// function of the index will do here, the byte table holds the distance to the exact value
	return __fmaf_rn(h, __fsub_rn(pw, toe), toe);
}

// exact table value from the byte table (shared or global memory)
// (bytes hold the distance + 128 so that the decode is one unsigned byte load and one 3-input add)
__device__ __forceinline__ float lut_decode(float fi, uint32_t idx, const uint8_t *d8, const LutParams &lp) {
	return __int_as_float(__float_as_int(lut_base(fi, lp)) + (int)d8[idx] - 128);
}

struct LutFitResult {
	int dmin, dmax;
	unsigned long long hash;   // order-independent content hash of the raw table
	int not_unit, pad;         // some entry lies outside [0, 1] (or is NaN)
};

// one candidate parameter set per blockIdx.y
static __global__ void lut_fit_kernel(const float *table, const LutParams *cands, uint8_t *d8_out, LutFitResult *res) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 65536) return;
	const LutParams lp = cands[blockIdx.y];
	const uint32_t tb = __float_as_uint(table[i]);
	const int d = (int)tb - __float_as_int(lut_base((float)i, lp));
	d8_out[(size_t)blockIdx.y * 65536 + i] = (uint8_t)(max(-128, min(127, d)) + 128);
	atomicMin(&res[blockIdx.y].dmin, d);
	atomicMax(&res[blockIdx.y].dmax, d);
	if (blockIdx.y == 0) {
		unsigned long long h = (unsigned long long)tb * 0x9E3779B97F4A7C15ull + i;
		h ^= h >> 29;
		h *= 0xBF58476D1CE4E5B9ull + 2ull * i;
		atomicAdd(&res[0].hash, h);
		const float tv = table[i];
		if (!(tv >= 0.0f && tv <= 1.0f)) atomicOr(&res[0].not_unit, 1);
	}
}

}  // namespace pb
