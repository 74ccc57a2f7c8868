// pb_device.cuh -- device-side building blocks (sm_100a) for the phaneron pixel path.
//
// Float semantics are the "canonical" ones documented in DESIGN.md: every operation
// is an explicit round-to-nearest intrinsic (__fmul_rn/__fadd_rn/__fmaf_rn/__fdiv_rn),
// so nvcc's -fmad contraction can never change a result; OpenCL dot() is LLVM's contracted
// chain (see dot3/dot4); convert_*_sat_rte is clamp + RNE (NaN -> 0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pb_desc.h"

namespace pb {

// ---- exact arithmetic helpers ------------------------------------------------------
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// OpenCL dot() as LLVM contracts it (read off the PTX NVIDIA's OpenCL compiler emits for the reference's
// kernels): a.x*b.x + a.y*b.y fuses to fma(a.x, b.x, a.y*b.y), the rest chain on.
__device__ __forceinline__ float dot3(float a0, float a1, float a2, const float *b) {
	float t = mul(a1, b[1]);
	t = fma_(a0, b[0], t);
	return fma_(a2, b[2], t);
}
__device__ __forceinline__ float dot4(float a0, float a1, float a2, float a3, const float *b) {
	float t = mul(a1, b[1]);
	t = fma_(a0, b[0], t);
	t = fma_(a2, b[2], t);
	return fma_(a3, b[3], t);
}

// convert_ushort_sat_rte (OpenCL 1.2 6.2.3.3): NaN -> 0, clamp to [0, 65535], round
// half to even.  x + 2^23 performs the RNE rounding in the FADD; the integer then sits
// in the low mantissa bits.  fmaxf(NaN, 0) = 0 gives the NaN rule.
__device__ __forceinline__ uint32_t sat_rte_u16(float x) {
	const float c = fminf(fmaxf(x, 0.0f), 65535.0f);
	return __float_as_uint(add(c, 8388608.0f)) & 0xFFFFu;
}
__device__ __forceinline__ uint32_t sat_rte_u8(float x) {
	const float c = fminf(fmaxf(x, 0.0f), 255.0f);
	return __float_as_uint(add(c, 8388608.0f)) & 0xFFu;
}
// convert_ushort_sat_rtz
__device__ __forceinline__ uint32_t sat_rtz_u16(float x) {
	const float c = fminf(fmaxf(x, 0.0f), 65535.0f);
	return (uint32_t)c;
}
// exact uint (< 2^23) -> float without the I2F pipe
__device__ __forceinline__ float u2f(uint32_t v) {
	return sub(__uint_as_float(0x4B000000u | v), 8388608.0f);
}

// ---- v210 (v210.ts:25-195) ---------------------------------------------------------
// 6 pixels per 16-byte group: three 10-bit codes per 32-bit word.
struct Ycc {
	uint32_t y, cb, cr;
};

__device__ __forceinline__ Ycc v210_px(const uint4 &w, int p) {
	// v210.ts:58-63
	Ycc o;
	switch (p) {
		case 0: o.y = (w.x >> 10) & 0x3ff; o.cb = w.x & 0x3ff; o.cr = (w.x >> 20) & 0x3ff; break;
		case 1: o.y = w.y & 0x3ff; o.cb = w.x & 0x3ff; o.cr = (w.x >> 20) & 0x3ff; break;
		case 2: o.y = (w.y >> 20) & 0x3ff; o.cb = (w.y >> 10) & 0x3ff; o.cr = w.z & 0x3ff; break;
		case 3: o.y = (w.z >> 10) & 0x3ff; o.cb = (w.y >> 10) & 0x3ff; o.cr = w.z & 0x3ff; break;
		case 4: o.y = w.w & 0x3ff; o.cb = (w.z >> 20) & 0x3ff; o.cr = (w.w >> 10) & 0x3ff; break;
		default: o.y = (w.w >> 20) & 0x3ff; o.cb = (w.z >> 20) & 0x3ff; o.cr = (w.w >> 10) & 0x3ff; break;
	}
	return o;
}

// YCbCr code triple -> linear RGB in the working gamut (v210.ts:65-77).
// alpha is the 4th component of yuva: 1 in the main loop, 0 in the line tail (Q1).
__device__ __forceinline__ float3 ycc_to_linear(const Ycc &c, float alpha, const ReadConsts &rc) {
	const float fy = u2f(c.y), fcb = u2f(c.cb), fcr = u2f(c.cr);
	const float r = __ldg(rc.lut + sat_rte_u16(mul(dot4(fy, fcb, fcr, alpha, rc.cm + 0), 65535.0f)));
	const float g = __ldg(rc.lut + sat_rte_u16(mul(dot4(fy, fcb, fcr, alpha, rc.cm + 4), 65535.0f)));
	const float b = __ldg(rc.lut + sat_rte_u16(mul(dot4(fy, fcb, fcr, alpha, rc.cm + 8), 65535.0f)));
	float3 o;
	o.x = dot3(r, g, b, rc.gamut + 0);
	o.y = dot3(r, g, b, rc.gamut + 3);
	o.z = dot3(r, g, b, rc.gamut + 6);
	return o;
}

// linear RGB -> 10-bit YCbCr codes (v210.ts:145-156)
__device__ __forceinline__ Ycc linear_to_ycc(float r, float g, float b, const WriteConsts &wc) {
	const float gr = __ldg(wc.lut + sat_rte_u16(mul(r, 65535.0f)));
	const float gg = __ldg(wc.lut + sat_rte_u16(mul(g, 65535.0f)));
	const float gb = __ldg(wc.lut + sat_rte_u16(mul(b, 65535.0f)));
	Ycc o;
	o.y = sat_rte_u16(dot4(gr, gg, gb, 1.0f, wc.cm + 0));
	o.cb = sat_rte_u16(dot4(gr, gg, gb, 1.0f, wc.cm + 4));
	o.cr = sat_rte_u16(dot4(gr, gg, gb, 1.0f, wc.cm + 8));
	return o;
}
// line-tail variant (Q2, v210.ts:173-184): _rtz LUT index, round() half away from zero
__device__ __forceinline__ Ycc linear_to_ycc_tail(float r, float g, float b, const WriteConsts &wc) {
	const float gr = __ldg(wc.lut + sat_rtz_u16(mul(r, 65535.0f)));
	const float gg = __ldg(wc.lut + sat_rtz_u16(mul(g, 65535.0f)));
	const float gb = __ldg(wc.lut + sat_rtz_u16(mul(b, 65535.0f)));
	Ycc o;
	o.y = sat_rtz_u16(roundf(dot4(gr, gg, gb, 1.0f, wc.cm + 0)));
	o.cb = sat_rtz_u16(roundf(dot4(gr, gg, gb, 1.0f, wc.cm + 4)));
	o.cr = sat_rtz_u16(roundf(dot4(gr, gg, gb, 1.0f, wc.cm + 8)));
	return o;
}

// v210.ts:158-163: chroma is taken from even pixels only (Q8)
__device__ __forceinline__ uint4 v210_pack(const Ycc *p) {
	uint4 w;
	w.x = p[0].cr << 20 | p[0].y << 10 | p[0].cb;
	w.y = p[2].y << 20 | p[2].cb << 10 | p[1].y;
	w.z = p[4].cb << 20 | p[3].y << 10 | p[2].cr;
	w.w = p[5].y << 20 | p[4].cr << 10 | p[4].y;
	return w;
}

// streaming 128-bit accesses: packed frames are touched exactly once
__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
	             : "l"(p));
	return r;
}
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 &v) {
	asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
	             "r"(v.z), "r"(v.w)
	             : "memory");
}

// ---- texels of the other packed formats (same arithmetic as the stand-alone read kernels in pb_kernels.cu) ----
// rgba8.ts:40-62 / bgra8.ts: 8-bit codes -> LUT index code * 65535 / 255; alpha goes through the LUT too (rgba8.ts:61)
__device__ __forceinline__ float4 rgba8_to_linear(uchar4 v, bool bgra, const ReadConsts &rc) {
	// convert_ushort_sat_rte(c * 65535.0f / 255.0f) for a code c in 0..255: c * 65535 < 2^24 is exact, 65535 / 255 = 257, so
	// the quotient is the integer c * 257 exactly and every rounding step is the identity: the index is c * 257
	const uint32_t c0 = bgra ? v.z : v.x, c1 = v.y, c2 = bgra ? v.x : v.z, c3 = v.w;
	const float r = __ldg(rc.lut + c0 * 257u);
	const float g = __ldg(rc.lut + c1 * 257u);
	const float b = __ldg(rc.lut + c2 * 257u);
	float4 o;
	o.x = dot3(r, g, b, rc.gamut + 0);
	o.y = dot3(r, g, b, rc.gamut + 3);
	o.z = dot3(r, g, b, rc.gamut + 6);
	o.w = __ldg(rc.lut + c3 * 257u);   // alpha goes through the LUT too (rgba8.ts:61)
	return o;
}

// texel (i, j) of an rgba8 / bgra8 / planar 4:2:2 / 4:2:0 leaf, inside the image
__device__ __forceinline__ float4 packed_texel(const Leaf &lf, const ReadConsts &rc, int i, int j) {
	if (lf.kind == LEAF_RGBA8 || lf.kind == LEAF_BGRA8)
		return rgba8_to_linear(__ldg(reinterpret_cast<const uchar4 *>(lf.ptr) + (size_t)j * lf.w + i), lf.kind == LEAF_BGRA8, rc);
	const int pitch = (lf.w + 7) / 8 * 8;   // samples per luma line (yuv422p10.ts:222, yuv420p.ts:240)
	Ycc c;
	if (lf.kind == LEAF_YUV422P10) {
		c.y = __ldg(reinterpret_cast<const uint16_t *>(lf.ptr) + (size_t)j * pitch + i);
		c.cb = __ldg(reinterpret_cast<const uint16_t *>(lf.ptr_u) + (size_t)j * (pitch / 2) + i / 2);
		c.cr = __ldg(reinterpret_cast<const uint16_t *>(lf.ptr_v) + (size_t)j * (pitch / 2) + i / 2);
	} else {
		const uint8_t *Y = reinterpret_cast<const uint8_t *>(lf.ptr), *U = reinterpret_cast<const uint8_t *>(lf.ptr_u),
		              *V = reinterpret_cast<const uint8_t *>(lf.ptr_v);
		c.y = __ldg(Y + (size_t)j * pitch + i);
		if (lf.kind == LEAF_YUV422P8) {
			c.cb = __ldg(U + (size_t)j * (pitch / 2) + i / 2);
			c.cr = __ldg(V + (size_t)j * (pitch / 2) + i / 2);
		} else if (lf.kind == LEAF_YUV420P) {   // one chroma line per line pair
			c.cb = __ldg(U + (size_t)(j / 2) * (pitch / 2) + i / 2);
			c.cr = __ldg(V + (size_t)(j / 2) * (pitch / 2) + i / 2);
		} else {   // LEAF_NV12: interleaved (U, V) pairs
			const uint8_t *cp = U + (size_t)(j / 2) * pitch + (i / 2) * 2;
			c.cb = __ldg(cp);
			c.cr = __ldg(cp + 1);
		}
	}
	const float3 rgb = ycc_to_linear(c, 1.0f, rc);
	return make_float4(rgb.x, rgb.y, rgb.z, 1.0f);
}

// ---- yadif: yadifCl.ts:28-167 ------------------------------------------------------------------------
__device__ __forceinline__ float half_sum(float a, float b) { return mul(add(a, b), 0.5f); }   // (a + b) / 2.0f, exact
__device__ __forceinline__ float ad(float a, float b) { return fabsf(sub(a, b)); }

__device__ __forceinline__ float spatial_predictor(float a, float b, float c, float d, float e, float f, float g, float h,
                                                    float i, float j, float k, float l, float m, float n) {
	float pred = half_sum(d, k);
	float best = add(add(ad(c, j), ad(d, k)), ad(e, l));
	float score = add(add(ad(b, k), ad(c, l)), ad(d, m));
	bool cmp = score < best;
	pred = cmp ? half_sum(c, l) : pred;
	best = cmp ? score : best;
	score = cmp ? add(add(ad(a, l), ad(b, m)), ad(c, n)) : score;
	cmp = cmp && (score < best);
	pred = cmp ? half_sum(b, m) : pred;
	best = cmp ? score : best;
	score = add(add(ad(d, i), ad(e, j)), ad(f, k));
	cmp = score < best;
	pred = cmp ? half_sum(e, j) : pred;
	best = cmp ? score : best;
	score = cmp ? add(add(ad(e, h), ad(f, i)), ad(g, j)) : score;
	cmp = cmp && (score < best);
	pred = cmp ? half_sum(f, i) : pred;
	return pred;
}

__device__ __forceinline__ float temporal_predictor(float A, float B, float C, float D, float E, float F, float G, float H,
                                                     float I, float J, float K, float L, float pred, int skip) {
	const float p0 = half_sum(C, H), p1 = F, p2 = half_sum(D, I), p3 = G, p4 = half_sum(E, J);
	const float t0 = ad(D, I);
	const float t1 = mul(add(ad(A, F), ad(B, G)), 0.5f);
	const float t2 = mul(add(ad(K, F), ad(G, L)), 0.5f);
	float diff = fmaxf(fmaxf(t0, t1), t2);
	if (!skip) {
		const float p2mp3 = sub(p2, p3), p2mp1 = sub(p2, p1), p0mp1 = sub(p0, p1), p4mp3 = sub(p4, p3);
		const float maxi = fmaxf(fmaxf(p2mp3, p2mp1), fminf(p0mp1, p4mp3));
		const float mini = fminf(fminf(p2mp3, p2mp1), fmaxf(p0mp1, p4mp3));
		diff = fmaxf(fmaxf(diff, mini), -maxi);
	}
	const float hi = add(p2, diff), lo = sub(p2, diff);
	pred = (pred > hi) ? hi : pred;
	pred = (pred < lo) ? lo : pred;
	return pred;
}

// one output pixel of the yadif kernel (yadifCl.ts:105-167) from three RGBA-f32 frames; reads clamp to the edge
// (CLK_ADDRESS_CLAMP_TO_EDGE).  Shared by the stand-alone kernel (pb_kernels.cu k_yadif) and the fused kernels' Yadif leaves.
__device__ __forceinline__ float4 yadif_texel(const float4 *__restrict__ prev, const float4 *__restrict__ cur, const float4 *__restrict__ next,
                                              int w, int h, int parity, int tff, int skip, int xo, int yo) {
	auto px = [&](const float4 *img, int x, int y) {
		x = min(max(x, 0), w - 1);
		y = min(max(y, 0), h - 1);
		return __ldg(img + (size_t)y * w + x);
	};
	if ((yo & 1) == parity) return px(cur, xo, yo);   // the primary field is not modified
	const int second = !(parity ^ tff);
	const float4 a = px(cur, xo - 3, yo - 1), b = px(cur, xo - 2, yo - 1), c = px(cur, xo - 1, yo - 1), d = px(cur, xo, yo - 1),
	             e = px(cur, xo + 1, yo - 1), f = px(cur, xo + 2, yo - 1), g = px(cur, xo + 3, yo - 1);
	const float4 hh = px(cur, xo - 3, yo + 1), i = px(cur, xo - 2, yo + 1), j = px(cur, xo - 1, yo + 1), k = px(cur, xo, yo + 1),
	             l = px(cur, xo + 1, yo + 1), m = px(cur, xo + 2, yo + 1), n = px(cur, xo + 3, yo + 1);
	const float4 A = px(prev, xo, yo - 1), B = px(prev, xo, yo + 1);
	const float4 C = px(second ? cur : prev, xo, yo - 2), D = px(second ? cur : prev, xo, yo), E = px(second ? cur : prev, xo, yo + 2);
	const float4 F = d, G = k;
	const float4 H = px(second ? next : cur, xo, yo - 2), I = px(second ? next : cur, xo, yo), J = px(second ? next : cur, xo, yo + 2);
	const float4 K = px(next, xo, yo - 1), L = px(next, xo, yo + 1);
	float4 o;
#define YADIF_CH(ch) \
	o.ch = temporal_predictor(A.ch, B.ch, C.ch, D.ch, E.ch, F.ch, G.ch, H.ch, I.ch, J.ch, K.ch, L.ch, \
	                          spatial_predictor(a.ch, b.ch, c.ch, d.ch, e.ch, f.ch, g.ch, hh.ch, i.ch, j.ch, k.ch, l.ch, m.ch, n.ch), skip)
	YADIF_CH(x);
	YADIF_CH(y);
	YADIF_CH(z);
#undef YADIF_CH
	o.w = px(cur, xo, yo).w;   // "Reset Alpha" (yadifCl.ts:164): the w channel's prediction is discarded
	return o;
}

// ---- leaves -------------------------------------------------------------------------
// One texel of a leaf as RGBA-f32; texels outside the image are the CLK_ADDRESS_CLAMP
// border colour (0,0,0,0).
__device__ __forceinline__ float4 leaf_texel(const Leaf &lf, const ReadConsts *rcs, int i, int j) {
	if (i < 0 || j < 0 || i >= lf.w || j >= lf.h) return make_float4(0.f, 0.f, 0.f, 0.f);
	if (lf.kind == LEAF_RGBA_F32) {
		return __ldg(reinterpret_cast<const float4 *>(lf.ptr) + (size_t)j * lf.w + i);
	}
	if (lf.kind == LEAF_YADIF)   // a de-interlaced field: the lines of its own parity are the current frame's (ptr), the interpolated ones
		// were computed by the launch's pre-pass (k_yadif_rows) into ptr_u, row j >> 1
		return (j & 1) == (lf.yadif & 1) ? __ldg(reinterpret_cast<const float4 *>(lf.ptr) + (size_t)j * lf.w + i)
		                                 : __ldg(reinterpret_cast<const float4 *>(lf.ptr_u) + (size_t)(j >> 1) * lf.w + i);
	if (lf.kind != LEAF_V210) return packed_texel(lf, rcs[lf.rc], i, j);
	const int g = i / 6, p = i - g * 6;
	const uint4 w = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(lf.ptr) + (size_t)j * lf.pitch) + g);
	const Ycc c = v210_px(w, p);
	// Q1: pixels of the partial last group are converted with yuva.w = 0
	const float alpha = (i >= lf.w - lf.w % 6) ? 0.0f : 1.0f;
	const float3 rgb = ycc_to_linear(c, alpha, rcs[lf.rc]);
	return make_float4(rgb.x, rgb.y, rgb.z, 1.0f);
}

// read_imagef(normalised, CLK_ADDRESS_CLAMP, CLK_FILTER_LINEAR): OpenCL 1.2 spec 8.2
template <typename Fetch>
__device__ __forceinline__ float4 sample_linear_clamp(int sw, int sh, float s, float t, Fetch fetch) {
	const float um = sub(mul(s, (float)sw), 0.5f), vm = sub(mul(t, (float)sh), 0.5f);
	const float fu = floorf(um), fv = floorf(vm);
	const float a = sub(um, fu), b = sub(vm, fv);
	int i0, j0;
	if (!(fu >= -2.0f)) i0 = -2; else if (fu > (float)sw) i0 = sw; else i0 = (int)fu;
	if (!(fv >= -2.0f)) j0 = -2; else if (fv > (float)sh) j0 = sh; else j0 = (int)fv;
	const float ra = sub(1.0f, a), rb = sub(1.0f, b);
	const float w00 = mul(ra, rb), w10 = mul(a, rb), w01 = mul(ra, b), w11 = mul(a, b);
	const float4 t00 = fetch(i0, j0), t10 = fetch(i0 + 1, j0), t01 = fetch(i0, j0 + 1), t11 = fetch(i0 + 1, j0 + 1);
	float4 r;
	r.x = fma_(w11, t11.x, fma_(w01, t01.x, fma_(w10, t10.x, mul(w00, t00.x))));
	r.y = fma_(w11, t11.y, fma_(w01, t01.y, fma_(w10, t10.y, mul(w00, t00.y))));
	r.z = fma_(w11, t11.z, fma_(w01, t01.z, fma_(w10, t10.z, mul(w00, t00.z))));
	r.w = fma_(w11, t11.w, fma_(w01, t01.w, fma_(w10, t10.w, mul(w00, t00.w))));
	return r;
}

// transform.ts:54-57: posIn = M(2x3) . (x/w - 1/2, y/h - 1/2, 1) + 1/2
__device__ __forceinline__ float2 transform_pos(const float *m, int x, int y, int w, int h) {
	const float ix = sub(__fdiv_rn((float)x, (float)w), 0.5f);
	const float iy = sub(__fdiv_rn((float)y, (float)h), 0.5f);
	float2 p;
	p.x = add(dot3(ix, iy, 1.0f, m + 0), 0.5f);
	p.y = add(dot3(ix, iy, 1.0f, m + 3), 0.5f);
	return p;
}

// Lanczos-filtered sample (definition: oracle/oracle.c): sum_j wy_j * (sum_i wx_i * T(i0 + i, j0 + j)), ascending
// fma chains from +0, weights from the host-built tables; texels outside the image are (0,0,0,0)
__device__ __forceinline__ float4 lanczos_sample(const Leaf &lf, const ReadConsts *rcs, int x, int y) {
	const int i0 = __ldg(lf.lz_i0 + x), j0 = __ldg(lf.lz_j0 + y);
	const float *wx = lf.lz_wx + (size_t)x * lf.lz_tx, *wy = lf.lz_wy + (size_t)y * lf.lz_ty;
	float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
	for (int j = 0; j < lf.lz_ty; ++j) {
		float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
		for (int i = 0; i < lf.lz_tx; ++i) {
			const float4 t = leaf_texel(lf, rcs, i0 + i, j0 + j);
			const float w = __ldg(wx + i);
			row.x = fma_(w, t.x, row.x);
			row.y = fma_(w, t.y, row.y);
			row.z = fma_(w, t.z, row.z);
			row.w = fma_(w, t.w, row.w);
		}
		const float w = __ldg(wy + j);
		acc.x = fma_(w, row.x, acc.x);
		acc.y = fma_(w, row.y, acc.y);
		acc.z = fma_(w, row.z, acc.z);
		acc.w = fma_(w, row.w, acc.w);
	}
	return acc;
}

// value of one leaf at output pixel (x, y)
__device__ __forceinline__ float4 leaf_value(const Leaf &lf, const ReadConsts *rcs, int x, int y) {
	if (lf.kind == LEAF_LANCZOS_V) {   // second pass of a separable Lanczos Transform over the filtered rows H
		const int j0 = __ldg(lf.lz_j0 + y);
		const float *wy = lf.lz_wy + (size_t)y * lf.lz_ty;
		const float4 *H = reinterpret_cast<const float4 *>(lf.ptr);
		float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
		for (int j = max(0, -j0); j < min(lf.lz_ty, lf.h - j0); ++j) {
			const float4 hv = __ldg(H + (size_t)(j0 + j) * lf.w + x);
			const float w = __ldg(wy + j);
			acc.x = fma_(w, hv.x, acc.x);
			acc.y = fma_(w, hv.y, acc.y);
			acc.z = fma_(w, hv.z, acc.z);
			acc.w = fma_(w, hv.w, acc.w);
		}
		return acc;
	}
	if (!lf.has_xf) return leaf_texel(lf, rcs, x, y);
	if (lf.lz_tx) return lanczos_sample(lf, rcs, x, y);
	const float2 p = transform_pos(lf.m, x, y, lf.xf_w, lf.xf_h);
	return sample_linear_clamp(lf.w, lf.h, p.x, p.y, [&](int i, int j) { return leaf_texel(lf, rcs, i, j); });
}

// transition.ts:60-73 / combine.ts:49-59
__device__ __forceinline__ float4 dissolve4(const float4 &in0, const float4 &in1, float mix) {
	const float rmix = sub(1.0f, mix);
	float4 o;
	o.x = fma_(in0.x, mix, mul(in1.x, rmix));
	o.y = fma_(in0.y, mix, mul(in1.y, rmix));
	o.z = fma_(in0.z, mix, mul(in1.z, rmix));
	o.w = fma_(in0.w, mix, mul(in1.w, rmix));
	return o;
}
__device__ __forceinline__ float4 wipe_mask4(const float4 &in0, const float4 &in1, float m) {
	const float rm = sub(1.0f, m);
	float4 o;
	o.x = fma_(in1.x, m, mul(in0.x, rm));
	o.y = fma_(in1.y, m, mul(in0.y, rm));
	o.z = fma_(in1.z, m, mul(in0.z, rm));
	o.w = fma_(in1.w, m, mul(in0.w, rm));
	return o;
}
__device__ __forceinline__ float4 over4(const float4 &acc, const float4 &l) {
	const float k = sub(1.0f, l.w);
	float4 o;
	o.x = fma_(acc.x, k, l.x);
	o.y = fma_(acc.y, k, l.y);
	o.z = fma_(acc.z, k, l.z);
	o.w = fma_(acc.w, 0.0f, l.w);
	return o;
}

__device__ __forceinline__ float4 layer_value(const Layer &ly, const ReadConsts *rcs, int x, int y) {
	const float4 a = leaf_value(ly.a, rcs, x, y);
	if (ly.kind == LAYER_DIRECT) return a;
	const float4 b = leaf_value(ly.b, rcs, x, y);
	if (ly.kind == LAYER_DISSOLVE) return dissolve4(a, b, ly.mix);
	const float4 m = leaf_value(ly.mask, rcs, x, y);
	return wipe_mask4(a, b, m.x);
}

}  // namespace pb
