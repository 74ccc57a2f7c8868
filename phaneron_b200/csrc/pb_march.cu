// pb_march.cu -- the fast fused kernel: warp-autonomous strip marching.
//
//   N layers x (v210 unpack -> YCbCr->R'G'B' -> gamma LUT -> gamut 3x3 -> bilinear Transform
//   -> dissolve | wipe) -> combine (premultiplied over) -> linear->gamma LUT -> RGB->YCbCr
//   -> 10-bit RTE -> v210 pack, one launch, every packed source byte read once from HBM.
// Reference stages replaced: v210.ts:25-195, transform.ts:36-59, transition.ts:60-73,
// combine.ts:24-68 and the RGBA-f32 frames between them.
//
// Shape of the kernel:
//   * one persistent CTA per SM, kMarchWarps warps; the gamma tables live in shared memory in the
//     lossless one-byte-per-entry form of pb_lut.cuh (128 KiB for a read + a write table);
//   * a work item is one output line of one strip (31 or 32 v210 groups = 186 / 192 px); items are
//     dealt round-robin to all warps of the grid, so a warp never waits on another warp: no
//     __syncthreads after the table load, only __syncwarp;
//   * per leaf and source row the warp converts the strip's source footprint ONCE (lane = v210
//     group: one 128-bit load, 6 texels) into its private planar row buffer, then every lane
//     takes its taps for 6 output pixels (lane = pixel, stride-1 conflict-free LDS) with the exact
//     {i0, a} / {j0, b} tables the host derived from the reference's float formula;
//   * the canonical FMA chain of the oracle (w00*t00 -> +w10*t10 -> +w01*t01 -> +w11*t11) is
//     evaluated row by row, so one row buffer per warp suffices;
//   * the 6-pixel / 4-word v210 regroup goes through the same buffer: 32 lanes x 6 rounds of
//     codes in, lane = group out, one coalesced 16-byte store per lane.
// Instruction economy (the kernel is issue / FMA-pipe / SFU bound, not HBM bound -- DESIGN.md 4.3):
//   * the two pixels of a 4:2:2 chroma pair are converted together with packed fp32x2
//     instructions (fma.rn.f32x2: same lane rate as FFMA, half the issue slots);
//   * 10-bit fields become floats with one LOP3 (mask | 2^23 exponent) and no shift where the
//     field sits at bit 10: it is read as 1024*v and meets a coefficient pre-divided by 1024;
//   * the luma bias of that trick is folded into the first FMA of the matrix row (ReadK::oY);
//   * the table index, the table address and the exact float index all come from ONE add of a
//     per-table magic constant 2^23 + (shared-memory address of the table);
//   * the toe / power select of the transfer function is arithmetic (saturating FMAs), keeping
//     the half-rate ALU pipe for the unpack masks and the final integer add.
// Every float operation is an explicit IEEE round-to-nearest op or an SFU approximation that the
// table fit has already absorbed, so results are bit-identical to the generic kernel
// (pb_fused.cu) and to the oracle.
#include <mutex>
#include <set>
#include <utility>

#include "pb_device.cuh"
#include "pb_launch.h"
#include "pb_lut.cuh"

namespace pb {

namespace {

constexpr int kRounds = 6;   // 192 px / 32 lanes
constexpr float kTwo23 = 8388608.0f;

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (one rounding instead of two), so a
// packed product that feeds a packed add is written as fma(a, b, +0): RN(a*b + 0) == RN(a*b) for
// the non-negative products used here, and an FFMA2 cannot be contracted any further.
__device__ __forceinline__ float2 mul2_unfusable(float2 a, float2 b) { return __ffma2_rn(a, b, f2s(0.0f)); }

__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) {
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
	return v;
}

// one gamma table as the kernel sees it
template <int kLutMode>
struct LutK {
	float magic;        // mode 1: 2^23 + shared-memory byte address of the d8 table (even); mode 0: 2^23
	const float *raw;   // mode 0: the raw table in global memory
};

// two saturated values -> two exact table values (v210.ts:68-70 / 148-150).
// convert_ushort_sat_rte(v * 65535) == RNE(sat(v) * 65535): both ends of the clamp are fixed points
// of the multiply, and NaN saturates to 0 either way.
template <int kLutMode>
__device__ __forceinline__ float2 lut2(float2 z, const LutK<kLutMode> &k, const LutParams &lp) {
	const float2 u = mul2_unfusable(z, f2s(65535.0f));
	const float2 v = __fadd2_rn(u, f2s(k.magic));   // RNE to an integer, held in the low mantissa bits
	if (kLutMode == 0) {
		return f2(__ldg(k.raw + (__float_as_uint(v.x) & 0xFFFFu)), __ldg(k.raw + (__float_as_uint(v.y) & 0xFFFFu)));
	}
	const uint32_t d0 = lds_u8(__float_as_uint(v.x) & 0x7FFFFFu), d1 = lds_u8(__float_as_uint(v.y) & 0x7FFFFFu);
	const float2 fi = __fadd2_rn(v, f2s(-k.magic));   // the index as an exact float
	// lut_base() of pb_lut.cuh, two lanes wide
	const float2 x = __ffma2_rn(fi, f2s(lp.p), f2s(lp.q));
	const float2 y = __fmul2_rn(f2(lg2_approx(x.x), lg2_approx(x.y)), f2s(lp.G));
	float2 pw = f2(ex2_approx(y.x), ex2_approx(y.y));
	if (lp.affine) pw = __ffma2_rn(pw, f2s(lp.s), f2s(lp.o));
	const float2 toe = __fmul2_rn(fi, f2s(lp.kt));
	const float h0 = __saturatef(add(fi.x, lp.cJ)), h1 = __saturatef(add(fi.y, lp.cJ));
	const float t0 = __saturatef(fma_(h0, -16.0f, toe.x)), t1 = __saturatef(fma_(h1, -16.0f, toe.y));
	const float2 base = __ffma2_rn(f2(h0, h1), pw, f2(t0, t1));
	return f2(__int_as_float(__float_as_int(base.x) + (int)d0 - 128), __int_as_float(__float_as_int(base.y) + (int)d1 - 128));
}

// Two horizontally adjacent pixels sharing one chroma pair -> linear RGB in the working gamut
// (v210.ts:65-77).  ya/yb/cb/cr are the raw exponent-trick floats 2^23 + s*v; SYA.. say whether
// s is 1 (0) or 1024 (1) for that field.
template <int kLutMode, bool kSparse, int SYA, int SYB, int SCB, int SCR>
__device__ __forceinline__ void convert_pair(uint32_t ya, uint32_t yb, uint32_t cb, uint32_t cr, const ReadConsts &rc, const ReadK &rk,
                                             const LutK<kLutMode> &lut, const LutParams &lp, float2 &R, float2 &G, float2 &B) {
	const float2 Y = f2(__uint_as_float(ya), __uint_as_float(yb));
	const float2 C = __fadd2_rn(f2(__uint_as_float(cb), __uint_as_float(cr)), f2s(-kTwo23));   // exact (scaled) chroma codes
	// dot(yuva, colMatrix row): mul, fma, fma, fma(1, m3, t) -- the last is RN(t + m3), fused with the saturate
	float2 tr = __ffma2_rn(Y, f2(rk.mY[0][SYA], rk.mY[0][SYB]), f2(rk.oY[0][SYA], rk.oY[0][SYB]));
	float2 tg = __ffma2_rn(Y, f2(rk.mY[1][SYA], rk.mY[1][SYB]), f2(rk.oY[1][SYA], rk.oY[1][SYB]));
	float2 tb = __ffma2_rn(Y, f2(rk.mY[2][SYA], rk.mY[2][SYB]), f2(rk.oY[2][SYA], rk.oY[2][SYB]));
	if (!kSparse) tr = f2(fma_(C.x, rk.mCb[0][SCB], tr.x), fma_(C.x, rk.mCb[0][SCB], tr.y));   // cm[1] == 0: fma(cb, 0, t) == t
	tr = f2(fma_(C.y, rk.mCr[0][SCR], tr.x), fma_(C.y, rk.mCr[0][SCR], tr.y));
	tg = f2(fma_(C.x, rk.mCb[1][SCB], tg.x), fma_(C.x, rk.mCb[1][SCB], tg.y));
	tg = f2(fma_(C.y, rk.mCr[1][SCR], tg.x), fma_(C.y, rk.mCr[1][SCR], tg.y));
	tb = f2(fma_(C.x, rk.mCb[2][SCB], tb.x), fma_(C.x, rk.mCb[2][SCB], tb.y));
	if (!kSparse) tb = f2(fma_(C.y, rk.mCr[2][SCR], tb.x), fma_(C.y, rk.mCr[2][SCR], tb.y));
	const float2 r = lut2<kLutMode>(f2(__saturatef(add(tr.x, rc.cm[3])), __saturatef(add(tr.y, rc.cm[3]))), lut, lp);
	const float2 g = lut2<kLutMode>(f2(__saturatef(add(tg.x, rc.cm[7])), __saturatef(add(tg.y, rc.cm[7]))), lut, lp);
	const float2 b = lut2<kLutMode>(f2(__saturatef(add(tb.x, rc.cm[11])), __saturatef(add(tb.y, rc.cm[11]))), lut, lp);
	// gamut 3x3: mul, fma, fma per row
	R = __ffma2_rn(b, f2s(rc.gamut[2]), __ffma2_rn(g, f2s(rc.gamut[1]), __fmul2_rn(r, f2s(rc.gamut[0]))));
	G = __ffma2_rn(b, f2s(rc.gamut[5]), __ffma2_rn(g, f2s(rc.gamut[4]), __fmul2_rn(r, f2s(rc.gamut[3]))));
	B = __ffma2_rn(b, f2s(rc.gamut[8]), __ffma2_rn(g, f2s(rc.gamut[7]), __fmul2_rn(r, f2s(rc.gamut[6]))));
}

// one v210 group (6 texels, v210.ts:58-63) -> the warp's planar row buffer at texel column 6g
template <int kLutMode, bool kSparse>
__device__ __forceinline__ void convert_group(const uint4 &w, int g, const ReadConsts &rc, const ReadK &rk, const LutK<kLutMode> &lut,
                                              const LutParams &lp, float *buf) {
	float2 *pr = reinterpret_cast<float2 *>(buf + 0 * kRowCap + g * 6);
	float2 *pg = reinterpret_cast<float2 *>(buf + 1 * kRowCap + g * 6);
	float2 *pb_ = reinterpret_cast<float2 *>(buf + 2 * kRowCap + g * 6);
	constexpr uint32_t E = 0x4B000000u, M0 = 0x3ffu, M10 = 0xffc00u;
	float2 R, G, B;
	// word 0: Cr0 | Y0 | Cb0     word 1: Y2 | Cb1 | Y1     word 2: Cb2 | Y3 | Cr1     word 3: Y5 | Cr2 | Y4
	convert_pair<kLutMode, kSparse, 1, 0, 0, 0>((w.x & M10) | E, (w.y & M0) | E, (w.x & M0) | E, ((w.x >> 20) & M0) | E, rc, rk, lut, lp, R, G, B);
	pr[0] = R; pg[0] = G; pb_[0] = B;
	convert_pair<kLutMode, kSparse, 0, 1, 1, 0>(((w.y >> 20) & M0) | E, (w.z & M10) | E, (w.y & M10) | E, (w.z & M0) | E, rc, rk, lut, lp, R, G, B);
	pr[1] = R; pg[1] = G; pb_[1] = B;
	convert_pair<kLutMode, kSparse, 0, 0, 0, 1>((w.w & M0) | E, ((w.w >> 20) & M0) | E, ((w.z >> 20) & M0) | E, (w.w & M10) | E, rc, rk, lut, lp, R, G, B);
	pr[2] = R; pg[2] = G; pb_[2] = B;
}

// value of one leaf at the 6 pixels of this lane -> p[r] = (r, g, b, alpha)
template <int kLutMode, bool kSparse, bool kSingleRc>
__device__ __forceinline__ void eval_leaf(const FusedDesc &d, const Leaf &lf, uint32_t lut_saddr, float *buf, int lane, int strip, int y,
                                          int x_first, int x_last, float4 (&p)[kRounds]) {
#pragma unroll
	for (int r = 0; r < kRounds; ++r) p[r] = make_float4(0.f, 0.f, 0.f, 0.f);
	const int4 si = __ldg(lf.strip_tab + strip);
	const int2 rt = __ldg(lf.row_tab + y);
	if (!(si.x & 1)) return;   // the strip does not touch this leaf's image: border colour everywhere
	const bool edge = (si.x & 2) != 0;
	const int g_lo = si.y, ng = si.z, origin = g_lo * 6, last = ng * 6 - 1;
	const int j0 = rt.x;
	const bool has_xf = lf.has_xf != 0;
	const bool ok0 = (unsigned)j0 < (unsigned)lf.h, ok1 = has_xf && (unsigned)(j0 + 1) < (unsigned)lf.h;
	if (!ok0 && !ok1) return;   // both rows are border rows

	// issue every HBM load of this leaf up front: <= 2 groups per lane per row (kRowGroups = 64)
	const uint4 *src0 = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(lf.ptr) + (size_t)j0 * lf.pitch) + g_lo + lane;
	const uint4 *src1 = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(src0) + lf.pitch);
	const uint4 z4 = make_uint4(0, 0, 0, 0);
	const uint4 a0 = (ok0 && lane < ng) ? ld_stream(src0) : z4, a1 = (ok0 && lane + 32 < ng) ? ld_stream(src0 + 32) : z4;
	const uint4 b0 = (ok1 && lane < ng) ? ld_stream(src1) : z4, b1 = (ok1 && lane + 32 < ng) ? ld_stream(src1 + 32) : z4;

	const float b = __int_as_float(rt.y), rb = sub(1.0f, b);
	const int rci = kSingleRc ? 0 : lf.rc;
	const ReadConsts &rc = d.rc[rci];
	const ReadK &rk = d.rk[rci];
	const int slot = kLutMode ? rc.lut_slot : 0;
	const LutParams &lp = d.luts[kSingleRc ? 0 : slot].lp;   // the host puts rc[0]'s table in slot 0
	LutK<kLutMode> lut;
	lut.raw = rc.lut;
	lut.magic = kLutMode ? kTwo23 + (float)(lut_saddr + (kSingleRc ? 0 : slot) * 65536) : kTwo23;

#pragma unroll 1
	for (int rr = 0; rr < 2; ++rr) {
		if (!(rr ? ok1 : ok0)) continue;   // border row: all its taps are (0,0,0,0)
#pragma unroll 1
		for (int it = 0; it < 2; ++it) {
			const int g = lane + it * 32;
			if (g < ng) {
				const uint4 w = rr ? (it ? b1 : b0) : (it ? a1 : a0);
				convert_group<kLutMode, kSparse>(w, g, rc, rk, lut, lp, buf);
			}
		}
		__syncwarp();
		if (!has_xf) {   // 1:1 read of texel (x, y): exact passthrough, alpha = 1 (leaf_value in pb_device.cuh)
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const float *t = buf + (min(x_first + r * 32 + lane, x_last) - origin);
				p[r] = make_float4(t[0], t[kRowCap], t[2 * kRowCap], 1.0f);
			}
		} else {
			const float wr = rr == 0 ? rb : b;
			const int2 *ctab = lf.col_tab + x_first + lane;
			const int xmax = x_last - x_first - lane;   // clamp for the ragged last strip
			if (!edge) {
#pragma unroll
				for (int r = 0; r < kRounds; ++r) {
					const int2 ct = __ldg(ctab + min(r * 32, xmax));
					const float *t = buf + (ct.x - origin);
					const float ca = __int_as_float(ct.y);
					const float w0 = mul(sub(1.0f, ca), wr), w1 = mul(ca, wr);   // w00|w01 , w10|w11
					p[r].x = fma_(w1, t[1], fma_(w0, t[0], p[r].x));
					p[r].y = fma_(w1, t[kRowCap + 1], fma_(w0, t[kRowCap], p[r].y));
					p[r].z = fma_(w1, t[2 * kRowCap + 1], fma_(w0, t[2 * kRowCap], p[r].z));
					// alpha taps are 1 inside the image: fma(w, 1, al) = RN(w + al)
					p[r].w = add(w1, add(w0, p[r].w));
				}
			} else {
#pragma unroll
				for (int r = 0; r < kRounds; ++r) {
					const int2 ct = __ldg(ctab + min(r * 32, xmax));
					const int i0 = ct.x, c0 = i0 - origin;
					const float ca = __int_as_float(ct.y);
					const bool f0 = (unsigned)i0 < (unsigned)lf.w, f1 = (unsigned)(i0 + 1) < (unsigned)lf.w;
					const float *t0 = buf + min(max(c0, 0), last), *t1 = buf + min(max(c0 + 1, 0), last);
					const float w0 = mul(sub(1.0f, ca), wr), w1 = mul(ca, wr);
					const float t0r = f0 ? t0[0] : 0.0f, t0g = f0 ? t0[kRowCap] : 0.0f, t0b = f0 ? t0[2 * kRowCap] : 0.0f;
					const float t1r = f1 ? t1[0] : 0.0f, t1g = f1 ? t1[kRowCap] : 0.0f, t1b = f1 ? t1[2 * kRowCap] : 0.0f;
					p[r].x = fma_(w1, t1r, fma_(w0, t0r, p[r].x));
					p[r].y = fma_(w1, t1g, fma_(w0, t0g, p[r].y));
					p[r].z = fma_(w1, t1b, fma_(w0, t0b, p[r].z));
					// border taps have alpha 0: fma(w, 0, al) = al
					float al = p[r].w;
					al = f0 ? add(w0, al) : al;
					al = f1 ? add(w1, al) : al;
					p[r].w = al;
				}
			}
		}
		__syncwarp();
	}
}

template <int kLutMode, bool kSparse, bool kSingleRc>
__global__ void __launch_bounds__(kMarchThreads, 1) k_fused_march(const __grid_constant__ FusedDesc d) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	uint8_t *lut_s = reinterpret_cast<uint8_t *>(smem_raw);
	const uint32_t lut_saddr = (uint32_t)__cvta_generic_to_shared(lut_s);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	float *buf = reinterpret_cast<float *>(smem_raw + (kLutMode ? (size_t)d.n_luts * 65536 : 0)) + warp * kRowFloats;

	if (kLutMode) {
		for (int t = 0; t < d.n_luts; ++t) {
			const uint4 *src = reinterpret_cast<const uint4 *>(d.luts[t].d8);
			uint4 *dst = reinterpret_cast<uint4 *>(lut_s + (size_t)t * 65536);
			for (int i = threadIdx.x; i < 65536 / 16; i += kMarchThreads) dst[i] = __ldg(src + i);
		}
		__syncthreads();
	}

	const int step = d.interlace == 0 ? 1 : 2;
	const int first_line = d.interlace == 3 ? 1 : 0;
	const int n_lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
	const int total = n_lines * d.n_strips;
	const int strip_px = d.strip_groups * 6;

	LutK<kLutMode> wlut;
	wlut.raw = d.wc.lut;
	wlut.magic = kLutMode ? kTwo23 + (float)(lut_saddr + d.wc.lut_slot * 65536) : kTwo23;
	const LutParams &wlp = d.wlp;

#pragma unroll 1
	for (int item = blockIdx.x * kMarchWarps + warp; item < total; item += gridDim.x * kMarchWarps) {
		const int k = item / d.n_strips, strip = item - k * d.n_strips;
		const int y = first_line + k * step;
		const int x_first = strip * strip_px;
		const int x_last = min(x_first + strip_px, d.out_w) - 1;

		float3 acc[kRounds];
#pragma unroll 1
		for (int l = 0; l < d.n_layers; ++l) {
			const Layer &ly = d.layers[l];
			float4 p[kRounds], t[kRounds];
			float m[kRounds];
			// evaluation order keeps at most {t, p} live: dissolve = b then a; wipe = mask, a, b
			const int nleaf = ly.kind == LAYER_DIRECT ? 1 : (ly.kind == LAYER_DISSOLVE ? 2 : 3);
#pragma unroll 1
			for (int q = 0; q < nleaf; ++q) {
				const Leaf &lf = ly.kind == LAYER_DIRECT ? ly.a
				                 : ly.kind == LAYER_DISSOLVE ? (q == 0 ? ly.b : ly.a)
				                                             : (q == 0 ? ly.mask : (q == 1 ? ly.a : ly.b));
				eval_leaf<kLutMode, kSparse, kSingleRc>(d, lf, lut_saddr, buf, lane, strip, y, x_first, x_last, p);
				if (ly.kind == LAYER_DISSOLVE) {   // transition.ts:60-65: fma(in0, mix, in1 * (1 - mix))
					if (q == 0) {
						const float rmix = sub(1.0f, ly.mix);
#pragma unroll
						for (int r = 0; r < kRounds; ++r) t[r] = make_float4(mul(p[r].x, rmix), mul(p[r].y, rmix), mul(p[r].z, rmix), mul(p[r].w, rmix));
					} else {
#pragma unroll
						for (int r = 0; r < kRounds; ++r)
							p[r] = make_float4(fma_(p[r].x, ly.mix, t[r].x), fma_(p[r].y, ly.mix, t[r].y), fma_(p[r].z, ly.mix, t[r].z),
							                   fma_(p[r].w, ly.mix, t[r].w));
					}
				} else if (ly.kind == LAYER_WIPE_MASK) {   // transition.ts:66-73: fma(in1, m, in0 * (1 - m)), m = mask.r
					if (q == 0) {
#pragma unroll
						for (int r = 0; r < kRounds; ++r) m[r] = p[r].x;
					} else if (q == 1) {
#pragma unroll
						for (int r = 0; r < kRounds; ++r) {
							const float rm = sub(1.0f, m[r]);
							t[r] = make_float4(mul(p[r].x, rm), mul(p[r].y, rm), mul(p[r].z, rm), mul(p[r].w, rm));
						}
					} else {
#pragma unroll
						for (int r = 0; r < kRounds; ++r)
							p[r] = make_float4(fma_(p[r].x, m[r], t[r].x), fma_(p[r].y, m[r], t[r].y), fma_(p[r].z, m[r], t[r].z),
							                   fma_(p[r].w, m[r], t[r].w));
					}
				}
			}
			if (l == 0) {
#pragma unroll
				for (int r = 0; r < kRounds; ++r) acc[r] = make_float3(p[r].x, p[r].y, p[r].z);
			} else {   // combine.ts:49-59: fma(prev, 1 - l.a, l)
#pragma unroll
				for (int r = 0; r < kRounds; ++r) {
					const float kk = sub(1.0f, p[r].w);
					acc[r] = make_float3(fma_(acc[r].x, kk, p[r].x), fma_(acc[r].y, kk, p[r].y), fma_(acc[r].z, kk, p[r].z));
				}
			}
		}

		// ---- encode (v210.ts:145-156), two rounds at a time, and regroup 6 pixels -> 4 words through the row buffer ----
		// The host has checked that every code lies in [0, 1023] for table values in [0, 1], so
		// convert_ushort_sat_rte reduces to the RNE add and the three codes share one word.
		uint32_t *stage = reinterpret_cast<uint32_t *>(buf);
#pragma unroll
		for (int r = 0; r < kRounds; r += 2) {
			const float2 gr = lut2<kLutMode>(f2(__saturatef(acc[r].x), __saturatef(acc[r + 1].x)), wlut, wlp);
			const float2 gg = lut2<kLutMode>(f2(__saturatef(acc[r].y), __saturatef(acc[r + 1].y)), wlut, wlp);
			const float2 gb = lut2<kLutMode>(f2(__saturatef(acc[r].z), __saturatef(acc[r + 1].z)), wlut, wlp);
			uint32_t code0 = 0, code1 = 0;
#pragma unroll
			for (int c = 0; c < 3; ++c) {   // dot(rgba, colMatrix row): mul, fma, fma, fma(1, m3, t) = RN(t + m3)
				float2 v = __ffma2_rn(gb, f2s(d.wc.cm[c * 4 + 2]), __ffma2_rn(gg, f2s(d.wc.cm[c * 4 + 1]), __fmul2_rn(gr, f2s(d.wc.cm[c * 4 + 0]))));
				v = __fadd2_rn(v, f2s(d.wc.cm[c * 4 + 3]));
				v = __fadd2_rn(v, f2s(kTwo23));
				code0 |= (__float_as_uint(v.x) & 0x3ffu) << (10 * c);
				code1 |= (__float_as_uint(v.y) & 0x3ffu) << (10 * c);
			}
			stage[r * 32 + lane] = code0;
			stage[(r + 1) * 32 + lane] = code1;
		}
		__syncwarp();
		if (x_first + lane * 6 <= x_last) {
			const uint32_t p0 = stage[lane * 6 + 0], p1 = stage[lane * 6 + 1], p2 = stage[lane * 6 + 2], p3 = stage[lane * 6 + 3],
			               p4 = stage[lane * 6 + 4], p5 = stage[lane * 6 + 5];
			uint4 w;   // v210.ts:158-163: chroma from even pixels only
			w.x = (p0 & 0x3ff00000u) | (p0 & 0x3ffu) << 10 | ((p0 >> 10) & 0x3ffu);
			w.y = (p2 & 0x3ffu) << 20 | (p2 & 0xffc00u) | (p1 & 0x3ffu);
			w.z = ((p4 >> 10) & 0x3ffu) << 20 | (p3 & 0x3ffu) << 10 | (p2 >> 20);
			w.w = (p5 & 0x3ffu) << 20 | ((p4 >> 20) << 10) | (p4 & 0x3ffu);
			st_stream(reinterpret_cast<uint4 *>(reinterpret_cast<char *>(d.out) + (size_t)y * d.out_pitch) + strip * d.strip_groups + lane, w);
		}
		__syncwarp();
	}
}

}  // namespace

cudaError_t launch_lut_fit(cudaStream_t s, const float *table, const LutParams *cands_dev, int n_cands, uint8_t *d8_out, void *results_dev) {
	lut_fit_kernel<<<dim3(65536 / 256, n_cands), 256, 0, s>>>(table, cands_dev, d8_out, reinterpret_cast<LutFitResult *>(results_dev));
	return cudaGetLastError();
}

size_t march_smem_bytes(const FusedDesc &d) { return (size_t)d.n_luts * 65536 + (size_t)kMarchWarps * kRowFloats * sizeof(float); }

cudaError_t launch_fused_march(cudaStream_t s, const FusedDesc &d, int num_sms) {
	const size_t smem = march_smem_bytes(d);
	auto launch = [&](void (*kernel)(const FusedDesc)) -> cudaError_t {
		// opt in to > 48 KiB of dynamic shared memory once per (kernel, device): the attribute is per context
		static std::mutex mu;
		static std::set<std::pair<const void *, int>> configured;
		int dev = 0;
		cudaGetDevice(&dev);
		{
			std::lock_guard<std::mutex> lk(mu);
			if (!configured.count({(const void *)kernel, dev})) {
				cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
				if (e != cudaSuccess) return e;
				configured.insert({(const void *)kernel, dev});
			}
		}
		const int n_lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
		const int total = n_lines * d.n_strips;
		const int grid = max(1, min(num_sms, (total + kMarchWarps - 1) / kMarchWarps));
		kernel<<<grid, kMarchThreads, smem, s>>>(d);
		return cudaGetLastError();
	};
	const bool single = d.n_rc == 1;
	if (d.n_luts > 0) {
		if (d.sparse_cm) return single ? launch(k_fused_march<1, true, true>) : launch(k_fused_march<1, true, false>);
		return single ? launch(k_fused_march<1, false, true>) : launch(k_fused_march<1, false, false>);
	}
	if (d.sparse_cm) return single ? launch(k_fused_march<0, true, true>) : launch(k_fused_march<0, true, false>);
	return single ? launch(k_fused_march<0, false, true>) : launch(k_fused_march<0, false, false>);
}

}  // namespace pb
