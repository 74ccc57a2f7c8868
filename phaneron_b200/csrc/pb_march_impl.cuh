// pb_march_impl.cuh -- the fast fused kernel (warp-autonomous strip marching): device code and kernel templates.
// Included by the translation units that instantiate them: pb_march.cu (the fast, dedicated and first-pass kernels),
// pb_march_general.cu (planar-format leaves / sinks) and pb_march_bigrows.cu (64-group rows, rgba8 / RGBA-f32 / Yadif leaves):
// the instantiations compile side by side.
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
//   * a work item is one output line of one strip (15 or 16 v210 groups = 90 / 96 px, 3 pixels per
//     lane); items are dealt round-robin to all warps of the grid, so a warp never waits on another
//     warp: no __syncthreads after the table load, only __syncwarp.  Narrow strips keep the per-lane
//     pixel state small (registers, not occupancy, were the limit of wider strips: profiles/);
//   * per leaf the warp converts the strip's source footprint ONCE (lane = v210 group: one 128-bit
//     load, 6 texels) into its private planar row buffer.  At scale >= ~1 a row needs <= 16 groups,
//     so lanes 0-15 convert row j0 and lanes 16-31 row j0+1 in ONE full-width pass and the 4-tap
//     bilinear chain runs at once; smaller scales (<= 32 groups per row) take one pass per row and
//     continue the oracle's canonical chain (w00*t00 -> +w10*t10 -> +w01*t01 -> +w11*t11) across them;
//   * every lane then takes its taps for 3 output pixels (lane = pixel, stride-1 conflict-free
//     LDS) with the exact {i0, a} / {j0, b} tables the host derived from the reference's float formula;
//   * the 6-pixel / 4-word v210 regroup goes through the same buffer: 32 lanes x 3 rounds of
//     codes in, lane = group out, one coalesced 16-byte store per lane.
// Instruction economy (the kernel is issue / FMA-pipe / SFU bound, not HBM bound -- DESIGN.md 4.3):
//   * the two pixels of a 4:2:2 chroma pair are converted together with packed fp32x2
//     instructions (fma.rn.f32x2: same lane rate as FFMA, half the issue slots);
//   * 10-bit fields become floats with one LOP3 (mask | 2^23 exponent); a chroma field at bit 10
//     needs no shift: it is read as 1024*c and meets a coefficient pre-divided by 1024;
//   * the 2^23 bias of the luma floats is folded into the first FMA of the matrix row (ReadK::oY);
//   * the table index, the table address and the exact float index all come from ONE add of a
//     per-table magic constant 2^23 + (shared-memory address of the table);
//   * the toe / power select of the transfer function is arithmetic (saturating FMAs), keeping
//     the half-rate ALU pipe for the unpack masks and the final integer add.
// Every float operation is an explicit IEEE round-to-nearest op or an SFU approximation that the
// table fit has already absorbed, so results are bit-identical to the generic kernel
// (pb_fused.cu) and to the oracle.
#pragma once
#include <cstdlib>
#include <mutex>
#include <set>
#include <utility>

#include "pb_device.cuh"
#include "pb_launch.h"
#include "pb_lut.cuh"


namespace pb {

namespace {

constexpr float kTwo23 = 8388608.0f;
// Keeps the three pixel pairs of a v210 group from being interleaved: fewer table lookups in flight at once means
// fewer live registers, and registers (not ILP) bound the number of resident warps (profiles/r01_kbench_warps_fence.txt).
#define PB_PAIR_FENCE() asm volatile("" ::: "memory")

// Programmatic dependent launch (PDL): consecutive frames are independent launches on one stream.  A kernel lets its successor's
// CTAs become resident as its own exit (pdl_trigger at entry), and does everything that does not depend on earlier launches --
// the TMA load of the gamma tables: 128 KiB per CTA, 19 MB per launch, constant data -- before pdl_wait, which returns when every
// earlier grid has completed and its writes are visible.  No frame data is read and nothing is written before it.  The
// successor's prologue and launch latency thus hide under the predecessor's tail.  (No-ops in a launch without the attribute.)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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

// A float* into the warp's row buffer, held as a 32-bit shared-window address.  A generic pointer makes the compiler
// re-derive the shared window base (S2UR SR_CgaCtaId ...) at every use site -- 3 % of the instructions and 8 % of
// the stall samples of the kernel (profiles/r01_march_ncu_lines.txt); an opaque 32-bit address costs one register.
struct SPtr {
	uint32_t a;
	__device__ __forceinline__ SPtr operator+(int i) const { return SPtr{a + 4u * (uint32_t)i}; }
	__device__ __forceinline__ float operator[](int i) const {
		float v;
		asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a + 4u * (uint32_t)i));
		return v;
	}
	__device__ __forceinline__ void st2(int i, float2 v) const {
		asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a + 4u * (uint32_t)i), "f"(v.x), "f"(v.y) : "memory");
	}
	__device__ __forceinline__ uint32_t ldu(int i) const {
		uint32_t v;
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a + 4u * (uint32_t)i));
		return v;
	}
	__device__ __forceinline__ void stu(int i, uint32_t v) const { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a + 4u * (uint32_t)i), "r"(v) : "memory"); }
};

// one gamma table as the kernel sees it
template <int kLutMode>
struct LutK {
	float magic;        // mode 1: 2^23 + shared-memory byte address of the d8 table (even); mode 0: 2^23
	const float *raw;   // mode 0: the raw table in global memory
	// mode 1: -0x4B000000 as run-time data (FusedDesc::lds_koff).  bits(RN(u + magic)) = 0x4B000000 + table address + index,
	// so the byte's address is bits + koff (32-bit wrap-around): a uniform-register operand of the load itself, LDS.U8 [R + UR],
	// where an immediate mask would cost a LOP3 per lookup (profiles/r02_kbench_lds_ur.txt)
	uint32_t koff;
};

// two saturated values -> two exact table values (v210.ts:68-70 / 148-150).
// convert_ushort_sat_rte(v * 65535) == RNE(sat(v) * 65535): both ends of the clamp are fixed points
// of the multiply, and NaN saturates to 0 either way.
// kAffine: the table model (LutParams::affine) as a compile-time constant: 2 = polynomial power segment (no MUFU),
// 1 = the table model has s != 1 or o != 0 (linear -> gamma direction), 0 = it has not (gamma -> linear: the
// predicated-off scale/offset FMA and its two constant loads cost issue slots all the same), -1 = decide at run time
template <int kLutMode, int kAffine = -1>
__device__ __forceinline__ float2 lut2(float2 z, const LutK<kLutMode> &k, const LutParams &lp) {
	const float2 u = mul2_unfusable(z, f2s(65535.0f));
	const float2 v = __fadd2_rn(u, f2s(k.magic));   // RNE to an integer, held in the low mantissa bits
	if (kLutMode == 0) {
		return f2(__ldg(k.raw + (__float_as_uint(v.x) & 0xFFFFu)), __ldg(k.raw + (__float_as_uint(v.y) & 0xFFFFu)));
	}
	const uint32_t d0 = lds_u8(__float_as_uint(v.x) + k.koff), d1 = lds_u8(__float_as_uint(v.y) + k.koff);
	const float2 fi = __fadd2_rn(v, f2s(-k.magic));   // the index as an exact float
	// lut_base() of pb_lut.cuh, two lanes wide
	const float2 x = __ffma2_rn(fi, f2s(lp.p), f2s(lp.q));
	float2 pw;
	if (kAffine < 0 ? lp.affine == 2 : kAffine == 2) {   // MUFU-free model: degree-7 Horner chain on the FMA pipe (LutParams::c)
		pw = f2s(lp.c[kLutPolyDeg]);
#pragma unroll
		for (int k = kLutPolyDeg - 1; k >= 0; --k) pw = __ffma2_rn(pw, x, f2s(lp.c[k]));
	} else {
		const float2 y = __fmul2_rn(f2(lg2_approx(x.x), lg2_approx(x.y)), f2s(lp.G));
		pw = f2(ex2_approx(y.x), ex2_approx(y.y));
		if (kAffine < 0 ? lp.affine == 1 : kAffine == 1) pw = __ffma2_rn(pw, f2s(lp.s), f2s(lp.o));
	}
	const float2 toe = mul2_unfusable(fi, f2s(lp.kt));   // feeds a packed add below
	const float h0 = __saturatef(add(fi.x, lp.cJ)), h1 = __saturatef(add(fi.y, lp.cJ));
	const float2 base = __ffma2_rn(f2(h0, h1), __fadd2_rn(pw, f2(-toe.x, -toe.y)), toe);
	return f2(__int_as_float(__float_as_int(base.x) + (int)d0 - 128), __int_as_float(__float_as_int(base.y) + (int)d1 - 128));
}

// Two horizontally adjacent pixels sharing one chroma pair -> linear RGB in the working gamut
// (v210.ts:65-77).  ya/yb are the exponent-trick floats 2^23 + y; cb/cr are 2^23 + s*c with s = 1
// (SCB/SCR = 0) or 1024 (= 1: a field at bit 10 taken without a shift, met by a coefficient / 1024).
template <int kLutMode, bool kSparse, int kReadAffine, int SCB, int SCR>
__device__ __forceinline__ void convert_pair(uint32_t ya, uint32_t yb, uint32_t cb, uint32_t cr, const ReadConsts &rc, const ReadK &rk,
                                             const LutK<kLutMode> &lut, const LutParams &lp, float2 &R, float2 &G, float2 &B) {
	// dot(yuva, colMatrix row) as LLVM contracts it: t = cb*m1; t = fma(y, m0, t); t = fma(cr, m2, t); t = fma(1, m3, t).
	// The cb product is shared by the two pixels; the last step is RN(t + m3), fused with the saturate below.
	const float2 Yb = f2(__uint_as_float(ya), __uint_as_float(yb));                                // 2^23 + y
	const float2 C = __fadd2_rn(f2(__uint_as_float(cb), __uint_as_float(cr)), f2s(-kTwo23));       // exact (scaled) chroma codes
	float2 tr, tg, tb;
	if (kSparse) {   // cm[1] == 0: cb*0 = 0 and fma(y, m0, 0) = RN(y*m0) = fma(2^23 + y, m0, -2^23*m0)
		tr = __ffma2_rn(Yb, f2s(rk.mY[0]), f2s(rk.oY[0]));
	} else {
		tr = __ffma2_rn(__fadd2_rn(Yb, f2s(-kTwo23)), f2s(rk.mY[0]), f2s(mul(C.x, rk.mCb[0][SCB])));
	}
	const float2 Y = __fadd2_rn(Yb, f2s(-kTwo23));
	tg = __ffma2_rn(Y, f2s(rk.mY[1]), f2s(mul(C.x, rk.mCb[1][SCB])));
	tb = __ffma2_rn(Y, f2s(rk.mY[2]), f2s(mul(C.x, rk.mCb[2][SCB])));
	// the Cr term of both pixels in one packed FMA (broadcast operands): RN(cr * m + t) per lane, as the scalar form
	tr = __ffma2_rn(f2s(C.y), f2s(rk.mCr[0][SCR]), tr);
	tg = __ffma2_rn(f2s(C.y), f2s(rk.mCr[1][SCR]), tg);
	if (!kSparse) tb = __ffma2_rn(f2s(C.y), f2s(rk.mCr[2][SCR]), tb);   // cm[10] == 0: fma(cr, 0, t) == t
	const float2 r = lut2<kLutMode, kReadAffine>(f2(__saturatef(add(tr.x, rc.cm[3])), __saturatef(add(tr.y, rc.cm[3]))), lut, lp);
	const float2 g = lut2<kLutMode, kReadAffine>(f2(__saturatef(add(tg.x, rc.cm[7])), __saturatef(add(tg.y, rc.cm[7]))), lut, lp);
	const float2 b = lut2<kLutMode, kReadAffine>(f2(__saturatef(add(tb.x, rc.cm[11])), __saturatef(add(tb.y, rc.cm[11]))), lut, lp);
	// gamut 3x3, dot(rgb, row): t = g*m1; t = fma(r, m0, t); t = fma(b, m2, t)
	R = __ffma2_rn(b, f2s(rc.gamut[2]), __ffma2_rn(r, f2s(rc.gamut[0]), __fmul2_rn(g, f2s(rc.gamut[1]))));
	G = __ffma2_rn(b, f2s(rc.gamut[5]), __ffma2_rn(r, f2s(rc.gamut[3]), __fmul2_rn(g, f2s(rc.gamut[4]))));
	B = __ffma2_rn(b, f2s(rc.gamut[8]), __ffma2_rn(r, f2s(rc.gamut[6]), __fmul2_rn(g, f2s(rc.gamut[7]))));
}

// (w & mask) | e in one LOP3: `e` (0x4B000000, FusedDesc::e_magic) arrives in a register so that
// ptxas does not split the operation around two immediates
__device__ __forceinline__ uint32_t mask_or(uint32_t w, uint32_t mask, uint32_t e) {
	uint32_t o;
	asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(o) : "r"(w), "r"(mask), "r"(e));
	return o;
}

// one v210 group (6 texels, v210.ts:58-63) -> planar row slot `row` (plane stride `cap` texels) at texel column 6g
template <int kLutMode, bool kSparse, int kReadAffine>
__device__ __forceinline__ void convert_group(const uint4 &w, int g, uint32_t E, const ReadConsts &rc, const ReadK &rk,
                                              const LutK<kLutMode> &lut, const LutParams &lp, SPtr row, int cap) {
	const SPtr pr = row + g * 6, pg = row + (cap + g * 6), pb_ = row + (2 * cap + g * 6);
	constexpr uint32_t M0 = 0x3ffu, M10 = 0xffc00u;
	float2 R, G, B;
	// word 0: Cr0 | Y0 | Cb0     word 1: Y2 | Cb1 | Y1     word 2: Cb2 | Y3 | Cr1     word 3: Y5 | Cr2 | Y4
	convert_pair<kLutMode, kSparse, kReadAffine, 0, 0>(mask_or(w.x >> 10, M0, E), mask_or(w.y, M0, E), mask_or(w.x, M0, E), mask_or(w.x >> 20, M0, E), rc, rk, lut, lp,
	                                      R, G, B);
	pr.st2(0, R); pg.st2(0, G); pb_.st2(0, B);
	PB_PAIR_FENCE();
	convert_pair<kLutMode, kSparse, kReadAffine, 1, 0>(mask_or(w.y >> 20, M0, E), mask_or(w.z >> 10, M0, E), mask_or(w.y, M10, E), mask_or(w.z, M0, E), rc, rk, lut, lp,
	                                      R, G, B);
	pr.st2(2, R); pg.st2(2, G); pb_.st2(2, B);
	PB_PAIR_FENCE();
	convert_pair<kLutMode, kSparse, kReadAffine, 0, 1>(mask_or(w.w, M0, E), mask_or(w.w >> 20, M0, E), mask_or(w.z >> 20, M0, E), mask_or(w.w, M10, E), rc, rk, lut, lp,
	                                      R, G, B);
	pr.st2(4, R); pg.st2(4, G); pb_.st2(4, B);
}

// ---- source group loads -----------------------------------------------------------------------------------------
// The conversion consumes a v210 group: 6 texels as three 10-bit fields in each of four words.  A v210 leaf loads it with
// one 128-bit access; a planar leaf (yuv422p10 / yuv422p8 / yuv420p / nv12: the FFmpegProducer formats) gathers the same 6
// luma + 3 + 3 chroma samples from its planes and lays them out the same way, so everything downstream is shared.  (8-bit
// samples simply occupy the low 8 bits of a field; the leaf's own colour matrix carries the 8-bit ranges.)
__device__ __forceinline__ uint32_t ldg_u16(const void *p) {
	uint32_t v;
	asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ uint32_t ldg_u8(const void *p) {
	uint32_t v;
	asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ uint32_t ldg_u32(const void *p) {
	uint32_t v;
	asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
// word 0: Cr0 | Y0 | Cb0     word 1: Y2 | Cb1 | Y1     word 2: Cb2 | Y3 | Cr1     word 3: Y5 | Cr2 | Y4   (v210.ts:58-63)
__device__ __forceinline__ uint4 as_v210_group(const uint32_t (&y)[6], const uint32_t (&cb)[3], const uint32_t (&cr)[3]) {
	uint4 w;
	w.x = cr[0] << 20 | y[0] << 10 | cb[0];
	w.y = y[2] << 20 | cb[1] << 10 | y[1];
	w.z = cb[2] << 20 | y[3] << 10 | cr[1];
	w.w = y[5] << 20 | cr[2] << 10 | y[4];
	return w;
}
// group g (texels 6g .. 6g+5) of source line j.  kPlanar = false: every leaf of the launch is v210 with a width that is a
// multiple of 6 (no format test, no flag).  Otherwise bit 31 of word 0 (unused by v210) flags a group that must be converted
// texel by texel with the readers' own code (convert_group_exact): the partial last group of a line whose width is not a
// multiple of 6 (1280-wide 720p: 213 groups + 2 pixels, read with the Q1 semantics of v210.ts:90-110) -- flagged WITHOUT
// touching memory, the group may straddle the end of the line -- and yuv422p10 groups holding words above 1023.
template <bool kPlanar>
__device__ __forceinline__ uint4 load_group(const Leaf &lf, int j, int g) {
	if (!kPlanar) return ld_stream(reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(lf.ptr) + (size_t)j * lf.pitch) + g);
	if (6 * g + 6 > lf.w) return make_uint4(0x80000000u, 0, 0, 0);
	if (lf.kind == LEAF_V210) {
		uint4 w = ld_stream(reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(lf.ptr) + (size_t)j * lf.pitch) + g);
		w.x &= 0x7fffffffu;   // the reference masks the fields, so stray top bits are legal input: keep them out of the flag
		return w;
	}
	const int pitch = (lf.w + 7) / 8 * 8;   // samples per luma line (yuv422p10.ts:222, yuv420p.ts:240)
	uint32_t y[6], cb[3], cr[3];
	if (lf.kind == LEAF_YUV422P10) {
		const char *Y = reinterpret_cast<const char *>(lf.ptr) + ((size_t)j * pitch + 6 * g) * 2;   // 12 bytes, 4-byte aligned
		const uint32_t y01 = ldg_u32(Y), y23 = ldg_u32(Y + 4), y45 = ldg_u32(Y + 8);
		y[0] = y01 & 0xffffu; y[1] = y01 >> 16; y[2] = y23 & 0xffffu; y[3] = y23 >> 16; y[4] = y45 & 0xffffu; y[5] = y45 >> 16;
		const char *U = reinterpret_cast<const char *>(lf.ptr_u) + ((size_t)j * (pitch / 2) + 3 * g) * 2;
		const char *V = reinterpret_cast<const char *>(lf.ptr_v) + ((size_t)j * (pitch / 2) + 3 * g) * 2;
#pragma unroll
		for (int k = 0; k < 3; ++k) { cb[k] = ldg_u16(U + 2 * k); cr[k] = ldg_u16(V + 2 * k); }
		// The planes hold 16-bit words and the reference's reader converts whatever is there (yuv422p10.ts:60-75).  A sample
		// above 1023 -- no legal stream carries one -- does not fit a 10-bit field: the group is flagged (bit 31 of word 0,
		// unused by v210) and converted sample by sample with the reader's own arithmetic (convert_group_exact).
		uint32_t top = y[0] | y[1] | y[2] | y[3] | y[4] | y[5] | cb[0] | cb[1] | cb[2] | cr[0] | cr[1] | cr[2];
		uint4 w = as_v210_group(y, cb, cr);
		if (top > 1023u) w.x = 0x80000000u;
		return w;
	}
	const char *Y = reinterpret_cast<const char *>(lf.ptr) + (size_t)j * pitch + 6 * g;   // 6 bytes, 2-byte aligned
	const uint32_t y01 = ldg_u16(Y), y23 = ldg_u16(Y + 2), y45 = ldg_u16(Y + 4);
	y[0] = y01 & 0xffu; y[1] = y01 >> 8; y[2] = y23 & 0xffu; y[3] = y23 >> 8; y[4] = y45 & 0xffu; y[5] = y45 >> 8;
	if (lf.kind == LEAF_NV12) {   // interleaved (U, V) pairs, one chroma line per line pair
		const char *C = reinterpret_cast<const char *>(lf.ptr_u) + (size_t)(j >> 1) * pitch + 6 * g;
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const uint32_t uv = ldg_u16(C + 2 * k);
			cb[k] = uv & 0xffu;
			cr[k] = uv >> 8;
		}
	} else {
		const size_t crow = (size_t)(lf.kind == LEAF_YUV420P ? (j >> 1) : j) * (pitch / 2) + 3 * g;
		const char *U = reinterpret_cast<const char *>(lf.ptr_u) + crow, *V = reinterpret_cast<const char *>(lf.ptr_v) + crow;
#pragma unroll
		for (int k = 0; k < 3; ++k) { cb[k] = ldg_u8(U + k); cr[k] = ldg_u8(V + k); }
	}
	return as_v210_group(y, cb, cr);
}

// flagged group (see load_group): every texel through pb_device.cuh leaf_texel -- the code of the stand-alone readers and
// of the generic fused kernel (Q1 for v210 line tails, raw 16-bit words for yuv422p10, zeros outside the image)
__device__ __noinline__ void convert_group_exact(const Leaf &lf, const ReadConsts *rcs, int j, int g, SPtr row, int cap, int local_g) {
	for (int p = 0; p < 6; ++p) {
		const float4 t = leaf_texel(lf, rcs, 6 * g + p, j);
		const uint32_t a = row.a + 4u * (uint32_t)(local_g * 6 + p);
		asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(t.x) : "memory");
		asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + 4u * (uint32_t)cap), "f"(t.y) : "memory");
		asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + 8u * (uint32_t)cap), "f"(t.z) : "memory");
	}
}

// ---- TMA row prefetch (fast variants) ------------------------------------------------------------------------------------
// A work item starts with two dependent round trips to memory: the leaf's table entries, then the packed source rows they
// point at.  Both are taken off the item's critical path, two items deep, without holding registers:
//   stage A (start of item i): which leaf item i + 1 starts with is known (its line mask was loaded during item i - 1);
//            lane 0 copies that leaf's {strip, row} table entries into scratch words of the warp (cp.async, LDGSTS);
//   stage B (item i, before its encode): the entries have arrived; lane 0 hands the two source rows they point at to the
//            TMA unit (cp.async.bulk into the warp's 1 KiB staging tile, completion on the warp's own mbarrier);
//   item i + 1: its first leaf finds its 128-bit groups in shared memory (and its table entries in L1).
constexpr int kPfRowBytes = 512;                 // 32 v210 groups
constexpr int kPfBytes = 2 * kPfRowBytes + 32;   // two source rows + the scratch words the table entries of the next item's first leaf land in
struct RowPf {
	uint32_t raw;      // shared-memory address of the warp's staging tile
	uint32_t bar;      // ... of its mbarrier
	uint32_t parity;   // phase the next wait is for
	int item, op;      // the (item, op) the tile was filled for; item < 0: nothing in flight
};
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	uint32_t done = 0;
	while (!done)
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint4 lds_u128(uint32_t a) {
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
	return v;
}

// value of one leaf at the 3 pixels of this lane -> p[r] = (r, g, b, alpha)
template <int kLutMode, bool kSparse, bool kSingleRc, int kReadAffine, bool kPlanar, bool kBigRows, bool kPf = false>
__device__ __forceinline__ void eval_leaf(const FusedDesc &d, const Leaf &lf, uint32_t lut_saddr, SPtr buf, int lane, int strip, int y,
                                          int x_first, int x_last, float4 (&p)[kRounds], RowPf *pf = nullptr, bool from_pf = false) {
	// (strips / lines where the whole layer is border colour never get here: FusedDesc::strip_ops & line_ops)
	const int4 si = __ldg(lf.strip_tab + strip);
	const int2 rt = __ldg(lf.row_tab + y);
	auto border = [&]() {
#pragma unroll
		for (int r = 0; r < kRounds; ++r) p[r] = make_float4(0.f, 0.f, 0.f, 0.f);
	};
	if (!(si.x & 1)) return border();   // the strip does not touch this leaf's image: border colour everywhere
	const bool edge = (si.x & 2) != 0;
	const int g_lo = si.y, ng = si.z, origin = g_lo * 6, last = ng * 6 - 1;
	const int j0 = rt.x;
	const bool has_xf = lf.has_xf != 0;
	const bool ok0 = (unsigned)j0 < (unsigned)lf.h, ok1 = has_xf && (unsigned)(j0 + 1) < (unsigned)lf.h;
	if (!ok0 && !ok1) return border();   // both rows are border rows
	const bool paired = ng <= 16;   // both rows fit one 32-lane pass

	// issue every HBM load of this leaf up front
	const uint4 z4 = make_uint4(0, 0, 0, 0);
	uint4 wa = z4, wb = z4;
	if (kPf && from_pf) {   // the rows were fetched by the TMA unit while the previous item was encoded
		mbar_wait(pf->bar, pf->parity);
		pf->parity ^= 1u;
		if (paired) {
			const int hi = lane >> 4, g = lane & 15;
			if (g < ng && (hi ? ok1 : ok0)) wa = lds_u128(pf->raw + hi * kPfRowBytes + g * 16);
		} else {
			if (lane < ng && ok0) wa = lds_u128(pf->raw + lane * 16);
			if (lane < ng && ok1) wb = lds_u128(pf->raw + kPfRowBytes + lane * 16);
		}
	} else if (paired) {   // lanes 0-15: row j0, lanes 16-31: row j0 + 1
		const int hi = lane >> 4, g = lane & 15;
		if (g < ng && (hi ? ok1 : ok0)) wa = load_group<kPlanar>(lf, j0 + hi, g_lo + g);
	} else {
		if (lane < ng && ok0) wa = load_group<kPlanar>(lf, j0, g_lo + lane);
		if (lane < ng && ok1) wb = load_group<kPlanar>(lf, j0 + 1, g_lo + lane);
	}
	// sampling columns of this lane's pixels (exact host tables): buffer column of tap 0 and the weight a
	int c0[kRounds];
	float ca[kRounds];
#pragma unroll
	for (int r = 0; r < kRounds; ++r) {
		const int x = min(x_first + r * 32 + lane, x_last);   // clamp for the ragged last strip / the unused lanes of round 2
		if (has_xf) {
			const int2 ct = __ldg(lf.col_tab + x);   // consumed after the conversion: the L2 latency hides behind it
			c0[r] = ct.x;
			ca[r] = __int_as_float(ct.y);
		} else {
			c0[r] = x;
			ca[r] = 0.0f;
		}
	}

	const float b = __int_as_float(rt.y), rb = sub(1.0f, b);
	const int rci = kSingleRc ? 0 : lf.rc;
	const ReadConsts &rc = d.rc[rci];
	const ReadK &rk = d.rk[rci];
	const int slot = kLutMode ? rc.lut_slot : 0;
	const LutParams &lp = d.luts[kSingleRc ? 0 : slot].lp;   // the host puts rc[0]'s table in slot 0
	LutK<kLutMode> lut;
	lut.raw = rc.lut;
	lut.magic = kLutMode ? kTwo23 + (float)(lut_saddr + (kSingleRc ? 0 : slot) * 65536) : kTwo23;
	lut.koff = d.lds_koff;
	const uint32_t E = d.e_magic;
	const SPtr bufo = buf + (-origin);   // row buffer addressed by source column

	if (paired) {
		constexpr int cap = 96, slot_floats = 3 * cap;   // two row slots of 16 groups
		{
			const int hi = lane >> 4, g = lane & 15;
			if (g < ng && (hi ? ok1 : ok0)) {
				if (kPlanar && (wa.x >> 31)) convert_group_exact(lf, d.rc, j0 + hi, g_lo + g, buf + hi * slot_floats, cap, g);
				else convert_group<kLutMode, kSparse, kReadAffine>(wa, g, E, rc, rk, lut, lp, buf + hi * slot_floats, cap);
			}
		}
		__syncwarp();
		if (!has_xf) {   // 1:1 read of texel (x, y): exact passthrough, alpha = 1 (leaf_value in pb_device.cuh)
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const SPtr t = bufo + c0[r];
				p[r] = make_float4(t[0], t[cap], t[2 * cap], 1.0f);
			}
		} else if (!edge && ok0 && ok1) {   // interior: all four taps are texels
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const SPtr t0 = bufo + c0[r], t1 = t0 + slot_floats;
				const float ra = sub(1.0f, ca[r]);
				const float w00 = mul(ra, rb), w10 = mul(ca[r], rb), w01 = mul(ra, b), w11 = mul(ca[r], b);
				p[r].x = fma_(w11, t1[1], fma_(w01, t1[0], fma_(w10, t0[1], mul(w00, t0[0]))));
				p[r].y = fma_(w11, t1[cap + 1], fma_(w01, t1[cap], fma_(w10, t0[cap + 1], mul(w00, t0[cap]))));
				p[r].z = fma_(w11, t1[2 * cap + 1], fma_(w01, t1[2 * cap], fma_(w10, t0[2 * cap + 1], mul(w00, t0[2 * cap]))));
				p[r].w = add(w11, add(w01, add(w10, w00)));   // alpha taps are all 1: fma(w, 1, al) = RN(w + al)
			}
		} else {   // some taps are border texels (0,0,0,0): fma(w, 0, x) = x
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const int i0 = c0[r];
				const bool fc0 = (unsigned)i0 < (unsigned)lf.w, fc1 = (unsigned)(i0 + 1) < (unsigned)lf.w;
				const bool f00 = fc0 && ok0, f10 = fc1 && ok0, f01 = fc0 && ok1, f11 = fc1 && ok1;
				const SPtr t0 = buf + min(max(i0 - origin, 0), last), t1 = buf + min(max(i0 - origin + 1, 0), last);
				const float ra = sub(1.0f, ca[r]);
				const float w00 = mul(ra, rb), w10 = mul(ca[r], rb), w01 = mul(ra, b), w11 = mul(ca[r], b);
#define PB_TAP(flag, ptr, off) ((flag) ? (ptr)[off] : 0.0f)
				p[r].x = fma_(w11, PB_TAP(f11, t1, slot_floats), fma_(w01, PB_TAP(f01, t0, slot_floats), fma_(w10, PB_TAP(f10, t1, 0), mul(w00, PB_TAP(f00, t0, 0)))));
				p[r].y = fma_(w11, PB_TAP(f11, t1, slot_floats + cap),
				              fma_(w01, PB_TAP(f01, t0, slot_floats + cap), fma_(w10, PB_TAP(f10, t1, cap), mul(w00, PB_TAP(f00, t0, cap)))));
				p[r].z = fma_(w11, PB_TAP(f11, t1, slot_floats + 2 * cap),
				              fma_(w01, PB_TAP(f01, t0, slot_floats + 2 * cap), fma_(w10, PB_TAP(f10, t1, 2 * cap), mul(w00, PB_TAP(f00, t0, 2 * cap)))));
#undef PB_TAP
				float al = f00 ? w00 : 0.0f;
				al = f10 ? add(w10, al) : al;
				al = f01 ? add(w01, al) : al;
				al = f11 ? add(w11, al) : al;
				p[r].w = al;
			}
		}
		__syncwarp();
		return;
	}

	// wide footprint (more than 16 groups per row): one pass per row, the canonical chain continues across them
	constexpr int cap = (kBigRows ? 2 : 1) * kRowGroups * 6;   // big rows: down-scales to ~0.24 (64 source groups per 90-px strip)
	border();
#pragma unroll 1
	for (int rr = 0; rr < 2; ++rr) {
		if (!(rr ? ok1 : ok0)) continue;   // border row: all its taps are (0,0,0,0)
		if (lane < ng) {
			const uint4 wr_ = rr ? wb : wa;
			if (kPlanar && (wr_.x >> 31)) convert_group_exact(lf, d.rc, j0 + rr, g_lo + lane, buf, cap, lane);
			else convert_group<kLutMode, kSparse, kReadAffine>(wr_, lane, E, rc, rk, lut, lp, buf, cap);
		}
#pragma unroll 1
		for (int g = lane + 32; g < ng; g += 32) {   // only strips wider than 96 px get here
			const uint4 w = load_group<kPlanar>(lf, j0 + rr, g_lo + g);
			if (kPlanar && (w.x >> 31)) convert_group_exact(lf, d.rc, j0 + rr, g_lo + g, buf, cap, g);
			else convert_group<kLutMode, kSparse, kReadAffine>(w, g, E, rc, rk, lut, lp, buf, cap);
		}
		__syncwarp();
		const float wr = rr == 0 ? rb : b;
		if (!edge) {
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const SPtr t = bufo + c0[r];
				const float w0 = mul(sub(1.0f, ca[r]), wr), w1 = mul(ca[r], wr);   // w00|w01 , w10|w11
				p[r].x = fma_(w1, t[1], fma_(w0, t[0], p[r].x));
				p[r].y = fma_(w1, t[cap + 1], fma_(w0, t[cap], p[r].y));
				p[r].z = fma_(w1, t[2 * cap + 1], fma_(w0, t[2 * cap], p[r].z));
				p[r].w = add(w1, add(w0, p[r].w));
			}
		} else {
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const int i0 = c0[r];
				const bool f0 = (unsigned)i0 < (unsigned)lf.w, f1 = (unsigned)(i0 + 1) < (unsigned)lf.w;
				const SPtr t0 = buf + min(max(i0 - origin, 0), last), t1 = buf + min(max(i0 - origin + 1, 0), last);
				const float w0 = mul(sub(1.0f, ca[r]), wr), w1 = mul(ca[r], wr);
				const float t0r = f0 ? t0[0] : 0.0f, t0g = f0 ? t0[cap] : 0.0f, t0b = f0 ? t0[2 * cap] : 0.0f;
				const float t1r = f1 ? t1[0] : 0.0f, t1g = f1 ? t1[cap] : 0.0f, t1b = f1 ? t1[2 * cap] : 0.0f;
				p[r].x = fma_(w1, t1r, fma_(w0, t0r, p[r].x));
				p[r].y = fma_(w1, t1g, fma_(w0, t0g, p[r].y));
				p[r].z = fma_(w1, t1b, fma_(w0, t0b, p[r].z));
				float al = p[r].w;
				al = f0 ? add(w0, al) : al;
				al = f1 ? add(w1, al) : al;
				p[r].w = al;
			}
		}
		__syncwarp();
	}
}

// ---- Lanczos leaves (extension; definition in oracle/oracle.c) ----------------------------------------------
// value = sum_j wy_j * (sum_i wx_i * T(i0 + i, j0 + j)): ascending fma chains from +0, weights and first taps from the
// host-built tables (Leaf::lz_*), texels outside the image (0,0,0,0) -- pb_device.cuh lanczos_sample, bit for bit.  One
// conversion pass per source row of the vertical support, like the wide bilinear form: the row is converted once into the
// warp's row buffer, every lane takes the horizontal taps of its pixels from it and continues its vertical chain.
// (A border texel leaves a chain as it is -- fma(w, 0, s) == s up to the sign of a zero -- so taps and rows outside the image
// are skipped.)
template <int kLutMode, bool kSparse, bool kSingleRc, int kReadAffine, bool kBigRows>
__device__ __noinline__ void eval_leaf_lanczos(const FusedDesc &d, const Leaf &lf, uint32_t lut_saddr, SPtr buf, int lane, int strip, int y, int x_first,
                                               int x_last, float4 (&p)[kRounds]) {
#pragma unroll
	for (int r = 0; r < kRounds; ++r) p[r] = make_float4(0.f, 0.f, 0.f, 0.f);
	const int4 si = __ldg(lf.strip_tab + strip);
	if (!(si.x & 1)) return;
	const int g_lo = si.y, ng = si.z, origin = g_lo * 6;
	const int tx = lf.lz_tx, ty = lf.lz_ty, j0 = __ldg(lf.lz_j0 + y);
	int i0[kRounds];
	const float *wxp[kRounds];   // this pixel's weights in the tap-major table: tap i at wxp[r][i * W]
	const int xw = lf.xf_w;
#pragma unroll
	for (int r = 0; r < kRounds; ++r) {
		const int x = min(x_first + r * 32 + lane, x_last);
		i0[r] = __ldg(lf.lz_i0 + x);
		wxp[r] = lf.lz_wxt + x;
	}
	const int rci = kSingleRc ? 0 : lf.rc;
	const ReadConsts &rc = d.rc[rci];
	const ReadK &rk = d.rk[rci];
	const int slot = kLutMode ? rc.lut_slot : 0;
	const LutParams &lp = d.luts[kSingleRc ? 0 : slot].lp;
	LutK<kLutMode> lut;
	lut.raw = rc.lut;
	lut.magic = kLutMode ? kTwo23 + (float)(lut_saddr + (kSingleRc ? 0 : slot) * 65536) : kTwo23;
	lut.koff = d.lds_koff;
	const uint32_t E = d.e_magic;
	constexpr int cap = (kBigRows ? 2 : 1) * kRowGroups * 6;
	const SPtr bufo = buf + (-origin);   // row buffer addressed by source column
	const float *wy = lf.lz_wy + (size_t)y * ty;
	const bool edge = (si.x & 2) != 0;   // some tap column of the strip lies outside the image
	// the alpha taps are all 1 inside the image: the horizontal alpha sum of a pixel is the same for every source row
	float hw[kRounds];
#pragma unroll
	for (int r = 0; r < kRounds; ++r) hw[r] = 0.f;
#pragma unroll 4
	for (int i = 0; i < tx; ++i) {
#pragma unroll
		for (int r = 0; r < kRounds; ++r)
			if (!edge || (unsigned)(i0[r] + i) < (unsigned)lf.w) hw[r] = fma_(__ldg(wxp[r] + (size_t)i * xw), 1.0f, hw[r]);
	}
	// software pipeline down the vertical support: the groups of the next source row are loaded while this row is converted
	// and sampled (a row costs one HBM round trip; with 12 of them per line the latency would otherwise add up)
	const int j_lo = max(0, -j0), j_hi = min(ty, lf.h - j0);   // rows of the support inside the image
	uint4 w_next = make_uint4(0, 0, 0, 0);
	if (j_lo < j_hi && lane < ng) w_next = load_group<true>(lf, j0 + j_lo, g_lo + lane);
#pragma unroll 1
	for (int j = j_lo; j < j_hi; ++j) {
		const int row = j0 + j;
		const uint4 w_cur = w_next;
		if (j + 1 < j_hi && lane < ng) w_next = load_group<true>(lf, row + 1, g_lo + lane);
		if (lane < ng) {
			if (w_cur.x >> 31) convert_group_exact(lf, d.rc, row, g_lo + lane, buf, cap, lane);
			else convert_group<kLutMode, kSparse, kReadAffine>(w_cur, lane, E, rc, rk, lut, lp, buf, cap);
		}
#pragma unroll 1
		for (int g = lane + 32; g < ng; g += 32) {   // (footprints wider than 32 groups: the big-row variants)
			const uint4 w = load_group<true>(lf, row, g_lo + g);
			if (w.x >> 31) convert_group_exact(lf, d.rc, row, g_lo + g, buf, cap, g);
			else convert_group<kLutMode, kSparse, kReadAffine>(w, g, E, rc, rk, lut, lp, buf, cap);
		}
		__syncwarp();
		const float wyj = __ldg(wy + j);
		float3 h[kRounds];
#pragma unroll
		for (int r = 0; r < kRounds; ++r) h[r] = make_float3(0.f, 0.f, 0.f);
		// the three pixels of a lane advance tap by tap together: three independent fma chains per channel in flight
		if (!edge) {
#pragma unroll 4
			for (int i = 0; i < tx; ++i) {
#pragma unroll
				for (int r = 0; r < kRounds; ++r) {
					const float w = __ldg(wxp[r] + (size_t)i * xw);
					const SPtr t = bufo + (i0[r] + i);
					h[r].x = fma_(w, t[0], h[r].x);
					h[r].y = fma_(w, t[cap], h[r].y);
					h[r].z = fma_(w, t[2 * cap], h[r].z);
				}
			}
		} else {
#pragma unroll 2
			for (int i = 0; i < tx; ++i) {
#pragma unroll
				for (int r = 0; r < kRounds; ++r) {
					const int cidx = i0[r] + i;
					if ((unsigned)cidx >= (unsigned)lf.w) continue;   // border texel
					const float w = __ldg(wxp[r] + (size_t)i * xw);
					const SPtr t = bufo + cidx;
					h[r].x = fma_(w, t[0], h[r].x);
					h[r].y = fma_(w, t[cap], h[r].y);
					h[r].z = fma_(w, t[2 * cap], h[r].z);
				}
			}
		}
#pragma unroll
		for (int r = 0; r < kRounds; ++r) {
			p[r].x = fma_(wyj, h[r].x, p[r].x);
			p[r].y = fma_(wyj, h[r].y, p[r].y);
			p[r].z = fma_(wyj, h[r].z, p[r].z);
			p[r].w = fma_(wyj, hw[r], p[r].w);
		}
		__syncwarp();
	}
}

// second pass of a separable Lanczos Transform: the vertical chain over the horizontally filtered rows H that k_lanczos_hpass
// wrote (one float4 per output column and source row).  No row buffer: every tap is one coalesced 16-byte load per lane.
__device__ __forceinline__ void eval_leaf_lanczos_v(const Leaf &lf, int lane, int y, int x_first, int x_last, float4 (&p)[kRounds]) {
	const int j0 = __ldg(lf.lz_j0 + y);
	const float *wy = lf.lz_wy + (size_t)y * lf.lz_ty;
	const float4 *H[kRounds];
#pragma unroll
	for (int r = 0; r < kRounds; ++r) {
		p[r] = make_float4(0.f, 0.f, 0.f, 0.f);
		H[r] = reinterpret_cast<const float4 *>(lf.ptr) + (size_t)j0 * lf.w + min(x_first + r * 32 + lane, x_last);
	}
	const int j_lo = max(0, -j0), j_hi = min(lf.lz_ty, lf.h - j0);
#pragma unroll 2
	for (int j = j_lo; j < j_hi; ++j) {
		const float w = __ldg(wy + j);
#pragma unroll
		for (int r = 0; r < kRounds; ++r) {
			const float4 hv = __ldg(H[r] + (size_t)j * lf.w);
			p[r].x = fma_(w, hv.x, p[r].x);
			p[r].y = fma_(w, hv.y, p[r].y);
			p[r].z = fma_(w, hv.z, p[r].z);
			p[r].w = fma_(w, hv.w, p[r].w);
		}
	}
}

// ---- RGBA-f32 leaves (routed channel frames, materialised sub-expressions, de-interlaced fields) ---------------------------
// Nothing to convert, so nothing to stage: every lane fetches the four taps of each of its pixels straight from global memory
// (L1 serves the overlap between neighbouring lanes and between the two rows), all twelve 16-byte loads of a leaf in flight at
// once.  The row-buffer form this replaces (four planes through shared memory, one pass per source row) exposed the load
// latency four times per leaf and item and needed the 64-group buffers: 220 us for the composite of a 2160p de-interlaced field.
// Same weights, same fma chain as eval_leaf_rgba: w(1-a)(1-b) t00 -> + a(1-b) t10 -> + (1-a)b t01 -> + ab t11, from +0.
// A de-interlaced field (yadifCl.ts:105-167) is two frames: the lines of its own parity are the current frame's (ptr), the
// interpolated ones come from the launch's pre-pass (k_yadif_rows), row j >> 1 of ptr_u.
__device__ __forceinline__ void eval_leaf_f32(const Leaf &lf, const ReadConsts *, int lane, int strip, int y, int x_first, int x_last, float4 (&p)[kRounds]) {
	if (lf.has_xf == 2) {   // rotation / shear: the position of every pixel from the matrix, as the generic kernel does (pb_device.cuh leaf_value)
#pragma unroll
		for (int r = 0; r < kRounds; ++r) {
			const float2 pos = transform_pos(lf.m, min(x_first + r * 32 + lane, x_last), y, lf.xf_w, lf.xf_h);
			p[r] = sample_linear_clamp(lf.w, lf.h, pos.x, pos.y, [&](int i, int j) {   // (leaf_texel restricted to the two kinds this function sees)
				if (i < 0 || j < 0 || i >= lf.w || j >= lf.h) return make_float4(0.f, 0.f, 0.f, 0.f);
				return (lf.kind == LEAF_YADIF && (j & 1) != (lf.yadif & 1)) ? __ldg(reinterpret_cast<const float4 *>(lf.ptr_u) + (size_t)(j >> 1) * lf.w + i)
				                                                            : __ldg(reinterpret_cast<const float4 *>(lf.ptr) + (size_t)j * lf.w + i);
			});
		}
		return;
	}
#pragma unroll
	for (int r = 0; r < kRounds; ++r) p[r] = make_float4(0.f, 0.f, 0.f, 0.f);
	const int4 si = __ldg(lf.strip_tab + strip);
	const int2 rt = __ldg(lf.row_tab + y);
	if (!(si.x & 1)) return;
	const int j0 = rt.x;
	const bool has_xf = lf.has_xf != 0;
	const bool ok0 = (unsigned)j0 < (unsigned)lf.h, ok1 = has_xf && (unsigned)(j0 + 1) < (unsigned)lf.h;
	if (!ok0 && !ok1) return;
	auto row = [&](int j) {
		return (lf.kind == LEAF_YADIF && (j & 1) != (lf.yadif & 1)) ? reinterpret_cast<const float4 *>(lf.ptr_u) + (size_t)(j >> 1) * lf.w
		                                                            : reinterpret_cast<const float4 *>(lf.ptr) + (size_t)j * lf.w;
	};
	const float4 *r0 = row(ok0 ? j0 : j0 + 1), *r1 = row(ok1 ? j0 + 1 : j0);
	const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
	if (!has_xf) {   // 1:1 read of texel (x, y): exact passthrough, alpha included
#pragma unroll
		for (int r = 0; r < kRounds; ++r) p[r] = __ldg(r0 + min(x_first + r * 32 + lane, x_last));
		return;
	}
	int i0[kRounds];
	float ca[kRounds];
	float4 t00[kRounds], t10[kRounds], t01[kRounds], t11[kRounds];
#pragma unroll
	for (int r = 0; r < kRounds; ++r) {
		const int2 ct = __ldg(lf.col_tab + min(x_first + r * 32 + lane, x_last));
		i0[r] = ct.x;
		ca[r] = __int_as_float(ct.y);
	}
#pragma unroll
	for (int r = 0; r < kRounds; ++r) {
		const bool f0 = (unsigned)i0[r] < (unsigned)lf.w, f1 = (unsigned)(i0[r] + 1) < (unsigned)lf.w;
		t00[r] = (ok0 && f0) ? __ldg(r0 + i0[r]) : zero;
		t10[r] = (ok0 && f1) ? __ldg(r0 + i0[r] + 1) : zero;
		t01[r] = (ok1 && f0) ? __ldg(r1 + i0[r]) : zero;
		t11[r] = (ok1 && f1) ? __ldg(r1 + i0[r] + 1) : zero;
	}
	const float b = __int_as_float(rt.y), rb = sub(1.0f, b);
#pragma unroll
	for (int r = 0; r < kRounds; ++r) {
		const float ra = sub(1.0f, ca[r]);
		if (ok0) {
			const float w0 = mul(ra, rb), w1 = mul(ca[r], rb);
			p[r].x = fma_(w1, t10[r].x, fma_(w0, t00[r].x, p[r].x));
			p[r].y = fma_(w1, t10[r].y, fma_(w0, t00[r].y, p[r].y));
			p[r].z = fma_(w1, t10[r].z, fma_(w0, t00[r].z, p[r].z));
			p[r].w = fma_(w1, t10[r].w, fma_(w0, t00[r].w, p[r].w));
		}
		if (ok1) {
			const float w0 = mul(ra, b), w1 = mul(ca[r], b);
			p[r].x = fma_(w1, t11[r].x, fma_(w0, t01[r].x, p[r].x));
			p[r].y = fma_(w1, t11[r].y, fma_(w0, t01[r].y, p[r].y));
			p[r].z = fma_(w1, t11[r].z, fma_(w0, t01[r].z, p[r].z));
			p[r].w = fma_(w1, t11[r].w, fma_(w0, t01[r].w, p[r].w));
		}
	}
}

// ---- rgba8 / bgra8 leaves (graphics with alpha: FFmpegProducer 'rgba' / 'bgra' / any rgb format, rgba8.ts) ----------
// Four planes (the alpha of these sources is data and is sampled like a colour channel), one pass per source row, every tap
// validated.  Conversion is lane = texel (coalesced 128-byte loads, 3 texels per lane in flight): a texel costs four table
// reads at 256 distinct indices and the 3x3 gamut matrix.
__device__ __forceinline__ void eval_leaf_rgba(const FusedDesc &d, const Leaf &lf, SPtr buf, uint32_t t256_saddr, int lane, int strip, int y, int x_first,
                                               int x_last, float4 (&p)[kRounds]) {
	constexpr int cap = kRowGroups * 6;   // 4 planes x 192 texels x 4 B = 3 KiB: the big row buffers (kernel variants with kBigRows)
#pragma unroll
	for (int r = 0; r < kRounds; ++r) p[r] = make_float4(0.f, 0.f, 0.f, 0.f);
	const int4 si = __ldg(lf.strip_tab + strip);
	const int2 rt = __ldg(lf.row_tab + y);
	if (!(si.x & 1)) return;
	const int g_lo = si.y, ng = si.z, origin = g_lo * 6, last = ng * 6 - 1;
	const int j0 = rt.x;
	const bool has_xf = lf.has_xf != 0;
	const bool ok0 = (unsigned)j0 < (unsigned)lf.h, ok1 = has_xf && (unsigned)(j0 + 1) < (unsigned)lf.h;
	if (!ok0 && !ok1) return;
	int c0[kRounds];
	float ca[kRounds];
#pragma unroll
	for (int r = 0; r < kRounds; ++r) {
		const int x = min(x_first + r * 32 + lane, x_last);
		if (has_xf) {
			const int2 ct = __ldg(lf.col_tab + x);
			c0[r] = ct.x;
			ca[r] = __int_as_float(ct.y);
		} else {
			c0[r] = x;
			ca[r] = 0.0f;
		}
	}
	const float b = __int_as_float(rt.y), rb = sub(1.0f, b);
	const ReadConsts &rc = d.rc[lf.rc];
	const bool bgra = lf.kind == LEAF_BGRA8;
	const SPtr tab{t256_saddr + (uint32_t)rc.t256_slot * 1024u};   // tab[c] = gammaLut[c * 257] (see rgba8_to_linear in pb_device.cuh)
	const int ntex = min(ng * 6, lf.w - origin);   // texels of the footprint that exist
#pragma unroll 1
	for (int rr = 0; rr < 2; ++rr) {
		if (!(rr ? ok1 : ok0)) continue;
		const uchar4 *line = reinterpret_cast<const uchar4 *>(lf.ptr) + (size_t)(j0 + rr) * lf.w + origin;   // (rgba8 / bgra8)
#pragma unroll 1
		for (int base = 0; base < ntex; base += 96) {
			uchar4 px[3];
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				const int t = base + k * 32 + lane;
				px[k] = t < ntex ? __ldg(line + t) : make_uchar4(0, 0, 0, 0);
			}
#pragma unroll
			for (int k = 0; k < 3; ++k) {
				const int t = base + k * 32 + lane;
				if (t < ntex) {
					const float lr = tab[bgra ? px[k].z : px[k].x], lg = tab[px[k].y], lb = tab[bgra ? px[k].x : px[k].z];
					float4 v;
					v.x = dot3(lr, lg, lb, rc.gamut + 0);
					v.y = dot3(lr, lg, lb, rc.gamut + 3);
					v.z = dot3(lr, lg, lb, rc.gamut + 6);
					v.w = tab[px[k].w];   // alpha goes through the LUT too (rgba8.ts:61)
					const uint32_t a = buf.a + 4u * (uint32_t)t;
					asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v.x) : "memory");
					asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + 4u * cap), "f"(v.y) : "memory");
					asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + 8u * cap), "f"(v.z) : "memory");
					asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + 12u * cap), "f"(v.w) : "memory");
				}
			}
		}
		__syncwarp();
		if (!has_xf) {   // 1:1 read of texel (x, y): exact passthrough, alpha included
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const SPtr t = buf + (c0[r] - origin);
				p[r] = make_float4(t[0], t[cap], t[2 * cap], t[3 * cap]);
			}
		} else {
			const float wr = rr == 0 ? rb : b;
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const int i0 = c0[r];
				const bool f0 = (unsigned)i0 < (unsigned)lf.w, f1 = (unsigned)(i0 + 1) < (unsigned)lf.w;
				const SPtr t0 = buf + min(max(i0 - origin, 0), last), t1 = buf + min(max(i0 - origin + 1, 0), last);
				const float w0 = mul(sub(1.0f, ca[r]), wr), w1 = mul(ca[r], wr);
				p[r].x = fma_(w1, f1 ? t1[0] : 0.0f, fma_(w0, f0 ? t0[0] : 0.0f, p[r].x));
				p[r].y = fma_(w1, f1 ? t1[cap] : 0.0f, fma_(w0, f0 ? t0[cap] : 0.0f, p[r].y));
				p[r].z = fma_(w1, f1 ? t1[2 * cap] : 0.0f, fma_(w0, f0 ? t0[2 * cap] : 0.0f, p[r].z));
				p[r].w = fma_(w1, f1 ? t1[3 * cap] : 0.0f, fma_(w0, f0 ? t0[3 * cap] : 0.0f, p[r].w));
			}
		}
		__syncwarp();
	}
}

// ---- one output group (6 pixels) from its staged codes: word k = Y | Cb << 10 | Cr << 20 of pixel k -----------------------
// kSinks = false: v210 only (the fast variants).  Otherwise also the planar consumer formats (FFmpegConsumer's yuv422p8 and its
// siblings): the same codes, stored by plane.  Chroma comes from the even pixels (v210.ts:158-163, yuv422p10.ts:170-171);
// 4:2:0 keeps one chroma line per line pair, written from the first line of the pair the launch processes (yuv420p.ts:160-200).
template <bool kSinks>
__device__ __forceinline__ void store_group(const FusedDesc &d, int y, int G, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t p4, uint32_t p5) {
	if (kSinks && d.sink != SINK_V210) {
		const int pitch = (d.out_w + 7) / 8 * 8;
		const uint32_t y0 = p0 & 0x3ffu, y1 = p1 & 0x3ffu, y2 = p2 & 0x3ffu, y3 = p3 & 0x3ffu, y4 = p4 & 0x3ffu, y5 = p5 & 0x3ffu;
		const uint32_t u0 = (p0 >> 10) & 0x3ffu, u1 = (p2 >> 10) & 0x3ffu, u2 = (p4 >> 10) & 0x3ffu;
		const uint32_t v0 = p0 >> 20, v1 = p2 >> 20, v2 = p4 >> 20;
		if (d.sink == SINK_YUV422P10) {
			uint32_t *Y = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(d.out) + ((size_t)y * pitch + 6 * G) * 2);
			Y[0] = y0 | y1 << 16; Y[1] = y2 | y3 << 16; Y[2] = y4 | y5 << 16;
			uint16_t *U = reinterpret_cast<uint16_t *>(d.out_u) + (size_t)y * (pitch / 2) + 3 * G;
			uint16_t *V = reinterpret_cast<uint16_t *>(d.out_v) + (size_t)y * (pitch / 2) + 3 * G;
			U[0] = (uint16_t)u0; U[1] = (uint16_t)u1; U[2] = (uint16_t)u2;
			V[0] = (uint16_t)v0; V[1] = (uint16_t)v1; V[2] = (uint16_t)v2;
		} else {
			uint16_t *Y = reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(d.out) + (size_t)y * pitch + 6 * G);
			Y[0] = (uint16_t)(y0 | y1 << 8); Y[1] = (uint16_t)(y2 | y3 << 8); Y[2] = (uint16_t)(y4 | y5 << 8);
			if (d.sink == SINK_YUV422P8) {
				uint8_t *U = reinterpret_cast<uint8_t *>(d.out_u) + (size_t)y * (pitch / 2) + 3 * G;
				uint8_t *V = reinterpret_cast<uint8_t *>(d.out_v) + (size_t)y * (pitch / 2) + 3 * G;
				U[0] = (uint8_t)u0; U[1] = (uint8_t)u1; U[2] = (uint8_t)u2;
				V[0] = (uint8_t)v0; V[1] = (uint8_t)v1; V[2] = (uint8_t)v2;
			} else if ((y & 1) == (d.interlace == 3 ? 1 : 0)) {   // 4:2:0: the pair's first processed line carries the chroma
				if (d.sink == SINK_YUV420P) {
					uint8_t *U = reinterpret_cast<uint8_t *>(d.out_u) + (size_t)(y >> 1) * (pitch / 2) + 3 * G;
					uint8_t *V = reinterpret_cast<uint8_t *>(d.out_v) + (size_t)(y >> 1) * (pitch / 2) + 3 * G;
					U[0] = (uint8_t)u0; U[1] = (uint8_t)u1; U[2] = (uint8_t)u2;
					V[0] = (uint8_t)v0; V[1] = (uint8_t)v1; V[2] = (uint8_t)v2;
				} else {   // SINK_NV12
					uint16_t *C = reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(d.out_u) + (size_t)(y >> 1) * pitch + 6 * G);
					C[0] = (uint16_t)(u0 | v0 << 8); C[1] = (uint16_t)(u1 | v1 << 8); C[2] = (uint16_t)(u2 | v2 << 8);
				}
			}
		}
		return;
	}
	uint4 w;   // v210.ts:158-163
	w.x = (p0 & 0x3ff00000u) | (p0 & 0x3ffu) << 10 | ((p0 >> 10) & 0x3ffu);
	w.y = (p2 & 0x3ffu) << 20 | (p2 & 0xffc00u) | (p1 & 0x3ffu);
	w.z = ((p4 >> 10) & 0x3ffu) << 20 | (p3 & 0x3ffu) << 10 | (p2 >> 20);
	w.w = (p5 & 0x3ffu) << 20 | ((p4 >> 20) << 10) | (p4 & 0x3ffu);
	st_stream(reinterpret_cast<uint4 *>(reinterpret_cast<char *>(d.out) + (size_t)y * d.out_pitch) + G, w);
}

// ---- k_march_single: ONE v210 layer through an axis-aligned Transform with vertical scale >= 1 into a v210 output -------
// (a channel playing one full-frame clip through its Mixer: mixer.ts always runs the Transform, identity included).
// The general kernel hands every output line of a 90-px strip to another warp, so each source row is converted twice (once
// as row j0 + 1 of line y, once as row j0 of line y + 1).  Here a warp walks down a block of consecutive lines of a 186-px
// strip and keeps the last two converted rows: every line costs ONE conversion pass with all 32 lanes busy, then the
// same taps (the edge-aware form of eval_leaf) and the same encode for its 6 pixels per lane.  Bit for bit the same results.
constexpr int kSingleRowFloats = 2 * 3 * 192 + 96;            // two row slots (3 planes x 192 texels) + 96 staging words
// the item loop of k_march_single.  kMasked: run as the background pass of the general kernel (same launch): only the
// lines FusedDesc::line_pairs marks for this strip pair -- those on which the bottom layer is the only live op of both strips
// kPlanarSrc: the layer is a planar 4:2:2 / 4:2:0 source (an FFmpegProducer clip): groups come through load_group<true>,
// flagged groups (yuv422p10 words above 1023) through convert_group_exact
template <bool kMasked, bool kPlanarSrc = false, int kReadMode = 0>
__device__ __forceinline__ void march_single_items(const FusedDesc &d, SPtr buf, uint32_t lut_saddr, int lane, int warp) {
	const Leaf &lf = d.layers[0].a;
	const ReadConsts &rc = d.rc[lf.rc];
	const ReadK &rk = d.rk[lf.rc];
	const LutParams &lp = d.luts[rc.lut_slot].lp;
	LutK<1> lut, wlut;
	lut.raw = rc.lut;
	lut.magic = kTwo23 + (float)(lut_saddr + rc.lut_slot * 65536);
	lut.koff = d.lds_koff;
	wlut.raw = d.wc.lut;
	wlut.magic = kTwo23 + (float)(lut_saddr + d.wc.lut_slot * 65536);
	wlut.koff = d.lds_koff;
	const LutParams &wlp = d.wlp;
	const uint32_t E = d.e_magic;
	constexpr int cap = 192, slot_floats = 3 * cap;
	const SPtr stage = buf + 2 * slot_floats;

	const int SG = d.single_strip_groups;   // 31 stand-alone, 30 (two strips of the general kernel) as its background pass
	const int groups = d.out_w / 6, n_strips = (groups + SG - 1) / SG;
	const int LB = d.single_lines, n_blocks = (d.out_h + LB - 1) / LB;
	const int total = n_strips * n_blocks;
	const int stride = gridDim.x * kMarchWarps;
	// Stand-alone the items are dealt round-robin.  As the background pass they are claimed from a counter in global memory:
	// warps reach this phase at different times and the blocks are coarse, so a static deal leaves a long tail.  The counter
	// belongs to this launch alone: the host zeroes it on the launching stream right before the launch (launch_compiled).
	int item = blockIdx.x * kMarchWarps + warp;
	auto claim = [&]() -> int {
		unsigned v = 0;
		if (lane == 0) v = atomicAdd(d.bg_counter, 1u) - d.bg_base;
		return (int)__shfl_sync(0xffffffffu, v, 0);
	};
	if (kMasked) item = claim();
#pragma unroll 1
	for (; (unsigned)item < (unsigned)total; item = kMasked ? claim() : item + stride) {
		const int blk = item / n_strips, strip = item - blk * n_strips;
		const int x_first = strip * (SG * 6), x_last = min(x_first + SG * 6, d.out_w) - 1;
		const int2 sg = d.single_strips[strip];   // first source group and group count of this strip's footprint (0 groups: all border)
		const int g_lo = sg.x, ng = sg.y & 0xff, origin = g_lo * 6, last = ng * 6 - 1;
		const bool strip_interior = (sg.y & 0x100) != 0;   // every tap column of the strip lies inside the image
		int have0 = -0x40000000, have1 = -0x40000000;   // source row held by slot 0 / slot 1
		const int y_end = min((blk + 1) * LB, d.out_h);
		// the sampling columns of this lane's 6 pixels do not change down the block
		int c0[2 * kRounds];
		float cw[2 * kRounds];
#pragma unroll
		for (int q = 0; q < 2 * kRounds; ++q) {
			const int x = min(x_first + (q / kRounds) * 96 + (q % kRounds) * 32 + lane, x_last);
			const int2 ct = __ldg(lf.col_tab + x);
			c0[q] = ct.x;
			cw[q] = __int_as_float(ct.y);
		}
		// software pipeline down the block: the row table entry of the next line and the source row that line will need
		// are loaded while this line is sampled and encoded
		int2 rt = __ldg(lf.row_tab + blk * LB);
		int2 rt_next = rt;
		uint4 w_pref = make_uint4(0, 0, 0, 0);
		int pref_row = -0x40000000;
#pragma unroll 1
		for (int y = blk * LB; y < y_end; ++y, rt = rt_next) {
			if (kMasked && !((__ldg(d.line_pairs + y) >> strip) & 1ull)) {   // not a background-only line of this strip pair
				if (y + 1 < y_end) rt_next = __ldg(lf.row_tab + y + 1);
				continue;
			}
			const int j0 = rt.x;
			const bool ok0 = ng > 0 && (unsigned)j0 < (unsigned)lf.h, ok1 = ng > 0 && (unsigned)(j0 + 1) < (unsigned)lf.h;
			if (y + 1 < y_end) rt_next = __ldg(lf.row_tab + y + 1);
#pragma unroll 1
			for (int rr = 0; rr < 2; ++rr) {   // bring in the rows this line needs and the slots do not hold yet
				const int row = j0 + rr, slot = row & 1;
				if (!(rr ? ok1 : ok0) || (slot ? have1 : have0) == row) continue;
				if (lane < ng) {
					uint4 w = w_pref;
					if (row != pref_row) w = load_group<kPlanarSrc>(lf, row, g_lo + lane);
					if (kPlanarSrc && (w.x >> 31)) convert_group_exact(lf, d.rc, row, g_lo + lane, buf + slot * slot_floats, cap, lane);
					else convert_group<1, true, kReadMode>(w, lane, E, rc, rk, lut, lp, buf + slot * slot_floats, cap);
				}
				if (slot) have1 = row; else have0 = row;
			}
			__syncwarp();
			if (y + 1 < y_end) {   // the one new row of the next line (vertical step <= 1), if any: its load flies over this line's arithmetic
				const int jn = rt_next.x;
				int want = -0x40000000;
				if ((unsigned)(jn + 1) < (unsigned)lf.h && have0 != jn + 1 && have1 != jn + 1) want = jn + 1;
				else if ((unsigned)jn < (unsigned)lf.h && have0 != jn && have1 != jn) want = jn;
				pref_row = want;
				if (want >= 0 && lane < ng) w_pref = load_group<kPlanarSrc>(lf, want, g_lo + lane);
			}
			const float b = __int_as_float(rt.y), rb = sub(1.0f, b);
			const SPtr s0 = buf + (j0 & 1) * slot_floats, s1 = buf + ((j0 + 1) & 1) * slot_floats;
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				float3 a3[kRounds];
				if (strip_interior && ok0 && ok1) {   // all four taps are texels (eval_leaf's interior form)
#pragma unroll
					for (int r = 0; r < kRounds; ++r) {
						const float ca = cw[h * kRounds + r], ra = sub(1.0f, ca);
						const SPtr t0 = s0 + (c0[h * kRounds + r] - origin), t1 = s1 + (c0[h * kRounds + r] - origin);
						const float w00 = mul(ra, rb), w10 = mul(ca, rb), w01 = mul(ra, b), w11 = mul(ca, b);
						float4 p;
						p.x = fma_(w11, t1[1], fma_(w01, t1[0], fma_(w10, t0[1], mul(w00, t0[0]))));
						p.y = fma_(w11, t1[cap + 1], fma_(w01, t1[cap], fma_(w10, t0[cap + 1], mul(w00, t0[cap]))));
						p.z = fma_(w11, t1[2 * cap + 1], fma_(w01, t1[2 * cap], fma_(w10, t0[2 * cap + 1], mul(w00, t0[2 * cap]))));
						p.w = add(w11, add(w01, add(w10, w00)));
						const float kk = sub(1.0f, p.w);   // combine.ts:49-59 over an empty frame: fma(0, 1 - alpha, p)
						a3[r] = make_float3(fma_(0.0f, kk, p.x), fma_(0.0f, kk, p.y), fma_(0.0f, kk, p.z));
					}
				} else {
#pragma unroll
				for (int r = 0; r < kRounds; ++r) {
					const int i0 = c0[h * kRounds + r];
					const float ca = cw[h * kRounds + r], ra = sub(1.0f, ca);
					const bool fc0 = (unsigned)i0 < (unsigned)lf.w, fc1 = (unsigned)(i0 + 1) < (unsigned)lf.w;
					const bool f00 = fc0 && ok0, f10 = fc1 && ok0, f01 = fc0 && ok1, f11 = fc1 && ok1;
					const int t0 = min(max(i0 - origin, 0), last), t1 = min(max(i0 - origin + 1, 0), last);
					const float w00 = mul(ra, rb), w10 = mul(ca, rb), w01 = mul(ra, b), w11 = mul(ca, b);
					const SPtr a0 = s0 + t0, a1 = s0 + t1, b0 = s1 + t0, b1 = s1 + t1;
#define PB_TAP(flag, ptr, off) ((flag) ? (ptr)[off] : 0.0f)
					float4 p;
					p.x = fma_(w11, PB_TAP(f11, b1, 0), fma_(w01, PB_TAP(f01, b0, 0), fma_(w10, PB_TAP(f10, a1, 0), mul(w00, PB_TAP(f00, a0, 0)))));
					p.y = fma_(w11, PB_TAP(f11, b1, cap), fma_(w01, PB_TAP(f01, b0, cap), fma_(w10, PB_TAP(f10, a1, cap), mul(w00, PB_TAP(f00, a0, cap)))));
					p.z = fma_(w11, PB_TAP(f11, b1, 2 * cap), fma_(w01, PB_TAP(f01, b0, 2 * cap), fma_(w10, PB_TAP(f10, a1, 2 * cap), mul(w00, PB_TAP(f00, a0, 2 * cap)))));
#undef PB_TAP
					float al = f00 ? w00 : 0.0f;
					al = f10 ? add(w10, al) : al;
					al = f01 ? add(w01, al) : al;
					al = f11 ? add(w11, al) : al;
					// combine.ts:49-59 over an empty frame: fma(0, 1 - alpha, p)
					const float kk = sub(1.0f, al);
					a3[r] = make_float3(fma_(0.0f, kk, p.x), fma_(0.0f, kk, p.y), fma_(0.0f, kk, p.z));
				}
				}
#pragma unroll
				for (int r = 0; r + 1 < kRounds; r += 2) {
					const float2 gr = lut2<1, 1>(f2(__saturatef(a3[r].x), __saturatef(a3[r + 1].x)), wlut, wlp);
					const float2 gg = lut2<1, 1>(f2(__saturatef(a3[r].y), __saturatef(a3[r + 1].y)), wlut, wlp);
					const float2 gb = lut2<1, 1>(f2(__saturatef(a3[r].z), __saturatef(a3[r + 1].z)), wlut, wlp);
					uint32_t code0 = 0, code1 = 0;
#pragma unroll
					for (int c = 0; c < 3; ++c) {
						float2 v = __ffma2_rn(gb, f2s(d.wc.cm[c * 4 + 2]), __ffma2_rn(gr, f2s(d.wc.cm[c * 4 + 0]), __fmul2_rn(gg, f2s(d.wc.cm[c * 4 + 1]))));
						v = __fadd2_rn(v, f2s(d.wc.cm[c * 4 + 3]));
						v = __fadd2_rn(v, f2s(kTwo23));
						code0 |= (__float_as_uint(v.x) & 0x3ffu) << (10 * c);
						code1 |= (__float_as_uint(v.y) & 0x3ffu) << (10 * c);
					}
					stage.stu(r * 32 + lane, code0);
					stage.stu((r + 1) * 32 + lane, code1);
				}
				if (kRounds & 1) {
					constexpr int r = kRounds - 1;
					const float2 hrg = lut2<1, 1>(f2(__saturatef(a3[r].x), __saturatef(a3[r].y)), wlut, wlp);
					const float2 hb = lut2<1, 1>(f2s(__saturatef(a3[r].z)), wlut, wlp);
					uint32_t code = 0;
#pragma unroll
					for (int c = 0; c < 3; ++c) {
						const float u = add(add(fma_(hb.x, d.wc.cm[c * 4 + 2], fma_(hrg.x, d.wc.cm[c * 4 + 0], mul(hrg.y, d.wc.cm[c * 4 + 1]))), d.wc.cm[c * 4 + 3]), kTwo23);
						code |= (__float_as_uint(u) & 0x3ffu) << (10 * c);
					}
					stage.stu(r * 32 + lane, code);
				}
				__syncwarp();
				const int xg = x_first + h * 96 + lane * 6;   // lanes 0-15 regroup and store this half's groups
				if (lane < 16 && xg <= x_last) {
					const SPtr sp = stage + lane * 6;
					const uint32_t p0 = sp.ldu(0), p1 = sp.ldu(1), p2 = sp.ldu(2), p3 = sp.ldu(3), p4 = sp.ldu(4), p5 = sp.ldu(5);
					store_group<kPlanarSrc>(d, y, xg / 6, p0, p1, p2, p3, p4, p5);   // (the <true> variant also writes the planar consumer formats)
				}
				__syncwarp();
			}
		}
	}
}

template <bool kPlanarSrc, int kReadMode>
__global__ void __launch_bounds__(kMarchThreads, 1) k_march_single(const __grid_constant__ FusedDesc d) {
	pdl_trigger();
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const uint32_t lut_saddr = (uint32_t)__cvta_generic_to_shared(smem_raw);
	uint32_t tid_x;
	asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_x));
	const int lane = tid_x & 31, warp = tid_x >> 5;
	SPtr buf;
	{
		const uint32_t addr = lut_saddr + (uint32_t)d.n_luts * 65536u + (uint32_t)warp * (kSingleRowFloats * 4u);
		asm volatile("mov.u32 %0, %1;" : "=r"(buf.a) : "r"(addr));
	}
	{
		__shared__ __align__(8) unsigned long long lut_bar;
		const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&lut_bar);
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(d.n_luts * 65536) : "memory");
			_Pragma("unroll 1") for (int tc = 0; tc < d.n_luts * 4; ++tc) {   // 16 KiB per copy; not unrolled: n_luts x 4 UBLKCP + ELECT blocks were ~15 % of the code
				const int t = tc >> 2, c = tc & 3;
					asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
					                 lut_saddr + t * 65536 + c * 16384),
					             "l"(d.luts[t].d8 + c * 16384), "r"(16384), "r"(bar)
					             : "memory");
			}
		}
		uint32_t done = 0;
		while (!done)
			asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
	}
	pdl_wait();   // from here on: frame data
	march_single_items<false, kPlanarSrc, kReadMode>(d, buf, lut_saddr, lane, warp);
}

// kPlain: 1 = every read table is a non-affine MUFU model and the write table an affine one (what colourMaths.ts produces:
// gamma -> linear on the way in, linear -> gamma on the way out), so both selects are compile-time constants; 2 = the same with
// every read table in the MUFU-free polynomial model (LutParams::affine == 2); 0 = decide per table at run time
// kPlanar: some leaf is a planar 4:2:2 / 4:2:0 source (load_group gathers it into the v210 group layout)
// kBigRows: the warps' row buffers hold 64 source groups instead of 32 (deep down-scales; needs <= 2 resident tables)
// kBg: the bottom layer is a full-frame-style v210 leaf (scale >= 1): the strip-pair lines on which it is the only live op are
// left to a second phase of the same launch, march_single_items<true> (every source row converted once: k_march_single)
// kFeat (general variants): the rarer abilities, compiled only into the instances a launch that needs one takes (FusedDesc::feat),
// so that FFmpeg-format / graphics scenes do not carry their registers: 1 = Lanczos leaves filtered inside the launch
// (eval_leaf_lanczos), 2 = RGBA-f32 / Yadif leaves (eval_leaf_f32), 4 = the RGBA-f32 sink and its alpha chain
template <int kLutMode, bool kSparse, bool kSingleRc, int kPlain = 0, bool kPlanar = false, bool kBigRows = false, bool kBg = false, int kFeat = 0>
__global__ void __launch_bounds__(((kPlanar || kBigRows) ? kGeneralWarps : kMarchWarps) * 32, 1) k_fused_march(const __grid_constant__ FusedDesc d) {
	constexpr int kW = (kPlanar || kBigRows) ? kGeneralWarps : kMarchWarps;   // warps of this variant (pb_desc.h)
	pdl_trigger();
	extern __shared__ __align__(128) unsigned char smem_raw[];
	uint8_t *lut_s = reinterpret_cast<uint8_t *>(smem_raw);
	const uint32_t lut_saddr = (uint32_t)__cvta_generic_to_shared(lut_s);
	// read %tid.x once through a volatile asm: the compiler otherwise re-reads the special register (S2R, ~20 cycles)
	// wherever lane / warp are needed again (profiles/r01_march_ncu_lines.txt)
	uint32_t tid_x;
	asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_x));
	const int lane = tid_x & 31, warp = tid_x >> 5;
	SPtr buf;   // this warp's row buffer
	{
		const uint32_t addr = lut_saddr + (kLutMode ? (uint32_t)d.n_luts * 65536u : 0u) + (uint32_t)warp * (kBg ? kSingleRowFloats * 4u : (kBigRows ? 2u : 1u) * kRowFloats * 4u);
		asm volatile("mov.u32 %0, %1;" : "=r"(buf.a) : "r"(addr));   // opaque: keep it in a register, do not re-derive it
	}

	const uint32_t t256_saddr = buf.a - (uint32_t)warp * ((kBigRows ? 2u : 1u) * kRowFloats * 4u) + (uint32_t)kW * ((kBigRows ? 2u : 1u) * kRowFloats * 4u);
	if (kBigRows && d.n_t256) {   // 256-entry tables of the rgba8 / bgra8 leaves: gammaLut[c * 257], behind the row buffers
		for (int i = 0; i < d.n_rc; ++i) {
			const int slot = d.rc[i].t256_slot;
			if (slot < 0) continue;
			for (uint32_t cidx = tid_x; cidx < 256u; cidx += (kW * 32)) {
				const float v = __ldg(d.rc[i].lut + cidx * 257u);
				asm volatile("st.shared.f32 [%0], %1;" ::"r"(t256_saddr + (uint32_t)slot * 1024u + cidx * 4u), "f"(v) : "memory");
			}
		}
		if (!kLutMode) __syncthreads();
	}
	if (kLutMode) {
		// The byte tables arrive by TMA bulk copies (cp.async.bulk, SASS UBLKCP) issued by one thread and
		// tracked by an mbarrier: 64 KiB per table without a register round trip or a per-thread loop.
		__shared__ __align__(8) unsigned long long lut_bar;
		const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&lut_bar);
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(d.n_luts * 65536) : "memory");
			_Pragma("unroll 1") for (int tc = 0; tc < d.n_luts * 4; ++tc) {   // 16 KiB per copy; not unrolled: n_luts x 4 UBLKCP + ELECT blocks were ~15 % of the code
				const int t = tc >> 2, c = tc & 3;
					asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
					                 lut_saddr + t * 65536 + c * 16384),
					             "l"(d.luts[t].d8 + c * 16384), "r"(16384), "r"(bar)
					             : "memory");
			}
		}
		uint32_t done = 0;
		while (!done)
			asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
	}
	pdl_wait();   // from here on: frame data

	const int step = d.interlace == 0 ? 1 : 2;
	const int first_line = d.interlace == 3 ? 1 : 0;
	const int n_lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
	const int total = n_lines * d.n_strips;
	const int strip_px = d.strip_groups * 6;

	LutK<kLutMode> wlut;
	wlut.raw = d.wc.lut;
	wlut.magic = kLutMode ? kTwo23 + (float)(lut_saddr + d.wc.lut_slot * 65536) : kTwo23;
	wlut.koff = d.lds_koff;
	const LutParams &wlp = d.wlp;

	// TMA row prefetch (see RowPf): the fast variants only (v210 leaves with whole groups, 32-group row buffers)
#ifdef PB_EXP_ROW_PREFETCH   // the TMA row prefetch experiment (exact, 6-9 % slower): compiled only into experiment builds
	constexpr bool kPf = kLutMode == 1 && !kPlanar && !kBigRows && !kBg;
#else
	constexpr bool kPf = false;
#endif
	__shared__ __align__(8) unsigned long long pf_bars[kPf ? kW : 1];
	RowPf pf;
	pf.raw = lut_saddr + (uint32_t)d.n_luts * 65536u + (uint32_t)kW * (kRowFloats * 4u) + (uint32_t)warp * kPfBytes;
	pf.bar = (uint32_t)__cvta_generic_to_shared(&pf_bars[kPf ? warp : 0]);
	pf.parity = 0;
	pf.item = -1;
	pf.op = 0;
	// Measured (profiles/r02_kbench_tma_prefetch_v*.txt): bit-exact, and 6-9 % SLOWER than without on the 2160p scenes -- other warps
	// already cover the load latency, what the prefetch adds is instructions.  Off unless PB_DBG=2 asks for it (A/B runs).
	const bool pf_on = kPf && (d.dbg & 2);
	uint32_t lo_next = 0;   // line mask of the next item (loaded one item ahead)
	if (kPf) {
		if (lane == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pf.bar));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncwarp();
	}

	// item -> (line k, strip) is kept incrementally: no integer division per item
	const int stride = gridDim.x * kW, stride_k = stride / d.n_strips, stride_s = stride - stride_k * d.n_strips;
	int k = (blockIdx.x * kW + warp) / d.n_strips, strip = (blockIdx.x * kW + warp) - k * d.n_strips;
	if (kPf && pf_on) {   // the line mask of this warp's second item
		int k2 = k + stride_k, strip2 = strip + stride_s;
		if (strip2 >= d.n_strips) ++k2;
		if (blockIdx.x * kW + warp + stride < total) lo_next = __ldg(d.line_ops + first_line + k2 * step);
	}
#pragma unroll 1
	for (int item = blockIdx.x * kW + warp; item < total; item += stride, k += stride_k, strip += stride_s) {
		if (strip >= d.n_strips) {
			strip -= d.n_strips;
			++k;
		}
		const int y = first_line + k * step;
		bool pf_stage = false;
		int pf_op2 = 0;   // first live op of the next item
		if (kPf && pf_on) {
			// ---- stage A: the table entries of the leaf the next item starts with; the line mask of the item after it ----
			pf.item = pf.item == item ? pf.item : -1;
			const int item2 = item + stride;
			if (item2 < total) {
				int k2 = k + stride_k, strip2 = strip + stride_s;
				if (strip2 >= d.n_strips) {
					strip2 -= d.n_strips;
					++k2;
				}
				const uint32_t both2 = d.strip_ops[strip2] & lo_next;
				uint32_t todo2 = both2 & 0xFFFFFFu;
				if (both2 >> 24) todo2 &= ~0u << d.layer_first_op[(31 - __clz(both2)) - 24];
				if (todo2) {
					const int oi2 = __ffs(todo2) - 1;
					const MarchOp &op2 = d.ops[oi2];
					const Leaf &lf2 = (&d.layers[op2.layer].a)[op2.which];
					if (lf2.kind == LEAF_V210 && lane == 0) {
						const uint32_t sc = pf.raw + 2 * kPfRowBytes;
						asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sc), "l"(lf2.strip_tab + strip2) : "memory");
						asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sc + 16), "l"(lf2.row_tab + (first_line + k2 * step)) : "memory");
					}
					pf_op2 = oi2;
					pf_stage = lf2.kind == LEAF_V210;
				}
				int k3 = k2 + stride_k, strip3 = strip2 + stride_s;
				if (strip3 >= d.n_strips) ++k3;
				if (item2 + stride < total) lo_next = __ldg(d.line_ops + first_line + k3 * step);
			}
		}
		if (kBg && ((__ldg(d.line_pairs + y) >> (strip >> 1)) & 1ull)) continue;   // a background-only line of this strip pair: second phase
		const int x_first = strip * strip_px;
		const int x_last = min(x_first + strip_px, d.march_w) - 1;   // whole output groups only: a ragged tail is the generic kernel's

		// The host flattened the layer graph into ops (a leaf evaluation + an action) and marked, per strip, the ops
		// that can touch it: leaves that lie elsewhere cost nothing here.  acc starts at 0 and every layer, the bottom
		// one included, is composited with `over`: fma(0, k, p) == p.
		float3 acc[kRounds];
		float4 p[kRounds], t[kRounds];
		float m[kRounds];
		float al[kRounds];   // general variants: the composite's alpha (combine.ts:49-59: fma(prev.a, 0, l.a)), for the RGBA-f32 sink
#pragma unroll
		for (int r = 0; r < kRounds; ++r) {
			acc[r] = make_float3(0.f, 0.f, 0.f);
			t[r] = make_float4(0.f, 0.f, 0.f, 0.f);
			m[r] = 0.f;
			al[r] = 0.f;
		}
		const uint32_t both = d.strip_ops[strip] & __ldg(d.line_ops + y);
		uint32_t todo = both & 0xFFFFFFu;
		// exact occlusion culling: the topmost layer that is opaque over this whole strip line hides all ops below it
		if (both >> 24) todo &= ~0u << d.layer_first_op[(31 - __clz(both)) - 24];
		const bool top_live = (todo >> (d.n_ops - 1)) & 1u;   // the top layer reaches this strip line (else it contributes (0,0,0,0))
#pragma unroll 1
		while (todo) {
			const int oi = __ffs(todo) - 1;
			todo &= todo - 1;
			const MarchOp &op = d.ops[oi];
			const Leaf &lf = (&d.layers[op.layer].a)[op.which];
			if (kPlanar && (kFeat & 2) && (lf.kind == LEAF_RGBA_F32 || lf.kind == LEAF_YADIF)) eval_leaf_f32(lf, d.rc, lane, strip, y, x_first, x_last, p);
			else if (kBigRows && (lf.kind == LEAF_RGBA8 || lf.kind == LEAF_BGRA8)) eval_leaf_rgba(d, lf, buf, t256_saddr, lane, strip, y, x_first, x_last, p);
			else if (lf.kind == LEAF_LANCZOS_V) eval_leaf_lanczos_v(lf, lane, y, x_first, x_last, p);
			else if (kPlanar && (kFeat & 1) && lf.lz_tx) eval_leaf_lanczos<kLutMode, kSparse, kSingleRc, (kPlain == 2 ? 2 : kPlain == 1 ? 0 : -1), kBigRows>(d, lf, lut_saddr, buf, lane, strip, y, x_first, x_last, p);
			else eval_leaf<kLutMode, kSparse, kSingleRc, (kPlain == 2 ? 2 : kPlain == 1 ? 0 : -1), kPlanar, kBigRows, kPf>(d, lf, lut_saddr, buf, lane, strip, y, x_first, x_last, p, &pf,
			                                                                                                                  kPf && pf.item == item && pf.op == oi);
			const int act = op.act;
			if (act == ACT_DIS_B) {   // transition.ts:60-65: fma(in0, mix, in1 * (1 - mix))
				const float rmix = sub(1.0f, op.mix);
#pragma unroll
				for (int r = 0; r < kRounds; ++r) t[r] = make_float4(mul(p[r].x, rmix), mul(p[r].y, rmix), mul(p[r].z, rmix), mul(p[r].w, rmix));
				continue;
			}
			if (act == ACT_WIPE_M) {   // transition.ts:66-73: fma(in1, m, in0 * (1 - m)), m = mask.r
#pragma unroll
				for (int r = 0; r < kRounds; ++r) m[r] = p[r].x;
				continue;
			}
			if (act == ACT_WIPE_A) {
#pragma unroll
				for (int r = 0; r < kRounds; ++r) {
					const float rm = sub(1.0f, m[r]);
					t[r] = make_float4(mul(p[r].x, rm), mul(p[r].y, rm), mul(p[r].z, rm), mul(p[r].w, rm));
				}
				continue;
			}
			if (act == ACT_DIS_A_OVER) {
				const float mix = op.mix;
#pragma unroll
				for (int r = 0; r < kRounds; ++r)
					p[r] = make_float4(fma_(p[r].x, mix, t[r].x), fma_(p[r].y, mix, t[r].y), fma_(p[r].z, mix, t[r].z), fma_(p[r].w, mix, t[r].w));
			} else if (act == ACT_WIPE_B_OVER) {
#pragma unroll
				for (int r = 0; r < kRounds; ++r)
					p[r] = make_float4(fma_(p[r].x, m[r], t[r].x), fma_(p[r].y, m[r], t[r].y), fma_(p[r].z, m[r], t[r].z), fma_(p[r].w, m[r], t[r].w));
			}
			// combine.ts:49-59: fma(prev, 1 - l.a, l)
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const float kk = sub(1.0f, p[r].w);
				acc[r] = make_float3(fma_(acc[r].x, kk, p[r].x), fma_(acc[r].y, kk, p[r].y), fma_(acc[r].z, kk, p[r].z));
				if (kPlanar && (kFeat & 4)) al[r] = fma_(al[r], 0.0f, p[r].w);
			}
		}

		if (kPf && pf_stage) {
			// ---- stage B: the table entries have arrived in the warp's scratch words: hand the row copies to the TMA unit ----
			if (lane == 0) asm volatile("cp.async.wait_all;" ::: "memory");
			__syncwarp();   // (also: every lane has taken its groups out of the tile)
			const uint32_t sc = pf.raw + 2 * kPfRowBytes;
			uint32_t sx, sy, sz, rx;
			asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(sx), "=r"(sy) : "r"(sc));
			asm volatile("ld.shared.u32 %0, [%1];" : "=r"(sz) : "r"(sc + 8));
			asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rx) : "r"(sc + 16));
			const Leaf &lf2 = (&d.layers[d.ops[pf_op2].layer].a)[d.ops[pf_op2].which];
			const int j2 = (int)rx;
			const bool ok0 = (unsigned)j2 < (unsigned)lf2.h, ok1 = lf2.has_xf != 0 && (unsigned)(j2 + 1) < (unsigned)lf2.h;
			if ((sx & 1u) && (ok0 || ok1)) {   // (else eval_leaf returns the border colour without loading anything)
				const uint32_t row_bytes = sz * 16u;
				const char *src = reinterpret_cast<const char *>(lf2.ptr) + (size_t)j2 * lf2.pitch + (size_t)sy * 16;
				if (lane == 0) {
					asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
					asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pf.bar), "r"(row_bytes * ((ok0 ? 1u : 0u) + (ok1 ? 1u : 0u))) : "memory");
					if (ok0)
						asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pf.raw), "l"(src), "r"(row_bytes), "r"(pf.bar) : "memory");
					if (ok1)
						asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pf.raw + kPfRowBytes), "l"(src + lf2.pitch), "r"(row_bytes), "r"(pf.bar) : "memory");
				}
				pf.item = item + stride;
				pf.op = pf_op2;
			}
		}

		if (kPlanar && (kFeat & 4) && d.sink == SINK_RGBA_F32) {   // the composite as it is: one float4 per pixel, 512 contiguous bytes per round
			float4 *o = reinterpret_cast<float4 *>(d.out) + (size_t)y * d.out_w;
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const int x = x_first + r * 32 + lane;
				if (x <= x_last) o[x] = make_float4(acc[r].x, acc[r].y, acc[r].z, top_live ? al[r] : fma_(al[r], 0.0f, 0.0f));
			}
			continue;
		}

		// ---- encode (v210.ts:145-156) and regroup 6 pixels -> 4 words through the row buffer ----
		// The host has checked that every code lies in [0, 1023] for table values in [0, 1], so
		// convert_ushort_sat_rte reduces to the RNE add and the three codes share one word.
		const SPtr stage = buf;
		const bool rgba_sink = kPlanar && (d.sink == SINK_RGBA8 || d.sink == SINK_BGRA8);   // ScreenConsumer: one word per pixel, no regroup
#pragma unroll
		for (int r = 0; r + 1 < kRounds; r += 2) {   // two rounds at a time
			const float2 gr = lut2<kLutMode, (kPlain ? 1 : -1)>(f2(__saturatef(acc[r].x), __saturatef(acc[r + 1].x)), wlut, wlp);
			const float2 gg = lut2<kLutMode, (kPlain ? 1 : -1)>(f2(__saturatef(acc[r].y), __saturatef(acc[r + 1].y)), wlut, wlp);
			const float2 gb = lut2<kLutMode, (kPlain ? 1 : -1)>(f2(__saturatef(acc[r].z), __saturatef(acc[r + 1].z)), wlut, wlp);
			if (kPlanar && rgba_sink) {   // rgba8.ts:83-101: convert_uchar_sat_rte(gamma * 255), alpha 255 (table values lie in [0, 1])
				const float2 r8 = __fadd2_rn(mul2_unfusable(gr, f2s(255.0f)), f2s(kTwo23)), g8 = __fadd2_rn(mul2_unfusable(gg, f2s(255.0f)), f2s(kTwo23)),
				             b8 = __fadd2_rn(mul2_unfusable(gb, f2s(255.0f)), f2s(kTwo23));
				const bool bgra = d.sink == SINK_BGRA8;
				uint32_t *o = reinterpret_cast<uint32_t *>(d.out) + (size_t)y * d.out_w;
				const int xa = x_first + r * 32 + lane, xb = xa + 32;
				const uint32_t ca = (__float_as_uint(bgra ? b8.x : r8.x) & 0xffu) | (__float_as_uint(g8.x) & 0xffu) << 8 |
				                    (__float_as_uint(bgra ? r8.x : b8.x) & 0xffu) << 16 | 0xff000000u;
				const uint32_t cb_ = (__float_as_uint(bgra ? b8.y : r8.y) & 0xffu) | (__float_as_uint(g8.y) & 0xffu) << 8 |
				                     (__float_as_uint(bgra ? r8.y : b8.y) & 0xffu) << 16 | 0xff000000u;
				if (xa <= x_last) o[xa] = ca;
				if (xb <= x_last) o[xb] = cb_;
				continue;
			}
			uint32_t code0 = 0, code1 = 0;
#pragma unroll
			for (int c = 0; c < 3; ++c) {   // dot(rgba, colMatrix row): t = g*m1; fma(r, m0, t); fma(b, m2, t); fma(1, m3, t) = RN(t + m3)
				float2 v = __ffma2_rn(gb, f2s(d.wc.cm[c * 4 + 2]), __ffma2_rn(gr, f2s(d.wc.cm[c * 4 + 0]), __fmul2_rn(gg, f2s(d.wc.cm[c * 4 + 1]))));
				v = __fadd2_rn(v, f2s(d.wc.cm[c * 4 + 3]));
				v = __fadd2_rn(v, f2s(kTwo23));
				code0 |= (__float_as_uint(v.x) & 0x3ffu) << (10 * c);
				code1 |= (__float_as_uint(v.y) & 0x3ffu) << (10 * c);
			}
			stage.stu(r * 32 + lane, code0);
			stage.stu((r + 1) * 32 + lane, code1);
		}
		if (kRounds & 1) {   // the odd round out: (r, g) as one pair, b alone
			constexpr int r = kRounds - 1;
			const float2 hrg = lut2<kLutMode, (kPlain ? 1 : -1)>(f2(__saturatef(acc[r].x), __saturatef(acc[r].y)), wlut, wlp);
			const float2 hb = lut2<kLutMode, (kPlain ? 1 : -1)>(f2s(__saturatef(acc[r].z)), wlut, wlp);
			if (kPlanar && rgba_sink) {
				const float r8 = add(mul(hrg.x, 255.0f), kTwo23), g8 = add(mul(hrg.y, 255.0f), kTwo23), b8 = add(mul(hb.x, 255.0f), kTwo23);
				const bool bgra = d.sink == SINK_BGRA8;
				const int xa = x_first + r * 32 + lane;
				if (xa <= x_last)
					reinterpret_cast<uint32_t *>(d.out)[(size_t)y * d.out_w + xa] = (__float_as_uint(bgra ? b8 : r8) & 0xffu) | (__float_as_uint(g8) & 0xffu) << 8 |
					                                                                 (__float_as_uint(bgra ? r8 : b8) & 0xffu) << 16 | 0xff000000u;
				continue;   // next item: nothing to regroup
			}
			uint32_t code = 0;
#pragma unroll
			for (int c = 0; c < 3; ++c) {
				const float u = add(add(fma_(hb.x, d.wc.cm[c * 4 + 2], fma_(hrg.x, d.wc.cm[c * 4 + 0], mul(hrg.y, d.wc.cm[c * 4 + 1]))), d.wc.cm[c * 4 + 3]), kTwo23);
				code |= (__float_as_uint(u) & 0x3ffu) << (10 * c);
			}
			stage.stu(r * 32 + lane, code);
		}
		__syncwarp();
		if (x_first + lane * 6 <= x_last) {
			const SPtr sp = stage + lane * 6;
			const uint32_t p0 = sp.ldu(0), p1 = sp.ldu(1), p2 = sp.ldu(2), p3 = sp.ldu(3), p4 = sp.ldu(4), p5 = sp.ldu(5);
			store_group<kPlanar>(d, y, strip * d.strip_groups + lane, p0, p1, p2, p3, p4, p5);
		}
		__syncwarp();
	}
	if (kBg) march_single_items<true, false, (kPlain == 2 ? 2 : 0)>(d, buf, lut_saddr, lane, warp);   // second phase: the background-only strip-pair lines
}


// ---- k_march_direct: one v210 source read 1:1 into a v210 output (ToRGBA -> FromRGBA, BASELINE.json config 2) ----------
// The general kernel converts a 96-px strip with lanes 0-15 only when the leaf is read 1:1 (one source row of 16 groups per
// output line).  Here a work item is one line of a 192-px strip: every lane converts one v210 group (same convert_group, same
// tables), then encodes its 3 + 3 pixels in two halves through the same staging words.  Same arithmetic, bit for bit.
#ifndef PB_DIRECT_WARPS
#define PB_DIRECT_WARPS 28
#endif
constexpr int kDirectWarps = PB_DIRECT_WARPS;   // 69 registers per thread: more resident warps than the general kernel's 20
// kRgbaOut: the converted pixels are written as an RGBA-f32 frame (a ToRGBA output made real).  kRgbaIn: the source already is
// an RGBA-f32 frame (a routed channel frame, a Yadif output, a host-written image) that FromRGBA packs: v210.ts:113-195 alone.
template <int kReadMode, bool kRgbaOut = false, bool kRgbaIn = false>
__global__ void __launch_bounds__(kDirectWarps * 32, 1) k_march_direct(const __grid_constant__ FusedDesc d) {
	pdl_trigger();
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const uint32_t lut_saddr = (uint32_t)__cvta_generic_to_shared(smem_raw);
	uint32_t tid_x;
	asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_x));
	const int lane = tid_x & 31, warp = tid_x >> 5;
	SPtr buf;
	{
		const uint32_t addr = lut_saddr + (uint32_t)d.n_luts * 65536u + (uint32_t)warp * (kRowFloats * 4u);
		asm volatile("mov.u32 %0, %1;" : "=r"(buf.a) : "r"(addr));
	}
	{
		__shared__ __align__(8) unsigned long long lut_bar;
		const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&lut_bar);
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(d.n_luts * 65536) : "memory");
			_Pragma("unroll 1") for (int tc = 0; tc < d.n_luts * 4; ++tc) {   // 16 KiB per copy; not unrolled: n_luts x 4 UBLKCP + ELECT blocks were ~15 % of the code
				const int t = tc >> 2, c = tc & 3;
					asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
					                 lut_saddr + t * 65536 + c * 16384),
					             "l"(d.luts[t].d8 + c * 16384), "r"(16384), "r"(bar)
					             : "memory");
			}
		}
		uint32_t done = 0;
		while (!done)
			asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
	}
	pdl_wait();   // from here on: frame data
	const Leaf &lf = d.layers[0].a;
	const ReadConsts &rc = d.rc[0];
	const ReadK &rk = d.rk[0];
	const LutParams &lp = d.luts[kRgbaIn ? 0 : rc.lut_slot].lp;
	LutK<1> lut, wlut;
	lut.raw = rc.lut;
	lut.magic = kTwo23 + (float)(lut_saddr + (kRgbaIn ? 0 : rc.lut_slot) * 65536);
	lut.koff = d.lds_koff;
	wlut.raw = d.wc.lut;
	wlut.magic = kTwo23 + (float)(lut_saddr + d.wc.lut_slot * 65536);
	wlut.koff = d.lds_koff;
	const LutParams &wlp = d.wlp;
	const uint32_t E = d.e_magic;
	constexpr int cap = kRowGroups * 6;   // 192 texels per plane

	const int step = d.interlace == 0 ? 1 : 2, first_line = d.interlace == 3 ? 1 : 0;
	const int n_lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
	const int groups = d.out_w / 6, n_strips = (groups + 31) / 32;
	const int total = n_lines * n_strips;
	const int stride = gridDim.x * kDirectWarps;
#pragma unroll 1
	for (int item = blockIdx.x * kDirectWarps + warp; item < total; item += stride) {
		const int k = item / n_strips, strip = item - k * n_strips;
		const int y = first_line + k * step;
		const int G = strip * 32 + lane;
		if (!kRgbaIn) {
			if (G < groups) {
				const uint4 w = ld_stream(reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(lf.ptr) + (size_t)y * lf.pitch) + G);
				convert_group<1, true, kReadMode>(w, lane, E, rc, rk, lut, lp, buf, cap);
			}
			__syncwarp();
		}
		if (kRgbaOut) {   // ToRGBA made real (v210.ts:25-111 as a frame in HBM): alpha 1, coalesced float4 stores
			float4 *o = reinterpret_cast<float4 *>(d.out) + (size_t)y * d.out_w + strip * 192;
#pragma unroll
			for (int q = 0; q < 6; ++q) {
				const int xs = q * 32 + lane;
				const SPtr t = buf + xs;
				if (strip * 192 + xs < d.out_w) o[xs] = make_float4(t[0], t[cap], t[2 * cap], 1.0f);
			}
			__syncwarp();
			continue;
		}
#pragma unroll 1
		for (int h = 0; h < 2; ++h) {
			// 1:1 read of texel (x, y): exact passthrough; over an empty frame fma(0, 0, p) == p.  The staging words below reuse
			// texels 0..95 of the first plane: each lane overwrites only what it has just read, and half 1 lives at 96..191.
			float3 a[kRounds];
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				if (kRgbaIn) {   // pixel (strip * 192 + h * 96 + r * 32 + lane) of the frame itself: 512 contiguous bytes per round
					const int x = min(strip * 192 + h * 96 + r * 32 + lane, d.out_w - 1);
					const float4 px = __ldg(reinterpret_cast<const float4 *>(lf.ptr) + (size_t)y * d.out_w + x);
					a[r] = make_float3(px.x, px.y, px.z);
				} else {
					const SPtr t = buf + (h * 96 + r * 32 + lane);
					a[r] = make_float3(t[0], t[cap], t[2 * cap]);
				}
			}
			const SPtr stage = buf;
#pragma unroll
			for (int r = 0; r + 1 < kRounds; r += 2) {
				const float2 gr = lut2<1, 1>(f2(__saturatef(a[r].x), __saturatef(a[r + 1].x)), wlut, wlp);
				const float2 gg = lut2<1, 1>(f2(__saturatef(a[r].y), __saturatef(a[r + 1].y)), wlut, wlp);
				const float2 gb = lut2<1, 1>(f2(__saturatef(a[r].z), __saturatef(a[r + 1].z)), wlut, wlp);
				uint32_t code0 = 0, code1 = 0;
#pragma unroll
				for (int c = 0; c < 3; ++c) {
					float2 v = __ffma2_rn(gb, f2s(d.wc.cm[c * 4 + 2]), __ffma2_rn(gr, f2s(d.wc.cm[c * 4 + 0]), __fmul2_rn(gg, f2s(d.wc.cm[c * 4 + 1]))));
					v = __fadd2_rn(v, f2s(d.wc.cm[c * 4 + 3]));
					v = __fadd2_rn(v, f2s(kTwo23));
					code0 |= (__float_as_uint(v.x) & 0x3ffu) << (10 * c);
					code1 |= (__float_as_uint(v.y) & 0x3ffu) << (10 * c);
				}
				stage.stu(r * 32 + lane, code0);
				stage.stu((r + 1) * 32 + lane, code1);
			}
			if (kRounds & 1) {
				constexpr int r = kRounds - 1;
				const float2 hrg = lut2<1, 1>(f2(__saturatef(a[r].x), __saturatef(a[r].y)), wlut, wlp);
				const float2 hb = lut2<1, 1>(f2s(__saturatef(a[r].z)), wlut, wlp);
				uint32_t code = 0;
#pragma unroll
				for (int c = 0; c < 3; ++c) {
					const float u = add(add(fma_(hb.x, d.wc.cm[c * 4 + 2], fma_(hrg.x, d.wc.cm[c * 4 + 0], mul(hrg.y, d.wc.cm[c * 4 + 1]))), d.wc.cm[c * 4 + 3]), kTwo23);
					code |= (__float_as_uint(u) & 0x3ffu) << (10 * c);
				}
				stage.stu(r * 32 + lane, code);
			}
			__syncwarp();
			const int Gh = strip * 32 + h * 16 + lane;   // lanes 0-15 regroup and store this half's 16 groups
			if (lane < 16 && Gh < groups) {
				const SPtr sp = stage + lane * 6;
				const uint32_t p0 = sp.ldu(0), p1 = sp.ldu(1), p2 = sp.ldu(2), p3 = sp.ldu(3), p4 = sp.ldu(4), p5 = sp.ldu(5);
				uint4 w;   // v210.ts:158-163: chroma from even pixels only
				w.x = (p0 & 0x3ff00000u) | (p0 & 0x3ffu) << 10 | ((p0 >> 10) & 0x3ffu);
				w.y = (p2 & 0x3ffu) << 20 | (p2 & 0xffc00u) | (p1 & 0x3ffu);
				w.z = ((p4 >> 10) & 0x3ffu) << 20 | (p3 & 0x3ffu) << 10 | (p2 >> 20);
				w.w = (p5 & 0x3ffu) << 20 | ((p4 >> 20) << 10) | (p4 & 0x3ffu);
				st_stream(reinterpret_cast<uint4 *>(reinterpret_cast<char *>(d.out) + (size_t)y * d.out_pitch) + Gh, w);
			}
			__syncwarp();
		}
	}
}


// ---- k_lanczos_hpass: first pass of a separable Lanczos Transform (HPassDesc) ---------------------------------------------------
// Work item = one source row x one strip of output columns: the warp converts the strip's source footprint once (lane = v210
// group, as everywhere), then every lane takes the horizontal taps of its 3 columns from the row buffer: the ascending fma
// chain of the filter's definition (oracle/oracle.c), border texels skipped (fma(w, 0, s) == s).  The result row goes to H.
template <int kReadMode>
__global__ void __launch_bounds__(kMarchThreads, 1) k_lanczos_hpass(const __grid_constant__ HPassDesc h) {
	pdl_trigger();
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const uint32_t lut_saddr = (uint32_t)__cvta_generic_to_shared(smem_raw);
	uint32_t tid_x;
	asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_x));
	const int lane = tid_x & 31, warp = tid_x >> 5;
	SPtr buf;
	{
		const uint32_t addr = lut_saddr + 65536u + (uint32_t)warp * (kRowFloats * 4u);
		asm volatile("mov.u32 %0, %1;" : "=r"(buf.a) : "r"(addr));
	}
	{
		__shared__ __align__(8) unsigned long long lut_bar;
		const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&lut_bar);
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(65536) : "memory");
			for (int c = 0; c < 4; ++c)
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(lut_saddr + c * 16384),
				             "l"(h.lut.d8 + c * 16384), "r"(16384), "r"(bar)
				             : "memory");
		}
		mbar_wait(bar, 0);
	}
	const Leaf &lf = h.lf;
	LutK<1> lut;
	lut.raw = h.rc.lut;
	lut.magic = kTwo23 + (float)lut_saddr;
	lut.koff = h.lds_koff;
	const LutParams &lp = h.lut.lp;
	const uint32_t E = h.e_magic;
	constexpr int cap = kRowGroups * 6;
	const int tx = lf.lz_tx, xw = h.xf_w, strip_px = h.strip_groups * 6;
	const int ns = h.s1 - h.s0 + 1, total = (h.j_hi - h.j_lo) * ns, stride = gridDim.x * kMarchWarps;
#pragma unroll 1
	for (int item = blockIdx.x * kMarchWarps + warp; item < total; item += stride) {
		const int jr = item / ns, strip = h.s0 + (item - jr * ns), row = h.j_lo + jr;
		const int x_first = strip * strip_px, x_last = min(x_first + strip_px, xw) - 1;
		const int4 si = __ldg(lf.strip_tab + strip);
		float4 *o = h.out + (size_t)row * xw;
		if (!(si.x & 1)) {   // no tap of this strip lies inside the image
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const int x = x_first + r * 32 + lane;
				if (x <= x_last) o[x] = make_float4(0.f, 0.f, 0.f, 0.f);
			}
			continue;
		}
		const int g_lo = si.y, ng = si.z, origin = g_lo * 6;
		const bool edge = (si.x & 2) != 0;
		if (lane < ng) {
			const uint4 w = load_group<true>(lf, row, g_lo + lane);
			if (w.x >> 31) convert_group_exact(lf, &h.rc, row, g_lo + lane, buf, cap, lane);
			else convert_group<1, true, kReadMode>(w, lane, E, h.rc, h.rk, lut, lp, buf, cap);
		}
		int i0[kRounds];
		const float *wxp[kRounds];
#pragma unroll
		for (int r = 0; r < kRounds; ++r) {
			const int x = min(x_first + r * 32 + lane, x_last);
			i0[r] = __ldg(lf.lz_i0 + x);
			wxp[r] = lf.lz_wxt + x;
		}
		__syncwarp();
		const SPtr bufo = buf + (-origin);
		float4 acc[kRounds];
#pragma unroll
		for (int r = 0; r < kRounds; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
		if (!edge) {
#pragma unroll 4
			for (int i = 0; i < tx; ++i) {
#pragma unroll
				for (int r = 0; r < kRounds; ++r) {
					const float w = __ldg(wxp[r] + (size_t)i * xw);
					const SPtr t = bufo + (i0[r] + i);
					acc[r].x = fma_(w, t[0], acc[r].x);
					acc[r].y = fma_(w, t[cap], acc[r].y);
					acc[r].z = fma_(w, t[2 * cap], acc[r].z);
					acc[r].w = fma_(w, 1.0f, acc[r].w);
				}
			}
		} else {
#pragma unroll 2
			for (int i = 0; i < tx; ++i) {
#pragma unroll
				for (int r = 0; r < kRounds; ++r) {
					const int cidx = i0[r] + i;
					if ((unsigned)cidx >= (unsigned)lf.w) continue;
					const float w = __ldg(wxp[r] + (size_t)i * xw);
					const SPtr t = bufo + cidx;
					acc[r].x = fma_(w, t[0], acc[r].x);
					acc[r].y = fma_(w, t[cap], acc[r].y);
					acc[r].z = fma_(w, t[2 * cap], acc[r].z);
					acc[r].w = fma_(w, 1.0f, acc[r].w);
				}
			}
		}
#pragma unroll
		for (int r = 0; r < kRounds; ++r) {
			const int x = x_first + r * 32 + lane;
			if (x <= x_last) o[x] = acc[r];
		}
		__syncwarp();
	}
}

}  // namespace

// opt in to > 48 KiB of dynamic shared memory once per (kernel, device) -- the attribute is per context -- and launch one
// a launch that may start while its predecessor on the stream drains (see pdl_trigger / pdl_wait); PB_NO_PDL=1 for A/B runs
template <typename Kernel>
inline cudaError_t launch_pdl(Kernel kernel, int grid, int block, size_t smem, cudaStream_t s, const FusedDesc &d) {
	static const bool pdl = getenv("PB_NO_PDL") == nullptr;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3((unsigned)grid);
	cfg.blockDim = dim3((unsigned)block);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = s;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = at;
	cfg.numAttrs = pdl ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kernel, d);
}

// persistent CTA per SM
template <typename Kernel>
inline cudaError_t march_launch(Kernel kernel, cudaStream_t s, const FusedDesc &d, int num_sms, size_t smem, int warps = kMarchWarps) {
	static std::mutex mu;
	static std::set<std::pair<const void *, int>> configured;
	int dev = 0;
	cudaGetDevice(&dev);
	{
		std::lock_guard<std::mutex> lk(mu);
		if (!configured.count({(const void *)kernel, dev})) {
			cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);   // 1 KiB left for static shared memory (the mbarriers)
			if (e != cudaSuccess) return e;
			configured.insert({(const void *)kernel, dev});
		}
	}
	const int n_lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
	const int total = n_lines * d.n_strips;
	const int grid = max(1, min(num_sms, (total + warps - 1) / warps));
	return launch_pdl(kernel, grid, warps * 32, smem, s, d);
}

// the general variants live in their own translation units
cudaError_t launch_fused_march_planar(cudaStream_t s, const FusedDesc &d, int num_sms, size_t smem, int plain, bool single);
cudaError_t launch_fused_march_bigrows(cudaStream_t s, const FusedDesc &d, int num_sms, size_t smem, int plain, bool single);

}  // namespace pb
