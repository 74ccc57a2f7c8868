// pb_strip.cu -- the fast fused kernel: "marching strips".
//
// One CTA owns a strip of 192 output pixels (32 v210 groups = one 512-byte coalesced
// store per line) and marches down a band of output lines.  For every leaf (packed source
// frame) it keeps a 2-row ring of CONVERTED source texels (linear RGB, fp32, planar) in
// shared memory:
//
//   phase C  rows of each leaf that the next output line samples and that are not in the
//            ring yet are converted: 128-bit loads of v210 groups, bit unpack, YCbCr->R'G'B'
//            (packed fp32x2 FMAs across the two pixels that share a chroma pair), gamma LUT,
//            gamut matrix, 64-bit conflict-free shared stores.  Each source texel is
//            converted ONCE per strip (the generic kernel converts it once per bilinear tap).
//   phase S  one thread per output pixel: bilinear taps from the rings with the exact
//            per-column / per-row {i0, a} / {j0, b} tables the host derived from the
//            reference's float formula, transition, N-layer over, linear->gamma LUT,
//            RGB->YCbCr, 10-bit RTE; codes are regrouped through shared memory so that 32
//            threads pack 6 pixels each into one coalesced 16-byte store per group.
//
// Results are bit-identical to the generic kernel (same canonical float semantics); the
// tables make that true by construction for the sampling positions and weights.
// Eligibility is decided on the host (pb_runtime.cu: prepare_strip).
#include "pb_device.cuh"
#include "pb_launch.h"

namespace pb {

struct StripLeafInfo {   // per CTA, per ring leaf (shared memory)
	int active;   // strip overlaps the source image horizontally
	int g_lo;     // first source group held by the ring rows
	int ng;       // groups per ring row
	int pad;
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }

// LUT index for a pair: sat(t) * 65535 -> RNE integer (sat before the multiply gives the
// same result as convert_ushort_sat_rte after it: both ends of the clamp are fixed points)
__device__ __forceinline__ void lut_pair(const float *__restrict__ lut, float t0, float t1, float &o0, float &o1) {
	const float2 u = __fmul2_rn(f2(__saturatef(t0), __saturatef(t1)), f2s(65535.0f));
	const float2 v = __fadd2_rn(u, f2s(8388608.0f));
	o0 = __ldg(lut + (__float_as_uint(v.x) & 0xFFFFu));
	o1 = __ldg(lut + (__float_as_uint(v.y) & 0xFFFFu));
}

// two horizontally adjacent pixels sharing one chroma pair -> linear RGB (v210.ts:65-77)
__device__ __forceinline__ void convert_pair(uint32_t y0, uint32_t y1, uint32_t cb, uint32_t cr, const ReadConsts &rc,
                                             float2 &R, float2 &G, float2 &B) {
	const float2 Y = f2(u2f(y0), u2f(y1));
	const float2 CB = f2s(u2f(cb)), CR = f2s(u2f(cr));
	float2 tr = __fmul2_rn(Y, f2s(rc.cm[0]));
	float2 tg = __fmul2_rn(Y, f2s(rc.cm[4]));
	float2 tb = __fmul2_rn(Y, f2s(rc.cm[8]));
	tr = __ffma2_rn(CB, f2s(rc.cm[1]), tr);
	tg = __ffma2_rn(CB, f2s(rc.cm[5]), tg);
	tb = __ffma2_rn(CB, f2s(rc.cm[9]), tb);
	tr = __ffma2_rn(CR, f2s(rc.cm[2]), tr);
	tg = __ffma2_rn(CR, f2s(rc.cm[6]), tg);
	tb = __ffma2_rn(CR, f2s(rc.cm[10]), tb);
	tr = __fadd2_rn(tr, f2s(rc.cm[3]));   // fma(1.0, m3, t) == RN(t + m3)
	tg = __fadd2_rn(tg, f2s(rc.cm[7]));
	tb = __fadd2_rn(tb, f2s(rc.cm[11]));
	float2 r, g, b;
	lut_pair(rc.lut, tr.x, tr.y, r.x, r.y);
	lut_pair(rc.lut, tg.x, tg.y, g.x, g.y);
	lut_pair(rc.lut, tb.x, tb.y, b.x, b.y);
	R = __ffma2_rn(b, f2s(rc.gamut[2]), __ffma2_rn(g, f2s(rc.gamut[1]), __fmul2_rn(r, f2s(rc.gamut[0]))));
	G = __ffma2_rn(b, f2s(rc.gamut[5]), __ffma2_rn(g, f2s(rc.gamut[4]), __fmul2_rn(r, f2s(rc.gamut[3]))));
	B = __ffma2_rn(b, f2s(rc.gamut[8]), __ffma2_rn(g, f2s(rc.gamut[7]), __fmul2_rn(r, f2s(rc.gamut[6]))));
}

// convert source row `row` of a leaf, groups [g_lo + gi] for gi = lane, lane+32, ..., into ring slot
__device__ __noinline__ void convert_row(const Leaf &lf, const ReadConsts &rc, int row, int g_lo, int ng, float *ring_slot, int lane,
                                         int gi_begin, int gi_step) {
	const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(lf.ptr) + (size_t)row * lf.pitch) + g_lo;
	for (int gi = gi_begin + lane; gi < ng; gi += gi_step) {
		const uint4 w = ld_stream(src + gi);
		float2 R0, G0, B0, R1, G1, B1, R2, G2, B2;
		convert_pair((w.x >> 10) & 0x3ff, w.y & 0x3ff, w.x & 0x3ff, (w.x >> 20) & 0x3ff, rc, R0, G0, B0);
		convert_pair((w.y >> 20) & 0x3ff, (w.z >> 10) & 0x3ff, (w.y >> 10) & 0x3ff, w.z & 0x3ff, rc, R1, G1, B1);
		convert_pair(w.w & 0x3ff, (w.w >> 20) & 0x3ff, (w.z >> 20) & 0x3ff, (w.w >> 10) & 0x3ff, rc, R2, G2, B2);
		float2 *pr = reinterpret_cast<float2 *>(ring_slot + 0 * kRingRow + gi * 6);
		float2 *pg = reinterpret_cast<float2 *>(ring_slot + 1 * kRingRow + gi * 6);
		float2 *pb = reinterpret_cast<float2 *>(ring_slot + 2 * kRingRow + gi * 6);
		pr[0] = R0; pr[1] = R1; pr[2] = R2;
		pg[0] = G0; pg[1] = G1; pg[2] = G2;
		pb[0] = B0; pb[1] = B1; pb[2] = B2;
	}
}

// bilinear sample of a ring leaf at one output pixel (OpenCL 1.2 8.2 formula, canonical order)
__device__ __forceinline__ float4 sample_ring(const Leaf &lf, const StripLeafInfo &li, const float *ring, int x, int j0, float b) {
	float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
	if (!li.active) return o;
	const int2 ct = __ldg(lf.col_tab + x);
	const int i0 = ct.x;
	const float a = __int_as_float(ct.y);
	const bool cx0 = (unsigned)i0 < (unsigned)lf.w, cx1 = (unsigned)(i0 + 1) < (unsigned)lf.w;
	const bool ry0 = (unsigned)j0 < (unsigned)lf.h, ry1 = (unsigned)(j0 + 1) < (unsigned)lf.h;
	if (!((cx0 || cx1) && (ry0 || ry1))) return o;   // all four taps are border texels
	const float ra = sub(1.0f, a), rb = sub(1.0f, b);
	const float w00 = mul(ra, rb), w10 = mul(a, rb), w01 = mul(ra, b), w11 = mul(a, b);
	// ring columns of the two taps, clamped into the row so that a border tap never addresses
	// outside the ring (its value is replaced by the border colour 0 below)
	const int last = li.ng * 6 - 1;
	const int c0 = min(max(i0 - li.g_lo * 6, 0), last), c1 = min(max(i0 + 1 - li.g_lo * 6, 0), last);
	const float *s0 = ring + (j0 & 1) * (3 * kRingRow);
	const float *s1 = ring + ((j0 + 1) & 1) * (3 * kRingRow);
	const bool f00 = cx0 && ry0, f10 = cx1 && ry0, f01 = cx0 && ry1, f11 = cx1 && ry1;
#define TAP(flag, ptr, plane, col) ((flag) ? (ptr)[(plane) * kRingRow + (col)] : 0.0f)
	o.x = fma_(w11, TAP(f11, s1, 0, c1), fma_(w01, TAP(f01, s1, 0, c0), fma_(w10, TAP(f10, s0, 0, c1), mul(w00, TAP(f00, s0, 0, c0)))));
	o.y = fma_(w11, TAP(f11, s1, 1, c1), fma_(w01, TAP(f01, s1, 1, c0), fma_(w10, TAP(f10, s0, 1, c1), mul(w00, TAP(f00, s0, 1, c0)))));
	o.z = fma_(w11, TAP(f11, s1, 2, c1), fma_(w01, TAP(f01, s1, 2, c0), fma_(w10, TAP(f10, s0, 2, c1), mul(w00, TAP(f00, s0, 2, c0)))));
#undef TAP
	// alpha taps are 1 inside the image, 0 on the border: w*1 = w, fma(w, 0, r) = r, fma(w, 1, r) = RN(w + r)
	float al = f00 ? w00 : 0.0f;
	al = f10 ? add(w10, al) : al;
	al = f01 ? add(w01, al) : al;
	al = f11 ? add(w11, al) : al;
	o.w = al;
	return o;
}

constexpr int kStripThreads = kStripPx;
constexpr int kStripWarps = kStripThreads / 32;

__global__ void __launch_bounds__(kStripThreads) k_fused_strip(const __grid_constant__ FusedDesc d) {
	extern __shared__ __align__(16) float smem[];
	float *rings = smem;                                                      // [n_ring][2][3][kRingRow]
	uint32_t *stage = reinterpret_cast<uint32_t *>(rings + (size_t)d.n_ring * 6 * kRingRow);   // [kStripPx]
	StripLeafInfo *info = reinterpret_cast<StripLeafInfo *>(stage + kStripPx);                 // [n_ring]

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int x_first = blockIdx.x * kStripPx;
	const int x_last = min(x_first + kStripPx, d.out_w) - 1;
	const int x = x_first + tid;
	const bool px_valid = x < d.out_w;
	const int xs = px_valid ? x : x_last;

	const int step = d.interlace == 0 ? 1 : 2;
	const int first_line = d.interlace == 3 ? 1 : 0;
	const int n_lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
	const int k_begin = blockIdx.y * d.band_lines;
	const int k_end = min(k_begin + d.band_lines, n_lines);

	// per-strip footprint of every ring leaf
	if (tid < 3 * d.n_layers) {
		const Layer &ly = d.layers[tid / 3];
		const int which = tid % 3;
		const Leaf &lf = which == 0 ? ly.a : (which == 1 ? ly.b : ly.mask);
		if (lf.kind != LEAF_NONE && lf.ring >= 0) {
			const int ia = __ldg(lf.col_tab + x_first).x, ib = __ldg(lf.col_tab + x_last).x;
			int c_lo = min(ia, ib), c_hi = max(ia, ib) + 1;
			StripLeafInfo li;
			li.active = !(c_hi < 0 || c_lo >= lf.w);
			c_lo = max(c_lo, 0);
			c_hi = min(c_hi, lf.w - 1);
			li.g_lo = c_lo / 6;
			li.ng = li.active ? (c_hi / 6 - li.g_lo + 1) : 0;
			li.pad = 0;
			info[lf.ring] = li;
		}
	}
	__syncthreads();

	for (int k = k_begin; k < k_end; ++k) {
		const int y = first_line + k * step;
		// ---- phase C: bring the source rows this line samples into the rings -------------------
		int chunk = 0;
		for (int l = 0; l < d.n_layers; ++l) {
			const Layer &ly = d.layers[l];
			const int nleaf = ly.kind == LAYER_DIRECT ? 1 : (ly.kind == LAYER_DISSOLVE ? 2 : 3);
			for (int q = 0; q < nleaf; ++q) {
				const Leaf &lf = q == 0 ? ly.a : (q == 1 ? ly.b : ly.mask);
				const StripLeafInfo li = info[lf.ring];
				if (!li.active) continue;
				const int j0 = __ldg(lf.row_tab + y).x;
				const int jp = (k > k_begin) ? __ldg(lf.row_tab + (y - step)).x : INT_MIN / 2;
				const int nchunks = (li.ng + 31) >> 5;
#pragma unroll
				for (int rr = 0; rr < 2; ++rr) {
					const int row = j0 + rr;
					if ((unsigned)row >= (unsigned)lf.h) continue;     // border row: never sampled with weight on a texel
					if (row == jp || row == jp + 1) continue;           // still resident from the previous line
					for (int c = 0; c < nchunks; ++c, ++chunk) {
						if (chunk % kStripWarps != warp) continue;
						convert_row(lf, d.rc[lf.rc], row, li.g_lo, li.ng, rings + ((size_t)lf.ring * 2 + (row & 1)) * 3 * kRingRow, lane, c * 32, 1 << 30);
					}
				}
			}
		}
		__syncthreads();
		// ---- phase S: one output pixel per thread ---------------------------------------------------
		float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
		for (int l = 0; l < d.n_layers; ++l) {
			const Layer &ly = d.layers[l];
			float4 v;
			{
				const int2 rt = __ldg(ly.a.row_tab + y);
				v = sample_ring(ly.a, info[ly.a.ring], rings + (size_t)ly.a.ring * 6 * kRingRow, xs, rt.x, __int_as_float(rt.y));
			}
			if (ly.kind != LAYER_DIRECT) {
				const int2 rtb = __ldg(ly.b.row_tab + y);
				const float4 vb = sample_ring(ly.b, info[ly.b.ring], rings + (size_t)ly.b.ring * 6 * kRingRow, xs, rtb.x, __int_as_float(rtb.y));
				if (ly.kind == LAYER_DISSOLVE) {
					v = dissolve4(v, vb, ly.mix);
				} else {
					const int2 rtm = __ldg(ly.mask.row_tab + y);
					const float4 vm = sample_ring(ly.mask, info[ly.mask.ring], rings + (size_t)ly.mask.ring * 6 * kRingRow, xs, rtm.x, __int_as_float(rtm.y));
					v = wipe_mask4(v, vb, vm.x);
				}
			}
			acc = (l == 0) ? v : over4(acc, v);
		}
		const Ycc c = linear_to_ycc(acc.x, acc.y, acc.z, d.wc);
		stage[tid] = c.y | (c.cb << 10) | (c.cr << 20);
		__syncthreads();
		// ---- pack: 32 threads, 6 pixels each -> one 16-byte store per v210 group ------------------------
		if (tid < 32 && x_first + tid * 6 < d.out_w) {
			const uint32_t p0 = stage[tid * 6 + 0], p1 = stage[tid * 6 + 1], p2 = stage[tid * 6 + 2], p3 = stage[tid * 6 + 3],
			               p4 = stage[tid * 6 + 4], p5 = stage[tid * 6 + 5];
			uint4 w;   // v210.ts:158-163: chroma from even pixels only
			w.x = ((p0 >> 20) & 0x3ff) << 20 | (p0 & 0x3ff) << 10 | ((p0 >> 10) & 0x3ff);
			w.y = (p2 & 0x3ff) << 20 | ((p2 >> 10) & 0x3ff) << 10 | (p1 & 0x3ff);
			w.z = ((p4 >> 10) & 0x3ff) << 20 | (p3 & 0x3ff) << 10 | ((p2 >> 20) & 0x3ff);
			w.w = (p5 & 0x3ff) << 20 | ((p4 >> 20) & 0x3ff) << 10 | (p4 & 0x3ff);
			st_stream(reinterpret_cast<uint4 *>(reinterpret_cast<char *>(d.out) + (size_t)y * d.out_pitch) + blockIdx.x * 32 + tid, w);
		}
		// the next phase C overwrites only ring rows this line no longer needed... but a fast warp
		// could start converting while a slow one still samples: the barrier above orders that,
		// because every warp has finished sampling before it arrives there.
	}
}

size_t strip_smem_bytes(const FusedDesc &d) {
	return (size_t)d.n_ring * 6 * kRingRow * sizeof(float) + kStripPx * sizeof(uint32_t) + (size_t)max(d.n_ring, 1) * sizeof(StripLeafInfo);
}

cudaError_t launch_fused_strip(cudaStream_t s, const FusedDesc &d) {
	static bool attr_set = false;
	const size_t smem = strip_smem_bytes(d);
	if (!attr_set || smem > 48 * 1024) {
		cudaError_t e = cudaFuncSetAttribute(k_fused_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		if (e != cudaSuccess) return e;
		attr_set = true;
	}
	const int n_lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
	dim3 grid((d.out_w + kStripPx - 1) / kStripPx, (n_lines + d.band_lines - 1) / d.band_lines);
	k_fused_strip<<<grid, kStripThreads, smem, s>>>(d);
	return cudaGetLastError();
}

}  // namespace pb
