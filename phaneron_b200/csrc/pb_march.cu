// pb_march.cu -- the fast fused kernel: warp-autonomous strip marching.
//
//   N layers x (v210 unpack -> YCbCr->R'G'B' -> gamma LUT -> gamut 3x3 -> bilinear Transform
//   -> dissolve | wipe) -> combine (premultiplied over) -> linear->gamma LUT -> RGB->YCbCr
//   -> 10-bit RTE -> v210 pack, one launch, every packed source byte read once from HBM.
// Reference stages replaced: v210.ts:25-195, transform.ts:36-59, transition.ts:60-73,
// combine.ts:24-68 and the RGBA-f32 frames between them.
//
// Shape of the kernel:
//   * one persistent CTA per SM, 16 warps; the gamma tables live in shared memory in the lossless
//     one-byte-per-entry form of pb_lut.cuh (128 KiB for a read + a write table);
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
// Results are bit-identical to the generic kernel (pb_fused.cu) and to the oracle.
#include <mutex>
#include <set>
#include <utility>

#include "pb_device.cuh"
#include "pb_launch.h"
#include "pb_lut.cuh"

namespace pb {

namespace {

constexpr int kRounds = 6;   // 192 px / 32 lanes

template <int kLutMode>
struct LutRef {
	const int8_t *d8;   // shared memory (mode 1)
	const float *raw;   // global memory (mode 0)
	LutParams lp;       // warp-uniform, from the constant bank
};

// v in [0, 1] (already saturated) -> table value.  convert_ushort_sat_rte(v * 65535) == RNE(sat(v) * 65535):
// both ends of the clamp are fixed points of the multiply, and NaN saturates to 0 either way.
template <int kLutMode>
__device__ __forceinline__ float lut_lookup(float v_sat, const LutRef<kLutMode> &lut) {
	const float u = add(mul(v_sat, 65535.0f), 8388608.0f);   // RNE to integer in the low mantissa bits
	const uint32_t idx = __float_as_uint(u) & 0xFFFFu;
	if (kLutMode == 0) return __ldg(lut.raw + idx);
	return lut_decode(sub(u, 8388608.0f), idx, lut.d8, lut.lp);
}

// one pixel: 10-bit codes (as exact floats) -> linear RGB in the working gamut (v210.ts:65-77)
template <int kLutMode, bool kSparse>
__device__ __forceinline__ void convert_px(float fy, float fcb, float fcr, const ReadConsts &rc, const LutRef<kLutMode> &lut, float &R,
                                           float &G, float &B) {
	float tr = mul(fy, rc.cm[0]);
	if (!kSparse) tr = fma_(fcb, rc.cm[1], tr);   // cm[1] == 0: fma(cb, 0, t) == t
	tr = fma_(fcr, rc.cm[2], tr);
	float tg = mul(fy, rc.cm[4]);
	tg = fma_(fcb, rc.cm[5], tg);
	tg = fma_(fcr, rc.cm[6], tg);
	float tb = mul(fy, rc.cm[8]);
	tb = fma_(fcb, rc.cm[9], tb);
	if (!kSparse) tb = fma_(fcr, rc.cm[10], tb);
	// fma(1.0, m3, t) == RN(t + m3); the saturate is the front half of convert_ushort_sat_rte
	const float r = lut_lookup<kLutMode>(__saturatef(add(tr, rc.cm[3])), lut);
	const float g = lut_lookup<kLutMode>(__saturatef(add(tg, rc.cm[7])), lut);
	const float b = lut_lookup<kLutMode>(__saturatef(add(tb, rc.cm[11])), lut);
	R = dot3(r, g, b, rc.gamut + 0);
	G = dot3(r, g, b, rc.gamut + 3);
	B = dot3(r, g, b, rc.gamut + 6);
}

// source groups [g_lo, g_lo + ng) of one row -> the warp's planar row buffer
template <int kLutMode, bool kSparse>
__device__ __forceinline__ void convert_row(const Leaf &lf, const ReadConsts &rc, const LutRef<kLutMode> &lut, int row, int g_lo, int ng,
                                            float *buf, int lane) {
	const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(lf.ptr) + (size_t)row * lf.pitch) + g_lo;
#pragma unroll 1
	for (int g = lane; g < ng; g += 32) {
		const uint4 w = ld_stream(src + g);
		float2 *pr = reinterpret_cast<float2 *>(buf + 0 * kRowCap + g * 6);
		float2 *pg = reinterpret_cast<float2 *>(buf + 1 * kRowCap + g * 6);
		float2 *pb_ = reinterpret_cast<float2 *>(buf + 2 * kRowCap + g * 6);
		float2 R, G, B;
		{   // pixels 0,1: Cb0 Y0 Cr0 | Y1
			const float cb = u2f(w.x & 0x3ff), cr = u2f((w.x >> 20) & 0x3ff);
			convert_px<kLutMode, kSparse>(u2f((w.x >> 10) & 0x3ff), cb, cr, rc, lut, R.x, G.x, B.x);
			convert_px<kLutMode, kSparse>(u2f(w.y & 0x3ff), cb, cr, rc, lut, R.y, G.y, B.y);
			pr[0] = R; pg[0] = G; pb_[0] = B;
		}
		{   // pixels 2,3: Cb2 Y2 | Cr2 Y3
			const float cb = u2f((w.y >> 10) & 0x3ff), cr = u2f(w.z & 0x3ff);
			convert_px<kLutMode, kSparse>(u2f((w.y >> 20) & 0x3ff), cb, cr, rc, lut, R.x, G.x, B.x);
			convert_px<kLutMode, kSparse>(u2f((w.z >> 10) & 0x3ff), cb, cr, rc, lut, R.y, G.y, B.y);
			pr[1] = R; pg[1] = G; pb_[1] = B;
		}
		{   // pixels 4,5: Cb4 | Y4 Cr4 Y5
			const float cb = u2f((w.z >> 20) & 0x3ff), cr = u2f((w.w >> 10) & 0x3ff);
			convert_px<kLutMode, kSparse>(u2f(w.w & 0x3ff), cb, cr, rc, lut, R.x, G.x, B.x);
			convert_px<kLutMode, kSparse>(u2f((w.w >> 20) & 0x3ff), cb, cr, rc, lut, R.y, G.y, B.y);
			pr[2] = R; pg[2] = G; pb_[2] = B;
		}
	}
}

// value of one leaf at the 6 pixels of this lane -> p[r] = (r, g, b, alpha)
template <int kLutMode, bool kSparse>
__device__ __forceinline__ void eval_leaf(const FusedDesc &d, const Leaf &lf, const int8_t *lut_s, float *buf, int lane, int strip, int y,
                                          int x_first, int x_last, float4 (&p)[kRounds]) {
#pragma unroll
	for (int r = 0; r < kRounds; ++r) p[r] = make_float4(0.f, 0.f, 0.f, 0.f);
	const int4 si = __ldg(lf.strip_tab + strip);
	if (!(si.x & 1)) return;   // the strip does not touch this leaf's image: border colour everywhere
	const bool edge = (si.x & 2) != 0;
	const int g_lo = si.y, ng = si.z, origin = g_lo * 6, last = ng * 6 - 1;
	const int2 rt = __ldg(lf.row_tab + y);
	const int j0 = rt.x;
	const float b = __int_as_float(rt.y), rb = sub(1.0f, b);
	const ReadConsts &rc = d.rc[lf.rc];
	LutRef<kLutMode> lut;
	lut.raw = rc.lut;
	lut.d8 = lut_s + (kLutMode ? rc.lut_slot * 65536 : 0);
	lut.lp = d.luts[kLutMode ? rc.lut_slot : 0].lp;
	const int nrows = lf.has_xf ? 2 : 1;
#pragma unroll 1
	for (int rr = 0; rr < nrows; ++rr) {
		const int row = j0 + rr;
		if ((unsigned)row >= (unsigned)lf.h) continue;   // border row: all its taps are (0,0,0,0)
		convert_row<kLutMode, kSparse>(lf, rc, lut, row, g_lo, ng, buf, lane);
		__syncwarp();
		if (!lf.has_xf) {   // 1:1 read of texel (x, y): exact passthrough, alpha = 1 (leaf_value in pb_device.cuh)
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const int c = min(x_first + r * 32 + lane, x_last) - origin;
				p[r] = make_float4(buf[c], buf[kRowCap + c], buf[2 * kRowCap + c], 1.0f);
			}
		} else {
			const float wr = rr == 0 ? rb : b;
#pragma unroll
			for (int r = 0; r < kRounds; ++r) {
				const int2 ct = __ldg(lf.col_tab + min(x_first + r * 32 + lane, x_last));
				const int i0 = ct.x;
				const float a = __int_as_float(ct.y);
				const float w0 = mul(sub(1.0f, a), wr), w1 = mul(a, wr);   // w00|w01 , w10|w11
				int c0 = i0 - origin, c1 = c0 + 1;
				bool f0 = true, f1 = true;
				if (edge) {
					f0 = (unsigned)i0 < (unsigned)lf.w;
					f1 = (unsigned)(i0 + 1) < (unsigned)lf.w;
					c0 = min(max(c0, 0), last);
					c1 = min(max(c1, 0), last);
				}
				const float t0r = f0 ? buf[c0] : 0.0f, t0g = f0 ? buf[kRowCap + c0] : 0.0f, t0b = f0 ? buf[2 * kRowCap + c0] : 0.0f;
				const float t1r = f1 ? buf[c1] : 0.0f, t1g = f1 ? buf[kRowCap + c1] : 0.0f, t1b = f1 ? buf[2 * kRowCap + c1] : 0.0f;
				p[r].x = fma_(w1, t1r, fma_(w0, t0r, p[r].x));
				p[r].y = fma_(w1, t1g, fma_(w0, t0g, p[r].y));
				p[r].z = fma_(w1, t1b, fma_(w0, t0b, p[r].z));
				// alpha taps are 1 inside the image and 0 on the border: fma(w, 1, al) = RN(w + al), fma(w, 0, al) = al
				float al = p[r].w;
				al = f0 ? add(w0, al) : al;
				al = f1 ? add(w1, al) : al;
				p[r].w = al;
			}
		}
		__syncwarp();
	}
}

template <int kLutMode, bool kSparse>
__global__ void __launch_bounds__(kMarchThreads, 1) k_fused_march(const __grid_constant__ FusedDesc d) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	int8_t *lut_s = reinterpret_cast<int8_t *>(smem_raw);
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

	LutRef<kLutMode> wlut;
	wlut.raw = d.wc.lut;
	wlut.d8 = lut_s + (kLutMode ? d.wc.lut_slot * 65536 : 0);
	wlut.lp = d.luts[kLutMode ? d.wc.lut_slot : 0].lp;

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
				eval_leaf<kLutMode, kSparse>(d, lf, lut_s, buf, lane, strip, y, x_first, x_last, p);
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

		// ---- encode (v210.ts:145-156) and regroup 6 pixels -> 4 words through the row buffer ----------------
		uint32_t *stage = reinterpret_cast<uint32_t *>(buf);
#pragma unroll
		for (int r = 0; r < kRounds; ++r) {
			const float gr = lut_lookup<kLutMode>(__saturatef(acc[r].x), wlut);
			const float gg = lut_lookup<kLutMode>(__saturatef(acc[r].y), wlut);
			const float gb = lut_lookup<kLutMode>(__saturatef(acc[r].z), wlut);
			const uint32_t cy = sat_rte_u16(dot4(gr, gg, gb, 1.0f, d.wc.cm + 0));
			const uint32_t cb = sat_rte_u16(dot4(gr, gg, gb, 1.0f, d.wc.cm + 4));
			const uint32_t cr = sat_rte_u16(dot4(gr, gg, gb, 1.0f, d.wc.cm + 8));
			stage[r * 32 + lane] = cy | (cb << 10) | (cr << 20);   // codes <= 1023 (checked on the host)
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

cudaError_t launch_lut_fit(cudaStream_t s, const float *table, const LutParams *cands_dev, int n_cands, int8_t *d8_out, void *results_dev) {
	lut_fit_kernel<<<dim3(65536 / 256, n_cands), 256, 0, s>>>(table, cands_dev, d8_out, reinterpret_cast<LutFitResult *>(results_dev));
	return cudaGetLastError();
}

size_t march_smem_bytes(const FusedDesc &d) { return (size_t)d.n_luts * 65536 + (size_t)kMarchWarps * kRowFloats * sizeof(float); }

cudaError_t launch_fused_march(cudaStream_t s, const FusedDesc &d, int num_sms) {
	const size_t smem = march_smem_bytes(d);
	auto launch = [&](auto kernel) -> cudaError_t {
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
	if (d.n_luts > 0) return d.sparse_cm ? launch(k_fused_march<1, true>) : launch(k_fused_march<1, false>);
	return d.sparse_cm ? launch(k_fused_march<0, true>) : launch(k_fused_march<0, false>);
}

}  // namespace pb
