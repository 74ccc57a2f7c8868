// pb_kernels.cu -- stand-alone ("materialising") kernels, one per reference kernel.
// They are what runs in eager mode and whenever a deferred RGBA frame must really exist
// in HBM (host read-back, ROUTE hand-off, yadif history, rotated transforms).  The fused
// chain lives in pb_fused.cu.
#include "pb_device.cuh"
#include "pb_writers.cuh"
#include "pb_launch.h"

namespace pb {

constexpr int kThreads = 256;
static inline unsigned blocks_for(size_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

// ---- v210 read: v210.ts:25-111 --------------------------------------------------------
// one thread per 16-byte group (6 pixels)
__global__ void __launch_bounds__(kThreads) k_v210_read(const uint4 *__restrict__ in, float4 *__restrict__ out,
                                                        int width, int height, int pitch16,
                                                        const __grid_constant__ ReadConsts rc) {
	const int groups = (width + 5) / 6;
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= (size_t)groups * height) return;
	const int line = (int)(tid / groups), g = (int)(tid - (size_t)line * groups);
	const uint4 w = ld_stream(in + (size_t)line * pitch16 + g);
	const int x0 = g * 6;
	const int n = min(6, width - x0);
	const float alpha = (n < 6) ? 0.0f : 1.0f;   // Q1
	float4 *o = out + (size_t)line * width + x0;
#pragma unroll
	for (int p = 0; p < 6; ++p) {
		if (p < n) {
			const float3 rgb = ycc_to_linear(v210_px(w, p), alpha, rc);
			o[p] = make_float4(rgb.x, rgb.y, rgb.z, 1.0f);
		}
	}
}

// ---- v210 write: v210.ts:113-195 ------------------------------------------------------
// one thread per 16-byte group of the destination pitch (padding groups are cleared as
// the reference's partial last work-item does, v210.ts:131-136)
__global__ void __launch_bounds__(kThreads) k_v210_write(const float4 *__restrict__ in, uint4 *__restrict__ out,
                                                         int width, int lines, int pitch16, int interlace,
                                                         const __grid_constant__ WriteConsts wc) {
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= (size_t)pitch16 * lines) return;
	const int gl = (int)(tid / pitch16), g = (int)(tid - (size_t)gl * pitch16);
	const int line = gl * (interlace == 0 ? 1 : 2) + (interlace == 3 ? 1 : 0);
	const int x0 = g * 6;
	uint4 w = make_uint4(0, 0, 0, 0);
	if (x0 < width) {
		const float4 *src = in + (size_t)line * width + x0;
		const int n = min(6, width - x0);
		Ycc px[6];
#pragma unroll
		for (int p = 0; p < 6; ++p) {
			px[p].y = px[p].cb = px[p].cr = 0;
			if (p < n) {
				const float4 v = __ldg(src + p);
				px[p] = (n == 6) ? linear_to_ycc(v.x, v.y, v.z, wc) : linear_to_ycc_tail(v.x, v.y, v.z, wc);
			}
		}
		if (n == 6) {
			w = v210_pack(px);
		} else {
			// v210.ts:186-192
			w.x = px[0].cr << 20 | px[0].y << 10 | px[0].cb;
			if (n == 2) w.y = px[1].y;
			else if (n == 4) {
				w.y = px[2].y << 20 | px[2].cb << 10 | px[1].y;
				w.z = px[3].y << 10 | px[2].cr;
			}
		}
	} else if (width % 48 == 0) {
		return;   // no padding exists
	}
	st_stream(out + (size_t)line * pitch16 + g, w);
}

// ---- rgba8 / bgra8: rgba8.ts:25-103, bgra8.ts:25-103 ------------------------------------
__global__ void __launch_bounds__(kThreads) k_rgba8_read(const uchar4 *__restrict__ in, float4 *__restrict__ out,
                                                         size_t n, int bgra, const __grid_constant__ ReadConsts rc) {
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= n) return;
	const uchar4 v = in[tid];
	const float c0 = u2f(bgra ? v.z : v.x), c1 = u2f(v.y), c2 = u2f(bgra ? v.x : v.z), c3 = u2f(v.w);
	const float r = __ldg(rc.lut + sat_rte_u16(__fdiv_rn(mul(c0, 65535.0f), 255.0f)));
	const float g = __ldg(rc.lut + sat_rte_u16(__fdiv_rn(mul(c1, 65535.0f), 255.0f)));
	const float b = __ldg(rc.lut + sat_rte_u16(__fdiv_rn(mul(c2, 65535.0f), 255.0f)));
	float4 o;
	o.x = dot3(r, g, b, rc.gamut + 0);
	o.y = dot3(r, g, b, rc.gamut + 3);
	o.z = dot3(r, g, b, rc.gamut + 6);
	o.w = __ldg(rc.lut + sat_rte_u16(__fdiv_rn(mul(c3, 65535.0f), 255.0f)));   // alpha goes through the LUT (rgba8.ts:61)
	out[tid] = o;
}

__global__ void __launch_bounds__(kThreads) k_rgba8_write(const float4 *__restrict__ in, uchar4 *__restrict__ out,
                                                          int width, int lines, int interlace, int bgra,
                                                          const __grid_constant__ WriteConsts wc) {
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= (size_t)width * lines) return;
	const int gl = (int)(tid / width), x = (int)(tid - (size_t)gl * width);
	const int line = gl * (interlace == 0 ? 1 : 2) + (interlace == 3 ? 1 : 0);
	rgba8_write_px([&](int px, int ly) { return __ldg(in + (size_t)ly * width + px); }, out, width, line, x, bgra, wc);
}

// ---- yuv422p10le / yuv422p8: yuv422p10.ts:25-219, yuv422p8.ts:25-219 ----------------------
// Three planes, line pitch = width rounded up to 8 samples (chroma: half).  The read kernel has no
// tail quirk (the partial last block converts like the others), so it is one thread per pixel.
template <int BITS>
__device__ __forceinline__ uint32_t ld_sample(const void *plane, size_t i) {
	return BITS == 8 ? (uint32_t) reinterpret_cast<const uint8_t *>(plane)[i] : (uint32_t) reinterpret_cast<const uint16_t *>(plane)[i];
}

template <int BITS>
__global__ void __launch_bounds__(kThreads) k_yuv422p_read(const void *__restrict__ Y, const void *__restrict__ U, const void *__restrict__ V,
                                                           float4 *__restrict__ out, int width, int height, const __grid_constant__ ReadConsts rc) {
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= (size_t)width * height) return;
	const int line = (int)(tid / width), x = (int)(tid - (size_t)line * width);
	const int pitch = (width + 7) / 8 * 8;
	Ycc c;
	c.y = ld_sample<BITS>(Y, (size_t)line * pitch + x);
	c.cb = ld_sample<BITS>(U, (size_t)line * (pitch / 2) + x / 2);
	c.cr = ld_sample<BITS>(V, (size_t)line * (pitch / 2) + x / 2);
	const float3 rgb = ycc_to_linear(c, 1.0f, rc);
	out[tid] = make_float4(rgb.x, rgb.y, rgb.z, 1.0f);
}

// one thread per block of 8 pixels (one ushort8 / uchar8 of luma, 4 + 4 chroma samples)
template <int BITS>
__global__ void __launch_bounds__(kThreads) k_yuv422p_write(const float4 *__restrict__ in, void *__restrict__ Y, void *__restrict__ U,
                                                            void *__restrict__ V, int width, int lines, int interlace,
                                                            const __grid_constant__ WriteConsts wc) {
	const int blocks = (width + 7) / 8;
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= (size_t)blocks * lines) return;
	const int gl = (int)(tid / blocks), bx = (int)(tid - (size_t)gl * blocks);
	const int line = gl * (interlace == 0 ? 1 : 2) + (interlace == 3 ? 1 : 0);
	yuv422p_write_block<BITS>([&](int px, int ly) { return __ldg(in + (size_t)ly * width + px); }, Y, U, V, width, line, bx, wc);
}

// ---- yuv420p / nv12: yuv420p.ts:25-238, nv12.ts:24-240 --------------------------------------
// 8-bit 4:2:0: a luma plane (pitch = width rounded up to 8) and the chroma of every line PAIR, as two planes of
// pitch/2 bytes (yuv420p) or one plane of interleaved (U, V) pairs, pitch bytes (nv12).  The read kernel has no tail
// quirk (one thread per pixel; both lines of a pair take the pair's chroma).
template <bool NV12>
__global__ void __launch_bounds__(kThreads) k_yuv420_read(const uint8_t *__restrict__ Y, const uint8_t *__restrict__ U, const uint8_t *__restrict__ V,
                                                          float4 *__restrict__ out, int width, int height, const __grid_constant__ ReadConsts rc) {
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= (size_t)width * height) return;   // height is even here (the reference launches height / 2 work-groups)
	const int line = (int)(tid / width), x = (int)(tid - (size_t)line * width);
	const int pitch = (width + 7) / 8 * 8;
	Ycc c;
	c.y = Y[(size_t)line * pitch + x];
	if (NV12) {
		const uint8_t *cp = U + (size_t)(line / 2) * pitch + (x / 2) * 2;
		c.cb = cp[0];
		c.cr = cp[1];
	} else {
		c.cb = U[(size_t)(line / 2) * (pitch / 2) + x / 2];
		c.cr = V[(size_t)(line / 2) * (pitch / 2) + x / 2];
	}
	const float3 rgb = ycc_to_linear(c, 1.0f, rc);
	out[tid] = make_float4(rgb.x, rgb.y, rgb.z, 1.0f);
}

// One thread per block of 8 pixels of one line pair (pb_writers.cuh yuv420_write_block)
template <bool NV12>
__global__ void __launch_bounds__(kThreads) k_yuv420_write(const float4 *__restrict__ in, uint8_t *__restrict__ Y, uint8_t *__restrict__ U,
                                                           uint8_t *__restrict__ V, int width, int pairs, int interlace,
                                                           const __grid_constant__ WriteConsts wc) {
	const int blocks = (width + 7) / 8;
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= (size_t)blocks * pairs) return;
	const int gid = (int)(tid / blocks), bx = (int)(tid - (size_t)gid * blocks);
	yuv420_write_block<NV12>([&](int px, int ly) { return __ldg(in + (size_t)ly * width + px); }, Y, U, V, width, gid, bx, interlace, wc);
}

// ---- combine_N: combine.ts:24-68 --------------------------------------------------------
struct CombineArgs {
	const float4 *in[kMaxLayers];
	int n;
};
__global__ void __launch_bounds__(kThreads) k_combine(const __grid_constant__ CombineArgs a, float4 *__restrict__ out, size_t npx) {
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= npx) return;
	float4 acc = __ldg(a.in[0] + tid);
	for (int i = 1; i < a.n; ++i) acc = over4(acc, __ldg(a.in[i] + tid));
	out[tid] = acc;
}

// ---- transition_dissolve / mixer: transition.ts:60-65, mix.ts:30-45 -----------------------
__global__ void __launch_bounds__(kThreads) k_dissolve(const float4 *__restrict__ in0, const float4 *__restrict__ in1, float mix,
                                                       float4 *__restrict__ out, size_t npx) {
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= npx) return;
	out[tid] = dissolve4(__ldg(in0 + tid), __ldg(in1 + tid), mix);
}

// ---- transition_wipe: transition.ts:66-73 ---------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_wipe_mask(const float4 *__restrict__ in0, const float4 *__restrict__ in1,
                                                        const float4 *__restrict__ mask, float4 *__restrict__ out, size_t npx) {
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= npx) return;
	out[tid] = wipe_mask4(__ldg(in0 + tid), __ldg(in1 + tid), __ldg(mask + tid).x);
}

// ---- wipe: wipe.ts:30-47 ---------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_wipe(const float4 *__restrict__ in0, const float4 *__restrict__ in1, float wipe,
                                                   float4 *__restrict__ out, int w, int h) {
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= (size_t)w * h) return;
	const int x = (int)(tid % w);
	out[tid] = ((float)x > mul((float)w, wipe)) ? __ldg(in1 + tid) : __ldg(in0 + tid);
}

// ---- transform: transform.ts:36-59 -------------------------------------------------------------
struct Mat6 {
	float m[6];
};
__global__ void __launch_bounds__(kThreads) k_transform(const float4 *__restrict__ in, int sw, int sh, const __grid_constant__ Mat6 mat,
                                                        float4 *__restrict__ out, int w, int h) {
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= (size_t)w * h) return;
	const int y = (int)(tid / w), x = (int)(tid - (size_t)y * w);
	const float2 p = transform_pos(mat.m, x, y, w, h);
	out[tid] = sample_linear_clamp(sw, sh, p.x, p.y, [&](int i, int j) {
		if (i < 0 || j < 0 || i >= sw || j >= sh) return make_float4(0.f, 0.f, 0.f, 0.f);
		return __ldg(in + (size_t)j * sw + i);
	});
}

// ---- resize: resize.ts:35-59 ---------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_resize(const float4 *__restrict__ in, int sw, int sh, float scale, float offsetX,
                                                     float offsetY, float f0, float f1, float f2, float f3,
                                                     float4 *__restrict__ out, int w, int h) {
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= (size_t)w * h) return;
	const int y = (int)(tid / w), x = (int)(tid - (size_t)y * w);
	const float cx = add(__fdiv_rn(sub(-0.5f, offsetX), scale), 0.5f);
	const float cy = add(__fdiv_rn(sub(-0.5f, offsetY), scale), 0.5f);
	const float offX = fma_(cx, f1, f0), offY = fma_(cy, f3, f2);
	const float mulX = __fdiv_rn(f1, scale), mulY = __fdiv_rn(f3, scale);
	const float px = fma_(__fdiv_rn((float)x, (float)w), mulX, offX);
	const float py = fma_(__fdiv_rn((float)y, (float)h), mulY, offY);
	out[tid] = sample_linear_clamp(sw, sh, px, py, [&](int i, int j) {
		if (i < 0 || j < 0 || i >= sw || j >= sh) return make_float4(0.f, 0.f, 0.f, 0.f);
		return __ldg(in + (size_t)j * sw + i);
	});
}

// ---- yadif: yadifCl.ts:28-167 ------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_yadif(const float4 *__restrict__ prev, const float4 *__restrict__ cur,
                                                    const float4 *__restrict__ next, int parity, int tff, int skip,
                                                    float4 *__restrict__ out, int w, int h) {
	const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
	if (tid >= (size_t)w * h) return;
	const int yo = (int)(tid / w), xo = (int)(tid - (size_t)yo * w);
	out[tid] = yadif_texel(prev, cur, next, w, h, parity, tff, skip, xo, yo);
}

// The interpolated lines only (the lines of the other parity are the current frame's own: "the primary field is not modified",
// yadifCl.ts:118-120): the pre-pass of a fused launch that samples a de-interlaced field.  Row r of out = line 2 r + (1 - parity).
//
// A pixel reads 27 float4 from five frames-rows-sets; one thread per pixel holding them in registers runs at 106 registers,
// 16 warps per SM and waits on memory (ncu: 74 % of the stall samples long-scoreboard, 162 us for a 2160p field at 390 MB of
// DRAM traffic = 37 % of the HBM rate).  Here a block stages the rows of a 64-column x 8-line tile once, by plane, with
// coalesced float4 loads (every coordinate clamped as the reference's CLK_ADDRESS_CLAMP_TO_EDGE sampler would: the tile holds
// exactly the values px() returns), and the predictors then run channel by channel on conflict-free LDS.32.
constexpr int kYTileW = 64, kYTileR = 8, kYHalo = 3;
constexpr int kYKeptRows = kYTileR + 1, kYOwnRows = kYTileR + 2, kYCurW = kYTileW + 2 * kYHalo;

__global__ void __launch_bounds__(kThreads) k_yadif_rows(const float4 *__restrict__ prev, const float4 *__restrict__ cur,
                                                         const float4 *__restrict__ next, int parity, int tff, int skip,
                                                         float4 *__restrict__ out, int w, int h) {
	// lines of the field's own parity (unmodified in the output) around the tile's interpolated lines: y - 1, y + 1
	__shared__ float ck[3][kYKeptRows][kYCurW];   // current frame, +-3 columns for the spatial predictor
	__shared__ float pk[3][kYKeptRows][kYTileW];   // previous frame
	__shared__ float nk[3][kYKeptRows][kYTileW];   // next frame
	// lines of the interpolated parity, y - 2, y, y + 2, of the two frames around the field in time (yadifCl.ts prev2 / next2)
	__shared__ float p2[4][kYOwnRows][kYTileW];
	__shared__ float n2[4][kYOwnRows][kYTileW];
	asm volatile("griddepcontrol.launch_dependents;");   // the fused launch behind this pre-pass may load its tables meanwhile (it waits before it reads)
	const int second = !(parity ^ tff);
	const float4 *f_p2 = second ? cur : prev, *f_n2 = second ? next : cur;
	const int x0 = blockIdx.x * kYTileW, r0 = blockIdx.y * kYTileR;
	const int y_first = 2 * r0 + (1 - parity);
	const int tid = threadIdx.x;
	auto row_ptr = [&](const float4 *img, int y) { return img + (size_t)min(max(y, 0), h - 1) * w; };
	auto col = [&](int x) { return min(max(x, 0), w - 1); };
#pragma unroll
	for (int it = 0; it < (kYKeptRows * kYCurW + kThreads - 1) / kThreads; ++it) {
		const int i = it * kThreads + tid;
		if (i < kYKeptRows * kYCurW) {
			const int r = i / kYCurW, c = i - r * kYCurW;
			const float4 v = __ldg(row_ptr(cur, y_first - 1 + 2 * r) + col(x0 - kYHalo + c));
			ck[0][r][c] = v.x;
			ck[1][r][c] = v.y;
			ck[2][r][c] = v.z;
		}
	}
#pragma unroll
	for (int it = 0; it < (kYKeptRows * kYTileW + kThreads - 1) / kThreads; ++it) {
		const int i = it * kThreads + tid;
		if (i < kYKeptRows * kYTileW) {
			const int r = i / kYTileW, c = i % kYTileW;
			const int y = y_first - 1 + 2 * r, x = col(x0 + c);
			const float4 a = __ldg(row_ptr(prev, y) + x), b = __ldg(row_ptr(next, y) + x);
			pk[0][r][c] = a.x;
			pk[1][r][c] = a.y;
			pk[2][r][c] = a.z;
			nk[0][r][c] = b.x;
			nk[1][r][c] = b.y;
			nk[2][r][c] = b.z;
		}
	}
#pragma unroll
	for (int it = 0; it < (kYOwnRows * kYTileW + kThreads - 1) / kThreads; ++it) {
		const int i = it * kThreads + tid;
		if (i < kYOwnRows * kYTileW) {
			const int r = i / kYTileW, c = i % kYTileW;
			const int y = y_first - 2 + 2 * r, x = col(x0 + c);
			const float4 a = __ldg(row_ptr(f_p2, y) + x), b = __ldg(row_ptr(f_n2, y) + x);
			p2[0][r][c] = a.x;
			p2[1][r][c] = a.y;
			p2[2][r][c] = a.z;
			p2[3][r][c] = a.w;
			n2[0][r][c] = b.x;
			n2[1][r][c] = b.y;
			n2[2][r][c] = b.z;
			n2[3][r][c] = b.w;
		}
	}
	__syncthreads();
	const int tx = tid % kYTileW;
#pragma unroll 1
	for (int rr = tid / kYTileW; rr < kYTileR; rr += kThreads / kYTileW) {
		const int xo = x0 + tx, yo = y_first + 2 * rr;
		if (xo >= w || yo >= h) continue;
		float o[3];
#pragma unroll
		for (int ch = 0; ch < 3; ++ch) {
			const float *up = &ck[ch][rr][tx], *dn = &ck[ch][rr + 1][tx];   // lines yo - 1 / yo + 1 of cur, columns xo - 3 ... xo + 3
			const float sp = spatial_predictor(up[0], up[1], up[2], up[3], up[4], up[5], up[6], dn[0], dn[1], dn[2], dn[3], dn[4], dn[5], dn[6]);
			o[ch] = temporal_predictor(pk[ch][rr][tx], pk[ch][rr + 1][tx], p2[ch][rr][tx], p2[ch][rr + 1][tx], p2[ch][rr + 2][tx], up[3], dn[3],
			                           n2[ch][rr][tx], n2[ch][rr + 1][tx], n2[ch][rr + 2][tx], nk[ch][rr][tx], nk[ch][rr + 1][tx], sp, skip);
		}
		// "Reset Alpha" (yadifCl.ts:164): the current frame's own alpha at (xo, yo)
		out[(size_t)(r0 + rr) * w + xo] = make_float4(o[0], o[1], o[2], second ? p2[3][rr + 1][tx] : n2[3][rr + 1][tx]);
	}
}

// ---- launchers -----------------------------------------------------------------------------------------
#define LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t launch_v210_read(cudaStream_t s, const void *in, void *out, int w, int h, const ReadConsts &rc) {
	const int pitch16 = ((w + 47) / 48) * 8;
	const size_t n = (size_t)((w + 5) / 6) * h;
	k_v210_read<<<blocks_for(n), kThreads, 0, s>>>((const uint4 *)in, (float4 *)out, w, h, pitch16, rc);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_v210_write(cudaStream_t s, const void *in, void *out, int w, int h, int interlace, const WriteConsts &wc) {
	const int pitch16 = ((w + 47) / 48) * 8;
	const int lines = interlace == 0 ? h : h / 2;
	k_v210_write<<<blocks_for((size_t)pitch16 * lines), kThreads, 0, s>>>((const float4 *)in, (uint4 *)out, w, lines, pitch16, interlace, wc);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_rgba8_read(cudaStream_t s, const void *in, void *out, int w, int h, int bgra, const ReadConsts &rc) {
	const size_t n = (size_t)w * h;
	k_rgba8_read<<<blocks_for(n), kThreads, 0, s>>>((const uchar4 *)in, (float4 *)out, n, bgra, rc);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_rgba8_write(cudaStream_t s, const void *in, void *out, int w, int h, int interlace, int bgra, const WriteConsts &wc) {
	const int lines = interlace == 0 ? h : h / 2;
	k_rgba8_write<<<blocks_for((size_t)w * lines), kThreads, 0, s>>>((const float4 *)in, (uchar4 *)out, w, lines, interlace, bgra, wc);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_yuv422p_read(cudaStream_t s, int bits, const void *y, const void *u, const void *v, void *out, int w, int h, const ReadConsts &rc) {
	const size_t n = (size_t)w * h;
	if (bits == 8) k_yuv422p_read<8><<<blocks_for(n), kThreads, 0, s>>>(y, u, v, (float4 *)out, w, h, rc);
	else k_yuv422p_read<10><<<blocks_for(n), kThreads, 0, s>>>(y, u, v, (float4 *)out, w, h, rc);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_yuv422p_write(cudaStream_t s, int bits, const void *in, void *y, void *u, void *v, int w, int h, int interlace, const WriteConsts &wc) {
	const int lines = interlace == 0 ? h : h / 2;
	const size_t n = (size_t)((w + 7) / 8) * lines;
	if (bits == 8) k_yuv422p_write<8><<<blocks_for(n), kThreads, 0, s>>>((const float4 *)in, y, u, v, w, lines, interlace, wc);
	else k_yuv422p_write<10><<<blocks_for(n), kThreads, 0, s>>>((const float4 *)in, y, u, v, w, lines, interlace, wc);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_yuv420_read(cudaStream_t s, int nv12, const void *y, const void *u, const void *v, void *out, int w, int h, const ReadConsts &rc) {
	const size_t n = (size_t)w * (h & ~1);
	if (nv12) k_yuv420_read<true><<<blocks_for(n), kThreads, 0, s>>>((const uint8_t *)y, (const uint8_t *)u, nullptr, (float4 *)out, w, h & ~1, rc);
	else k_yuv420_read<false><<<blocks_for(n), kThreads, 0, s>>>((const uint8_t *)y, (const uint8_t *)u, (const uint8_t *)v, (float4 *)out, w, h & ~1, rc);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_yuv420_write(cudaStream_t s, int nv12, const void *in, void *y, void *u, void *v, int w, int h, int interlace, const WriteConsts &wc) {
	const int pairs = h / 2;
	const size_t n = (size_t)((w + 7) / 8) * pairs;
	if (nv12) k_yuv420_write<true><<<blocks_for(n), kThreads, 0, s>>>((const float4 *)in, (uint8_t *)y, (uint8_t *)u, nullptr, w, pairs, interlace, wc);
	else k_yuv420_write<false><<<blocks_for(n), kThreads, 0, s>>>((const float4 *)in, (uint8_t *)y, (uint8_t *)u, (uint8_t *)v, w, pairs, interlace, wc);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_combine(cudaStream_t s, const void *const *in, int n, void *out, int w, int h) {
	CombineArgs a;
	a.n = n;
	for (int i = 0; i < kMaxLayers; ++i) a.in[i] = (const float4 *)(i < n ? in[i] : nullptr);
	const size_t npx = (size_t)w * h;
	k_combine<<<blocks_for(npx), kThreads, 0, s>>>(a, (float4 *)out, npx);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_dissolve(cudaStream_t s, const void *in0, const void *in1, float mix, void *out, int w, int h) {
	const size_t npx = (size_t)w * h;
	k_dissolve<<<blocks_for(npx), kThreads, 0, s>>>((const float4 *)in0, (const float4 *)in1, mix, (float4 *)out, npx);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_wipe_mask(cudaStream_t s, const void *in0, const void *in1, const void *mask, void *out, int w, int h) {
	const size_t npx = (size_t)w * h;
	k_wipe_mask<<<blocks_for(npx), kThreads, 0, s>>>((const float4 *)in0, (const float4 *)in1, (const float4 *)mask, (float4 *)out, npx);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_wipe(cudaStream_t s, const void *in0, const void *in1, float wipe, void *out, int w, int h) {
	k_wipe<<<blocks_for((size_t)w * h), kThreads, 0, s>>>((const float4 *)in0, (const float4 *)in1, wipe, (float4 *)out, w, h);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_transform(cudaStream_t s, const void *in, int sw, int sh, const float *mat6, void *out, int w, int h) {
	Mat6 m;
	for (int i = 0; i < 6; ++i) m.m[i] = mat6[i];
	k_transform<<<blocks_for((size_t)w * h), kThreads, 0, s>>>((const float4 *)in, sw, sh, m, (float4 *)out, w, h);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_resize(cudaStream_t s, const void *in, int sw, int sh, float scale, float ox, float oy, const float *flip4,
                          void *out, int w, int h) {
	k_resize<<<blocks_for((size_t)w * h), kThreads, 0, s>>>((const float4 *)in, sw, sh, scale, ox, oy, flip4[0], flip4[1], flip4[2],
	                                                        flip4[3], (float4 *)out, w, h);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_yadif_rows(cudaStream_t s, const void *prev, const void *cur, const void *next, int parity, int tff, int skip,
                              void *out, int w, int h) {
	const int rows = (h - (1 - parity) + 1) / 2;   // lines 1 - parity, 3 - parity, ... below h
	if (rows <= 0 || w <= 0) return cudaSuccess;
	k_yadif_rows<<<dim3((w + kYTileW - 1) / kYTileW, (rows + kYTileR - 1) / kYTileR), kThreads, 0, s>>>((const float4 *)prev, (const float4 *)cur, (const float4 *)next, parity,
	                                                                          tff, skip, (float4 *)out, w, h);
	LAUNCH_CHECK();
	return cudaSuccess;
}
cudaError_t launch_yadif(cudaStream_t s, const void *prev, const void *cur, const void *next, int parity, int tff, int skip,
                         void *out, int w, int h) {
	k_yadif<<<blocks_for((size_t)w * h), kThreads, 0, s>>>((const float4 *)prev, (const float4 *)cur, (const float4 *)next, parity, tff,
	                                                       skip, (float4 *)out, w, h);
	LAUNCH_CHECK();
	return cudaSuccess;
}

}  // namespace pb
