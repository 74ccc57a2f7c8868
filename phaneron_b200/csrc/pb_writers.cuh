// pb_writers.cuh -- the Writer kernels of the packed formats other than v210, as device templates over a pixel
// source `px(x, line) -> float4` (linear RGBA).  The stand-alone kernels (pb_kernels.cu, the reference's launch
// structure) read `px` from an RGBA-f32 frame; the fused sink kernels (pb_fused.cu) evaluate the layer graph
// instead, so FFmpegConsumer / ScreenConsumer formats are written without an RGBA-f32 intermediate.
// Reference kernels: rgba8.ts:69-103, bgra8.ts:69-103, yuv422p10.ts:126-219, yuv422p8.ts:126-219,
// yuv420p.ts:142-238, nv12.ts:134-240.
#pragma once
#include "pb_device.cuh"

namespace pb {

// rgba8.ts:83-101: one pixel; alpha is written as 255
template <class Src>
__device__ __forceinline__ void rgba8_write_px(Src px, uchar4 *__restrict__ out, int width, int line, int x, int bgra, const WriteConsts &wc) {
	const float4 v = px(x, line);
	const float r = __ldg(wc.lut + sat_rte_u16(mul(v.x, 65535.0f)));
	const float g = __ldg(wc.lut + sat_rte_u16(mul(v.y, 65535.0f)));
	const float b = __ldg(wc.lut + sat_rte_u16(mul(v.z, 65535.0f)));
	uchar4 o;
	const unsigned char r8 = (unsigned char)sat_rte_u8(mul(r, 255.0f)), g8 = (unsigned char)sat_rte_u8(mul(g, 255.0f)),
	                    b8 = (unsigned char)sat_rte_u8(mul(b, 255.0f));
	o.x = bgra ? b8 : r8;
	o.y = g8;
	o.z = bgra ? r8 : b8;
	o.w = 255;
	out[(size_t)line * width + x] = o;
}

// codes of the partial last block of a line: round() (half away from zero) before the saturating conversion
__device__ __forceinline__ void ycc_tail_round(const float4 &l, const WriteConsts &wc, uint32_t &y, uint32_t &u, uint32_t &v) {
	const float gr = __ldg(wc.lut + sat_rte_u16(mul(l.x, 65535.0f))), gg = __ldg(wc.lut + sat_rte_u16(mul(l.y, 65535.0f))),
	            gb = __ldg(wc.lut + sat_rte_u16(mul(l.z, 65535.0f)));
	y = sat_rte_u16(roundf(dot4(gr, gg, gb, 1.0f, wc.cm + 0)));
	u = sat_rte_u16(roundf(dot4(gr, gg, gb, 1.0f, wc.cm + 4)));
	v = sat_rte_u16(roundf(dot4(gr, gg, gb, 1.0f, wc.cm + 8)));
}

template <int BITS>
__device__ __forceinline__ void st_sample(void *plane, size_t i, uint32_t v) {
	if (BITS == 8) reinterpret_cast<uint8_t *>(plane)[i] = (uint8_t)v;   // Q11: the 16-bit conversion result wraps into a uchar
	else reinterpret_cast<uint16_t *>(plane)[i] = (uint16_t)v;
}

// planar 4:2:2: block bx (8 pixels) of `line`
template <int BITS, class Src>
__device__ __forceinline__ void yuv422p_write_block(Src px, void *__restrict__ Y, void *__restrict__ U, void *__restrict__ V, int width, int line,
                                                    int bx, const WriteConsts &wc) {
	const int blocks = (width + 7) / 8;
	const int x0 = bx * 8, n = min(8, width - x0);
	const size_t yo = ((size_t)line * blocks + bx) * 8, co = ((size_t)line * blocks + bx) * 4;
	uint32_t y[8], u[4], v[4];
	if (n == 8) {   // yuv422p10.ts:155-178
#pragma unroll
		for (int p = 0; p < 8; ++p) {
			const float4 l = px(x0 + p, line);
			const Ycc c = linear_to_ycc(l.x, l.y, l.z, wc);
			y[p] = c.y;
			if (!(p & 1)) { u[p / 2] = c.cb; v[p / 2] = c.cr; }   // chroma from even pixels only
		}
	} else {   // the partial last block of a line, yuv422p10.ts:180-218
#pragma unroll
		for (int k = 0; k < 8; ++k) y[k] = BITS == 8 ? 16 : 64;
#pragma unroll
		for (int k = 0; k < 4; ++k) u[k] = v[k] = BITS == 8 ? 128 : 512;
		uint32_t ty[6], tu[6], tv[6];
#pragma unroll
		for (int p = 0; p < 6; ++p) {
			ty[p] = tu[p] = tv[p] = 0;
			if (p < n) ycc_tail_round(px(x0 + p, line), wc, ty[p], tu[p], tv[p]);
		}
		y[0] = ty[0]; y[1] = ty[1]; u[0] = tu[0]; v[0] = tv[0];
		if (n > 2) {
			y[2] = ty[2]; y[3] = ty[3]; u[1] = tu[2]; v[1] = tv[2];
			if (n > 4) {
				y[4] = ty[4]; y[5] = ty[5];
				u[1] = tu[4]; v[1] = tv[4];   // Q12: .s1 where .s2 is meant (yuv422p10.ts:210-211)
			}
		}
	}
#pragma unroll
	for (int p = 0; p < 8; ++p) st_sample<BITS>(Y, yo + p, y[p]);
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		st_sample<BITS>(U, co + k, u[k]);
		st_sample<BITS>(V, co + k, v[k]);
	}
}

// 8-bit 4:2:0: block bx (8 pixels) of line pair gid: the luma of one line (a field launch, interlace 1 / 3) or of both
// (progressive), and the pair's chroma taken from the even pixels of the first line processed (yuv420p.ts:160-200)
template <bool NV12, class Src>
__device__ __forceinline__ void yuv420_write_block(Src px, uint8_t *__restrict__ Y, uint8_t *__restrict__ U, uint8_t *__restrict__ V, int width,
                                                   int gid, int bx, int interlace, const WriteConsts &wc) {
	const int blocks = (width + 7) / 8;
	const int line0 = gid * 2 + (interlace == 3 ? 1 : 0), n_lines = interlace == 0 ? 2 : 1;
	const int x0 = bx * 8, n = min(8, width - x0);
	for (int l = 0; l < n_lines; ++l) {
		const int line = line0 + l;
		uint32_t y[8], u[4], v[4];
		if (n == 8) {
#pragma unroll
			for (int p = 0; p < 8; ++p) {
				const float4 q = px(x0 + p, line);
				const Ycc c = linear_to_ycc(q.x, q.y, q.z, wc);
				y[p] = c.y;
				if (!(p & 1)) { u[p / 2] = c.cb; v[p / 2] = c.cr; }
			}
		} else {   // the partial last block of a line: round() before the conversion, unwritten samples 16 / 128 (yuv420p.ts:204-236)
#pragma unroll
			for (int k = 0; k < 8; ++k) y[k] = 16;
#pragma unroll
			for (int k = 0; k < 4; ++k) u[k] = v[k] = 128;
			uint32_t ty[6], tu[6], tv[6];
#pragma unroll
			for (int p = 0; p < 6; ++p) {
				ty[p] = tu[p] = tv[p] = 0;
				if (p < n) ycc_tail_round(px(x0 + p, line), wc, ty[p], tu[p], tv[p]);
			}
			y[0] = ty[0]; y[1] = ty[1]; u[0] = tu[0]; v[0] = tv[0];
			if (n > 2) {
				y[2] = ty[2]; y[3] = ty[3]; u[1] = tu[2]; v[1] = tv[2];
				if (n > 4) { y[4] = ty[4]; y[5] = ty[5]; u[2] = tu[4]; v[2] = tv[4]; }
			}
		}
		// uchar stores: the 16-bit conversion results wrap (Q11, as in yuv422p8)
		uint2 yw;
		yw.x = (y[0] & 255u) | (y[1] & 255u) << 8 | (y[2] & 255u) << 16 | (y[3] & 255u) << 24;
		yw.y = (y[4] & 255u) | (y[5] & 255u) << 8 | (y[6] & 255u) << 16 | (y[7] & 255u) << 24;
		*reinterpret_cast<uint2 *>(Y + ((size_t)line * blocks + bx) * 8) = yw;
		if (l == 0) {
			if (NV12) {
				uint2 cw;
				cw.x = (u[0] & 255u) | (v[0] & 255u) << 8 | (u[1] & 255u) << 16 | (v[1] & 255u) << 24;
				cw.y = (u[2] & 255u) | (v[2] & 255u) << 8 | (u[3] & 255u) << 16 | (v[3] & 255u) << 24;
				*reinterpret_cast<uint2 *>(U + ((size_t)gid * blocks + bx) * 8) = cw;
			} else {
				const size_t co = ((size_t)gid * blocks + bx) * 4;
				*reinterpret_cast<uint32_t *>(U + co) = (u[0] & 255u) | (u[1] & 255u) << 8 | (u[2] & 255u) << 16 | (u[3] & 255u) << 24;
				*reinterpret_cast<uint32_t *>(V + co) = (v[0] & 255u) | (v[1] & 255u) << 8 | (v[2] & 255u) << 16 | (v[3] & 255u) << 24;
			}
		}
	}
}

}  // namespace pb
