// pb_fused.cu -- the fused frame-expression kernels:
//   N layers x (v210 unpack -> YCbCr->linear RGB -> [transform/bilinear] -> [dissolve|wipe])
//   -> combine (premultiplied over) -> linear->gamma -> RGB->YCbCr -> v210 pack
// in ONE launch, reading each packed source once from HBM and writing the packed
// output once.  Reference stages replaced: v210.ts:25-195, transform.ts:36-59,
// transition.ts:60-73, combine.ts:24-68 (and the RGBA-f32 round trips between them).
//
// Variants (picked by launch_fused):
//   generic : any transform (incl. rotation), any leaf kind; every bilinear tap converts
//             its own texel.  Correct everywhere, used as the fallback.
#include "pb_device.cuh"
#include "pb_launch.h"

namespace pb {

constexpr int kFusedThreads = 128;

template <bool kToRgba>
__global__ void __launch_bounds__(kFusedThreads) k_fused_generic(const __grid_constant__ FusedDesc d, float4 *__restrict__ out_rgba) {
	const int pitch16 = kToRgba ? (d.out_w + 5) / 6 : d.out_pitch / 16;
	const int lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
	const size_t tid = (size_t)blockIdx.x * kFusedThreads + threadIdx.x;
	if (tid >= (size_t)pitch16 * lines) return;
	const int gl = (int)(tid / pitch16), g = (int)(tid - (size_t)gl * pitch16);
	const int line = gl * (d.interlace == 0 ? 1 : 2) + (d.interlace == 3 ? 1 : 0);
	const int x0 = g * 6;
	uint4 w = make_uint4(0, 0, 0, 0);
	if (x0 < d.out_w) {
		const int n = min(6, d.out_w - x0);
#pragma unroll 1
		for (int p = 0; p < n; ++p) {
			const int x = x0 + p;
			float4 acc = layer_value(d.layers[0], d.rc, x, line);
#pragma unroll 1
			for (int l = 1; l < d.n_layers; ++l) acc = over4(acc, layer_value(d.layers[l], d.rc, x, line));
			if (kToRgba) {
				out_rgba[(size_t)line * d.out_w + x] = acc;
			} else {
				const Ycc c = (n == 6) ? linear_to_ycc(acc.x, acc.y, acc.z, d.wc) : linear_to_ycc_tail(acc.x, acc.y, acc.z, d.wc);
				switch (p) {   // v210.ts:158-163 / 186-192
					case 0: w.x = c.cr << 20 | c.y << 10 | c.cb; break;
					case 1: w.y = c.y; break;
					case 2: w.y |= c.y << 20 | c.cb << 10; w.z = c.cr; break;
					case 3: w.z |= c.y << 10; break;
					case 4: w.z |= c.cb << 20; w.w = c.cr << 10 | c.y; break;
					default: w.w |= c.y << 20; break;
				}
			}
		}
	} else if (kToRgba || d.out_w % 48 == 0) {
		return;
	}
	if (!kToRgba) st_stream(reinterpret_cast<uint4 *>(reinterpret_cast<char *>(d.out) + (size_t)line * d.out_pitch) + g, w);
}

const char *fused_variant(const FusedDesc &) { return "generic"; }

cudaError_t launch_fused(cudaStream_t s, const FusedDesc &d, void *out_rgba) {
	const int lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
	if (out_rgba) {
		const size_t n = (size_t)((d.out_w + 5) / 6) * lines;
		k_fused_generic<true><<<(unsigned)((n + kFusedThreads - 1) / kFusedThreads), kFusedThreads, 0, s>>>(d, (float4 *)out_rgba);
	} else {
		const size_t n = (size_t)(d.out_pitch / 16) * lines;
		k_fused_generic<false><<<(unsigned)((n + kFusedThreads - 1) / kFusedThreads), kFusedThreads, 0, s>>>(d, nullptr);
	}
	return cudaGetLastError();
}

}  // namespace pb
