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
#include "pb_writers.cuh"
#include "pb_launch.h"

namespace pb {

constexpr int kFusedThreads = 128;

template <bool kToRgba>
__global__ void __launch_bounds__(kFusedThreads) k_fused_generic(const __grid_constant__ FusedDesc d, float4 *__restrict__ out_rgba) {
	asm volatile("griddepcontrol.launch_dependents;");   // a march launch behind this one may load its tables meanwhile (pb_march_impl.cuh pdl_wait)
	const int pitch16 = kToRgba ? (d.out_w + 5) / 6 : d.out_pitch / 16;
	const int g_first = kToRgba ? 0 : d.g_first, cols = pitch16 - g_first;   // g_first > 0: only the ragged tail columns of each line
	const int lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
	const size_t tid = (size_t)blockIdx.x * kFusedThreads + threadIdx.x;
	if (tid >= (size_t)cols * lines) return;
	const int gl = (int)(tid / cols), g = g_first + (int)(tid - (size_t)gl * cols);
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

// the layer graph at one output pixel: bottom layer, then `over` for each layer above (combine.ts:49-59)
__device__ __forceinline__ float4 composite_px(const FusedDesc &d, int x, int line) {
	float4 acc = layer_value(d.layers[0], d.rc, x, line);
#pragma unroll 1
	for (int l = 1; l < d.n_layers; ++l) acc = over4(acc, layer_value(d.layers[l], d.rc, x, line));
	return acc;
}

// Fused sinks for the other Writer PackImpls: the same decomposition into threads as the stand-alone writer kernels
// (pb_kernels.cu), the pixel source being the layer graph instead of an RGBA-f32 frame.
template <int kSink>
__global__ void __launch_bounds__(kFusedThreads) k_fused_sink(const __grid_constant__ FusedDesc d) {
	asm volatile("griddepcontrol.launch_dependents;");   // a march launch behind this one may load its tables meanwhile (pb_march_impl.cuh pdl_wait)
	const size_t tid = (size_t)blockIdx.x * kFusedThreads + threadIdx.x;
	auto px = [&](int x, int line) { return composite_px(d, x, line); };
	if (kSink == SINK_RGBA8 || kSink == SINK_BGRA8) {
		const int lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
		if (tid >= (size_t)d.out_w * lines) return;
		const int gl = (int)(tid / d.out_w), x = (int)(tid - (size_t)gl * d.out_w);
		const int line = gl * (d.interlace == 0 ? 1 : 2) + (d.interlace == 3 ? 1 : 0);
		rgba8_write_px(px, reinterpret_cast<uchar4 *>(d.out), d.out_w, line, x, kSink == SINK_BGRA8, d.wc);
	} else if (kSink == SINK_YUV422P10 || kSink == SINK_YUV422P8) {
		const int blocks = (d.out_w + 7) / 8, lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
		if (tid >= (size_t)blocks * lines) return;
		const int gl = (int)(tid / blocks), bx = (int)(tid - (size_t)gl * blocks);
		const int line = gl * (d.interlace == 0 ? 1 : 2) + (d.interlace == 3 ? 1 : 0);
		yuv422p_write_block<kSink == SINK_YUV422P8 ? 8 : 10>(px, d.out, d.out_u, d.out_v, d.out_w, line, bx, d.wc);
	} else {
		const int blocks = (d.out_w + 7) / 8, pairs = d.out_h / 2;
		if (tid >= (size_t)blocks * pairs) return;
		const int gid = (int)(tid / blocks), bx = (int)(tid - (size_t)gid * blocks);
		yuv420_write_block<kSink == SINK_NV12>(px, reinterpret_cast<uint8_t *>(d.out), reinterpret_cast<uint8_t *>(d.out_u),
		                                       reinterpret_cast<uint8_t *>(d.out_v), d.out_w, gid, bx, d.interlace, d.wc);
	}
}

const char *fused_variant(const FusedDesc &) { return "generic"; }

// a graph with a Yadif leaf takes the kernel instances that can evaluate one (see leaf_texel)
cudaError_t launch_fused(cudaStream_t s, const FusedDesc &d, void *out_rgba) {
	const int lines = d.interlace == 0 ? d.out_h : d.out_h / 2;
	if (!out_rgba && d.sink != SINK_V210) {
		const size_t blocks8 = (size_t)((d.out_w + 7) / 8);
		size_t n;
		auto grid = [&](size_t threads) { return (unsigned)((threads + kFusedThreads - 1) / kFusedThreads); };
		switch (d.sink) {
			case SINK_RGBA8: n = (size_t)d.out_w * lines; k_fused_sink<SINK_RGBA8><<<grid(n), kFusedThreads, 0, s>>>(d); break;
			case SINK_BGRA8: n = (size_t)d.out_w * lines; k_fused_sink<SINK_BGRA8><<<grid(n), kFusedThreads, 0, s>>>(d); break;
			case SINK_YUV422P10: n = blocks8 * lines; k_fused_sink<SINK_YUV422P10><<<grid(n), kFusedThreads, 0, s>>>(d); break;
			case SINK_YUV422P8: n = blocks8 * lines; k_fused_sink<SINK_YUV422P8><<<grid(n), kFusedThreads, 0, s>>>(d); break;
			case SINK_YUV420P: n = blocks8 * (d.out_h / 2); k_fused_sink<SINK_YUV420P><<<grid(n), kFusedThreads, 0, s>>>(d); break;
			case SINK_NV12: n = blocks8 * (d.out_h / 2); k_fused_sink<SINK_NV12><<<grid(n), kFusedThreads, 0, s>>>(d); break;
			default: return cudaErrorInvalidValue;
		}
		return cudaGetLastError();
	}
	if (out_rgba) {
		const size_t n = (size_t)((d.out_w + 5) / 6) * lines;
		k_fused_generic<true><<<(unsigned)((n + kFusedThreads - 1) / kFusedThreads), kFusedThreads, 0, s>>>(d, (float4 *)out_rgba);
	} else {
		const size_t n = (size_t)(d.out_pitch / 16 - d.g_first) * lines;
		k_fused_generic<false><<<(unsigned)((n + kFusedThreads - 1) / kFusedThreads), kFusedThreads, 0, s>>>(d, nullptr);
	}
	return cudaGetLastError();
}

}  // namespace pb
