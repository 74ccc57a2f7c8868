// pb_launch.h -- host-callable launchers (defined in pb_kernels.cu / pb_fused.cu)
#pragma once
#include <cuda_runtime.h>

#include "pb_desc.h"

namespace pb {

cudaError_t launch_v210_read(cudaStream_t s, const void *in, void *out, int w, int h, const ReadConsts &rc);
cudaError_t launch_v210_write(cudaStream_t s, const void *in, void *out, int w, int h, int interlace, const WriteConsts &wc);
cudaError_t launch_rgba8_read(cudaStream_t s, const void *in, void *out, int w, int h, int bgra, const ReadConsts &rc);
cudaError_t launch_rgba8_write(cudaStream_t s, const void *in, void *out, int w, int h, int interlace, int bgra, const WriteConsts &wc);
// planar 4:2:2, bits = 8 or 10 (yuv422p8.ts / yuv422p10.ts)
cudaError_t launch_yuv422p_read(cudaStream_t s, int bits, const void *y, const void *u, const void *v, void *out, int w, int h, const ReadConsts &rc);
cudaError_t launch_yuv422p_write(cudaStream_t s, int bits, const void *in, void *y, void *u, void *v, int w, int h, int interlace, const WriteConsts &wc);
// 8-bit 4:2:0: yuv420p.ts (three planes) / nv12.ts (nv12 = 1: u is the interleaved chroma plane, v unused)
cudaError_t launch_yuv420_read(cudaStream_t s, int nv12, const void *y, const void *u, const void *v, void *out, int w, int h, const ReadConsts &rc);
cudaError_t launch_yuv420_write(cudaStream_t s, int nv12, const void *in, void *y, void *u, void *v, int w, int h, int interlace, const WriteConsts &wc);
cudaError_t launch_combine(cudaStream_t s, const void *const *in, int n, void *out, int w, int h);
cudaError_t launch_dissolve(cudaStream_t s, const void *in0, const void *in1, float mix, void *out, int w, int h);
cudaError_t launch_wipe_mask(cudaStream_t s, const void *in0, const void *in1, const void *mask, void *out, int w, int h);
cudaError_t launch_wipe(cudaStream_t s, const void *in0, const void *in1, float wipe, void *out, int w, int h);
cudaError_t launch_transform(cudaStream_t s, const void *in, int sw, int sh, const float *mat6, void *out, int w, int h);
cudaError_t launch_resize(cudaStream_t s, const void *in, int sw, int sh, float scale, float ox, float oy, const float *flip4,
                          void *out, int w, int h);
cudaError_t launch_yadif(cudaStream_t s, const void *prev, const void *cur, const void *next, int parity, int tff, int skip,
                         void *out, int w, int h);

// the interpolated lines of a de-interlaced field only, packed: row r of out = line 2 r + (1 - parity)  (pre-pass of fused launches)
cudaError_t launch_yadif_rows(cudaStream_t s, const void *prev, const void *cur, const void *next, int parity, int tff, int skip,
                              void *out, int w, int h);

// Fused chain: N layers of (leaf | dissolve | wipe) -> combine -> v210 pack, one launch.
// out_rgba != nullptr writes the composite as RGBA-f32 instead of packing (materialise).
cudaError_t launch_fused(cudaStream_t s, const FusedDesc &d, void *out_rgba);
// March kernel (pb_march.cu); the descriptor must have been prepared (sampling tables, LUT slots).
cudaError_t launch_fused_march(cudaStream_t s, const FusedDesc &d, int num_sms);
size_t march_smem_bytes(const FusedDesc &d);
// first pass of a separable Lanczos Transform (see HPassDesc)
cudaError_t launch_lanczos_hpass(cudaStream_t s, const HPassDesc &h, int num_sms);
// gamma table -> one-byte-per-entry form (pb_lut.cuh): n_cands candidate models evaluated in one launch
struct LutFitResult;
cudaError_t launch_lut_fit(cudaStream_t s, const float *table, const LutParams *cands_dev, int n_cands, uint8_t *d8_out, void *results_dev);
// name of the kernel variant launch_fused would pick (for stats / tests)
const char *fused_variant(const FusedDesc &d);

}  // namespace pb
