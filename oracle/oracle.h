/*
 * oracle.h -- CPU restatement of the Streampunk/phaneron pixel path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in phaneron_b200/ (the product) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * Every function cites the reference file:line (paths under /root/reference/)
 * whose behaviour it restates.  The reference's kernels are OpenCL C strings;
 * where OpenCL leaves float evaluation implementation-defined we fix ONE
 * member of the family ("canonical semantics", see oracle.c header).
 *
 * Pinning status: the reference has no golden vectors for this path other
 * than the fillBuf round-trip invariant (src/process/v210.ts:206-236 and the
 * src/process/test scripts) and the formulas themselves; SURVEY.md section 8c
 * lists known-answers derived from those formulas.  tests/test_oracle.py
 * checks both.  Everything else is "parity unpinned by reference tests".
 */
#ifndef PHANERON_ORACLE_H
#define PHANERON_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

void orc_set_threads(int n);
int orc_get_threads(void);

/* colourMaths.ts */
int orc_gamma2linear_lut(const char *colspec, float *out65536);
int orc_linear2gamma_lut(const char *colspec, float *out65536);
int orc_ycbcr2rgb_matrix(const char *colspec, int num_bits, int luma_black, int luma_white,
                         int chr_range, float *out12);
int orc_rgb2ycbcr_matrix(const char *colspec, int num_bits, int luma_black, int luma_white,
                         int chr_range, float *out12);
int orc_rgb2rgb_matrix(const char *src, const char *dst, float *out9);
/* transform.ts:119-171 */
void orc_transform_matrix(int width, int height, int flip_h, int flip_v, double anchor_x,
                          double anchor_y, double scale_x, double scale_y, double offset_x,
                          double offset_y, double rotate, float *out9);

/* v210.ts */
uint32_t orc_v210_pitch(uint32_t width);       /* pixels */
uint32_t orc_v210_pitch_bytes(uint32_t width); /* bytes */
void orc_v210_fill(uint8_t *buf, uint32_t width, uint32_t height);
void orc_v210_read(const uint32_t *input, float *output, uint32_t width, uint32_t height,
                   const float *col_matrix12, const float *gamma_lut, const float *gamut9);
/* q3_literal != 0 reproduces `outOff = width*line/6` (v210.ts:129) serially;
   0 uses the pitch-correct offset (identical when width % 48 == 0). */
void orc_v210_write(const float *input, uint32_t *output, uint32_t width, uint32_t height,
                    uint32_t interlace, const float *col_matrix12, const float *gamma_lut,
                    int q3_literal);

/* image ops: all images are RGBA float32, row-major, w*h*4 floats */
void orc_combine(const float *const *layers, int num_layers, float *out, int w, int h);
void orc_dissolve(const float *in0, const float *in1, float mix, float *out, int w, int h);
void orc_wipe_mask(const float *in0, const float *in1, const float *mask, float *out, int w, int h);
void orc_transform(const float *in, int sw, int sh, const float *mat9, float *out, int w, int h);
void orc_yadif(const float *prev, const float *cur, const float *next, int parity, int tff,
               int skip_spatial, float *out, int w, int h);
void orc_mix(const float *in0, const float *in1, float mix, float *out, int w, int h);
void orc_wipe(const float *in0, const float *in1, float wipe, float *out, int w, int h);
void orc_resize(const float *in, int sw, int sh, float scale, float offset_x, float offset_y,
                const float *flip4, float *out, int w, int h);

/* other packers */
void orc_rgba8_read(const uint8_t *input, float *output, uint32_t width, uint32_t height,
                    const float *gamma_lut, const float *gamut9, int bgra);
void orc_rgba8_write(const float *input, uint8_t *output, uint32_t width, uint32_t height,
                     uint32_t interlace, const float *gamma_lut, int bgra);

/* yuv422p10le / yuv422p8: bits = 10 or 8; planes are byte pointers (little-endian samples) */
uint32_t orc_yuv422p_pitch(uint32_t width); /* pixels */
void orc_yuv422p_fill(int bits, uint8_t *buf, uint32_t width, uint32_t height);
void orc_yuv422p_read(int bits, const uint8_t *inY, const uint8_t *inU, const uint8_t *inV, float *output,
                      uint32_t width, uint32_t height, const float *col_matrix12, const float *gamma_lut,
                      const float *gamut9);
void orc_yuv422p_write(int bits, const float *input, uint8_t *outY, uint8_t *outU, uint8_t *outV,
                       uint32_t width, uint32_t height, uint32_t interlace, const float *col_matrix12,
                       const float *gamma_lut);

/* yuv420p / nv12: 8-bit 4:2:0; nv12 = 1: inU is the interleaved chroma plane, inV / outV unused */
void orc_yuv420_plane_bytes(int nv12, uint32_t width, uint32_t height, uint32_t out[3]);
void orc_yuv420_fill(int nv12, uint8_t *buf, uint32_t width, uint32_t height);
void orc_yuv420_read(int nv12, const uint8_t *inY, const uint8_t *inU, const uint8_t *inV, float *output,
                     uint32_t width, uint32_t height, const float *col_matrix12, const float *gamma_lut,
                     const float *gamut9);
void orc_yuv420_write(int nv12, const float *input, uint8_t *outY, uint8_t *outU, uint8_t *outV,
                      uint32_t width, uint32_t height, uint32_t interlace, const float *col_matrix12,
                      const float *gamma_lut);

/* Lanczos filter for axis-aligned Transforms: NOT in the reference (BASELINE.json config 5 / SURVEY 8f row 4); the
   definition lives in oracle.c.  Returns 0, or -1 for rotated transforms / more than 64 taps per axis. */
int orc_transform_lanczos(const float *in, int sw, int sh, const float *mat9, int lobes, float *out, int w, int h);

#ifdef __cplusplus
}
#endif
#endif
