/*
 * oracle.c -- CPU restatement of the Streampunk/phaneron pixel path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Plain C99, no dependencies.
 * Build: see oracle/Makefile (-O2 -ffp-contract=off; OpenMP optional).
 *
 * Canonical float semantics (the one member of the OpenCL-permitted family
 * that both this oracle and the CUDA kernels implement, bit for bit):
 *   - all arithmetic IEEE-754 binary32, round-to-nearest-even, no implicit
 *     contraction (this file is compiled with -ffp-contract=off and every
 *     fused operation is an explicit fmaf());
 *   - OpenCL dot(a,b) = the expansion LLVM produces for
 *     a.x*b.x + a.y*b.y + a.z*b.z (+ a.w*b.w) with contraction on (the OpenCL default),
 *     read off the PTX NVIDIA's OpenCL compiler emits for the reference's own kernels on the
 *     B200 (profiles/r01_reference_opencl_ptx_excerpt.txt): the first sum fuses as
 *     fma(a.x, b.x, a.y*b.y), so the .y product is the one plain multiply:
 *         t = a.y*b.y; t = fma(a.x,b.x,t); t = fma(a.z,b.z,t); [t = fma(a.w,b.w,t)]
 *   - fma() in kernel source is a single correctly rounded fma;
 *   - float division is correctly rounded;
 *   - convert_T_sat_rte: NaN -> 0, clamp, round half to even;
 *   - CLK_FILTER_LINEAR follows the OpenCL 1.2 spec section 8.2 formula in
 *     binary32 with full-precision weights:
 *         i0 = floor(u-0.5), a = (u-0.5) - floor(u-0.5)
 *         T = w00*T00 + w10*T10 + w01*T01 + w11*T11   evaluated as
 *         r = w00*T00; r = fma(w10,T10,r); r = fma(w01,T01,r); r = fma(w11,T11,r)
 *         with w00=(1-a)*(1-b), w10=a*(1-b), w01=(1-a)*b, w11=a*b
 *     CLK_ADDRESS_CLAMP border colour = (0,0,0,0) (CL_RGBA / CL_FLOAT).
 *   - host-side matrix algebra follows colourMaths.ts exactly: Float32Array
 *     storage rounding, accumulation in double.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int orc_get_threads(void) { return g_threads; }

#ifdef _OPENMP
#define PAR_FOR _Pragma("omp parallel for schedule(static) num_threads(g_threads)")
#else
#define PAR_FOR
#endif

/* ------------------------------------------------------------------------- */
/* colour constants: src/process/colourMaths.ts:42-128                        */
/* ------------------------------------------------------------------------- */
typedef struct {
	const char *name;
	double kR, kB, rx, ry, gx, gy, bx, by, wx, wy, alpha, beta, gamma, delta;
} ColParam;

static const ColParam COL_PARAMS[] = {
	{"601-625", 0.299, 0.114, 0.64, 0.33, 0.29, 0.6, 0.15, 0.06, 0.3127, 0.329, 1.099, 0.018, 0.45, 4.5},
	{"601_525", 0.299, 0.114, 0.63, 0.34, 0.31, 0.595, 0.155, 0.07, 0.3127, 0.329, 1.099, 0.018, 0.45, 4.5},
	{"709", 0.2126, 0.0722, 0.64, 0.33, 0.3, 0.6, 0.15, 0.06, 0.3127, 0.329, 1.099, 0.018, 0.45, 4.5},
	{"2020", 0.2627, 0.0593, 0.708, 0.292, 0.17, 0.797, 0.131, 0.046, 0.3127, 0.329, 1.099, 0.018, 0.45, 4.5},
	{"sRGB", 0.0, 0.0, 0.64, 0.33, 0.3, 0.6, 0.15, 0.06, 0.3127, 0.329, 1.055, 0.0031308, 1.0 / 2.4, 12.92},
};

/* unknown colSpec falls back to '709' (colourMaths.ts:131-134 and siblings) */
static const ColParam *col_find(const char *spec, int *known) {
	for (size_t i = 0; i < sizeof(COL_PARAMS) / sizeof(COL_PARAMS[0]); ++i)
		if (spec && 0 == strcmp(spec, COL_PARAMS[i].name)) {
			if (known) *known = 1;
			return &COL_PARAMS[i];
		}
	if (known) *known = 0;
	return &COL_PARAMS[2];
}

/* colourMaths.ts:130-149 */
int orc_gamma2linear_lut(const char *colspec, float *out) {
	int known;
	const ColParam *p = col_find(colspec, &known);
	const double alpha = p->alpha, delta = p->delta, beta = p->beta * delta, gamma = p->gamma;
	const int n = 1 << 16;
	for (int i = 0; i < n; ++i) {
		const double fi = (double)i / (double)(n - 1);
		if (fi < beta) out[i] = (float)(fi / delta);
		else out[i] = (float)pow((fi + (alpha - 1)) / alpha, 1 / gamma);
	}
	return known;
}

/* colourMaths.ts:151-169 */
int orc_linear2gamma_lut(const char *colspec, float *out) {
	int known;
	const ColParam *p = col_find(colspec, &known);
	const double alpha = p->alpha, beta = p->beta, gamma = p->gamma, delta = p->delta;
	const int n = 1 << 16;
	for (int i = 0; i < n; ++i) {
		const double fi = (double)i / (double)(n - 1);
		if (fi < beta) out[i] = (float)(fi * delta);
		else out[i] = (float)(alpha * pow(fi, gamma) - (alpha - 1));
	}
	return known;
}

/* Float32Array matrices, row-major, up to 3x4.  colourMaths.ts:171-178:
   result[i][j] = f32( sum_k (double)a[i][k] * (double)b[k][j] ), sum from 0.0 */
typedef struct {
	int r, c;
	float v[3][4];
} Mat;

static Mat mat_mul(const Mat *a, const Mat *b) {
	Mat o;
	memset(&o, 0, sizeof o);
	o.r = a->r;
	o.c = b->c;
	for (int i = 0; i < a->r; ++i)
		for (int j = 0; j < b->c; ++j) {
			double sum = 0.0;
			for (int k = 0; k < a->c; ++k) sum = sum + (double)a->v[i][k] * (double)b->v[k][j];
			o.v[i][j] = (float)sum;
		}
	return o;
}

/* colourMaths.ts:180-187 */
static Mat mat_scale(const Mat *a, double c) {
	Mat o = *a;
	for (int i = 0; i < a->r; ++i)
		for (int j = 0; j < a->c; ++j) o.v[i][j] = (float)((double)a->v[i][j] * c);
	return o;
}

/* colourMaths.ts:199-238 */
static Mat mat_invert3(const Mat *a) {
	Mat minors, cof, adj;
	memset(&minors, 0, sizeof minors);
	minors.r = minors.c = 3;
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) {
			int y[2], x[2];
			if (i == 1) { y[0] = 0; y[1] = 2; } else { y[0] = (i + 1) % 3; y[1] = (i + 2) % 3; }
			if (j == 1) { x[0] = 0; x[1] = 2; } else { x[0] = (j + 1) % 3; x[1] = (j + 2) % 3; }
			const double m00 = a->v[y[0]][x[0]], m01 = a->v[y[0]][x[1]];
			const double m10 = a->v[y[1]][x[0]], m11 = a->v[y[1]][x[1]];
			minors.v[i][j] = (float)(m00 * m11 - m01 * m10);
		}
	cof = minors;
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			cof.v[i][j] = (float)((double)minors.v[i][j] * (((i + j) & 1) ? -1.0 : 1.0));
	adj = cof;
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) adj.v[c][r] = cof.v[r][c];
	const double det = (double)a->v[0][0] * (double)minors.v[0][0] -
	                   (double)a->v[0][1] * (double)minors.v[0][1] +
	                   (double)a->v[0][2] * (double)minors.v[0][2];
	return mat_scale(&adj, 1.0 / det);
}

/* colourMaths.ts:240-266 */
static Mat rgb2xyz(const ColParam *p) {
	Mat w, xyz, scale;
	memset(&w, 0, sizeof w);
	memset(&xyz, 0, sizeof xyz);
	memset(&scale, 0, sizeof scale);
	w.r = 3; w.c = 1;
	w.v[0][0] = (float)p->wx;
	w.v[1][0] = (float)p->wy;
	w.v[2][0] = (float)(1.0 - p->wx - p->wy);
	Mat W = mat_scale(&w, 1.0 / (double)w.v[1][0]);
	xyz.r = xyz.c = 3;
	xyz.v[0][0] = (float)p->rx; xyz.v[0][1] = (float)p->gx; xyz.v[0][2] = (float)p->bx;
	xyz.v[1][0] = (float)p->ry; xyz.v[1][1] = (float)p->gy; xyz.v[1][2] = (float)p->by;
	xyz.v[2][0] = (float)(1.0 - p->rx - p->ry);
	xyz.v[2][1] = (float)(1.0 - p->gx - p->gy);
	xyz.v[2][2] = (float)(1.0 - p->bx - p->by);
	Mat inv = mat_invert3(&xyz);
	Mat f = mat_mul(&inv, &W);
	scale.r = scale.c = 3;
	scale.v[0][0] = f.v[0][0];
	scale.v[1][1] = f.v[1][0];
	scale.v[2][2] = f.v[2][0];
	return mat_mul(&xyz, &scale);
}

/* colourMaths.ts:392-394 (xyz2rgb: 268-274) */
int orc_rgb2rgb_matrix(const char *src, const char *dst, float *out9) {
	int k1, k2;
	const ColParam *ps = col_find(src, &k1), *pd = col_find(dst, &k2);
	Mat d = rgb2xyz(pd);
	Mat dinv = mat_invert3(&d);
	Mat s = rgb2xyz(ps);
	Mat m = mat_mul(&dinv, &s);
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) out9[i * 3 + j] = m.v[i][j];
	return k1 && k2;
}

/* colourMaths.ts:276-332 */
int orc_ycbcr2rgb_matrix(const char *colspec, int num_bits, int luma_black, int luma_white,
                         int chr_range, float *out12) {
	int known;
	const ColParam *p = col_find(colspec, &known);
	const double chrNull = (double)(128 << (num_bits - 8));
	const double lumaRange = luma_white - luma_black;
	const double kR = p->kR, kB = p->kB, kG = 1.0 - kR - kB;
	Mat col, sc;
	memset(&col, 0, sizeof col);
	memset(&sc, 0, sizeof sc);
	col.r = 3; col.c = 3;
	col.v[0][0] = 1.0f; col.v[0][1] = 0.0f; col.v[0][2] = (float)(1.0 - kR);
	col.v[1][0] = 1.0f; col.v[1][1] = (float)((-(1.0 - kB) * kB) / kG);
	col.v[1][2] = (float)((-(1.0 - kR) * kR) / kG);
	col.v[2][0] = 1.0f; col.v[2][1] = (float)(1.0 - kB); col.v[2][2] = 0.0f;
	sc.r = 3; sc.c = 4;
	sc.v[0][0] = (float)(1.0 / lumaRange); sc.v[0][3] = (float)(-(double)luma_black / lumaRange);
	sc.v[1][1] = (float)((1.0 / chr_range) * 2); sc.v[1][3] = (float)(-(chrNull / chr_range) * 2);
	sc.v[2][2] = (float)((1.0 / chr_range) * 2); sc.v[2][3] = (float)(-(chrNull / chr_range) * 2);
	Mat m = mat_mul(&col, &sc);
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 4; ++j) out12[i * 4 + j] = m.v[i][j];
	return known;
}

/* colourMaths.ts:334-390 */
int orc_rgb2ycbcr_matrix(const char *colspec, int num_bits, int luma_black, int luma_white,
                         int chr_range, float *out12) {
	int known;
	const ColParam *p = col_find(colspec, &known);
	const double chrNull = (double)(128 << (num_bits - 8));
	const double lumaRange = luma_white - luma_black;
	const double kR = p->kR, kB = p->kB, kG = 1.0 - kR - kB;
	Mat sc, col;
	memset(&col, 0, sizeof col);
	memset(&sc, 0, sizeof sc);
	sc.r = 3; sc.c = 3;
	sc.v[0][0] = (float)lumaRange;
	sc.v[1][1] = (float)(chr_range / 2.0);
	sc.v[2][2] = (float)(chr_range / 2.0);
	col.r = 3; col.c = 4;
	col.v[0][0] = (float)kR; col.v[0][1] = (float)kG; col.v[0][2] = (float)kB;
	col.v[0][3] = (float)((double)luma_black / lumaRange);
	col.v[1][0] = (float)(-kR / (1.0 - kB)); col.v[1][1] = (float)(-kG / (1.0 - kB));
	col.v[1][2] = (float)((1.0 - kB) / (1.0 - kB)); col.v[1][3] = (float)((chrNull / chr_range) * 2.0);
	col.v[2][0] = (float)((1.0 - kR) / (1.0 - kR)); col.v[2][1] = (float)(-kG / (1.0 - kR));
	col.v[2][2] = (float)(-kB / (1.0 - kR)); col.v[2][3] = (float)((chrNull / chr_range) * 2.0);
	Mat m = mat_mul(&sc, &col);
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 4; ++j) out12[i * 4 + j] = m.v[i][j];
	return known;
}

/* transform.ts:119-171 (host side of Transform.getKernelParams) */
void orc_transform_matrix(int width, int height, int flip_h, int flip_v, double anchorX,
                          double anchorY, double scale_x, double scale_y, double offsetX,
                          double offsetY, double rotate_turns, float *out9) {
	const double aspect = (double)width / (double)height;
	const double flipX = flip_h ? -1.0 : 1.0, flipY = flip_v ? -1.0 : 1.0;
	/* `(params.scaleX as number) || 1.0`: 0 / NaN / undefined become 1.0 */
	const double scaleX = ((scale_x != 0.0 && scale_x == scale_x) ? scale_x : 1.0) * flipX;
	const double scaleY = ((scale_y != 0.0 && scale_y == scale_y) ? scale_y : 1.0) * flipY;
	const double rotate = rotate_turns * 2 * 3.141592653589793;
	Mat ai, sc, ro, tr, ao, pr;
#define ID3(m) do { memset(&(m), 0, sizeof(m)); (m).r = (m).c = 3; (m).v[0][0] = (m).v[1][1] = (m).v[2][2] = 1.0f; } while (0)
	ID3(ai); ID3(sc); ID3(ro); ID3(tr); ID3(ao); ID3(pr);
	ai.v[0][2] = (float)anchorX; ai.v[1][2] = (float)anchorY;
	sc.v[0][0] = (float)(1.0 / (scaleX * aspect)); sc.v[1][1] = (float)(1.0 / scaleY);
	ro.v[0][0] = (float)cos(rotate); ro.v[0][1] = (float)(-sin(rotate));
	ro.v[1][0] = (float)sin(rotate); ro.v[1][1] = (float)cos(rotate);
	tr.v[0][2] = (float)(offsetX * aspect); tr.v[1][2] = (float)offsetY;
	ao.v[0][2] = (float)(-anchorX * aspect); ao.v[1][2] = (float)(-anchorY);
	pr.v[0][0] = (float)aspect;
	Mat m = mat_mul(&ai, &sc);
	m = mat_mul(&m, &ro);
	m = mat_mul(&m, &tr);
	m = mat_mul(&m, &ao);
	m = mat_mul(&m, &pr);
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) out9[i * 3 + j] = m.v[i][j];
}

/* ------------------------------------------------------------------------- */
/* OpenCL built-ins under the canonical semantics                            */
/* ------------------------------------------------------------------------- */
static inline float dot4(const float a[4], const float b[4]) {
	float t = a[1] * b[1];   /* a.x*b.x + a.y*b.y contracts to fma(a.x, b.x, a.y*b.y): the .y product is the plain multiply */
	t = fmaf(a[0], b[0], t);
	t = fmaf(a[2], b[2], t);
	t = fmaf(a[3], b[3], t);
	return t;
}
static inline float dot3(const float a[3], const float b[3]) {
	float t = a[1] * b[1];
	t = fmaf(a[0], b[0], t);
	t = fmaf(a[2], b[2], t);
	return t;
}
/* convert_ushort_sat_rte / convert_uchar_sat_rte (OpenCL 1.2 6.2.3.3): NaN->0 */
static inline uint32_t sat_rte(float x, float hi) {
	if (!(x > 0.0f)) return 0;
	if (x >= hi) return (uint32_t)hi;
	return (uint32_t)nearbyintf(x); /* default rounding mode: nearest-even */
}
static inline uint32_t sat_rtz(float x, float hi) {
	if (!(x > 0.0f)) return 0;
	if (x >= hi) return (uint32_t)hi;
	return (uint32_t)x;
}

/* ------------------------------------------------------------------------- */
/* v210: src/process/v210.ts                                                  */
/* ------------------------------------------------------------------------- */
/* v210.ts:198-204 */
uint32_t orc_v210_pitch(uint32_t width) { return width + 47 - ((width - 1) % 48); }
uint32_t orc_v210_pitch_bytes(uint32_t width) { return orc_v210_pitch(width) * 8 / 3; }

static inline void wr32(uint8_t *b, size_t off, uint32_t v) {
	b[off] = (uint8_t)v; b[off + 1] = (uint8_t)(v >> 8); b[off + 2] = (uint8_t)(v >> 16); b[off + 3] = (uint8_t)(v >> 24);
}

/* v210.ts:206-236 fillBuf */
void orc_v210_fill(uint8_t *buf, uint32_t width, uint32_t height) {
	const uint32_t pitchBytes = orc_v210_pitch_bytes(width);
	memset(buf, 0, (size_t)pitchBytes * height);
	uint32_t Y = 64;
	const uint32_t Cb = 512, Cr = 512;
	size_t yOff = 0;
	for (uint32_t y = 0; y < height; ++y) {
		size_t xOff = 0;
		for (uint32_t x = 0; x < (width - (width % 6)) / 6; ++x) {
			wr32(buf, yOff + xOff, (Cr << 20) | (Y << 10) | Cb);
			wr32(buf, yOff + xOff + 4, (Y << 20) | (Cb << 10) | Y);
			wr32(buf, yOff + xOff + 8, (Cb << 20) | (Y << 10) | Cr);
			wr32(buf, yOff + xOff + 12, (Y << 20) | (Cr << 10) | Y);
			xOff += 16;
			Y = (940 == Y) ? 64 : Y + 1;
		}
		const uint32_t remain = width % 6;
		if (remain) {
			wr32(buf, yOff + xOff, (Cr << 20) | (Y << 10) | Cb);
			if (2 == remain) {
				wr32(buf, yOff + xOff + 4, Y);
			} else if (4 == remain) {
				wr32(buf, yOff + xOff + 4, (Y << 20) | (Cb << 10) | Y);
				wr32(buf, yOff + xOff + 8, (Y << 10) | Cr);
			}
		}
		yOff += pitchBytes;
	}
}

/* one pixel of the read kernel, v210.ts:65-78 (alpha=1) / 96-109 (alpha=0) */
static inline void read_px(const uint32_t yuv[3], float alpha, const float *cm, const float *lut,
                           const float *gamut, float *out) {
	const float yuva_f[4] = {(float)yuv[0], (float)yuv[1], (float)yuv[2], alpha};
	float rgb[3];
	rgb[0] = lut[sat_rte(dot4(yuva_f, cm + 0) * 65535.0f, 65535.0f)];
	rgb[1] = lut[sat_rte(dot4(yuva_f, cm + 4) * 65535.0f, 65535.0f)];
	rgb[2] = lut[sat_rte(dot4(yuva_f, cm + 8) * 65535.0f, 65535.0f)];
	out[0] = dot3(rgb, gamut + 0);
	out[1] = dot3(rgb, gamut + 3);
	out[2] = dot3(rgb, gamut + 6);
	out[3] = 1.0f;
}

/* v210.ts:25-111.  One work-group per line, one work-item per 48 pixels. */
void orc_v210_read(const uint32_t *input, float *output, uint32_t width, uint32_t height,
                   const float *cm, const float *lut, const float *gamut) {
	const uint32_t itemsPerLine = orc_v210_pitch(width) / 48;
	PAR_FOR
	for (int64_t line = 0; line < (int64_t)height; ++line) {
		for (uint32_t lid = 0; lid < itemsPerLine; ++lid) {
			const uint32_t item = (uint32_t)line * itemsPerLine + lid;
			const int last = lid == itemsPerLine - 1;
			const uint32_t numPixels = (last && (0 != width % 48)) ? width % 48 : 48;
			const uint32_t numLoops = numPixels / 6, remain = numPixels % 6;
			size_t inOff = (size_t)8 * item;
			size_t outOff = (size_t)width * line + (size_t)lid * 48;
			for (uint32_t i = 0; i < numLoops; ++i) {
				const uint32_t *w = input + inOff * 4;
				uint32_t yuva[6][3];
				yuva[0][0] = (w[0] >> 10) & 0x3ff; yuva[0][1] = w[0] & 0x3ff; yuva[0][2] = (w[0] >> 20) & 0x3ff;
				yuva[1][0] = w[1] & 0x3ff; yuva[1][1] = yuva[0][1]; yuva[1][2] = yuva[0][2];
				yuva[2][0] = (w[1] >> 20) & 0x3ff; yuva[2][1] = (w[1] >> 10) & 0x3ff; yuva[2][2] = w[2] & 0x3ff;
				yuva[3][0] = (w[2] >> 10) & 0x3ff; yuva[3][1] = yuva[2][1]; yuva[3][2] = yuva[2][2];
				yuva[4][0] = w[3] & 0x3ff; yuva[4][1] = (w[2] >> 20) & 0x3ff; yuva[4][2] = (w[3] >> 10) & 0x3ff;
				yuva[5][0] = (w[3] >> 20) & 0x3ff; yuva[5][1] = yuva[4][1]; yuva[5][2] = yuva[4][2];
				for (uint32_t p = 0; p < 6; ++p)
					read_px(yuva[p], 1.0f, cm, lut, gamut, output + (outOff + p) * 4);
				inOff++;
				outOff += 6;
			}
			if (remain > 0) {
				const uint32_t *w = input + inOff * 4;
				uint32_t yuva[4][3] = {{0}};
				yuva[0][0] = (w[0] >> 10) & 0x3ff; yuva[0][1] = w[0] & 0x3ff; yuva[0][2] = (w[0] >> 20) & 0x3ff;
				yuva[1][0] = w[1] & 0x3ff; yuva[1][1] = yuva[0][1]; yuva[1][2] = yuva[0][2];
				if (4 == remain) {
					yuva[2][0] = (w[1] >> 20) & 0x3ff; yuva[2][1] = (w[1] >> 10) & 0x3ff; yuva[2][2] = w[2] & 0x3ff;
					yuva[3][0] = (w[2] >> 10) & 0x3ff; yuva[3][1] = yuva[2][1]; yuva[3][2] = yuva[2][2];
				}
				/* Q1: tail builds yuva with alpha 0, dropping the matrix offset column */
				for (uint32_t p = 0; p < remain; ++p)
					read_px(yuva[p], 0.0f, cm, lut, gamut, output + (outOff + p) * 4);
			}
		}
	}
}

/* v210.ts:113-195 */
void orc_v210_write(const float *input, uint32_t *output, uint32_t width, uint32_t height,
                    uint32_t interlace, const float *cm, const float *lut, int q3_literal) {
	const uint32_t pitch = orc_v210_pitch(width);
	const uint32_t itemsPerLine = pitch / 48;
	const uint32_t groups = (0 == interlace) ? height : height / 2;
	/* with q3_literal the work-items of different lines can overlap (the
	   reference races); run serially in ascending order to be deterministic */
	const int par = !q3_literal || (width % 48 == 0);
	(void)par;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(g_threads) if (par)
#endif
	for (int64_t gid = 0; gid < (int64_t)groups; ++gid) {
		for (uint32_t lid = 0; lid < itemsPerLine; ++lid) {
			const int last = lid == itemsPerLine - 1;
			const uint32_t numPixels = (last && (0 != width % 48)) ? width % 48 : 48;
			const uint32_t numLoops = numPixels / 6, remain = numPixels % 6;
			const uint32_t interlaceOff = (3 == interlace) ? 1 : 0;
			const uint32_t line = (uint32_t)gid * ((0 == interlace) ? 1 : 2) + interlaceOff;
			size_t inOff = (size_t)width * line + (size_t)lid * 48;
			size_t outOff = (q3_literal ? (size_t)width * line / 6 : (size_t)pitch * line / 6) + (size_t)lid * 8;
			if (48 != numPixels) {
				size_t clearOff = outOff;
				for (uint32_t i = 0; i < 8; ++i) {
					memset(output + clearOff * 4, 0, 16);
					clearOff++;
				}
			}
			for (uint32_t i = 0; i < numLoops; ++i) {
				uint32_t yuv[6][3];
				for (uint32_t p = 0; p < 6; ++p) {
					const float *l = input + (inOff + p) * 4;
					float rgba[4];
					rgba[0] = lut[sat_rte(l[0] * 65535.0f, 65535.0f)];
					rgba[1] = lut[sat_rte(l[1] * 65535.0f, 65535.0f)];
					rgba[2] = lut[sat_rte(l[2] * 65535.0f, 65535.0f)];
					rgba[3] = 1.0f;
					yuv[p][0] = sat_rte(dot4(rgba, cm + 0), 65535.0f);
					yuv[p][1] = sat_rte(dot4(rgba, cm + 4), 65535.0f);
					yuv[p][2] = sat_rte(dot4(rgba, cm + 8), 65535.0f);
				}
				uint32_t *w = output + outOff * 4;
				w[0] = yuv[0][2] << 20 | yuv[0][0] << 10 | yuv[0][1];
				w[1] = yuv[2][0] << 20 | yuv[2][1] << 10 | yuv[1][0];
				w[2] = yuv[4][1] << 20 | yuv[3][0] << 10 | yuv[2][2];
				w[3] = yuv[5][0] << 20 | yuv[4][2] << 10 | yuv[4][0];
				inOff += 6;
				outOff++;
			}
			if (remain > 0) {
				uint32_t w[4] = {0, 0, 0, 0};
				uint32_t yuv[4][3] = {{0}};
				for (uint32_t p = 0; p < remain; ++p) {
					const float *l = input + (inOff + p) * 4;
					float rgba[4];
					/* Q2: tail uses _rtz for the LUT index and round() (half away) */
					rgba[0] = lut[sat_rtz(l[0] * 65535.0f, 65535.0f)];
					rgba[1] = lut[sat_rtz(l[1] * 65535.0f, 65535.0f)];
					rgba[2] = lut[sat_rtz(l[2] * 65535.0f, 65535.0f)];
					rgba[3] = 1.0f;
					yuv[p][0] = sat_rtz(roundf(dot4(rgba, cm + 0)), 65535.0f);
					yuv[p][1] = sat_rtz(roundf(dot4(rgba, cm + 4)), 65535.0f);
					yuv[p][2] = sat_rtz(roundf(dot4(rgba, cm + 8)), 65535.0f);
				}
				w[0] = yuv[0][2] << 20 | yuv[0][0] << 10 | yuv[0][1];
				if (2 == remain) {
					w[1] = yuv[1][0];
				} else if (4 == remain) {
					w[1] = yuv[2][0] << 20 | yuv[2][1] << 10 | yuv[1][0];
					w[2] = yuv[3][0] << 10 | yuv[2][2];
				}
				memcpy(output + outOff * 4, w, 16);
			}
		}
	}
}

/* ------------------------------------------------------------------------- */
/* image ops                                                                 */
/* ------------------------------------------------------------------------- */
/* combine.ts:24-68: out = fma(prev, (k,k,k,0), lN), k = 1 - lN.a */
void orc_combine(const float *const *layers, int n, float *out, int w, int h) {
	PAR_FOR
	for (int64_t y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x) {
			const size_t o = ((size_t)y * w + x) * 4;
			float acc[4];
			memcpy(acc, layers[0] + o, 16);
			for (int i = 1; i < n; ++i) {
				const float *l = layers[i] + o;
				const float k = 1.0f - l[3];
				acc[0] = fmaf(acc[0], k, l[0]);
				acc[1] = fmaf(acc[1], k, l[1]);
				acc[2] = fmaf(acc[2], k, l[2]);
				acc[3] = fmaf(acc[3], 0.0f, l[3]);
			}
			memcpy(out + o, acc, 16);
		}
}

/* transition.ts:60-65: out = fma(in0, mix4, in1 * rmix) */
void orc_dissolve(const float *in0, const float *in1, float mix, float *out, int w, int h) {
	const float rmix = 1.0f - mix;
	PAR_FOR
	for (int64_t i = 0; i < (int64_t)w * h * 4; ++i) out[i] = fmaf(in0[i], mix, in1[i] * rmix);
}

/* mix.ts:30-45 (dead in the reference): same arithmetic as dissolve */
void orc_mix(const float *in0, const float *in1, float mix, float *out, int w, int h) {
	orc_dissolve(in0, in1, mix, out, w, h);
}

/* transition.ts:66-73: m = mask.r; out = fma(in1, m4, in0 * (1-m)) */
void orc_wipe_mask(const float *in0, const float *in1, const float *mask, float *out, int w, int h) {
	PAR_FOR
	for (int64_t p = 0; p < (int64_t)w * h; ++p) {
		const float m = mask[p * 4], rm = 1.0f - m;
		for (int c = 0; c < 4; ++c) out[p * 4 + c] = fmaf(in1[p * 4 + c], m, in0[p * 4 + c] * rm);
	}
}

/* wipe.ts:30-47 (dead): out = x > w*wipe ? in1 : in0 */
void orc_wipe(const float *in0, const float *in1, float wipe, float *out, int w, int h) {
	const float edge = (float)w * wipe;
	PAR_FOR
	for (int64_t y = 0; y < h; ++y)
		for (int x = 0; x < w; ++x) {
			const size_t o = ((size_t)y * w + x) * 4;
			memcpy(out + o, ((float)x > edge ? in1 : in0) + o, 16);
		}
}

/* read_imagef(CLK_NORMALIZED_COORDS_TRUE | CLK_ADDRESS_CLAMP | CLK_FILTER_LINEAR),
   OpenCL 1.2 spec 8.2, canonical evaluation order (see header) */
static inline void sample_linear_clamp(const float *in, int sw, int sh, float s, float t, float *out) {
	const float u = s * (float)sw, v = t * (float)sh;
	const float um = u - 0.5f, vm = v - 0.5f;
	const float fu = floorf(um), fv = floorf(vm);
	const float a = um - fu, b = vm - fv;
	/* coordinates far outside the image (or NaN) address only border texels */
	int i0, j0;
	if (!(fu >= -2.0f)) i0 = -2; else if (fu > (float)sw) i0 = sw; else i0 = (int)fu;
	if (!(fv >= -2.0f)) j0 = -2; else if (fv > (float)sh) j0 = sh; else j0 = (int)fv;
	const int i1 = i0 + 1, j1 = j0 + 1;
	const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
	static const float border[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#define TEXEL(i, j) (((i) < 0 || (i) >= sw || (j) < 0 || (j) >= sh) ? border : in + ((size_t)(j) * sw + (i)) * 4)
	const float *t00 = TEXEL(i0, j0), *t10 = TEXEL(i1, j0), *t01 = TEXEL(i0, j1), *t11 = TEXEL(i1, j1);
#undef TEXEL
	for (int c = 0; c < 4; ++c) {
		float r = w00 * t00[c];
		r = fmaf(w10, t10[c], r);
		r = fmaf(w01, t01[c], r);
		r = fmaf(w11, t11[c], r);
		out[c] = r;
	}
}

/* transform.ts:36-59 */
void orc_transform(const float *in, int sw, int sh, const float *mat9, float *out, int w, int h) {
	const float mat0[3] = {mat9[0], mat9[1], mat9[2]};
	const float mat1[3] = {mat9[3], mat9[4], mat9[5]};
	PAR_FOR
	for (int64_t outY = 0; outY < h; ++outY)
		for (int outX = 0; outX < w; ++outX) {
			const float inPos[3] = {(float)outX / (float)w - 0.5f, (float)outY / (float)h - 0.5f, 1.0f};
			const float px = dot3(mat0, inPos) + 0.5f, py = dot3(mat1, inPos) + 0.5f;
			sample_linear_clamp(in, sw, sh, px, py, out + ((size_t)outY * w + outX) * 4);
		}
}

/* resize.ts:35-59 (dead) */
void orc_resize(const float *in, int sw, int sh, float scale, float offsetX, float offsetY,
                const float *flip, float *out, int w, int h) {
	const float centreOffX = (-0.5f - offsetX) / scale + 0.5f;
	const float centreOffY = (-0.5f - offsetY) / scale + 0.5f;
	const float offX = fmaf(centreOffX, flip[1], flip[0]), offY = fmaf(centreOffY, flip[3], flip[2]);
	const float mulX = flip[1] / scale, mulY = flip[3] / scale;
	PAR_FOR
	for (int64_t outY = 0; outY < h; ++outY)
		for (int outX = 0; outX < w; ++outX) {
			const float ix = (float)outX / (float)w, iy = (float)outY / (float)h;
			sample_linear_clamp(in, sw, sh, fmaf(ix, mulX, offX), fmaf(iy, mulY, offY),
			                    out + ((size_t)outY * w + outX) * 4);
		}
}

/* yadifCl.ts:34-62; every relational / select is component-wise */
static inline float spatial_predictor(float a, float b, float c, float d, float e, float f, float g,
                                      float h, float i, float j, float k, float l, float m, float n) {
	float spatialPred = (d + k) / 2.0f;
	float spatialScore = fabsf(c - j) + fabsf(d - k) + fabsf(e - l);

	float score = fabsf(b - k) + fabsf(c - l) + fabsf(d - m);
	int cmp = score < spatialScore;
	spatialPred = cmp ? (c + l) / 2.0f : spatialPred;
	spatialScore = cmp ? score : spatialScore;
	score = cmp ? fabsf(a - l) + fabsf(b - m) + fabsf(c - n) : score;
	cmp = cmp && (score < spatialScore);
	spatialPred = cmp ? (b + m) / 2.0f : spatialPred;
	spatialScore = cmp ? score : spatialScore;

	score = fabsf(d - i) + fabsf(e - j) + fabsf(f - k);
	cmp = score < spatialScore;
	spatialPred = cmp ? (e + j) / 2.0f : spatialPred;
	spatialScore = cmp ? score : spatialScore;
	score = cmp ? fabsf(e - h) + fabsf(f - i) + fabsf(g - j) : score;
	cmp = cmp && (score < spatialScore);
	spatialPred = cmp ? (f + i) / 2.0f : spatialPred;
	spatialScore = cmp ? score : spatialScore;
	(void)spatialScore;
	return spatialPred;
}

static inline float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
static inline float fmin3(float a, float b, float c) { return fminf(fminf(a, b), c); }

/* yadifCl.ts:72-103 */
static inline float temporal_predictor(float A, float B, float C, float D, float E, float F, float G,
                                       float H, float I, float J, float K, float L,
                                       float spatialPred, int skipCheck) {
	const float p0 = (C + H) / 2.0f, p1 = F, p2 = (D + I) / 2.0f, p3 = G, p4 = (E + J) / 2.0f;
	const float tdiff0 = fabsf(D - I);
	const float tdiff1 = (fabsf(A - F) + fabsf(B - G)) / 2.0f;
	const float tdiff2 = (fabsf(K - F) + fabsf(G - L)) / 2.0f;
	float diff = fmax3(tdiff0, tdiff1, tdiff2);
	if (!skipCheck) {
		const float p2mp3 = p2 - p3, p2mp1 = p2 - p1, p0mp1 = p0 - p1, p4mp3 = p4 - p3;
		const float maxi = fmax3(p2mp3, p2mp1, fminf(p0mp1, p4mp3));
		const float mini = fmin3(p2mp3, p2mp1, fmaxf(p0mp1, p4mp3));
		diff = fmax3(diff, mini, -maxi);
	}
	spatialPred = (spatialPred > (p2 + diff)) ? p2 + diff : spatialPred;
	spatialPred = (spatialPred < (p2 - diff)) ? p2 - diff : spatialPred;
	return spatialPred;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* yadifCl.ts:105-167; sampler = unnormalised, CLAMP_TO_EDGE, nearest */
void orc_yadif(const float *prev, const float *cur, const float *next, int parity, int tff,
               int skipSpatial, float *out, int w, int h) {
#define PX(img, x, y) ((img) + ((size_t)clampi((y), 0, h - 1) * w + clampi((x), 0, w - 1)) * 4)
	PAR_FOR
	for (int64_t yy = 0; yy < h; ++yy) {
		const int yo = (int)yy;
		for (int xo = 0; xo < w; ++xo) {
			float *o = out + ((size_t)yo * w + xo) * 4;
			if (yo % 2 == parity) {
				memcpy(o, PX(cur, xo, yo), 16);
				continue;
			}
			const int isSecondField = !(parity ^ tff);
			const float *a = PX(cur, xo - 3, yo - 1), *b = PX(cur, xo - 2, yo - 1), *c = PX(cur, xo - 1, yo - 1),
			            *d = PX(cur, xo, yo - 1), *e = PX(cur, xo + 1, yo - 1), *f = PX(cur, xo + 2, yo - 1),
			            *g = PX(cur, xo + 3, yo - 1);
			const float *hh = PX(cur, xo - 3, yo + 1), *i = PX(cur, xo - 2, yo + 1), *j = PX(cur, xo - 1, yo + 1),
			            *k = PX(cur, xo, yo + 1), *l = PX(cur, xo + 1, yo + 1), *m = PX(cur, xo + 2, yo + 1),
			            *n = PX(cur, xo + 3, yo + 1);
			const float *A = PX(prev, xo, yo - 1), *B = PX(prev, xo, yo + 1);
			const float *C = isSecondField ? PX(cur, xo, yo - 2) : PX(prev, xo, yo - 2);
			const float *D = isSecondField ? PX(cur, xo, yo) : PX(prev, xo, yo);
			const float *E = isSecondField ? PX(cur, xo, yo + 2) : PX(prev, xo, yo + 2);
			const float *F = PX(cur, xo, yo - 1), *G = PX(cur, xo, yo + 1);
			const float *H = isSecondField ? PX(next, xo, yo - 2) : PX(cur, xo, yo - 2);
			const float *I = isSecondField ? PX(next, xo, yo) : PX(cur, xo, yo);
			const float *J = isSecondField ? PX(next, xo, yo + 2) : PX(cur, xo, yo + 2);
			const float *K = PX(next, xo, yo - 1), *L = PX(next, xo, yo + 1);
			for (int ch = 0; ch < 4; ++ch) {
				float sp = spatial_predictor(a[ch], b[ch], c[ch], d[ch], e[ch], f[ch], g[ch], hh[ch], i[ch],
				                             j[ch], k[ch], l[ch], m[ch], n[ch]);
				o[ch] = temporal_predictor(A[ch], B[ch], C[ch], D[ch], E[ch], F[ch], G[ch], H[ch], I[ch],
				                           J[ch], K[ch], L[ch], sp, skipSpatial);
			}
			o[3] = PX(cur, xo, yo)[3];
		}
	}
#undef PX
}

/* ------------------------------------------------------------------------- */
/* rgba8 / bgra8: src/process/rgba8.ts:25-103, bgra8.ts:25-103                */
/* ------------------------------------------------------------------------- */
void orc_rgba8_read(const uint8_t *input, float *output, uint32_t width, uint32_t height,
                    const float *lut, const float *gamut, int bgra) {
	PAR_FOR
	for (int64_t p = 0; p < (int64_t)width * height; ++p) {
		const uint8_t *in = input + p * 4;
		const float c0 = (float)in[bgra ? 2 : 0], c1 = (float)in[1], c2 = (float)in[bgra ? 0 : 2], c3 = (float)in[3];
		float rgb[3];
		rgb[0] = lut[sat_rte(c0 * 65535.0f / 255.0f, 65535.0f)];
		rgb[1] = lut[sat_rte(c1 * 65535.0f / 255.0f, 65535.0f)];
		rgb[2] = lut[sat_rte(c2 * 65535.0f / 255.0f, 65535.0f)];
		float *o = output + p * 4;
		o[0] = dot3(rgb, gamut + 0);
		o[1] = dot3(rgb, gamut + 3);
		o[2] = dot3(rgb, gamut + 6);
		o[3] = lut[sat_rte(c3 * 65535.0f / 255.0f, 65535.0f)];
	}
}

void orc_rgba8_write(const float *input, uint8_t *output, uint32_t width, uint32_t height,
                     uint32_t interlace, const float *lut, int bgra) {
	const uint32_t groups = (0 == interlace) ? height : height / 2;
	PAR_FOR
	for (int64_t gid = 0; gid < (int64_t)groups; ++gid) {
		const uint32_t line = (uint32_t)gid * ((0 == interlace) ? 1 : 2) + ((3 == interlace) ? 1 : 0);
		for (uint32_t x = 0; x < width; ++x) {
			const float *l = input + ((size_t)line * width + x) * 4;
			uint8_t *o = output + ((size_t)line * width + x) * 4;
			const float r = lut[sat_rte(l[0] * 65535.0f, 65535.0f)];
			const float g = lut[sat_rte(l[1] * 65535.0f, 65535.0f)];
			const float b = lut[sat_rte(l[2] * 65535.0f, 65535.0f)];
			o[bgra ? 2 : 0] = (uint8_t)sat_rte(r * 255.0f, 255.0f);
			o[1] = (uint8_t)sat_rte(g * 255.0f, 255.0f);
			o[bgra ? 0 : 2] = (uint8_t)sat_rte(b * 255.0f, 255.0f);
			o[3] = 255;
		}
	}
}

/* ------------------------------------------------------------------------- */
/* yuv422p10le / yuv422p8: src/process/yuv422p10.ts, src/process/yuv422p8.ts  */
/* The two files differ only in the sample type (ushort / uchar), the tail    */
/* fill values and the host-side range constants; `bits` selects.            */
/* ------------------------------------------------------------------------- */
uint32_t orc_yuv422p_pitch(uint32_t width) { return width + 7 - ((width - 1) % 8); } /* yuv422p10.ts:222, pixels */

static inline uint32_t ld_s(const uint8_t *plane, size_t i, int bits) {
	return bits == 8 ? plane[i] : (uint32_t)plane[2 * i] | (uint32_t)plane[2 * i + 1] << 8;
}
static inline void st_s(uint8_t *plane, size_t i, uint32_t v, int bits) {
	if (bits == 8) {
		plane[i] = (uint8_t)v; /* Q11: convert_ushort_sat_rte() assigned to a uchar field wraps (yuv422p8.ts:153,163-165) */
	} else {
		plane[2 * i] = (uint8_t)v;
		plane[2 * i + 1] = (uint8_t)(v >> 8);
	}
}

/* fillBuf: yuv422p10.ts:225-255 / yuv422p8.ts:225-253 (one buffer: Y plane | U plane | V plane) */
void orc_yuv422p_fill(int bits, uint8_t *buf, uint32_t width, uint32_t height) {
	const uint32_t pitch = orc_yuv422p_pitch(width);
	const uint32_t black = bits == 8 ? 16 : 64, grey = bits == 8 ? 128 : 512, wrap = bits == 8 ? 234 : 938;
	uint8_t *py = buf, *pu = buf + (size_t)pitch * height * (bits == 8 ? 1 : 2), *pv = pu + (size_t)(pitch / 2) * height * (bits == 8 ? 1 : 2);
	for (size_t i = 0; i < (size_t)pitch * height; ++i) st_s(py, i, black, bits);
	for (size_t i = 0; i < (size_t)(pitch / 2) * height; ++i) {
		st_s(pu, i, grey, bits);
		st_s(pv, i, grey, bits);
	}
	uint32_t Y = black;
	for (uint32_t y = 0; y < height; ++y) {
		for (uint32_t x = 0; x < width; x += 2) {
			st_s(py, (size_t)y * pitch + x, Y, bits);
			st_s(py, (size_t)y * pitch + x + 1, Y + 1, bits);
			st_s(pu, (size_t)y * (pitch / 2) + x / 2, grey, bits);
			st_s(pv, (size_t)y * (pitch / 2) + x / 2, grey, bits);
			Y = (wrap == Y) ? black : Y + 2;
		}
	}
}

/* read kernel: yuv422p10.ts:25-124.  One work-group per line, 64 pixels per work-item, blocks of 8. */
void orc_yuv422p_read(int bits, const uint8_t *inY, const uint8_t *inU, const uint8_t *inV, float *output, uint32_t width,
                      uint32_t height, const float *cm, const float *lut, const float *gamut) {
	const uint32_t itemsPerLine = (orc_yuv422p_pitch(width) + 63) / 64; /* Math.ceil(getPitch/64), yuv422p10.ts:308 */
	const uint32_t pitchReads = (width + 7) / 8;
	PAR_FOR
	for (int64_t gid = 0; gid < (int64_t)height; ++gid) {
		for (uint32_t lid = 0; lid < itemsPerLine; ++lid) {
			const int last = lid == itemsPerLine - 1;
			const uint32_t numPixels = (last && (0 != width % 64)) ? width % 64 : 64;
			const uint32_t numLoops = numPixels / 8, remain = numPixels % 8;
			size_t inOff = 8 * (size_t)lid + (size_t)pitchReads * gid;
			size_t outOff = (size_t)width * gid + (size_t)lid * 64;
			for (uint32_t i = 0; i <= numLoops; ++i) {
				const uint32_t n = i < numLoops ? 8 : remain; /* the tail block converts `remain` pixels the same way (:90-123) */
				for (uint32_t p = 0; p < n; ++p) {
					const uint32_t yuv[3] = {ld_s(inY, inOff * 8 + p, bits), ld_s(inU, inOff * 4 + p / 2, bits), ld_s(inV, inOff * 4 + p / 2, bits)};
					read_px(yuv, 1.0f, cm, lut, gamut, output + (outOff + p) * 4);
				}
				inOff++;
				outOff += 8;
			}
		}
	}
}

/* write kernel: yuv422p10.ts:126-219 */
void orc_yuv422p_write(int bits, const float *input, uint8_t *outY, uint8_t *outU, uint8_t *outV, uint32_t width, uint32_t height,
                       uint32_t interlace, const float *cm, const float *lut) {
	const uint32_t itemsPerLine = (orc_yuv422p_pitch(width) + 63) / 64;
	const uint32_t pitchReads = (width + 7) / 8;
	const uint32_t groups = (0 == interlace) ? height : height / 2;
	PAR_FOR
	for (int64_t gid = 0; gid < (int64_t)groups; ++gid) {
		for (uint32_t lid = 0; lid < itemsPerLine; ++lid) {
			const int last = lid == itemsPerLine - 1;
			const uint32_t numPixels = (last && (0 != width % 64)) ? width % 64 : 64;
			const uint32_t numLoops = numPixels / 8, remain = numPixels % 8;
			const uint32_t line = (uint32_t)gid * ((0 == interlace) ? 1 : 2) + ((3 == interlace) ? 1 : 0);
			size_t inOff = (size_t)width * line + (size_t)lid * 64;
			size_t outOff = (size_t)pitchReads * line + (size_t)lid * 8;
			for (uint32_t i = 0; i < numLoops; ++i) {
				uint32_t yuv[8][3];
				for (uint32_t p = 0; p < 8; ++p) {
					const float *l = input + (inOff + p) * 4;
					const float rgba[4] = {lut[sat_rte(l[0] * 65535.0f, 65535.0f)], lut[sat_rte(l[1] * 65535.0f, 65535.0f)],
					                       lut[sat_rte(l[2] * 65535.0f, 65535.0f)], 1.0f};
					yuv[p][0] = sat_rte(dot4(rgba, cm + 0), 65535.0f);
					yuv[p][1] = sat_rte(dot4(rgba, cm + 4), 65535.0f);
					yuv[p][2] = sat_rte(dot4(rgba, cm + 8), 65535.0f);
				}
				for (uint32_t p = 0; p < 8; ++p) st_s(outY, outOff * 8 + p, yuv[p][0], bits);
				for (uint32_t c = 0; c < 4; ++c) { /* chroma from even pixels only (:170-171) */
					st_s(outU, outOff * 4 + c, yuv[2 * c][1], bits);
					st_s(outV, outOff * 4 + c, yuv[2 * c][2], bits);
				}
				inOff += 8;
				outOff++;
			}
			if (remain > 0) { /* :180-218 */
				uint32_t y[8], u[4], v[4], yuv[6][3] = {{0}};
				for (int k = 0; k < 8; ++k) y[k] = bits == 8 ? 16 : 64;
				for (int k = 0; k < 4; ++k) u[k] = v[k] = bits == 8 ? 128 : 512;
				for (uint32_t p = 0; p < remain && p < 6; ++p) {
					const float *l = input + (inOff + p) * 4;
					const float rgba[4] = {lut[sat_rte(l[0] * 65535.0f, 65535.0f)], lut[sat_rte(l[1] * 65535.0f, 65535.0f)],
					                       lut[sat_rte(l[2] * 65535.0f, 65535.0f)], 1.0f};
					yuv[p][0] = sat_rte(roundf(dot4(rgba, cm + 0)), 65535.0f); /* round(): half away from zero */
					yuv[p][1] = sat_rte(roundf(dot4(rgba, cm + 4)), 65535.0f);
					yuv[p][2] = sat_rte(roundf(dot4(rgba, cm + 8)), 65535.0f);
				}
				y[0] = yuv[0][0]; y[1] = yuv[1][0]; u[0] = yuv[0][1]; v[0] = yuv[0][2];
				if (remain > 2) {
					y[2] = yuv[2][0]; y[3] = yuv[3][0]; u[1] = yuv[2][1]; v[1] = yuv[2][2];
					if (remain > 4) {
						y[4] = yuv[4][0]; y[5] = yuv[5][0];
						u[1] = yuv[4][1]; v[1] = yuv[4][2]; /* Q12: .s1 where .s2 is meant (yuv422p10.ts:210-211) */
					}
				}
				for (uint32_t p = 0; p < 8; ++p) st_s(outY, outOff * 8 + p, y[p], bits);
				for (uint32_t c = 0; c < 4; ++c) {
					st_s(outU, outOff * 4 + c, u[c], bits);
					st_s(outV, outOff * 4 + c, v[c], bits);
				}
			}
		}
	}
}

/* ------------------------------------------------------------------------- */
/* yuv420p / nv12: src/process/yuv420p.ts, src/process/nv12.ts                */
/* 8-bit 4:2:0.  The two files differ only in the chroma layout: two planes  */
/* of pitch/2 bytes per line pair (yuv420p) or one plane of interleaved U,V  */
/* pairs, pitch bytes per line pair (nv12); `nv12` selects.  A work-group is */
/* one PAIR of lines; the chroma of a pair is read for both lines and        */
/* written from the first line of the pair that the launch processes.        */
/* ------------------------------------------------------------------------- */
void orc_yuv420_plane_bytes(int nv12, uint32_t width, uint32_t height, uint32_t out[3]) {
	const uint32_t luma = orc_yuv422p_pitch(width) * height; /* yuv420p.ts:240-241,332-333; nv12.ts:322-323 */
	out[0] = luma;
	out[1] = nv12 ? luma / 2 : luma / 4;
	out[2] = nv12 ? 0 : luma / 4;
}

/* fillBuf: yuv420p.ts:243-280 / nv12.ts:246-281 (one buffer: Y plane | U plane | V plane, or Y plane | C plane) */
void orc_yuv420_fill(int nv12, uint8_t *buf, uint32_t width, uint32_t height) {
	const uint32_t lumaPitch = orc_yuv422p_pitch(width);
	const uint32_t chromaPitch = nv12 ? lumaPitch : lumaPitch / 2;
	size_t lOff = 0, uOff = (size_t)lumaPitch * height;
	size_t vOff = uOff + (size_t)chromaPitch * ((height + 1) / 2); /* yuv420p only */
	const size_t total = nv12 ? uOff + (size_t)chromaPitch * (height / 2) : uOff + 2 * (size_t)(lumaPitch * height / 4);
	memset(buf, 16, uOff);
	memset(buf + uOff, 128, total - uOff);
	uint32_t Y0 = 16, Y1 = 234;
	for (uint32_t y = 0; y < height; y += 2) {
		size_t xl = 0, xc = 0;
		for (uint32_t x = 0; x < width; x += 2) {
			buf[lOff + xl] = (uint8_t)Y0;
			buf[lOff + xl + 1] = (uint8_t)(Y0 + 1);
			buf[lumaPitch + lOff + xl] = (uint8_t)(Y1 + 1);
			buf[lumaPitch + lOff + xl + 1] = (uint8_t)Y1;
			xl += 2;
			if (nv12) {
				buf[uOff + xc] = 128;
				buf[uOff + xc + 1] = 128;
				xc += 2;
			} else {
				buf[uOff + xc] = 128;
				buf[vOff + xc] = 128;
				xc++;
			}
			Y0 = (234 == Y0) ? 16 : Y0 + 2;
			Y1 = (16 == Y1) ? 234 : Y1 - 2;
		}
		lOff += (size_t)lumaPitch * 2;
		uOff += chromaPitch;
		vOff += chromaPitch;
	}
}

/* chroma sample pair of pixel pair `c` (0..3) in 8-pixel block `blk` of line pair `gid` */
static inline void ld_chroma420(int nv12, const uint8_t *inU, const uint8_t *inV, size_t blk, uint32_t c, uint32_t *u, uint32_t *v) {
	if (nv12) { /* uchar8 c: (.s0,.s1) (.s2,.s3) ... = (U,V) pairs, nv12.ts:65-72 */
		*u = inU[blk * 8 + 2 * c];
		*v = inU[blk * 8 + 2 * c + 1];
	} else { /* uchar4 u, v: yuv420p.ts:67-78 */
		*u = inU[blk * 4 + c];
		*v = inV[blk * 4 + c];
	}
}

/* read kernel: yuv420p.ts:25-140 / nv12.ts:24-132.  One work-group per line pair, 64 pixels per work-item, blocks of 8. */
void orc_yuv420_read(int nv12, const uint8_t *inY, const uint8_t *inU, const uint8_t *inV, float *output, uint32_t width,
                     uint32_t height, const float *cm, const float *lut, const float *gamut) {
	const uint32_t itemsPerLine = (orc_yuv422p_pitch(width) + 63) / 64; /* yuv420p.ts:334 */
	const uint32_t pitchReads = (width + 7) / 8;
	PAR_FOR
	for (int64_t gid = 0; gid < (int64_t)(height / 2); ++gid) { /* globalWorkItems = items * height / 2 (:336) */
		for (uint32_t lid = 0; lid < itemsPerLine; ++lid) {
			const int last = lid == itemsPerLine - 1;
			const uint32_t numPixels = (last && (0 != width % 64)) ? width % 64 : 64;
			const uint32_t numLoops = numPixels / 8, remain = numPixels % 8;
			size_t inOffY[2], outOff[2];
			inOffY[0] = 8 * (size_t)lid + (size_t)pitchReads * gid * 2;
			inOffY[1] = inOffY[0] + pitchReads;
			size_t inOffUV = 8 * (size_t)lid + (size_t)pitchReads * gid;
			outOff[0] = 64 * (size_t)lid + (size_t)width * gid * 2;
			outOff[1] = outOff[0] + width;
			for (uint32_t i = 0; i <= numLoops; ++i) {
				const uint32_t n = i < numLoops ? 8 : remain; /* the tail block converts `remain` pixels the same way */
				if (n == 0) break;
				for (uint32_t l = 0; l < 2; ++l) {
					for (uint32_t p = 0; p < n; ++p) {
						uint32_t yuv[3];
						yuv[0] = inY[inOffY[l] * 8 + p];
						ld_chroma420(nv12, inU, inV, inOffUV, p / 2, &yuv[1], &yuv[2]);
						read_px(yuv, 1.0f, cm, lut, gamut, output + (outOff[l] + p) * 4);
					}
					inOffY[l]++;
					outOff[l] += 8;
				}
				inOffUV++;
			}
		}
	}
}

/* write kernel: yuv420p.ts:142-238 / nv12.ts:134-240.  Always height/2 work-groups (yuv420p.ts:358); a field launch
 * (interlace 1 / 3) writes one luma line per group and the group's chroma from that line, so after both fields the
 * chroma plane holds the bottom field's values. */
void orc_yuv420_write(int nv12, const float *input, uint8_t *outY, uint8_t *outU, uint8_t *outV, uint32_t width, uint32_t height,
                      uint32_t interlace, const float *cm, const float *lut) {
	const uint32_t itemsPerLine = (orc_yuv422p_pitch(width) + 63) / 64;
	const uint32_t pitchReads = (width + 7) / 8;
	PAR_FOR
	for (int64_t gid = 0; gid < (int64_t)(height / 2); ++gid) {
		for (uint32_t lid = 0; lid < itemsPerLine; ++lid) {
			const int last = lid == itemsPerLine - 1;
			const uint32_t numPixels = (last && (0 != width % 64)) ? width % 64 : 64;
			const uint32_t numLoops = numPixels / 8, remain = numPixels % 8;
			const uint32_t line = (uint32_t)gid * 2 + ((3 == interlace) ? 1 : 0);
			const uint32_t numLines = (0 == interlace) ? 2 : 1;
			size_t inOff[2], outOffY[2];
			inOff[0] = 64 * (size_t)lid + (size_t)width * line;
			inOff[1] = inOff[0] + width;
			outOffY[0] = 8 * (size_t)lid + (size_t)pitchReads * line;
			outOffY[1] = outOffY[0] + pitchReads;
			size_t outOffUV = 8 * (size_t)lid + (size_t)pitchReads * gid;
			for (uint32_t l = 0; l < numLines; ++l) {
				for (uint32_t i = 0; i < numLoops; ++i) {
					uint32_t yuv[8][3];
					for (uint32_t p = 0; p < 8; ++p) {
						const float *px = input + (inOff[l] + p) * 4;
						const float rgba[4] = {lut[sat_rte(px[0] * 65535.0f, 65535.0f)], lut[sat_rte(px[1] * 65535.0f, 65535.0f)],
						                       lut[sat_rte(px[2] * 65535.0f, 65535.0f)], 1.0f};
						/* uchar3 yuv[p]: the ushort conversion result wraps into a uchar (yuv420p.ts:186-188) */
						yuv[p][0] = (uint8_t)sat_rte(dot4(rgba, cm + 0), 65535.0f);
						yuv[p][1] = (uint8_t)sat_rte(dot4(rgba, cm + 4), 65535.0f);
						yuv[p][2] = (uint8_t)sat_rte(dot4(rgba, cm + 8), 65535.0f);
					}
					inOff[l] += 8;
					for (uint32_t p = 0; p < 8; ++p) outY[outOffY[l] * 8 + p] = (uint8_t)yuv[p][0];
					outOffY[l]++;
					if (l == 0) { /* chroma from even pixels of the first line only (:192-200) */
						for (uint32_t c = 0; c < 4; ++c) {
							if (nv12) {
								outU[outOffUV * 8 + 2 * c] = (uint8_t)yuv[2 * c][1];
								outU[outOffUV * 8 + 2 * c + 1] = (uint8_t)yuv[2 * c][2];
							} else {
								outU[outOffUV * 4 + c] = (uint8_t)yuv[2 * c][1];
								outV[outOffUV * 4 + c] = (uint8_t)yuv[2 * c][2];
							}
						}
						outOffUV++;
					}
				}
			}
			if (remain > 0) { /* yuv420p.ts:204-236 / nv12.ts:199-238 */
				for (uint32_t l = 0; l < numLines; ++l) {
					uint8_t y[8], u[4], v[4];
					uint32_t yuv[6][3] = {{0}};
					memset(y, 16, 8);
					memset(u, 128, 4);
					memset(v, 128, 4);
					for (uint32_t p = 0; p < remain && p < 6; ++p) {
						const float *px = input + (inOff[l] + p) * 4;
						const float rgba[4] = {lut[sat_rte(px[0] * 65535.0f, 65535.0f)], lut[sat_rte(px[1] * 65535.0f, 65535.0f)],
						                       lut[sat_rte(px[2] * 65535.0f, 65535.0f)], 1.0f};
						yuv[p][0] = (uint8_t)sat_rte(roundf(dot4(rgba, cm + 0)), 65535.0f); /* round(): half away from zero */
						yuv[p][1] = (uint8_t)sat_rte(roundf(dot4(rgba, cm + 4)), 65535.0f);
						yuv[p][2] = (uint8_t)sat_rte(roundf(dot4(rgba, cm + 8)), 65535.0f);
					}
					y[0] = (uint8_t)yuv[0][0]; y[1] = (uint8_t)yuv[1][0]; u[0] = (uint8_t)yuv[0][1]; v[0] = (uint8_t)yuv[0][2];
					if (remain > 2) {
						y[2] = (uint8_t)yuv[2][0]; y[3] = (uint8_t)yuv[3][0]; u[1] = (uint8_t)yuv[2][1]; v[1] = (uint8_t)yuv[2][2];
						if (remain > 4) {
							y[4] = (uint8_t)yuv[4][0]; y[5] = (uint8_t)yuv[5][0]; u[2] = (uint8_t)yuv[4][1]; v[2] = (uint8_t)yuv[4][2];
						}
					}
					memcpy(outY + outOffY[l] * 8, y, 8);
					if (l == 0) {
						if (nv12) {
							for (uint32_t c = 0; c < 4; ++c) {
								outU[outOffUV * 8 + 2 * c] = u[c];
								outU[outOffUV * 8 + 2 * c + 1] = v[c];
							}
						} else {
							memcpy(outU + outOffUV * 4, u, 4);
							memcpy(outV + outOffUV * 4, v, 4);
						}
					}
				}
			}
		}
	}
}

/* ------------------------------------------------------------------------- */
/* Lanczos Transform filter -- NOT in the reference (BASELINE.json config 5   */
/* asks for it; SURVEY 8f row 4: "parity unpinned", own definition).  This is */
/* the definition; phaneron_b200 (pb_runtime.cu lanczos tables + pb_device.cuh */
/* lanczos_sample) implements the same arithmetic and is tested bit-exact.    */
/*                                                                           */
/*  per axis (x shown), for an axis-aligned transformMatrix (m[1] = m[3] = 0): */
/*   p  = dot3(m_row, (x/w - 1/2, y/h - 1/2, 1)) + 1/2      transform.ts:54-57  */
/*   um = p * sw - 1/2 ; fu = floor(um) ; a = um - fu       (binary32, as the  */
/*                                          bilinear sampler places its taps) */
/*   fs = max(1, |m0| * sw / w)       source texels per output pixel (double) */
/*   R  = ceil(lobes * fs) ; taps k = -R+1 .. R             (2R taps)          */
/*   w_k = L((a - k) / fs), L(t) = sinc(t) sinc(t / lobes), |t| < lobes       */
/*   weights normalised to sum 1 (double), then rounded to binary32           */
/*  value = sum_j wy_j * (sum_i wx_i * T(i0 + i, j0 + j)), both sums in       */
/*  ascending tap order as fma chains starting from +0; texels outside the    */
/*  image are the CLK_ADDRESS_CLAMP border colour (0,0,0,0).                  */
/* ------------------------------------------------------------------------- */
#define ORC_LANCZOS_MAX_TAPS 64

static double lanczos_kernel(double t, int lobes) {
	if (t == 0.0) return 1.0;
	if (fabs(t) >= (double)lobes) return 0.0;
	const double pt = 3.14159265358979323846 * t;
	return (double)lobes * sin(pt) * sin(pt / (double)lobes) / (pt * pt);
}

/* taps of output coordinate o along one axis; returns the tap count (0: unsupported) */
static int lanczos_axis(int o, int out_n, int src_n, float p, float m_scale, int lobes, int *first, float *w) {
	(void)o;
	const float um = p * (float)src_n - 0.5f;
	const float fu = floorf(um);
	const float a = um - fu;
	const double step = fabs((double)m_scale) * (double)src_n / (double)out_n;
	const double fs = step > 1.0 ? step : 1.0;
	const int R = (int)ceil((double)lobes * fs);
	if (2 * R > ORC_LANCZOS_MAX_TAPS) return 0;
	float fuc = fu;
	if (!(fuc >= -1.0e6f)) fuc = -1.0e6f; /* NaN / far outside: only border texels */
	if (fuc > 1.0e6f) fuc = 1.0e6f;
	*first = (int)fuc - R + 1;
	double wd[ORC_LANCZOS_MAX_TAPS], sum = 0.0;
	for (int k = 0; k < 2 * R; ++k) {
		wd[k] = lanczos_kernel(((double)a - (double)(k - R + 1)) / fs, lobes);
		sum += wd[k];
	}
	for (int k = 0; k < 2 * R; ++k) w[k] = (float)(wd[k] / sum);
	return 2 * R;
}

int orc_transform_lanczos(const float *in, int sw, int sh, const float *mat9, int lobes, float *out, int w, int h) {
	if (mat9[1] != 0.0f || mat9[3] != 0.0f || lobes < 1 || lobes > 8) return -1; /* axis-aligned transforms only */
	const float mat0[3] = {mat9[0], mat9[1], mat9[2]};
	const float mat1[3] = {mat9[3], mat9[4], mat9[5]};
	int *i0 = (int *)malloc(sizeof(int) * (size_t)(w + h));
	int *nt = (int *)malloc(sizeof(int) * 2);
	float *wx = (float *)malloc(sizeof(float) * ORC_LANCZOS_MAX_TAPS * (size_t)(w + h));
	int *j0 = i0 + w;
	float *wy = wx + (size_t)ORC_LANCZOS_MAX_TAPS * w;
	int ok = 1;
	nt[0] = nt[1] = 0;
	for (int x = 0; x < w && ok; ++x) {
		const float inPos[3] = {(float)x / (float)w - 0.5f, (float)0 / (float)h - 0.5f, 1.0f};
		const float px = dot3(mat0, inPos) + 0.5f; /* the iy term is iy * 0 */
		nt[0] = lanczos_axis(x, w, sw, px, mat9[0], lobes, &i0[x], wx + (size_t)ORC_LANCZOS_MAX_TAPS * x);
		ok = nt[0] > 0;
	}
	for (int y = 0; y < h && ok; ++y) {
		const float inPos[3] = {(float)0 / (float)w - 0.5f, (float)y / (float)h - 0.5f, 1.0f};
		const float py = dot3(mat1, inPos) + 0.5f;
		nt[1] = lanczos_axis(y, h, sh, py, mat9[4], lobes, &j0[y], wy + (size_t)ORC_LANCZOS_MAX_TAPS * y);
		ok = nt[1] > 0;
	}
	if (ok) {
		const int tx = nt[0], ty = nt[1];
		PAR_FOR
		for (int64_t y = 0; y < h; ++y)
			for (int x = 0; x < w; ++x) {
				float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
				for (int j = 0; j < ty; ++j) {
					const int sy = j0[y] + j;
					float row[4] = {0.0f, 0.0f, 0.0f, 0.0f};
					for (int i = 0; i < tx; ++i) {
						const int sx = i0[x] + i;
						const float wgt = wx[(size_t)ORC_LANCZOS_MAX_TAPS * x + i];
						if (sx < 0 || sx >= sw || sy < 0 || sy >= sh) {
							for (int c = 0; c < 4; ++c) row[c] = fmaf(wgt, 0.0f, row[c]);
						} else {
							const float *t = in + ((size_t)sy * sw + sx) * 4;
							for (int c = 0; c < 4; ++c) row[c] = fmaf(wgt, t[c], row[c]);
						}
					}
					const float wgt = wy[(size_t)ORC_LANCZOS_MAX_TAPS * y + j];
					for (int c = 0; c < 4; ++c) acc[c] = fmaf(wgt, row[c], acc[c]);
				}
				memcpy(out + ((size_t)y * w + x) * 4, acc, sizeof acc);
			}
	}
	free(i0);
	free(nt);
	free(wx);
	return ok ? 0 : -1;
}
