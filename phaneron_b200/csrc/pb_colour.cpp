// pb_colour.cpp -- host-side colour science of the product: the same tables and matrices
// phaneron's src/process/colourMaths.ts hands to its kernels, reproduced bit for bit so
// that hosts without the TypeScript (Python, C++) feed the CUDA kernels identical constants.
//
// The TypeScript works on Float32Array rows and accumulates in JS doubles
// (colourMaths.ts:171-178), so "F32" below marks every point where a value is rounded
// to binary32; everything between two such points is double arithmetic.
#include <array>
#include <cmath>
#include <cstring>
#include <string>

#include "../../include/phaneron_b200.h"

namespace {

struct Primaries {
	const char *name;
	double kR, kB;
	double rx, ry, gx, gy, bx, by, wx, wy;
	double alpha, beta, gamma, delta;
};

// colourMaths.ts:42-128
constexpr Primaries kSpecs[] = {
	{"601-625", 0.299, 0.114, 0.64, 0.33, 0.29, 0.6, 0.15, 0.06, 0.3127, 0.329, 1.099, 0.018, 0.45, 4.5},
	{"601_525", 0.299, 0.114, 0.63, 0.34, 0.31, 0.595, 0.155, 0.07, 0.3127, 0.329, 1.099, 0.018, 0.45, 4.5},
	{"709", 0.2126, 0.0722, 0.64, 0.33, 0.3, 0.6, 0.15, 0.06, 0.3127, 0.329, 1.099, 0.018, 0.45, 4.5},
	{"2020", 0.2627, 0.0593, 0.708, 0.292, 0.17, 0.797, 0.131, 0.046, 0.3127, 0.329, 1.099, 0.018, 0.45, 4.5},
	{"sRGB", 0.0, 0.0, 0.64, 0.33, 0.3, 0.6, 0.15, 0.06, 0.3127, 0.329, 1.055, 0.0031308, 1.0 / 2.4, 12.92},
};

const Primaries &spec_of(const char *name, bool *known) {
	for (const auto &s : kSpecs)
		if (name && std::strcmp(name, s.name) == 0) {
			*known = true;
			return s;
		}
	*known = false;   // "Unrecognised colourspace ... defaulting to BT.709"
	return kSpecs[2];
}

inline float F32(double v) { return static_cast<float>(v); }

// R x C matrix of binary32 values
template <int R, int C>
struct M {
	std::array<std::array<float, C>, R> v{};
	float &operator()(int r, int c) { return v[r][c]; }
	float operator()(int r, int c) const { return v[r][c]; }
};

template <int R, int K, int C>
M<R, C> product(const M<R, K> &a, const M<K, C> &b) {
	M<R, C> out;
	for (int r = 0; r < R; ++r)
		for (int c = 0; c < C; ++c) {
			double acc = 0.0;
			for (int k = 0; k < K; ++k) acc += static_cast<double>(a(r, k)) * static_cast<double>(b(k, c));
			out(r, c) = F32(acc);
		}
	return out;
}

template <int R, int C>
M<R, C> scaled(const M<R, C> &a, double s) {
	M<R, C> out;
	for (int r = 0; r < R; ++r)
		for (int c = 0; c < C; ++c) out(r, c) = F32(static_cast<double>(a(r, c)) * s);
	return out;
}

// colourMaths.ts:199-238: minors -> cofactors -> adjugate -> * 1/det, each stage stored as f32
M<3, 3> inverse(const M<3, 3> &a) {
	auto others = [](int i, int out[2]) {
		if (i == 1) { out[0] = 0; out[1] = 2; }
		else { out[0] = (i + 1) % 3; out[1] = (i + 2) % 3; }
	};
	M<3, 3> minors;
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) {
			int ys[2], xs[2];
			others(i, ys);
			others(j, xs);
			const double p = static_cast<double>(a(ys[0], xs[0])) * a(ys[1], xs[1]);
			const double q = static_cast<double>(a(ys[0], xs[1])) * a(ys[1], xs[0]);
			minors(i, j) = F32(p - q);
		}
	M<3, 3> adjugate;
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) {
			const float cof = F32(static_cast<double>(minors(i, j)) * (((i + j) % 2) ? -1.0 : 1.0));
			adjugate(j, i) = cof;
		}
	const double det = static_cast<double>(a(0, 0)) * minors(0, 0) - static_cast<double>(a(0, 1)) * minors(0, 1) +
	                   static_cast<double>(a(0, 2)) * minors(0, 2);
	return scaled(adjugate, 1.0 / det);
}

// colourMaths.ts:240-266
M<3, 3> rgb_to_xyz(const Primaries &p) {
	M<3, 1> white;
	white(0, 0) = F32(p.wx);
	white(1, 0) = F32(p.wy);
	white(2, 0) = F32(1.0 - p.wx - p.wy);
	const M<3, 1> W = scaled(white, 1.0 / static_cast<double>(white(1, 0)));
	M<3, 3> xyz;
	xyz(0, 0) = F32(p.rx); xyz(0, 1) = F32(p.gx); xyz(0, 2) = F32(p.bx);
	xyz(1, 0) = F32(p.ry); xyz(1, 1) = F32(p.gy); xyz(1, 2) = F32(p.by);
	xyz(2, 0) = F32(1.0 - p.rx - p.ry); xyz(2, 1) = F32(1.0 - p.gx - p.gy); xyz(2, 2) = F32(1.0 - p.bx - p.by);
	const M<3, 1> factors = product(inverse(xyz), W);
	M<3, 3> diag;
	for (int i = 0; i < 3; ++i) diag(i, i) = factors(i, 0);
	return product(xyz, diag);
}

template <int R, int C>
void flatten(const M<R, C> &m, float *out) {
	for (int r = 0; r < R; ++r)
		for (int c = 0; c < C; ++c) out[r * C + c] = m(r, c);
}

M<3, 3> identity3() {
	M<3, 3> m;
	m(0, 0) = m(1, 1) = m(2, 2) = 1.0f;
	return m;
}

}  // namespace

extern "C" {

int pb_gamma2linear_lut(const char *colspec, float *out) {
	bool known;
	const Primaries &p = spec_of(colspec, &known);
	const double knee = p.beta * p.delta, inv_gamma = 1 / p.gamma;
	for (int i = 0; i < 65536; ++i) {
		const double fi = i / 65535.0;
		out[i] = (fi < knee) ? F32(fi / p.delta) : F32(std::pow((fi + (p.alpha - 1)) / p.alpha, inv_gamma));
	}
	return known ? 1 : 0;
}

int pb_linear2gamma_lut(const char *colspec, float *out) {
	bool known;
	const Primaries &p = spec_of(colspec, &known);
	for (int i = 0; i < 65536; ++i) {
		const double fi = i / 65535.0;
		out[i] = (fi < p.beta) ? F32(fi * p.delta) : F32(p.alpha * std::pow(fi, p.gamma) - (p.alpha - 1));
	}
	return known ? 1 : 0;
}

int pb_ycbcr2rgb_matrix(const char *colspec, int num_bits, int luma_black, int luma_white, int chr_range, float *out12) {
	bool known;
	const Primaries &p = spec_of(colspec, &known);
	const double chr_null = static_cast<double>(128 << (num_bits - 8));
	const double luma_range = luma_white - luma_black;
	const double kG = 1.0 - p.kR - p.kB;
	M<3, 3> colour;
	colour(0, 0) = 1.0f; colour(0, 2) = F32(1.0 - p.kR);
	colour(1, 0) = 1.0f; colour(1, 1) = F32((-(1.0 - p.kB) * p.kB) / kG); colour(1, 2) = F32((-(1.0 - p.kR) * p.kR) / kG);
	colour(2, 0) = 1.0f; colour(2, 1) = F32(1.0 - p.kB);
	M<3, 4> range;
	range(0, 0) = F32(1.0 / luma_range); range(0, 3) = F32(-luma_black / luma_range);
	range(1, 1) = F32((1.0 / chr_range) * 2); range(1, 3) = F32(-(chr_null / chr_range) * 2);
	range(2, 2) = F32((1.0 / chr_range) * 2); range(2, 3) = F32(-(chr_null / chr_range) * 2);
	flatten(product(colour, range), out12);
	return known ? 1 : 0;
}

int pb_rgb2ycbcr_matrix(const char *colspec, int num_bits, int luma_black, int luma_white, int chr_range, float *out12) {
	bool known;
	const Primaries &p = spec_of(colspec, &known);
	const double chr_null = static_cast<double>(128 << (num_bits - 8));
	const double luma_range = luma_white - luma_black;
	const double kG = 1.0 - p.kR - p.kB;
	M<3, 3> range;
	range(0, 0) = F32(luma_range);
	range(1, 1) = F32(chr_range / 2.0);
	range(2, 2) = F32(chr_range / 2.0);
	M<3, 4> colour;
	colour(0, 0) = F32(p.kR); colour(0, 1) = F32(kG); colour(0, 2) = F32(p.kB); colour(0, 3) = F32(luma_black / luma_range);
	colour(1, 0) = F32(-p.kR / (1.0 - p.kB)); colour(1, 1) = F32(-kG / (1.0 - p.kB));
	colour(1, 2) = F32((1.0 - p.kB) / (1.0 - p.kB)); colour(1, 3) = F32((chr_null / chr_range) * 2.0);
	colour(2, 0) = F32((1.0 - p.kR) / (1.0 - p.kR)); colour(2, 1) = F32(-kG / (1.0 - p.kR));
	colour(2, 2) = F32(-p.kB / (1.0 - p.kR)); colour(2, 3) = F32((chr_null / chr_range) * 2.0);
	flatten(product(range, colour), out12);
	return known ? 1 : 0;
}

int pb_rgb2rgb_matrix(const char *src, const char *dst, float *out9) {
	bool ks, kd;
	const Primaries &ps = spec_of(src, &ks);
	const Primaries &pd = spec_of(dst, &kd);
	flatten(product(inverse(rgb_to_xyz(pd)), rgb_to_xyz(ps)), out9);
	return (ks && kd) ? 1 : 0;
}

// transform.ts:119-171
int pb_transform_matrix(int width, int height, int flip_h, int flip_v, double anchor_x, double anchor_y, double scale_x,
                        double scale_y, double offset_x, double offset_y, double rotate_turns, float *out9) {
	if (width <= 0 || height <= 0 || !out9) return PB_ERR_ARG;
	const double aspect = static_cast<double>(width) / height;
	auto or_one = [](double v) { return (v == 0.0 || v != v) ? 1.0 : v; };   // `(x as number) || 1.0`
	const double sx = or_one(scale_x) * (flip_h ? -1.0 : 1.0), sy = or_one(scale_y) * (flip_v ? -1.0 : 1.0);
	const double angle = rotate_turns * 2 * 3.141592653589793;
	M<3, 3> anchor_in = identity3(), scale = identity3(), rot = identity3(), shift = identity3(), anchor_out = identity3(),
	        project = identity3();
	anchor_in(0, 2) = F32(anchor_x); anchor_in(1, 2) = F32(anchor_y);
	scale(0, 0) = F32(1.0 / (sx * aspect)); scale(1, 1) = F32(1.0 / sy);
	rot(0, 0) = F32(std::cos(angle)); rot(0, 1) = F32(-std::sin(angle));
	rot(1, 0) = F32(std::sin(angle)); rot(1, 1) = F32(std::cos(angle));
	shift(0, 2) = F32(offset_x * aspect); shift(1, 2) = F32(offset_y);
	anchor_out(0, 2) = F32(-anchor_x * aspect); anchor_out(1, 2) = F32(-anchor_y);
	project(0, 0) = F32(aspect);
	const M<3, 3> m = product(product(product(product(product(anchor_in, scale), rot), shift), anchor_out), project);
	flatten(m, out9);
	return PB_OK;
}

}  // extern "C"
