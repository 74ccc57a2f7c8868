// pb_lut_cache.cu -- gamma tables of a context: one device copy per distinct content and, where a transfer-function
// model describes the table to within a byte, the lossless one-byte form the march kernel keeps in shared memory
// (pb_lut.cuh).
#include "pb_internal.h"

namespace pbrt {

// ---- gamma tables: content-deduplicated, with a lossless one-byte form for the march kernel -------
// Every Loader/Saver uploads its own copy of a 65536-float table (loadSave.ts:69-77, 155-164); five
// sources mean five identical buffers.  The context keeps ONE device copy per distinct content
// (keyed by a device-computed hash) plus, when the table is one of the transfer functions of
// colourMaths.ts:42-128, the d8 form of pb_lut.cuh.  Fitting is a tiny kernel + one blocking
// read-back, paid once per uploaded table.
struct TransferSet {
	double alpha, beta, gamma, delta;
};
constexpr TransferSet kTransferSets[] = {
	{1.099, 0.018, 0.45, 4.5},                   // 601 / 709 / 2020
	{1.055, 0.0031308, 1.0 / 2.4, 12.92},        // sRGB
};
constexpr int kLutCands = 6;   // per transfer set: gamma->linear as a polynomial (MUFU-free), gamma->linear and linear->gamma as MUFU models

// Coefficients of the MUFU-free model (pb_desc.h LutParams::affine == 2): the power segment of gamma2linearLUT,
// ((i / 65535 + alpha - 1) / alpha) ^ (1 / gamma) for i in [J, 65535], as a degree-7 polynomial in x = i * p + q in [-1, 1].
// Least squares on 1024 Chebyshev nodes in the Chebyshev basis, weighted by 1 / f (relative error, i.e. roughly ulps), then
// converted to monomials.  Any deterministic coefficients would do: the byte table holds the distance to the exact table
// value and lut_fit_kernel verifies that it fits a byte for all 65536 entries (else the MUFU model of the same curve is taken).
void fit_power_poly(double alpha, double gamma, int J, pb::LutParams *g) {
	constexpr int N = pb::kLutPolyDeg + 1, K = 1024;
	const double pp = 2.0 / (65535.0 - J), qq = -1.0 - pp * J;
	long double A[N][N + 1] = {};
	for (int k = 0; k < K; ++k) {
		const double t = std::cos(M_PI * (k + 0.5) / K);
		const double i = (t - qq) / pp;
		const double f = std::pow((i / 65535.0 + alpha - 1.0) / alpha, 1.0 / gamma);
		double T[N];
		T[0] = 1.0;
		T[1] = t;
		for (int n = 2; n < N; ++n) T[n] = 2.0 * t * T[n - 1] - T[n - 2];
		for (int r = 0; r < N; ++r) {
			for (int cidx = 0; cidx < N; ++cidx) A[r][cidx] += (long double)(T[r] / f) * (T[cidx] / f);
			A[r][N] += (long double)(T[r] / f);
		}
	}
	for (int col = 0; col < N; ++col) {   // Gauss-Jordan with partial pivoting
		int piv = col;
		for (int r = col + 1; r < N; ++r)
			if (fabsl(A[r][col]) > fabsl(A[piv][col])) piv = r;
		for (int k = 0; k <= N; ++k) std::swap(A[col][k], A[piv][k]);
		for (int r = 0; r < N; ++r) {
			if (r == col) continue;
			const long double m = A[r][col] / A[col][col];
			for (int k = col; k <= N; ++k) A[r][k] -= m * A[col][k];
		}
	}
	long double mono[N] = {}, Tm2[N] = {1}, Tm1[N] = {0, 1};   // Chebyshev -> monomial: T_n = 2 t T_{n-1} - T_{n-2}
	for (int n = 0; n < N; ++n) {
		long double Tn[N] = {};
		if (n == 0) Tn[0] = 1;
		else if (n == 1) Tn[1] = 1;
		else {
			for (int k = 0; k + 1 < N; ++k) Tn[k + 1] += 2 * Tm1[k];
			for (int k = 0; k < N; ++k) Tn[k] -= Tm2[k];
			for (int k = 0; k < N; ++k) { Tm2[k] = Tm1[k]; Tm1[k] = Tn[k]; }
		}
		const long double cn = A[n][N] / A[n][n];
		for (int k = 0; k < N; ++k) mono[k] += cn * Tn[k];
	}
	g->p = (float)pp;
	g->q = (float)qq;
	for (int k = 0; k < N; ++k) g->c[k] = (float)mono[k];
	g->affine = 2;
}

void lut_candidates(pb::LutParams *out) {
	int n = 0;
	// The MUFU-free polynomial model is exact (the whole GPU suite passes with it) but costs one more issue slot per lookup than
	// MUFU.LG2 + MUFU.EX2, and the kernels are issue-bound, not XU-bound: 170.0 vs 164.5 us on the 2160p bench scene
	// (profiles/r02_kbench_poly_ab.txt).  Opt-in for A/B runs.
	const bool no_poly = getenv("PB_LUT_POLY") == nullptr;
	for (const TransferSet &t : kTransferSets) {
		pb::LutParams g{};   // gamma2linearLUT (colourMaths.ts:130-149)
		g.p = (float)(1.0 / (65535.0 * t.alpha));
		g.q = (float)((t.alpha - 1.0) / t.alpha);
		g.G = (float)(1.0 / t.gamma);
		g.s = 1.0f;
		g.o = 0.0f;
		g.kt = (float)(1.0 / (65535.0 * t.delta));
		int J = 0;
		while (J < 65536 && J / 65535.0 < t.beta * t.delta) ++J;
		g.cJ = (float)(1 - J);
		g.affine = 0;
		pb::LutParams gp = g;   // the same curve, MUFU-free: preferred when it fits (listed first)
		fit_power_poly(t.alpha, t.gamma, J, &gp);
		if (no_poly) gp = g;
		out[n++] = gp;
		out[n++] = g;
		pb::LutParams l{};   // linear2gammaLUT (colourMaths.ts:151-169)
		l.p = (float)(1.0 / 65535.0);
		l.q = 0.0f;
		l.G = (float)t.gamma;
		l.s = (float)t.alpha;
		l.o = (float)(-(t.alpha - 1.0));
		l.kt = (float)(t.delta / 65535.0);
		J = 0;
		while (J < 65536 && J / 65535.0 < t.beta) ++J;
		l.cJ = (float)(1 - J);
		l.affine = 1;
		out[n++] = l;
	}
}

struct FitResultHost {   // mirrors pb::LutFitResult
	int dmin, dmax;
	unsigned long long hash;
	int not_unit, pad;
};

int lut_table_of(pb_ctx *c, pb_buf *lut, int *table_out) {
	for (const auto &f : c->lut_fits)
		if (lut->version != 0 && f.version == lut->version) {
			*table_out = f.table;
			return PB_OK;
		}
	cudaStream_t s = c->q[PB_QUEUE_PROCESS];
	if (!c->lut_scratch) {
		pb::LutParams cands[kLutCands];
		lut_candidates(cands);
		memcpy(c->lut_cands, cands, sizeof cands);
		CU(cudaMalloc(&c->lut_cands_dev, sizeof cands));
		CU(cudaMemcpyAsync(c->lut_cands_dev, cands, sizeof cands, cudaMemcpyHostToDevice, s));
		CU(cudaMalloc(&c->lut_res_dev, kLutCands * sizeof(FitResultHost)));
		CU(cudaMalloc(&c->lut_scratch, (size_t)kLutCands * 65536));
	}
	FitResultHost res[kLutCands];
	for (auto &r : res) r = FitResultHost{INT32_MAX, INT32_MIN, 0ull, 0, 0};
	CU(cudaMemcpyAsync(c->lut_res_dev, res, sizeof res, cudaMemcpyHostToDevice, s));
	cudaError_t e = pb::launch_lut_fit(s, (const float *)lut->dev, (const pb::LutParams *)c->lut_cands_dev, kLutCands, (uint8_t *)c->lut_scratch, c->lut_res_dev);
	if (e != cudaSuccess) return fail(PB_ERR_CUDA, "lut fit launch: %s", cudaGetErrorString(e));
	CU(cudaMemcpyAsync(res, c->lut_res_dev, sizeof res, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	int table = -1;
	for (size_t i = 0; i < c->lut_tables.size() && table < 0; ++i) {
		if (c->lut_tables[i].hash != res[0].hash) continue;
		// same 64-bit hash: confirm the contents before sharing the table (a collision would mean wrong colours, silently)
		std::vector<float> a(65536), b(65536);
		CU(cudaMemcpyAsync(a.data(), c->lut_tables[i].raw, 65536 * sizeof(float), cudaMemcpyDeviceToHost, s));
		CU(cudaMemcpyAsync(b.data(), lut->dev, 65536 * sizeof(float), cudaMemcpyDeviceToHost, s));
		CU(cudaStreamSynchronize(s));
		if (0 == memcmp(a.data(), b.data(), 65536 * sizeof(float))) table = (int)i;
	}
	if (table < 0) {
		pb_ctx::LutTable t;
		t.hash = res[0].hash;
		t.unit_range = res[0].not_unit == 0;
		CU(cudaMalloc(&t.raw, 65536 * sizeof(float)));
		CU(cudaMemcpyAsync(t.raw, lut->dev, 65536 * sizeof(float), cudaMemcpyDeviceToDevice, s));
		for (int k = 0; k < kLutCands && !t.d8; ++k)
			if (res[k].dmin >= -128 && res[k].dmax <= 127) {
				CU(cudaMalloc(&t.d8, 65536));
				CU(cudaMemcpyAsync(t.d8, (const char *)c->lut_scratch + (size_t)k * 65536, 65536, cudaMemcpyDeviceToDevice, s));
				t.lp = c->lut_cands[k];
				t.model = k;
				t.dmin = res[k].dmin;
				t.dmax = res[k].dmax;
			}
		CU(cudaStreamSynchronize(s));
		c->lut_tables.push_back(t);
		table = (int)c->lut_tables.size() - 1;
	}
	if (c->lut_fits.size() >= 4096) c->lut_fits.erase(c->lut_fits.begin(), c->lut_fits.begin() + 2048);
	if (lut->version != 0) c->lut_fits.push_back({lut->version, table});
	*table_out = table;
	return PB_OK;
}

int lut_table_by_raw(pb_ctx *c, const float *raw) {
	for (size_t i = 0; i < c->lut_tables.size(); ++i)
		if (c->lut_tables[i].raw == raw) return (int)i;
	return -1;
}

}  // namespace pbrt
