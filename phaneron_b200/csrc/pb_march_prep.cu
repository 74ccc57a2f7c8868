// pb_march_prep.cu -- host side of the march kernel (pb_march.cu): exact sampling tables from the reference's float
// formula (transform.ts:54-57 + OpenCL 1.2 section 8.2), occlusion analysis, Lanczos tap tables, kernel-variant
// conditions and the launch of a flattened frame expression (march kernel when eligible, else the generic fused kernel).
#include "pb_internal.h"

namespace pbrt {

// ---- Lanczos tap tables (definition: oracle/oracle.c "Lanczos Transform filter"; same arithmetic, same libm) ----
constexpr int kLanczosMaxTaps = 64;

double lanczos_kernel(double t, int lobes) {
	if (t == 0.0) return 1.0;
	if (fabs(t) >= (double)lobes) return 0.0;
	const double pt = 3.14159265358979323846 * t;
	return (double)lobes * sin(pt) * sin(pt / (double)lobes) / (pt * pt);
}

// taps of one output coordinate along one axis from its sampling position p (normalised source coordinate)
int lanczos_axis(int out_n, int src_n, float p, float m_scale, int lobes, int *first, float *w) {
	const float um = p * (float)src_n - 0.5f;
	const float fu = floorf(um);
	const float a = um - fu;
	const double step = fabs((double)m_scale) * (double)src_n / (double)out_n;
	const double fs = step > 1.0 ? step : 1.0;
	const int R = (int)ceil((double)lobes * fs);
	if (2 * R > kLanczosMaxTaps) return 0;
	float fuc = fu;
	if (!(fuc >= -1.0e6f)) fuc = -1.0e6f;
	if (fuc > 1.0e6f) fuc = 1.0e6f;
	*first = (int)fuc - R + 1;
	double wd[kLanczosMaxTaps], sum = 0.0;
	for (int k = 0; k < 2 * R; ++k) {
		wd[k] = lanczos_kernel(((double)a - (double)(k - R + 1)) / fs, lobes);
		sum += wd[k];
	}
	for (int k = 0; k < 2 * R; ++k) w[k] = (float)(wd[k] / sum);
	return 2 * R;
}

// dot3(m_row, (ix, iy, 1)) + 1/2 with the cross term exactly zero, as pb_device.cuh transform_pos evaluates it
inline float lanczos_pos(int o, int out_n, float m_scale, float m_off, bool is_x) {
	const float ic = (float)o / (float)out_n - 0.5f;
	float t;
	if (is_x) {
		t = -0.5f * 0.0f;                 // iy * m1 (m1 == 0; any finite iy gives a zero)
		t = fmaf(ic, m_scale, t);
	} else {
		t = ic * m_scale;                 // iy * m4
		t = fmaf(-0.5f, 0.0f, t);         // ix * m3 (m3 == 0)
	}
	t = fmaf(1.0f, m_off, t);
	return t + 0.5f;
}

int attach_lanczos(pb_ctx *c, pb::Leaf *lf, int lobes) {
	if (lobes < 1 || lobes > 8) return fail(PB_ERR_ARG, "lanczos lobes must be 1..8, found %d", lobes);
	if (lf->m[1] != 0.0f || lf->m[3] != 0.0f) return fail(PB_ERR_ARG, "the lanczos filter needs an axis-aligned transform (no rotation)");
	const int W = lf->xf_w, H = lf->xf_h;
	const float key[4] = {lf->m[0], lf->m[2], lf->m[4], lf->m[5]};
	pb_ctx::LanczosTab *t = nullptr;
	for (auto &e : c->lanczos_tabs)
		if (e.sw == lf->w && e.sh == lf->h && e.W == W && e.H == H && e.lobes == lobes && 0 == memcmp(e.m, key, sizeof key)) t = &e;
	if (!t) {
		if (c->lanczos_tabs.size() >= 64) {   // parameters are animating: start over
			CU(cudaStreamSynchronize(c->q[PB_QUEUE_PROCESS]));
			for (auto &e : c->lanczos_tabs) {
				cudaFree(e.dev);
				cudaFree(e.dstrip);
			}
			c->lanczos_tabs.clear();
		}
		std::vector<int> i0((size_t)W + H);
		std::vector<float> wx((size_t)kLanczosMaxTaps * W), wy((size_t)kLanczosMaxTaps * H);
		int tx = 0, ty = 0;
		for (int x = 0; x < W; ++x) {
			tx = lanczos_axis(W, lf->w, lanczos_pos(x, W, lf->m[0], lf->m[2], true), lf->m[0], lobes, &i0[x], &wx[(size_t)kLanczosMaxTaps * x]);
			if (!tx) return fail(PB_ERR_ARG, "lanczos: more than %d taps per axis (scale too small)", kLanczosMaxTaps);
		}
		for (int y = 0; y < H; ++y) {
			ty = lanczos_axis(H, lf->h, lanczos_pos(y, H, lf->m[4], lf->m[5], false), lf->m[4], lobes, &i0[(size_t)W + y], &wy[(size_t)kLanczosMaxTaps * y]);
			if (!ty) return fail(PB_ERR_ARG, "lanczos: more than %d taps per axis (scale too small)", kLanczosMaxTaps);
		}
		// compact: [i0 (W) | j0 (H)] ints, then wx (W * tx), wy (H * ty) floats
		// ... and wx once more tap-major (W per tap) for the march kernel's coalesced reads
		std::vector<float> packed((size_t)2 * W * tx + (size_t)H * ty);
		for (int x = 0; x < W; ++x) memcpy(&packed[(size_t)x * tx], &wx[(size_t)kLanczosMaxTaps * x], sizeof(float) * tx);
		for (int y = 0; y < H; ++y) memcpy(&packed[(size_t)W * tx + (size_t)y * ty], &wy[(size_t)kLanczosMaxTaps * y], sizeof(float) * ty);
		float *wxt = &packed[(size_t)W * tx + (size_t)H * ty];
		for (int x = 0; x < W; ++x)
			for (int i = 0; i < tx; ++i) wxt[(size_t)i * W + x] = wx[(size_t)kLanczosMaxTaps * x + i];
		pb_ctx::LanczosTab e;
		memcpy(e.m, key, sizeof key);
		e.sw = lf->w; e.sh = lf->h; e.W = W; e.H = H; e.lobes = lobes; e.tx = tx; e.ty = ty;
		const size_t ib = i0.size() * sizeof(int), fb = packed.size() * sizeof(float);
		CU(cudaMalloc(&e.dev, ib + fb));
		CU(cudaMemcpyAsync(e.dev, i0.data(), ib, cudaMemcpyHostToDevice, c->q[PB_QUEUE_PROCESS]));
		CU(cudaMemcpyAsync((char *)e.dev + ib, packed.data(), fb, cudaMemcpyHostToDevice, c->q[PB_QUEUE_PROCESS]));
		CU(cudaStreamSynchronize(c->q[PB_QUEUE_PROCESS]));   // the staging vectors are locals; once per new transform
		e.i0 = (int *)e.dev;
		e.j0 = e.i0 + W;
		e.wx = (float *)((char *)e.dev + ib);
		e.wy = e.wx + (size_t)W * tx;
		e.wxt = e.wy + (size_t)H * ty;
		e.h_i0.assign(i0.begin(), i0.begin() + W);
		e.h_j0.assign(i0.begin() + W, i0.end());
		c->lanczos_tabs.push_back(e);
		t = &c->lanczos_tabs.back();
	}
	lf->lz_tx = t->tx; lf->lz_ty = t->ty;
	lf->lz_i0 = t->i0; lf->lz_j0 = t->j0;
	lf->lz_wx = t->wx; lf->lz_wy = t->wy;
	lf->lz_wxt = t->wxt;
	return PB_OK;
}

// march kernel: per-strip source footprints of a Lanczos leaf (the counterpart of get_tabs for the bilinear sampler)
int get_lanczos_tabs(pb_ctx *c, pb::Leaf *lf, int W, int H, int strip_groups, pb_ctx::LanczosTab **out) {
	pb_ctx::LanczosTab *t = nullptr;
	for (auto &e : c->lanczos_tabs)
		if (e.i0 == lf->lz_i0) t = &e;
	if (!t) return fail(PB_ERR_STATE, "lanczos tables of a leaf are gone");
	*out = t;
	if (t->strip_groups == strip_groups && t->dstrip) return PB_OK;
	if (t->dstrip) {
		CU(cudaStreamSynchronize(c->q[PB_QUEUE_PROCESS]));
		cudaFree(t->dstrip);
		t->dstrip = nullptr;
	}
	const int strip_px = strip_groups * 6, n_strips = (W + strip_px - 1) / strip_px;
	std::vector<int4> hstrip((size_t)n_strips);
	auto o = std::make_shared<pb_ctx::SampleTab::Opq>();
	o->id = c->next_tab_id++;
	o->strip_full.assign((size_t)n_strips, 0);
	o->row_full.assign((size_t)H, 0);
	o->row_j0 = t->h_j0;
	o->strip_ng.assign((size_t)n_strips, 0);
	o->rows_per_line = t->ty;
	o->src_h = lf->h;
	t->fits = 1;
	t->s0 = 0; t->s1 = -1; t->y0 = 0; t->y1 = -1;
	for (int sidx = 0; sidx < n_strips; ++sidx) {
		const int x0 = sidx * strip_px, x1 = std::min(x0 + strip_px, W) - 1;
		int lo = INT32_MAX, hi = INT32_MIN;
		for (int x = x0; x <= x1; ++x) {
			lo = std::min(lo, t->h_i0[x]);
			hi = std::max(hi, t->h_i0[x] + t->tx - 1);
		}
		int4 e = make_int4(0, 0, 0, 0);
		if (!(hi < 0 || lo >= lf->w)) {
			e.x = 1 | ((lo < 0 || hi >= lf->w) ? 2 : 0);
			lo = std::max(lo, 0);
			hi = std::min(hi, lf->w - 1);
			e.y = lo / 6;
			e.z = hi / 6 - e.y + 1;
			if (e.z > 2 * pb::kRowGroups) t->fits = 0;
			else if (e.z > pb::kRowGroups && t->fits) t->fits = 2;
			if (t->s1 < t->s0) t->s0 = sidx;
			t->s1 = sidx;
		}
		hstrip[sidx] = e;
		o->strip_ng[sidx] = e.z;
	}
	for (int y = 0; y < H; ++y)
		if (t->h_j0[y] + t->ty - 1 >= 0 && t->h_j0[y] < lf->h) {
			if (t->y1 < t->y0) t->y0 = y;
			t->y1 = y;
		}
	CU(cudaMalloc(&t->dstrip, hstrip.size() * sizeof(int4)));
	CU(cudaMemcpyAsync(t->dstrip, hstrip.data(), hstrip.size() * sizeof(int4), cudaMemcpyHostToDevice, c->q[PB_QUEUE_PROCESS]));
	CU(cudaStreamSynchronize(c->q[PB_QUEUE_PROCESS]));   // `hstrip` is a local; once per new transform
	t->strip_groups = strip_groups;
	t->opq = o;
	return PB_OK;
}

// ---- march kernel preparation -------------------------------------------------------------------
// Exact host evaluation of the sampling position of transform.ts:54-57 followed by the
// OpenCL 1.2 8.2 linear-filter prologue, for one axis of a separable (no rotation / shear)
// transform.  Same operations, same order, same rounding as pb_device.cuh transform_pos() +
// sample_linear_clamp(); the host compiler runs with -ffp-contract=off.
inline int2 axis_entry(int o, int out_n, int src_n, float m_scale, float m_other, float m_off, bool is_x, bool has_xf) {
	int2 e;
	if (!has_xf) {   // direct read of texel o
		e.x = o;
		e.y = 0;
		return e;
	}
	const float ic = (float)o / (float)out_n - 0.5f;
	// dot3(ix, iy, 1, m): t = ix*m[0]; t = fma(iy, m[1], t); t = fma(1, m[2], t), with the cross term exactly zero
	float t;
	if (is_x) {
		t = ic * m_scale;                 // ix * m0
		t = fmaf(0.0f, m_other, t);       // iy * 0 (m_other == 0 is an eligibility condition; iy is finite)
	} else {
		t = 0.0f * m_other;               // ix * 0
		t = fmaf(ic, m_scale, t);         // iy * m4
	}
	t = fmaf(1.0f, m_off, t);
	const float p = t + 0.5f;
	const float um = p * (float)src_n - 0.5f;
	const float fu = floorf(um);
	const float a = um - fu;
	int i0;
	if (!(fu >= -2.0f)) i0 = -2; else if (fu > (float)src_n) i0 = src_n; else i0 = (int)fu;
	e.x = i0;
	memcpy(&e.y, &a, 4);
	return e;
}

// Exact occlusion culling.  combine.ts:49-59 composites `fma(prev, 1 - l.a, l)`: where a layer's
// alpha is EXACTLY 1.0f, k = 0 and fma(prev, 0, l) == l for every finite prev, so nothing below that layer can
// reach the output and the march kernel need not evaluate it there.  The alpha of a v210 leaf seen through a
// Transform is the float sum of its in-image tap weights, w11 + (w01 + (w10 + w00)) with w00 = (1-a)(1-b) ...
// (pb_march.cu eval_leaf, same chain in pb_device.cuh and the oracle); whether that sum rounds to exactly 1
// depends on the fractional weights a (per column) and b (per line).  This routine evaluates the very same
// float chain for every distinct (a, b) pair of the leaf and keeps a separable set columns x lines on which
// all pairs give 1.0f; strips made of such columns and lines made of such rows are "full".
std::shared_ptr<const pb_ctx::SampleTab::Opq> leaf_opacity(int id, const pb::Leaf &lf, int W, int H, int strip_px, int n_strips,
                                                           const int2 *hcol, const int2 *hrow, const int4 *hstrip) {
	auto o = std::make_shared<pb_ctx::SampleTab::Opq>();
	o->id = id;
	o->rows_per_line = lf.has_xf ? 2 : 1;
	o->src_h = lf.h;
	o->strip_ng.resize(n_strips);
	for (int sidx = 0; sidx < n_strips; ++sidx) o->strip_ng[sidx] = (hstrip[sidx].x & 1) ? hstrip[sidx].z : 0;
	o->strip_full.assign(n_strips, 0);
	o->row_full.assign(H, 0);
	o->row_j0.resize(H);
	for (int y = 0; y < H; ++y) o->row_j0[y] = hrow[y].x;
	o->col_i0.resize(W);
	for (int x = 0; x < W; ++x) o->col_i0[x] = hcol[x].x;
	if (!lf.has_xf) {   // 1:1 read: alpha = 1 everywhere (the leaf has the output's dimensions)
		o->strip_full.assign(n_strips, 1);
		o->row_full.assign(H, 1);
		return o;
	}
	// candidates: all four taps inside the image
	std::vector<uint8_t> col_ok(W), row_ok(H);
	std::map<uint32_t, int> a_ids, b_ids;
	std::vector<int> col_a(W, -1), row_b(H, -1);
	for (int x = 0; x < W; ++x) {
		col_ok[x] = hcol[x].x >= 0 && hcol[x].x + 1 < lf.w;
		if (col_ok[x]) col_a[x] = a_ids.emplace((uint32_t)hcol[x].y, (int)a_ids.size()).first->second;
	}
	for (int y = 0; y < H; ++y) {
		row_ok[y] = hrow[y].x >= 0 && hrow[y].x + 1 < lf.h;
		if (row_ok[y]) row_b[y] = b_ids.emplace((uint32_t)hrow[y].y, (int)b_ids.size()).first->second;
	}
	const size_t na = a_ids.size(), nb = b_ids.size();
	if (na == 0 || nb == 0 || na * nb > (size_t)(1u << 21)) return o;   // nothing opaque / too many weight pairs to certify
	std::vector<float> av(na), bv(nb);
	for (auto &kv : a_ids) memcpy(&av[kv.second], &kv.first, 4);
	for (auto &kv : b_ids) memcpy(&bv[kv.second], &kv.first, 4);
	// cost of giving a value up: an a value takes its strips with it (for every line), a b value only its lines
	std::vector<int> a_strips(na, 0), b_rows(nb, 0);
	{
		std::vector<int> last(na, -1);
		for (int x = 0; x < W; ++x)
			if (col_a[x] >= 0 && last[col_a[x]] != x / strip_px) { last[col_a[x]] = x / strip_px; a_strips[col_a[x]]++; }
		for (int y = 0; y < H; ++y)
			if (row_b[y] >= 0) b_rows[row_b[y]]++;
	}
	std::vector<uint8_t> a_keep(na, 1), b_keep(nb, 1);
	for (size_t i = 0; i < na; ++i) {
		const float a = av[i], ra = 1.0f - a;
		for (size_t j = 0; j < nb; ++j) {
			if (!b_keep[j]) continue;
			const float b = bv[j], rb = 1.0f - b;
			const float w00 = ra * rb, w10 = a * rb, w01 = ra * b, w11 = a * b;
			const float alpha = w11 + (w01 + (w10 + w00));
			if (alpha == 1.0f) continue;
			if ((long long)a_strips[i] * H < (long long)b_rows[j] * n_strips) { a_keep[i] = 0; break; }
			b_keep[j] = 0;
		}
	}
	for (int y = 0; y < H; ++y) o->row_full[y] = row_b[y] >= 0 && b_keep[row_b[y]];
	for (int sidx = 0; sidx < n_strips; ++sidx) {
		const int x0 = sidx * strip_px, x1 = std::min(x0 + strip_px, W) - 1;
		bool full = true;
		for (int x = x0; x <= x1 && full; ++x) full = col_a[x] >= 0 && a_keep[col_a[x]];
		o->strip_full[sidx] = full;
	}
	return o;
}

// sampling tables of one leaf; *fits = 0 if some strip's source footprint exceeds a row buffer
int get_tabs(pb_ctx *c, const pb::Leaf &lf, int W, int H, int strip_groups, pb_ctx::SampleTab **out, int *fits) {
	for (auto &t : c->tabs)
		if (t.sw == lf.w && t.sh == lf.h && t.W == W && t.H == H && t.has_xf == lf.has_xf && t.strip_groups == strip_groups &&
		    (!lf.has_xf || 0 == memcmp(t.m, lf.m, sizeof t.m))) {
			*out = &t;
			*fits = t.fits;
			return PB_OK;
		}
	if (c->tabs.size() >= 256) {   // parameters are animating: start over (rare; tables are tiny)
		CU(cudaStreamSynchronize(c->q[PB_QUEUE_PROCESS]));
		for (auto &t : c->tabs) cudaFree(t.dev);
		c->tabs.clear();
	}
	pb_ctx::SampleTab t;
	memcpy(t.m, lf.m, sizeof t.m);
	t.sw = lf.w; t.sh = lf.h; t.W = W; t.H = H; t.has_xf = lf.has_xf; t.strip_groups = strip_groups;
	const int strip_px = strip_groups * 6;
	const int n_strips = (W + strip_px - 1) / strip_px;
	// one allocation: int4 strip[n_strips] | int2 col[W] | int2 row[H]  (the 16-byte entries first: W + H may be odd)
	const size_t bytes = ((size_t)W + H) * sizeof(int2) + (size_t)n_strips * sizeof(int4);
	std::vector<int4> host_store((bytes + sizeof(int4) - 1) / sizeof(int4));
	struct { char *p; char *data() const { return p; } } host{reinterpret_cast<char *>(host_store.data())};
	int4 *hstrip = reinterpret_cast<int4 *>(host.data());
	int2 *hcol = reinterpret_cast<int2 *>(hstrip + n_strips);
	int2 *hrow = hcol + W;
	for (int x = 0; x < W; ++x) hcol[x] = axis_entry(x, W, lf.w, lf.m[0], lf.m[1], lf.m[2], true, lf.has_xf != 0);
	for (int y = 0; y < H; ++y) hrow[y] = axis_entry(y, H, lf.h, lf.m[4], lf.m[3], lf.m[5], false, lf.has_xf != 0);
	t.fits = 1;
	for (int sidx = 0; sidx < n_strips; ++sidx) {
		const int x0 = sidx * strip_px, x1 = std::min(x0 + strip_px, W) - 1;
		int lo = INT32_MAX, hi = INT32_MIN;
		for (int x = x0; x <= x1; ++x) {   // not assumed monotone (flips, degenerate scales)
			lo = std::min(lo, hcol[x].x);
			hi = std::max(hi, hcol[x].x + (lf.has_xf ? 1 : 0));
		}
		int4 e = make_int4(0, 0, 0, 0);
		if (!(hi < 0 || lo >= lf.w)) {
			e.x = 1 | ((lo < 0 || hi >= lf.w) ? 2 : 0);
			lo = std::max(lo, 0);
			hi = std::min(hi, lf.w - 1);
			e.y = lo / 6;
			e.z = hi / 6 - e.y + 1;
			if (e.z > 2 * pb::kRowGroups) t.fits = 0;   // beyond even the big row buffers
			else if (e.z > pb::kRowGroups && t.fits) t.fits = 2;   // needs the big row buffers
		}
		hstrip[sidx] = e;
		if (e.x & 1) {
			if (t.s1 < t.s0) t.s0 = sidx;
			t.s1 = sidx;
		}
	}
	for (int y = 0; y < H; ++y) {
		const int j0 = hrow[y].x;
		const bool ok = (j0 >= 0 && j0 < lf.h) || (lf.has_xf && j0 + 1 >= 0 && j0 + 1 < lf.h);
		if (ok) {
			if (t.y1 < t.y0) t.y0 = y;
			t.y1 = y;
		}
	}
	t.opq = leaf_opacity(c->next_tab_id++, lf, W, H, strip_px, n_strips, hcol, hrow, hstrip);
	CU(cudaMalloc(&t.dev, bytes));
	CU(cudaMemcpyAsync(t.dev, host.data(), bytes, cudaMemcpyHostToDevice, c->q[PB_QUEUE_PROCESS]));
	CU(cudaStreamSynchronize(c->q[PB_QUEUE_PROCESS]));   // `host` is a local; once per new transform only
	t.dstrip = reinterpret_cast<int4 *>(t.dev);
	t.dcol = reinterpret_cast<int2 *>(t.dstrip + n_strips);
	t.drow = t.dcol + W;
	c->tabs.push_back(std::move(t));
	*out = &c->tabs.back();
	*fits = c->tabs.back().fits;
	return PB_OK;
}

// Decide whether the march kernel can evaluate this descriptor and, if so, attach the sampling
// tables and gamma-table slots.  Returns 1 = march, 0 = use the generic kernel, <0 = error.
int prepare_march(pb_ctx *c, pb::FusedDesc &d) {
	if (!c->allow_march) return 0;
	// sinks: v210, the planar YCbCr formats (the same 3 codes per pixel, stored by plane) and rgba8 / bgra8 (one word per pixel)
	// ... and the composite itself as an RGBA-f32 frame (deferred frames made real: ROUTE payloads, Yadif inputs, host reads)
	const bool rgba_f32_sink = d.sink == pb::SINK_RGBA_F32;
	const bool planar_sink = d.sink == pb::SINK_YUV422P10 || d.sink == pb::SINK_YUV422P8 || d.sink == pb::SINK_YUV420P || d.sink == pb::SINK_NV12 ||
	                         d.sink == pb::SINK_RGBA8 || d.sink == pb::SINK_BGRA8 || rgba_f32_sink;
	if (d.sink != pb::SINK_V210 && !planar_sink) return 0;
	if ((d.sink == pb::SINK_YUV420P || d.sink == pb::SINK_NV12) && (d.out_h & 1)) return 0;
	// Widths: the march kernel writes whole v210 groups.  With a v210 sink a ragged width (1280-wide 720p = 213 groups + 2 pixels,
	// 27 x 128-byte pitch) is split: the march kernel takes the whole groups, a second small launch of the generic kernel the
	// tail columns of every line (the partial group with its Q2 semantics, v210.ts:166-192, and the padding groups).  The other
	// sinks have their own 8-pixel tail quirks: whole multiples of 48 only.
	if (d.out_h < 1 || d.out_w < 6 || (d.out_w & 1)) return 0;
	if (d.sink != pb::SINK_V210 && d.out_w % 48 != 0) return 0;
	if (d.interlace != 0 && d.out_h < 2) return 0;
	bool any_xf = false, any_planar = planar_sink, any_rgba = false, any_f32 = false;
	const std::vector<uint32_t> *line_ops_host = nullptr;
	std::vector<int> line_ops_key;
	bool rc_ycc[pb::kMaxReadConsts] = {};   // read constants used by some YCbCr leaf (their tables go to shared memory)
	pb::Leaf *leaves[3 * pb::kMaxLayers];
	int n_leaves = 0;
	for (int l = 0; l < d.n_layers; ++l) {
		pb::Layer &ly = d.layers[l];
		pb::Leaf *ll[3] = {&ly.a, &ly.b, &ly.mask};
		const int nleaf = ly.kind == pb::LAYER_DIRECT ? 1 : (ly.kind == pb::LAYER_DISSOLVE ? 2 : 3);
		for (int q = 0; q < nleaf; ++q) {
			pb::Leaf &lf = *ll[q];
			// packed 4:2:2 / 4:2:0 YCbCr sources convert through the v210 group path; rgba8 / bgra8 (alpha) and RGBA-f32 leaves do not
			const bool ycc = lf.kind == pb::LEAF_V210 || lf.kind == pb::LEAF_YUV422P10 || lf.kind == pb::LEAF_YUV422P8 ||
			                 lf.kind == pb::LEAF_YUV420P || lf.kind == pb::LEAF_NV12;
			// graphics with alpha, and RGBA-f32 frames (Yadif outputs, materialised sub-expressions): pb_march.cu eval_leaf_rgba
			const bool rgba = lf.kind == pb::LEAF_RGBA8 || lf.kind == pb::LEAF_BGRA8 || lf.kind == pb::LEAF_RGBA_F32 || lf.kind == pb::LEAF_YADIF;
			if (lf.kind == pb::LEAF_RGBA_F32 || lf.kind == pb::LEAF_YADIF) any_f32 = true;
			if (!(ycc || rgba) || lf.w < 6 || (lf.lz_tx && !ycc)) return 0;
			// (a Lanczos leaf -- the extension of DESIGN.md 4.6 -- is evaluated separably where its footprints allow: the kernel then
			// sees only the vertical pass, LEAF_LANCZOS_V, which every variant evaluates; else inside the launch, by the general
			// variants: eval_leaf_lanczos.  Decided below, once the strip width is known.)
			lf.lz_sep = 0;
			if (lf.kind != pb::LEAF_V210 || lf.w % 6 != 0) any_planar = true;   // general load path (formats, partial last groups)
			if (rgba) any_rgba = true;
			else rc_ycc[lf.rc] = true;
			if (lf.has_xf) {
				for (float v : lf.m)
					if (!(v == v) || v > 1e30f || v < -1e30f) return 0;
				if (lf.m[1] != 0.0f || lf.m[3] != 0.0f) {   // rotation / shear: RGBA-f32 frames only (sampled per pixel, no row buffers, no tables)
					if (!(lf.kind == pb::LEAF_RGBA_F32 || lf.kind == pb::LEAF_YADIF) || lf.lz_tx) return 0;
					lf.has_xf = 2;
				}
				if (lf.xf_w != d.out_w || lf.xf_h != d.out_h) return 0;
				any_xf = true;
			} else if (lf.w != d.out_w || lf.h != d.out_h) {
				return 0;
			}
			leaves[n_leaves++] = &lf;
		}
	}
	if (const char *dbg = getenv("PB_DBG")) d.dbg = atoi(dbg);
	d.e_magic = 0x4B000000u;
	d.lds_koff = 0u - 0x4B000000u;
	d.march_w = d.out_w / 6 * 6;
	d.g_first = 0;
	d.strip_groups = any_xf ? pb::kStripGroupsXf : pb::kStripGroupsDirect;
	// Lanczos leaves: a strip's source footprint is its taps wider than the bilinear one, and a row of more than 32 groups
	// costs a second conversion pass with one or two lanes busy.  Narrow the strips until every Lanczos footprint fits one pass
	// (0.5x with 12 taps: 14 groups = 84 px -> at most 31 source groups); if nothing fits, stay wide and take the big rows.
	{
		bool any_lz = false;
		for (int i = 0; i < n_leaves; ++i) any_lz = any_lz || leaves[i]->lz_tx;
		if (any_lz) {
			int pick = 0;
			for (int sg = d.strip_groups; sg >= 8 && !pick; --sg) {
				if ((d.out_w / 6 + sg - 1) / sg > pb::kMaxStrips) break;
				bool ok = true;
				for (int i = 0; i < n_leaves && ok; ++i) {
					if (!leaves[i]->lz_tx) continue;
					pb_ctx::LanczosTab *lt;
					int r = get_lanczos_tabs(c, leaves[i], d.out_w, d.out_h, sg, &lt);
					if (r) return r;
					ok = lt->fits == 1;
				}
				if (ok) pick = sg;
			}
			if (pick) d.strip_groups = pick;
		}
	}
	d.n_strips = (d.out_w / 6 + d.strip_groups - 1) / d.strip_groups;
	if (d.n_strips > pb::kMaxStrips) return 0;
	std::shared_ptr<const pb_ctx::SampleTab::Opq> opq[3 * pb::kMaxLayers], tab_of[3 * pb::kMaxLayers];
	bool big_rows = false;
	for (int i = 0; i < n_leaves; ++i) {
		if (leaves[i]->lz_tx) {   // separable Lanczos taps: footprints from the tap tables
			pb_ctx::LanczosTab *lt;
			int r = get_lanczos_tabs(c, leaves[i], d.out_w, d.out_h, d.strip_groups, &lt);
			if (r) return r;
			if (!lt->fits) return 0;
			if (lt->fits == 2) big_rows = true;
			leaves[i]->lz_sep = (lt->fits == 1 && lt->s1 >= lt->s0 && lt->y1 >= lt->y0 && !getenv("PB_LANCZOS_ONE_PASS")) ? 1 : 0;
			if (!leaves[i]->lz_sep) any_planar = true;   // evaluated inside the launch: the general variants
			opq[i] = nullptr;
			tab_of[i] = lt->opq;
			leaves[i]->col_tab = nullptr;
			leaves[i]->row_tab = nullptr;
			leaves[i]->strip_tab = lt->dstrip;
			leaves[i]->s0 = lt->s0; leaves[i]->s1 = lt->s1; leaves[i]->y0 = lt->y0; leaves[i]->y1 = lt->y1;
			continue;
		}
		if (leaves[i]->has_xf == 2) {   // general affine position of an RGBA-f32 frame: only the bounding box of where it can be non-zero
			// transform.ts:54-57: (s, t) = M . (x/W - 1/2, y/H - 1/2, 1) + 1/2; a tap is inside the image only for s in
			// [-1/2sw, 1 + 1/2sw) (t alike).  Map the corners of that square, widened to two texels, back to output pixels.
			const pb::Leaf &lf = *leaves[i];
			const double det = (double)lf.m[0] * lf.m[4] - (double)lf.m[1] * lf.m[3];
			int x_lo = 0, x_hi = d.out_w - 1, y_lo = 0, y_hi = d.out_h - 1;
			if (std::fabs(det) > 1e-9) {
				double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
				for (int k = 0; k < 4; ++k) {
					const double sN = (k & 1) ? 1.0 + 2.0 / lf.w : -2.0 / lf.w, tN = (k & 2) ? 1.0 + 2.0 / lf.h : -2.0 / lf.h;
					const double bs = sN - 0.5 - lf.m[2], bt = tN - 0.5 - lf.m[5];
					const double ix = (bs * lf.m[4] - lf.m[1] * bt) / det, iy = (lf.m[0] * bt - lf.m[3] * bs) / det;
					const double x = (ix + 0.5) * d.out_w, y = (iy + 0.5) * d.out_h;
					xmin = std::min(xmin, x); xmax = std::max(xmax, x);
					ymin = std::min(ymin, y); ymax = std::max(ymax, y);
				}
				xmin = std::max(xmin - 3.0, -1.0); ymin = std::max(ymin - 3.0, -1.0);
				xmax = std::min(xmax + 3.0, (double)d.out_w); ymax = std::min(ymax + 3.0, (double)d.out_h);
				x_lo = std::max(0, (int)std::floor(xmin)); x_hi = std::min(d.out_w - 1, (int)std::ceil(xmax));
				y_lo = std::max(0, (int)std::floor(ymin)); y_hi = std::min(d.out_h - 1, (int)std::ceil(ymax));
			}
			opq[i] = nullptr;
			tab_of[i] = nullptr;
			leaves[i]->col_tab = nullptr;
			leaves[i]->row_tab = nullptr;
			leaves[i]->strip_tab = nullptr;
			if (x_hi < x_lo || y_hi < y_lo) { leaves[i]->s0 = 1; leaves[i]->s1 = 0; leaves[i]->y0 = 1; leaves[i]->y1 = 0; }   // nowhere
			else { leaves[i]->s0 = x_lo / (d.strip_groups * 6); leaves[i]->s1 = x_hi / (d.strip_groups * 6); leaves[i]->y0 = y_lo; leaves[i]->y1 = y_hi; }
			continue;
		}
		pb_ctx::SampleTab *t;
		int fits = 0;
		int r = get_tabs(c, *leaves[i], d.out_w, d.out_h, d.strip_groups, &t, &fits);
		if (r) return r;
		if (!fits) return 0;
		const bool leaf_f32 = leaves[i]->kind == pb::LEAF_RGBA_F32 || leaves[i]->kind == pb::LEAF_YADIF;   // taps straight from global memory: no row buffer
		const bool leaf_rgba = leaves[i]->kind == pb::LEAF_RGBA8 || leaves[i]->kind == pb::LEAF_BGRA8 || leaf_f32;
		if (leaf_rgba && !leaf_f32 && fits != 1) return 0;   // four planes: 32 source groups per row at most
		if (!leaf_f32 && (fits == 2 || leaf_rgba)) big_rows = true;
		opq[i] = leaf_rgba ? nullptr : t->opq;   // the alpha of an rgba8 / RGBA-f32 leaf is data: never certifiably opaque
		tab_of[i] = t->opq;
		leaves[i]->col_tab = t->dcol;
		leaves[i]->row_tab = t->drow;
		leaves[i]->strip_tab = t->dstrip;
		leaves[i]->s0 = t->s0; leaves[i]->s1 = t->s1; leaves[i]->y0 = t->y0; leaves[i]->y1 = t->y1;
	}
	// flatten the layer graph: evaluation order keeps at most {t, p} live (dissolve = b then a; wipe = mask, a, b)
	d.n_ops = 0;
	int layer_first_op[pb::kMaxLayers], layer_n_ops[pb::kMaxLayers];
	for (int l = 0; l < d.n_layers; ++l) {
		const pb::Layer &ly = d.layers[l];
		layer_first_op[l] = d.n_ops;
		auto push = [&](int which, int act) { d.ops[d.n_ops++] = pb::MarchOp{l, which, act, ly.mix}; };
		if (ly.kind == pb::LAYER_DIRECT) {
			push(0, pb::ACT_OVER);
		} else if (ly.kind == pb::LAYER_DISSOLVE) {
			push(1, pb::ACT_DIS_B);
			push(0, pb::ACT_DIS_A_OVER);
		} else {
			push(2, pb::ACT_WIPE_M);
			push(0, pb::ACT_WIPE_A);
			push(1, pb::ACT_WIPE_B_OVER);
		}
		layer_n_ops[l] = d.n_ops - layer_first_op[l];
		d.layer_first_op[l] = layer_first_op[l];
	}
	// exact occlusion culling: which layers are opaque (alpha == 1.0f) over whole strips / whole lines.
	// Needs finite values below (NaN * 0 != 0): every read table must lie in [0, 1].
	bool cull = !(c->flags & PB_CTX_NO_CULL);
	for (int i = 0; i < n_leaves && cull && any_f32; ++i) {   // an RGBA-f32 frame may hold NaN / inf (NaN * 0 != 0) unless the library made it
		if (leaves[i]->kind != pb::LEAF_RGBA_F32 && leaves[i]->kind != pb::LEAF_YADIF) continue;   // from a packed source itself
		const int t = leaves[i]->finite_lut ? lut_table_by_raw(c, leaves[i]->finite_lut) : -1;
		cull = t >= 0 && c->lut_tables[t].unit_range;
	}
	for (int i = 0; i < d.n_rc && cull; ++i) {
		const int t = lut_table_by_raw(c, d.rc[i].lut);
		cull = t >= 0 && c->lut_tables[t].unit_range;
	}
	const pb_ctx::SampleTab::Opq *lopq[pb::kMaxLayers][2] = {};   // per layer: the leaves that must all be full
	{
		int li = 0;
		for (int l = 0; l < d.n_layers; ++l) {
			const pb::Layer &ly = d.layers[l];
			const int nleaf = ly.kind == pb::LAYER_DIRECT ? 1 : (ly.kind == pb::LAYER_DISSOLVE ? 2 : 3);
			if (cull && ly.kind == pb::LAYER_DIRECT) {
				lopq[l][0] = opq[li].get();
			} else if (cull && ly.kind == pb::LAYER_DISSOLVE) {
				// transition.ts:60-65 on two alphas of 1: fma(1, mix, 1 * (1 - mix)) = RN(mix + RN(1 - mix))
				const float rmix = 1.0f - ly.mix;
				if (ly.mix + rmix == 1.0f && opq[li] && opq[li + 1]) { lopq[l][0] = opq[li].get(); lopq[l][1] = opq[li + 1].get(); }
			}   // wipe: alpha depends on the mask picture
			li += nleaf;
		}
	}
	auto layer_full = [&](int l, bool strips, int idx) {
		if (!lopq[l][0]) return false;
		for (int q = 0; q < 2; ++q)
			if (lopq[l][q] && !(strips ? lopq[l][q]->strip_full[idx] : lopq[l][q]->row_full[idx])) return false;
		return true;
	};
	for (int sidx = 0; sidx < d.n_strips; ++sidx) {
		uint32_t mask = 0;
		for (int l = 0; l < d.n_layers; ++l) {
			const pb::Layer &ly = d.layers[l];
			const pb::Leaf *ll[3] = {&ly.a, &ly.b, &ly.mask};
			const int nleaf = ly.kind == pb::LAYER_DIRECT ? 1 : (ly.kind == pb::LAYER_DISSOLVE ? 2 : 3);
			bool any = false;
			for (int q = 0; q < nleaf; ++q) any = any || (sidx >= ll[q]->s0 && sidx <= ll[q]->s1);
			// a transition whose leaves are all elsewhere yields (0,0,0,0): `over` leaves acc untouched, skip the layer;
			// otherwise all of its ops run (a leaf that is elsewhere evaluates to the border colour by itself)
			if (any) mask |= ((1u << layer_n_ops[l]) - 1u) << layer_first_op[l];
			if (layer_full(l, true, sidx)) mask |= 1u << (24 + l);
		}
		d.strip_ops[sidx] = mask;
	}
	{   // the same per output line; cached per (layer structure, leaf line ranges, height)
		std::vector<int> key = {d.out_h, d.n_layers};
		for (int l = 0; l < d.n_layers; ++l) {
			const pb::Layer &ly = d.layers[l];
			key.push_back(ly.kind);
			const pb::Leaf *ll[3] = {&ly.a, &ly.b, &ly.mask};
			for (int q = 0; q < 3; ++q) { key.push_back(ll[q]->y0); key.push_back(ll[q]->y1); }
			for (int q = 0; q < 2; ++q) key.push_back(lopq[l][q] ? lopq[l][q]->id : -1);
		}
		pb_ctx::LineOps *found = nullptr;
		for (auto &lo : c->line_ops)
			if (lo.key == key) found = &lo;
		if (!found) {
			if (c->line_ops.size() >= 64) {
				CU(cudaStreamSynchronize(c->q[PB_QUEUE_PROCESS]));
				for (auto &lo : c->line_ops) cudaFree(lo.dev);
				c->line_ops.clear();
			}
			std::vector<uint32_t> host((size_t)d.out_h);
			for (int y = 0; y < d.out_h; ++y) {
				uint32_t mask = 0;
				for (int l = 0; l < d.n_layers; ++l) {
					const pb::Layer &ly = d.layers[l];
					const pb::Leaf *ll[3] = {&ly.a, &ly.b, &ly.mask};
					const int nleaf = ly.kind == pb::LAYER_DIRECT ? 1 : (ly.kind == pb::LAYER_DISSOLVE ? 2 : 3);
					bool any = false;
					for (int q = 0; q < nleaf; ++q) any = any || (y >= ll[q]->y0 && y <= ll[q]->y1);
					if (any) mask |= ((1u << layer_n_ops[l]) - 1u) << layer_first_op[l];
					if (layer_full(l, false, y)) mask |= 1u << (24 + l);
				}
				host[y] = mask;
			}
			pb_ctx::LineOps lo;
			lo.key = key;
			lo.host = host;
			CU(cudaMalloc(&lo.dev, host.size() * sizeof(uint32_t)));
			CU(cudaMemcpyAsync(lo.dev, host.data(), host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->q[PB_QUEUE_PROCESS]));
			CU(cudaStreamSynchronize(c->q[PB_QUEUE_PROCESS]));   // `host` is a local
			c->line_ops.push_back(std::move(lo));
			found = &c->line_ops.back();
		}
		d.line_ops = found->dev;
		line_ops_host = &found->host;
		line_ops_key = key;
	}
	// the write side packs three codes into one word while regrouping: they must fit 10 bits (8 for the 8-bit sinks, whose
	// uchar stores would otherwise wrap, Q11)
	const int wt = rgba_f32_sink ? -1 : lut_table_by_raw(c, d.wc.lut);   // (an RGBA-f32 sink has no write side)
	if (!rgba_f32_sink && (wt < 0 || !c->lut_tables[wt].unit_range)) return 0;
	// (and, being inside the code range, need no saturation: the encoder drops the clamp of convert_ushort_sat_rte)
	const double code_max = (d.sink == pb::SINK_V210 || d.sink == pb::SINK_YUV422P10) ? 1023.0 : 255.0;
	for (int row = 0; row < 3 && !rgba_f32_sink; ++row) {
		double hi = d.wc.cm[row * 4 + 3], lo = hi;
		for (int k = 0; k < 3; ++k) {
			hi += std::max(0.0, (double)d.wc.cm[row * 4 + k]);
			lo += std::min(0.0, (double)d.wc.cm[row * 4 + k]);
		}
		if (!(hi < code_max + 0.25 && lo > -0.25)) return 0;
	}
	// gamma tables: all in the one-byte form (shared memory) or all raw (global memory)
	d.sparse_cm = 1;
	int slots[pb::kMaxLuts], n_slots = 0;
	bool all_d8 = !(c->flags & PB_CTX_RAW_LUT);
	auto slot_of = [&](int table) -> int {
		if (table < 0 || !c->lut_tables[table].d8) return -1;
		for (int i = 0; i < n_slots; ++i)
			if (slots[i] == table) return i;
		if (n_slots >= pb::kMaxLuts) return -1;
		slots[n_slots] = table;
		return n_slots++;
	};
	int n_t256 = 0;
	for (int i = 0; i < d.n_rc; ++i) {   // rc[0]'s table takes slot 0
		if (d.rc[i].cm[1] != 0.0f || d.rc[i].cm[10] != 0.0f) d.sparse_cm = 0;
		d.rc[i].t256_slot = -1;
		if (!rc_ycc[i]) {   // constants of rgba8 / bgra8 leaves only: 256 distinct table entries, staged as a 1 KiB table
			d.rc[i].lut_slot = -1;
			if (n_t256 >= 4) return 0;
			d.rc[i].t256_slot = n_t256++;
			continue;
		}
		d.rc[i].lut_slot = slot_of(lut_table_by_raw(c, d.rc[i].lut));
		if (d.rc[i].lut_slot < 0) all_d8 = false;
		for (int ch = 0; ch < 3; ++ch) {
			d.rk[i].mY[ch] = d.rc[i].cm[ch * 4 + 0];
			d.rk[i].oY[ch] = -8388608.0f * d.rk[i].mY[ch];
			for (int sc = 0; sc < 2; ++sc) {
				const float k = sc ? 1.0f / 1024.0f : 1.0f;   // exact scalings
				d.rk[i].mCb[ch][sc] = d.rc[i].cm[ch * 4 + 1] * k;
				d.rk[i].mCr[ch][sc] = d.rc[i].cm[ch * 4 + 2] * k;
			}
		}
	}
	d.wc.lut_slot = rgba_f32_sink ? -1 : slot_of(wt);
	if (d.wc.lut_slot < 0 && !rgba_f32_sink) all_d8 = false;
	d.n_luts = all_d8 ? n_slots : 0;
	d.n_t256 = n_t256;
	d.big_rows = big_rows;
	// The bottom layer as a row-reuse pass.  One v210 layer through an axis-aligned Transform alone (a channel playing one clip
	// through its Mixer) takes k_march_single: blocks of lines of 186-px strips, every source row converted once.  With more
	// layers on top, the same item loop runs as the second phase of the general kernel (bg_single) over the strip-pair lines on
	// which the bottom layer is the only live op after bounding boxes and occlusion culling.  Needs <= 32 source groups per strip
	// row (horizontal scale >= ~1) and pays when consecutive lines share a source row (vertical step <= 1).
	d.single_lines = 0;
	d.bg_single = 0;
	d.line_pairs = nullptr;
	d.single_strip_groups = 31;
	const bool write_plain = rgba_f32_sink || (d.wc.lut_slot >= 0 && c->lut_tables[slots[d.wc.lut_slot]].lp.affine == 1);
	const bool plain_tables = all_d8 && n_slots <= 2 && d.sparse_cm && write_plain && !rgba_f32_sink;
	bool reads_plain = plain_tables;   // all read tables in the same non-affine model: MUFU (0) or polynomial (2)
	for (int i = 0; i < d.n_rc && reads_plain; ++i)
		reads_plain = d.rc[i].lut_slot >= 0 && c->lut_tables[slots[d.rc[i].lut_slot]].lp.affine != 1 &&
		              c->lut_tables[slots[d.rc[i].lut_slot]].lp.affine == c->lut_tables[slots[d.rc[0].lut_slot]].lp.affine;
	const int k0 = d.layers[0].a.kind;
	const bool l0_v210 = k0 == pb::LEAF_V210;
	const bool l0_planar = k0 == pb::LEAF_YUV422P10 || k0 == pb::LEAF_YUV422P8 || k0 == pb::LEAF_YUV420P || k0 == pb::LEAF_NV12;
	// (stand-alone: v210 or a planar FFmpegProducer clip; as a background pass of the fast variant: v210 only)
	if (reads_plain && d.n_ops >= 1 && d.layers[0].kind == pb::LAYER_DIRECT && d.layers[0].a.has_xf &&
	    ((d.n_ops == 1 && (l0_v210 || l0_planar)) || (l0_v210 && !any_planar && d.sink == pb::SINK_V210)) &&
	    d.layers[0].a.w % 6 == 0 && d.sink != pb::SINK_RGBA8 && d.sink != pb::SINK_BGRA8 && d.out_w % 48 == 0 && d.interlace == 0 && !big_rows &&
	    !(c->flags & PB_CTX_NO_DIRECT) && tab_of[0] && !tab_of[0]->col_i0.empty() && line_ops_host) {
		const auto &tb = *tab_of[0];
		const pb::Leaf &lf = d.layers[0].a;
		const bool bg = d.n_ops > 1;
		const int kSG = bg ? 2 * d.strip_groups : 31;   // as a background pass: two strips of the general kernel
		const int groups = d.out_w / 6, n_strips = (groups + kSG - 1) / kSG;
		bool ok = n_strips <= 64 && (!bg || d.strip_groups == pb::kStripGroupsXf);
		for (int y = 0; y + 1 < d.out_h && ok; ++y) ok = std::abs(tb.row_j0[y + 1] - tb.row_j0[y]) <= 1;
		for (int sidx = 0; sidx < n_strips && ok; ++sidx) {
			const int x0 = sidx * kSG * 6, x1 = std::min(x0 + kSG * 6, d.out_w) - 1;
			int lo = INT32_MAX, hi = INT32_MIN;
			for (int x = x0; x <= x1; ++x) {
				lo = std::min(lo, tb.col_i0[x]);
				hi = std::max(hi, tb.col_i0[x] + 1);
			}
			int2 e = make_int2(0, 0);
			if (!(hi < 0 || lo >= lf.w)) {
				const bool interior = lo >= 0 && hi < lf.w;   // (before the clamp below) no tap column outside the image
				lo = std::max(lo, 0);
				hi = std::min(hi, lf.w - 1);
				e.x = lo / 6;
				e.y = hi / 6 - e.x + 1;
				ok = e.y <= 32;
				if (interior) e.y |= 0x100;
			}
			d.single_strips[sidx] = e;
		}
		if (ok && bg) {
			// which strip-pair lines are background-only: todo == {op 0} on both strips, evaluated as the kernel evaluates it
			pb_ctx::LinePairs *found = nullptr;
			std::vector<uint32_t> sops(d.strip_ops, d.strip_ops + d.n_strips);
			for (auto &lp_ : c->line_pairs)
				if (lp_.key == line_ops_key && lp_.strip_ops == sops) found = &lp_;
			if (!found) {
				if (c->line_pairs.size() >= 64) {
					CU(cudaStreamSynchronize(c->q[PB_QUEUE_PROCESS]));
					for (auto &lp_ : c->line_pairs) cudaFree(lp_.dev);
					c->line_pairs.clear();
				}
				std::vector<unsigned long long> host((size_t)d.out_h, 0ull);
				auto todo_of = [&](int sidx, int y) {
					const uint32_t both = d.strip_ops[sidx] & (*line_ops_host)[y];
					uint32_t todo = both & 0xFFFFFFu;
					if (both >> 24) todo &= ~0u << d.layer_first_op[(31 - __builtin_clz(both)) - 24];
					return todo;
				};
				long long marked = 0;
				for (int y = 0; y < d.out_h; ++y) {
					unsigned long long m = 0;
					for (int pr = 0; pr < n_strips; ++pr) {
						const int sa = 2 * pr, sb = 2 * pr + 1;
						if (todo_of(sa, y) == 1u && (sb >= d.n_strips || todo_of(sb, y) == 1u)) {
							m |= 1ull << pr;
							++marked;
						}
					}
					host[y] = m;
				}
				pb_ctx::LinePairs e;
				e.key = line_ops_key;
				e.strip_ops = sops;
				// Worth a second phase only when it has a few blocks for every warp of the grid.  Measured on B200 with the items
				// claimed dynamically: 4320p two layers 426 -> 361 us, 2160p two layers 113 -> 106 us; the 2160p four-layer bench
				// scene (17 k marked lines) is neutral, smaller frames lose
				const long long warps_ = (long long)c->prop.multiProcessorCount * pb::kMarchWarps;
				long long min_marked = 8 * warps_;
				if (const char *ov = getenv("PB_BG_MIN")) min_marked = atoll(ov);   // tests force the pass on small frames
				if (marked > 0 && marked >= min_marked) {
					CU(cudaMalloc(&e.dev, host.size() * sizeof(unsigned long long)));
					CU(cudaMemcpyAsync(e.dev, host.data(), host.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->q[PB_QUEUE_PROCESS]));
					CU(cudaStreamSynchronize(c->q[PB_QUEUE_PROCESS]));   // `host` is a local
				}
				c->line_pairs.push_back(std::move(e));
				found = &c->line_pairs.back();
			}
			ok = found->dev != nullptr;   // no background-only line anywhere: nothing to gain
			d.line_pairs = found->dev;
		}
		if (ok) {
			// Lines per work item: taller blocks reuse more rows (a block of L lines costs L + 1 conversion passes) but leave fewer
			// items to spread over the grid's warps.  Pick the L that minimises the longest warp's work: rounds x (L lines of
			// sampling + encoding (~1060 issue slots per 186-px line) + L + 1 ... conversion passes (~360 each)).
			const long long warps = (long long)c->prop.multiProcessorCount * pb::kMarchWarps;
			long long best = -1;
			for (int L = 1; L <= 16; ++L) {
				const long long items = (long long)n_strips * ((d.out_h + L - 1) / L), rounds = (items + warps - 1) / warps;
				const long long cost = rounds * (L * 1060LL + 360LL);
				if (best < 0 || cost < best) { best = cost; d.single_lines = L; }
			}
			if (bg) d.single_lines = 6;   // (items are claimed dynamically: moderately tall blocks balance and still reuse 5 rows of 6)
			d.single_strip_groups = kSG;
			d.bg_single = bg ? 1 : 0;
		}
	}
	// ToRGBA -> FromRGBA of one v210 source with colourMaths-style tables: the dedicated direct kernel
	// (or into an RGBA-f32 frame: a ToRGBA output made real, k_march_direct<.., true>)
	d.direct_mode = d.n_ops == 1 && d.layers[0].kind == pb::LAYER_DIRECT && !d.layers[0].a.has_xf && d.layers[0].a.kind == pb::LEAF_V210 &&
	                d.layers[0].a.w % 6 == 0 && (d.sink == pb::SINK_V210 || rgba_f32_sink) && d.out_w % 48 == 0 && all_d8 && n_slots <= 2 && d.sparse_cm &&
	                (!any_planar || rgba_f32_sink) && !big_rows && d.n_rc == 1 &&
	                d.rc[0].lut_slot >= 0 && c->lut_tables[slots[d.rc[0].lut_slot]].lp.affine != 1 && write_plain && !(c->flags & PB_CTX_NO_DIRECT);
	// FromRGBA of a frame that is real memory (one RGBA-f32 leaf read 1:1 into a v210 output: a routed channel frame, a Yadif
	// output): the direct kernel with the frame itself as the pixel source
	if (!d.direct_mode && d.n_ops == 1 && d.layers[0].kind == pb::LAYER_DIRECT && !d.layers[0].a.has_xf && d.layers[0].a.kind == pb::LEAF_RGBA_F32 &&
	    d.sink == pb::SINK_V210 && d.out_w % 48 == 0 && all_d8 && n_slots == 1 && d.wc.lut_slot == 0 && write_plain && !(c->flags & PB_CTX_NO_DIRECT))
		d.direct_mode = 2;
	if (big_rows && d.direct_mode != 2) {   // 2 x 64 KiB of tables + 20 x 4.5 KiB of rows is what an SM holds
		if (d.n_luts > 2) return 0;
		any_planar = true;
	}
	if (any_rgba && d.n_luts == 0) return 0;   // rgba8 leaves ride on the big-row variants, which exist for shared-memory tables
	d.any_planar = any_planar;
	if (any_planar && !(d.n_luts > 0 && d.sparse_cm)) return 0;   // planar variants exist for the common configuration only
	d.feat = rgba_f32_sink ? 4 : 0;
	for (int i = 0; i < n_leaves; ++i) {
		if (leaves[i]->lz_tx && !leaves[i]->lz_sep) d.feat |= 1;
		if (leaves[i]->kind == pb::LEAF_RGBA_F32 || leaves[i]->kind == pb::LEAF_YADIF) d.feat |= 2;
	}
	for (int i = 0; i < n_leaves; ++i)   // the first pass of a separable Lanczos leaf decodes its table from shared memory
		if (leaves[i]->lz_sep && !(d.n_luts > 0 && d.sparse_cm && d.rc[leaves[i]->rc].lut_slot >= 0)) return 0;
	if (rgba_f32_sink && !d.direct_mode) {
		// A frame made real.  The march kernel pays where packed leaves are sampled through a Transform (every texel converted once
		// instead of once per tap); a graph of 1:1 packed reads and RGBA-f32 frames is a few gathers per pixel, which the generic
		// kernel does faster than the row-buffer round trip (a routed channel frame: 17 vs 52 us, profiles/r02_route_launches.txt)
		bool sampled_packed = false;
		for (int i = 0; i < n_leaves; ++i) sampled_packed = sampled_packed || (leaves[i]->kind != pb::LEAF_RGBA_F32 && leaves[i]->has_xf);
		if (!sampled_packed && !getenv("PB_RGBA_SINK_MARCH")) return 0;
	}
	for (int i = 0; i < d.n_luts; ++i) {
		d.luts[i].d8 = c->lut_tables[slots[i]].d8;
		d.luts[i].lp = c->lut_tables[slots[i]].lp;
	}
	if (d.n_luts && d.wc.lut_slot >= 0) d.wlp = d.luts[d.wc.lut_slot].lp;
	else if (rgba_f32_sink) d.wlp.affine = 1;   // (no write table: the kernel variants are picked by the read tables alone)
	if (c->flags & PB_CTX_FOOTPRINT) {
		// distinct packed source bytes this launch reads (after bounding-box masks and occlusion culling): per
		// leaf and strip, the distinct source rows of the lines on which the leaf's op survives
		std::vector<uint32_t> lines((size_t)d.out_h);
		CU(cudaMemcpy(lines.data(), d.line_ops, lines.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
		const int step = d.interlace == 0 ? 1 : 2, first = d.interlace == 3 ? 1 : 0;
		uint64_t bytes = 0;
		std::vector<int> seen;
		for (int oi = 0; oi < d.n_ops; ++oi) {
			int li = 0;   // index of the op's leaf in leaves[] / opq[]
			for (int l = 0; l < d.ops[oi].layer; ++l) li += layer_n_ops[l];
			li += d.ops[oi].which;
			if (!tab_of[li]) continue;   // (a rotated RGBA-f32 leaf: no tables; not a packed source either)
			const auto &o = *tab_of[li];
			for (int sidx = 0; sidx < d.n_strips; ++sidx) {
				if (!o.strip_ng[sidx]) continue;
				seen.assign((size_t)o.src_h, 0);
				for (int y = first; y < d.out_h; y += step) {
					const uint32_t both = d.strip_ops[sidx] & lines[y];
					uint32_t todo = both & 0xFFFFFFu;
					if (both >> 24) todo &= ~0u << d.layer_first_op[(31 - __builtin_clz(both)) - 24];
					if (!((todo >> oi) & 1u)) continue;
					for (int r = 0; r < o.rows_per_line; ++r) {
						const int j = o.row_j0[y] + r;
						if (j >= 0 && j < o.src_h) seen[j] = 1;
					}
				}
				uint64_t rows = 0;
				for (int v : seen) rows += v;
				bytes += rows * (uint64_t)o.strip_ng[sidx] * 16u;
			}
		}
		c->stats.march_src_bytes = bytes;
	}
	return 1;
}

// issue the launch(es) of a prepared descriptor: the march or the generic kernel, plus -- after a march launch on a ragged
// v210 width -- the generic kernel on the tail columns (prepare_march).  Also the replay path of recorded chains.
int launch_compiled(pb_ctx *c, cudaStream_t s, const pb::FusedDesc &d_in, bool march, void *out_rgba, const std::vector<pb::HPassDesc> *pre) {
	if (pre)
		for (const pb::HPassDesc &h : *pre) {   // first passes of separable Lanczos leaves, interpolated lines of Yadif leaves
			cudaError_t e = h.pre_kind == 1 ? pb::launch_yadif_rows(s, h.lf.ptr_u, h.lf.ptr, h.lf.ptr_v, h.lf.yadif & 1, (h.lf.yadif >> 1) & 1,
			                                                        (h.lf.yadif >> 2) & 1, h.out, h.lf.w, h.lf.h)
			                                : pb::launch_lanczos_hpass(s, h, c->march_sms);
			if (e != cudaSuccess) return fail(PB_ERR_CUDA, "%s: %s", h.pre_kind == 1 ? "yadif lines" : "lanczos first pass", cudaGetErrorString(e));
			c->stats.kernel_launches++;   // the caller counts the main launch
		}
	pb::FusedDesc bg_copy;
	const pb::FusedDesc *dp = &d_in;
	if (march && d_in.bg_single) {
		// The second phase claims its items from a counter in global memory (pb_march.cu march_single_items).  Every launch gets
		// a counter of its own out of a ring, zeroed on the launching stream just before the launch: nothing to predict on the
		// host, nothing shared between launches on different queues, and a failed launch leaves no state behind.
		constexpr unsigned kRing = 256;
		if (!c->bg_counter) CU(cudaMalloc(&c->bg_counter, kRing * sizeof(unsigned int)));
		unsigned int *ctr = c->bg_counter + (c->bg_next_base++ % kRing);
		CU(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), s));
		bg_copy = d_in;
		bg_copy.bg_counter = ctr;
		bg_copy.bg_base = 0;
		dp = &bg_copy;
	}
	const pb::FusedDesc &d = *dp;
	cudaError_t e = march ? pb::launch_fused_march(s, d, c->march_sms) : pb::launch_fused(s, d, out_rgba);
	if (e != cudaSuccess) return fail(PB_ERR_CUDA, "fused launch (%s): %s", march ? "march" : "generic", cudaGetErrorString(e));
	if (march && d.sink == pb::SINK_V210 && d.out_w % 48 != 0) {
		pb::FusedDesc tail = d;
		tail.g_first = d.march_w / 6;
		e = pb::launch_fused(s, tail, nullptr);
		if (e != cudaSuccess) return fail(PB_ERR_CUDA, "fused launch (line tails): %s", cudaGetErrorString(e));
		c->stats.kernel_launches++;   // the caller counts the main launch
	}
	return PB_OK;
}

// launch a compiled descriptor (march kernel when eligible); *march_out reports the choice
int launch_desc(pb_ctx *c, cudaStream_t s, pb::FusedDesc &d, void *out_rgba, bool *march_out) {
	bool march = false;
	if (!out_rgba || d.sink == pb::SINK_RGBA_F32) {   // (an RGBA-f32 destination: the march kernel where it can, else the generic one into out_rgba)
		int r = prepare_march(c, d);
		if (r < 0) return r;
		march = r == 1;
	}
	// Lanczos leaves of a march launch are evaluated separably: a first launch converts every source row the filter reaches
	// ONCE and filters it horizontally into H (RGBA-f32 rows in HBM / L2); the fused launch then runs the vertical chains over
	// H.  Evaluating the filter inside the one launch converts every source row once per output line that reaches it (six
	// times at 0.5x with three lobes): 2.8 ms against 0.x ms for BASELINE.json's config 5 (profiles/r02_bench_config5*.json).
	// Same fma chains in the same order, so the same bits.  PB_LANCZOS_ONE_PASS=1 keeps the single launch (A/B).
	for (auto &sc : c->pending_scratch) c->pool.dev_put(sc.second, sc.first);   // (left over by a caller that did not record)
	c->pending_scratch.clear();
	c->pending_pre.clear();
	// Yadif leaves (march and generic kernel alike): the interpolated lines of the field -- half of its lines; 27 float4 reads and
	// two predictors per channel each -- are computed ONCE by a pre-pass into a half-height RGBA-f32 block; the lines of the
	// field's own parity are read from the current frame where they lie.  Evaluating the filter where the field is sampled
	// costs it once per tap row (twice per pixel through the Mixer's Transform): 727 against xxx us per 2160p field
	// (profiles/r02_kbench_yadif.txt).  The consuming launch sees two RGBA-f32 frames: ptr = cur, ptr_u = the block.
	for (int l = 0; l < d.n_layers; ++l) {
		pb::Layer &ly = d.layers[l];
		pb::Leaf *ll[3] = {&ly.a, &ly.b, &ly.mask};
		const int nleaf = ly.kind == pb::LAYER_DIRECT ? 1 : (ly.kind == pb::LAYER_DISSOLVE ? 2 : 3);
		for (int q = 0; q < nleaf; ++q) {
			pb::Leaf &lf = *ll[q];
			if (lf.kind != pb::LEAF_YADIF || (lf.yadif & 8)) continue;   // (bit 3: ptr_u already is the block of interpolated lines)
			void *rows = nullptr;
			for (const pb::HPassDesc &h : c->pending_pre)   // the same field sampled by another leaf of this launch
				if (h.pre_kind == 1 && h.lf.ptr == lf.ptr && h.lf.ptr_u == lf.ptr_u && h.lf.ptr_v == lf.ptr_v && h.lf.yadif == lf.yadif) rows = h.out;
			if (!rows) {
				pb::HPassDesc h{};
				h.pre_kind = 1;
				h.lf = lf;
				const size_t bytes = (size_t)((lf.h + 1) / 2) * lf.w * sizeof(float4);
				CU(c->pool.dev_get(bytes, &rows));
				h.out = (float4 *)rows;
				c->pending_pre.push_back(h);
				c->pending_scratch.push_back({rows, bytes});
			}
			lf.ptr_u = rows;
			lf.ptr_v = nullptr;
			lf.yadif |= 8;
		}
	}
	if (march) {
		for (int l = 0; l < d.n_layers; ++l) {
			pb::Layer &ly = d.layers[l];
			pb::Leaf *ll[3] = {&ly.a, &ly.b, &ly.mask};
			const int nleaf = ly.kind == pb::LAYER_DIRECT ? 1 : (ly.kind == pb::LAYER_DISSOLVE ? 2 : 3);
			for (int q = 0; q < nleaf; ++q) {
				pb::Leaf &lf = *ll[q];
				if (!lf.lz_tx || !lf.lz_sep || lf.kind == pb::LEAF_LANCZOS_V) continue;
				pb_ctx::LanczosTab *lt = nullptr;
				for (auto &e : c->lanczos_tabs)
					if (e.i0 == lf.lz_i0) lt = &e;
				const int slot = d.rc[lf.rc].lut_slot;
				if (!lt || lt->strip_groups != d.strip_groups || slot < 0 || d.n_luts <= slot || !d.sparse_cm)
					return fail(PB_ERR_STATE, "separable lanczos leaf without its tables");
				pb::HPassDesc h{};
				h.lf = lf;
				h.rc = d.rc[lf.rc];
				h.rk = d.rk[lf.rc];
				h.lut = d.luts[slot];
				h.xf_w = lf.xf_w;
				h.strip_groups = d.strip_groups;
				h.s0 = lt->s0;
				h.s1 = lt->s1;
				h.j_lo = std::max(0, lt->h_j0[lt->y0]);
				h.j_hi = std::min(lf.h, lt->h_j0[lt->y1] + lt->ty);
				for (int y = lt->y0; y <= lt->y1; ++y) {   // (not assumed monotone: flips)
					h.j_lo = std::min(h.j_lo, std::max(0, lt->h_j0[y]));
					h.j_hi = std::max(h.j_hi, std::min(lf.h, lt->h_j0[y] + lt->ty));
				}
				h.e_magic = d.e_magic;
				h.lds_koff = d.lds_koff;
				const size_t bytes = (size_t)lf.h * lf.xf_w * sizeof(float4);
				void *H = nullptr;
				CU(c->pool.dev_get(bytes, &H));
				h.out = (float4 *)H;
				c->pending_pre.push_back(h);
				c->pending_scratch.push_back({H, bytes});
				// the consuming launch sees the second pass
				lf.kind = pb::LEAF_LANCZOS_V;
				lf.ptr = H;
				lf.w = lf.xf_w;   // H: source rows x output columns
			}
		}
	}
	int r = launch_compiled(c, s, d, march, out_rgba, c->pending_pre.empty() ? nullptr : &c->pending_pre);
	if (!c->recording) {   // (a recording keeps the intermediates: record_launch takes them over)
		for (auto &sc : c->pending_scratch) c->pool.dev_put(sc.second, sc.first);   // stream-ordered: the next user of a block waits for this launch
		c->pending_scratch.clear();
		c->pending_pre.clear();
	}
	if (r) return r;
	if (march_out) *march_out = march;
	return PB_OK;
}

}  // namespace pbrt
