// pb_runtime.cu -- the C ABI (include/phaneron_b200.h): contexts, refcounted buffers with a
// pinned host face, op-enum "programs", and the frame-expression recorder that turns
// phaneron's per-stage jobs (clJobQueue.ts:122-128) into one fused launch per output frame.
//
// Design: in PB_CTX_DEFER mode a job whose output is an RGBA-f32 image
// is not executed; its output buffer gets an expression node that references the input
// expressions (and holds references on the packed leaves).  A packed writer is a sink: it
// flattens the expression into a pb::FusedDesc and launches once.  Anything the fused kernel
// cannot express is materialised bottom-up with the stand-alone kernels and re-enters as an
// RGBA leaf.  Host reads of a deferred frame materialise it on demand.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/phaneron_b200.h"
#include "pb_desc.h"
#include "pb_launch.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
	char tmp[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(tmp, sizeof tmp, fmt, ap);
	va_end(ap);
	g_err = tmp;
	return code;
}

#define CU(call)                                                                                   \
	do {                                                                                           \
		cudaError_t e__ = (call);                                                                  \
		if (e__ != cudaSuccess) return fail(PB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
	} while (0)

struct Node;
using NodeP = std::shared_ptr<Node>;

// size-keyed free lists; frames of one format recycle the same few blocks, so steady state
// performs no cudaMalloc/cudaFree (the reference allocates a fresh SVM buffer per stage per
// frame: mixer.ts:196, transitioner.ts:152, combiner.ts:230)
struct Pool {
	std::unordered_map<size_t, std::vector<void *>> dev, host;
	size_t dev_pooled = 0, dev_live = 0;
	static constexpr size_t kMaxPooled = size_t(24) << 30;

	cudaError_t dev_get(size_t n, void **p) {
		auto &v = dev[n];
		if (!v.empty()) {
			*p = v.back();
			v.pop_back();
			dev_pooled -= n;
			dev_live += n;
			return cudaSuccess;
		}
		cudaError_t e = cudaMalloc(p, n);
		if (e != cudaSuccess) {   // give pooled memory back and retry once
			trim();
			e = cudaMalloc(p, n);
		}
		if (e == cudaSuccess) dev_live += n;
		return e;
	}
	void dev_put(size_t n, void *p) {
		if (!p) return;
		dev_live -= n;
		if (dev_pooled + n > kMaxPooled) {
			cudaFree(p);
			return;
		}
		dev[n].push_back(p);
		dev_pooled += n;
	}
	cudaError_t host_get(size_t n, void **p) {
		auto &v = host[n];
		if (!v.empty()) {
			*p = v.back();
			v.pop_back();
			return cudaSuccess;
		}
		return cudaMallocHost(p, n);
	}
	void host_put(size_t n, void *p) {
		if (p) host[n].push_back(p);
	}
	void trim() {
		for (auto &kv : dev)
			for (void *p : kv.second) cudaFree(p);
		dev.clear();
		dev_pooled = 0;
	}
	void destroy() {
		trim();
		for (auto &kv : host)
			for (void *p : kv.second) cudaFreeHost(p);
		host.clear();
	}
};

}  // namespace

struct pb_ctx {
	int dev = 0;
	unsigned flags = 0;
	cudaStream_t q[3] = {nullptr, nullptr, nullptr};
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_x = nullptr;
	std::recursive_mutex mu;
	Pool pool;
	pb_stats stats{};
	cudaDeviceProp prop{};
	struct pb_chain *recording = nullptr;
	// sampling tables of the march kernel, cached per (transform, source dims, output dims, strip width)
	struct SampleTab {
		float m[6];
		int sw, sh, W, H, has_xf, strip_groups, fits = 1;
		int s0 = 0, s1 = -1, y0 = 0, y1 = -1;   // active strips / lines
		void *dev = nullptr;
		int2 *dcol = nullptr, *drow = nullptr;
		int4 *dstrip = nullptr;
		// where the leaf is exactly opaque (alpha == 1.0f bit for bit): whole strips x lines; see leaf_opacity()
		struct Opq {
			int id;
			std::vector<uint8_t> strip_full, row_full;
			std::vector<int> row_j0, strip_ng;   // first source row each output line reads, source groups per strip (footprint accounting)
			std::vector<int> col_i0;             // first source column each output column reads (k_march_single strip footprints)
			int rows_per_line = 1, src_h = 0;
		};
		std::shared_ptr<const Opq> opq;
	};
	std::vector<SampleTab> tabs;
	int next_tab_id = 0;
	// Lanczos tap tables, cached per (matrix, source dims, output dims, lobes)
	struct LanczosTab {
		float m[4];   // m0, m2, m4, m5
		int sw, sh, W, H, lobes, tx, ty;
		void *dev = nullptr;
		int *i0 = nullptr, *j0 = nullptr;
		float *wx = nullptr, *wy = nullptr;
	};
	std::vector<LanczosTab> lanczos_tabs;
	// blocking-sync events for waits on the copy queues: a host thread waiting for a frame-sized DMA sleeps instead of
	// spinning, and waits for ITS copy only, not for whatever other producers have queued behind it
	std::vector<cudaEvent_t> copy_events;
	bool allow_march = true;
	// gamma tables by content (see lut_table_of)
	struct LutTable {
		unsigned long long hash = 0;
		float *raw = nullptr;      // context-owned copy every ReadConsts/WriteConsts points at
		uint8_t *d8 = nullptr;      // one-byte form, null if no model fits
		pb::LutParams lp{};
		int model = -1, dmin = 0, dmax = 0;
		bool unit_range = false;
	};
	struct LutFit {
		uint64_t version;
		int table;
	};
	std::vector<LutTable> lut_tables;
	std::vector<LutFit> lut_fits;
	pb::LutParams lut_cands[6];
	void *lut_cands_dev = nullptr, *lut_res_dev = nullptr, *lut_scratch = nullptr;
	uint64_t version_counter = 0;
	struct LineOps {   // per-line op masks of the march kernel
		std::vector<int> key;
		uint32_t *dev = nullptr;
		std::vector<uint32_t> host;   // the same, for the host passes that need it (background-pass masks)
	};
	struct LinePairs {   // per-line strip-pair masks of the background pass (FusedDesc::line_pairs)
		std::vector<int> key;
		std::vector<uint32_t> strip_ops;
		unsigned long long *dev = nullptr;
	};
	std::vector<LinePairs> line_pairs;
	unsigned int *bg_counter = nullptr;   // device counter of the background pass (never reset) and the value the next launch starts from
	unsigned int bg_next_base = 0;
	std::vector<LineOps> line_ops;
};

struct pb_buf {
	pb_ctx *ctx = nullptr;
	size_t bytes = 0;
	int dir = 0, svm = 0, w = 0, h = 0;
	std::atomic<int> refs{1};
	void *dev = nullptr;
	bool dev_external = false;
	void *host = nullptr;
	bool host_dirty = false;   // host face written since the last upload
	uint64_t version = 0;      // unique id of the device contents (bumped on every upload)
	NodeP expr;                // non-null: frame exists only as an expression
	std::string owner;
};

struct pb_prog {
	pb_ctx *ctx;
	int op, w, h;
};

struct pb_chain {
	pb_ctx *ctx = nullptr;
	struct Item {
		pb::FusedDesc d;
		bool march = false;
		void *out_rgba;
		std::vector<std::shared_ptr<void>> keep;   // expression nodes (hold the leaf buffers)
		pb_buf *out_buf;                           // addref'd destination
	};
	std::vector<Item> items;
	bool complete = true;
};

namespace {

enum NodeKind { N_LEAF_V210, N_LEAF_RGBA, N_TRANSFORM, N_DISSOLVE, N_WIPE_MASK, N_COMBINE, N_LEAF_PACKED };

struct Node {
	NodeKind kind;
	pb_ctx *ctx;
	int w = 0, h = 0;            // dimensions of the image this node produces
	std::vector<NodeP> in;
	pb_buf *src = nullptr;       // leaves: referenced source buffer
	pb_buf *src_u = nullptr, *src_v = nullptr;   // N_LEAF_PACKED, planar formats: chroma planes
	int leaf_kind = 0;           // N_LEAF_PACKED: pb::LeafKind (rgba8, bgra8, yuv422p10/8, yuv420p, nv12)
	pb_buf *lut_buf = nullptr;   // packed leaves: referenced gamma LUT buffer
	pb::ReadConsts rc{};         // packed leaves
	float mat[6] = {0};          // transform
	int lanczos = 0;             // transform: 0 = the reference's bilinear sampler, else Lanczos lobes
	float mix = 0.f;             // dissolve
	void *mat_dev = nullptr;     // RGBA-f32 copy if this node had to be materialised
	~Node();
};

void buf_release_locked(pb_buf *b);

Node::~Node() {
	std::lock_guard<std::recursive_mutex> lk(ctx->mu);
	if (mat_dev) ctx->pool.dev_put((size_t)w * h * 16, mat_dev);
	if (src) buf_release_locked(src);
	if (src_u) buf_release_locked(src_u);
	if (src_v) buf_release_locked(src_v);
	if (lut_buf) buf_release_locked(lut_buf);
}

void buf_free(pb_buf *b) {
	pb_ctx *c = b->ctx;
	b->expr.reset();
	if (b->dev && !b->dev_external) c->pool.dev_put(b->bytes, b->dev);
	if (b->host) c->pool.host_put(b->bytes, b->host);
	delete b;
}

void buf_release_locked(pb_buf *b) {
	if (b->refs.fetch_sub(1) == 1) buf_free(b);
}

int ensure_dev(pb_buf *b) {
	if (b->dev) return PB_OK;
	CU(b->ctx->pool.dev_get(b->bytes, &b->dev));
	return PB_OK;
}

int ensure_host(pb_buf *b) {
	if (b->host) return PB_OK;
	CU(b->ctx->pool.host_get(b->bytes, &b->host));
	return PB_OK;
}

// push pending host writes (hostAccess('writeonly') without a source) to the device
int flush_host(pb_buf *b, cudaStream_t s) {
	if (!b->host_dirty) return PB_OK;
	int r = ensure_dev(b);
	if (r) return r;
	CU(cudaMemcpyAsync(b->dev, b->host, b->bytes, cudaMemcpyHostToDevice, s));
	b->ctx->stats.h2d_bytes += b->bytes;
	b->host_dirty = false;
	b->version = ++b->ctx->version_counter;
	return PB_OK;
}

const pb_param *find(const pb_param *p, int n, const char *name) {
	for (int i = 0; i < n; ++i)
		if (p[i].name && 0 == strcmp(p[i].name, name)) return &p[i];
	return nullptr;
}

int need_buf(const pb_param *p, int n, const char *name, pb_buf **out) {
	const pb_param *q = find(p, n, name);
	if (!q || q->kind != PB_PARAM_BUF || !q->buf) return fail(PB_ERR_ARG, "missing buffer parameter '%s'", name);
	if (q->buf->refs.load() <= 0) return fail(PB_ERR_STATE, "parameter '%s' is a released buffer", name);
	*out = q->buf;
	return PB_OK;
}

int need_num(const pb_param *p, int n, const char *name, double *out) {
	const pb_param *q = find(p, n, name);
	if (!q || q->kind != PB_PARAM_NUM) return fail(PB_ERR_ARG, "missing numeric parameter '%s'", name);
	*out = q->num;
	return PB_OK;
}

// small constant buffers (matrices) are read from their host face
int host_floats(pb_buf *b, int count, float *out, const char *what) {
	if (b->bytes < (size_t)count * 4) return fail(PB_ERR_ARG, "%s buffer holds %zu bytes, need %d", what, b->bytes, count * 4);
	if (!b->host) return fail(PB_ERR_STATE, "%s buffer was never written by the host", what);
	memcpy(out, b->host, (size_t)count * 4);
	return PB_OK;
}

int lut_table_of(pb_ctx *c, pb_buf *lut, int *table_out);

int make_read_consts(pb_ctx *c, const pb_param *p, int n, bool ycbcr, pb::ReadConsts *rc, pb_buf **lut_out) {
	pb_buf *lut, *gamut, *cm = nullptr;
	int r;
	if ((r = need_buf(p, n, "gammaLut", &lut))) return r;
	if ((r = need_buf(p, n, "gamutMatrix", &gamut))) return r;
	if (ycbcr && (r = need_buf(p, n, "colMatrix", &cm))) return r;
	memset(rc, 0, sizeof *rc);
	if (cm && (r = host_floats(cm, 12, rc->cm, "colMatrix"))) return r;
	if ((r = host_floats(gamut, 9, rc->gamut, "gamutMatrix"))) return r;   // Q4: only 9 floats are meaningful
	if (lut->bytes < 65536 * 4) return fail(PB_ERR_ARG, "gammaLut must hold 65536 floats");
	if ((r = flush_host(lut, c->q[PB_QUEUE_PROCESS]))) return r;
	if (!lut->dev) return fail(PB_ERR_STATE, "gammaLut was never written");
	int table;
	if ((r = lut_table_of(c, lut, &table))) return r;
	rc->lut = c->lut_tables[table].raw;   // one pointer per distinct table content
	rc->lut_slot = -1;
	rc->t256_slot = -1;
	*lut_out = lut;
	return PB_OK;
}

int make_write_consts(pb_ctx *c, const pb_param *p, int n, bool ycbcr, pb::WriteConsts *wc) {
	pb_buf *lut, *cm = nullptr;
	int r;
	if ((r = need_buf(p, n, "gammaLut", &lut))) return r;
	if (ycbcr && (r = need_buf(p, n, "colMatrix", &cm))) return r;
	memset(wc, 0, sizeof *wc);
	if (cm && (r = host_floats(cm, 12, wc->cm, "colMatrix"))) return r;
	if (lut->bytes < 65536 * 4) return fail(PB_ERR_ARG, "gammaLut must hold 65536 floats");
	if ((r = flush_host(lut, c->q[PB_QUEUE_PROCESS]))) return r;
	if (!lut->dev) return fail(PB_ERR_STATE, "gammaLut was never written");
	int table;
	if ((r = lut_table_of(c, lut, &table))) return r;
	wc->lut = c->lut_tables[table].raw;
	wc->lut_slot = -1;
	return PB_OK;
}

inline int v210_pitch_bytes(int w) { return ((w + 47) / 48) * 128; }

// ---- expression handling -------------------------------------------------------------------

// the expression an RGBA input buffer stands for (its recorded node, or itself as a leaf)
int input_expr(pb_buf *b, NodeP *out) {
	if (b->expr) {
		*out = b->expr;
		return PB_OK;
	}
	if (b->w <= 0 || b->h <= 0) return fail(PB_ERR_ARG, "image input '%s' was created without imageDims", b->owner.c_str());
	int r = flush_host(b, b->ctx->q[PB_QUEUE_PROCESS]);
	if (r) return r;
	if (!b->dev) return fail(PB_ERR_STATE, "image input '%s' has no contents", b->owner.c_str());
	auto n = std::make_shared<Node>();
	n->kind = N_LEAF_RGBA;
	n->ctx = b->ctx;
	n->w = b->w;
	n->h = b->h;
	n->src = b;
	b->refs.fetch_add(1);
	*out = n;
	return PB_OK;
}

int materialise_node(pb_ctx *c, const NodeP &n, const void **dev_out);

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
			for (auto &e : c->lanczos_tabs) cudaFree(e.dev);
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
		std::vector<float> packed((size_t)W * tx + (size_t)H * ty);
		for (int x = 0; x < W; ++x) memcpy(&packed[(size_t)x * tx], &wx[(size_t)kLanczosMaxTaps * x], sizeof(float) * tx);
		for (int y = 0; y < H; ++y) memcpy(&packed[(size_t)W * tx + (size_t)y * ty], &wy[(size_t)kLanczosMaxTaps * y], sizeof(float) * ty);
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
		c->lanczos_tabs.push_back(e);
		t = &c->lanczos_tabs.back();
	}
	lf->lz_tx = t->tx; lf->lz_ty = t->ty;
	lf->lz_i0 = t->i0; lf->lz_j0 = t->j0;
	lf->lz_wx = t->wx; lf->lz_wy = t->wy;
	return PB_OK;
}

struct Compiler {
	pb_ctx *c;
	pb::FusedDesc d;
	std::vector<NodeP> keep;

	int rc_index(const pb::ReadConsts &rc, int *idx) {
		for (int i = 0; i < d.n_rc; ++i)
			if (0 == memcmp(&d.rc[i], &rc, sizeof rc)) {
				*idx = i;
				return PB_OK;
			}
		if (d.n_rc >= pb::kMaxReadConsts) return 1;   // caller materialises instead
		d.rc[d.n_rc] = rc;
		*idx = d.n_rc++;
		return PB_OK;
	}

	int set_basic_leaf(const NodeP &n, pb::Leaf *lf) {
		if (n->kind == N_LEAF_V210) {
			int idx;
			if (rc_index(n->rc, &idx)) return 1;
			lf->kind = pb::LEAF_V210;
			lf->ptr = n->src->dev;
			lf->w = n->w;
			lf->h = n->h;
			lf->pitch = v210_pitch_bytes(n->w);
			lf->rc = idx;
			return PB_OK;
		}
		if (n->kind == N_LEAF_RGBA) {
			lf->kind = pb::LEAF_RGBA_F32;
			lf->ptr = n->src->dev;
			lf->w = n->w;
			lf->h = n->h;
			lf->pitch = n->w * 16;
			return PB_OK;
		}
		if (n->kind == N_LEAF_PACKED) {   // rgba8 / bgra8 / planar 4:2:2 / 4:2:0 sources, read in place by the fused kernel
			int idx;
			if (rc_index(n->rc, &idx)) return 1;
			lf->kind = n->leaf_kind;
			lf->ptr = n->src->dev;
			lf->ptr_u = n->src_u ? n->src_u->dev : nullptr;
			lf->ptr_v = n->src_v ? n->src_v->dev : nullptr;
			lf->w = n->w;
			lf->h = n->h;
			lf->rc = idx;
			return PB_OK;
		}
		return 1;
	}

	int as_rgba_leaf(const NodeP &n, pb::Leaf *lf) {
		const void *p;
		int r = materialise_node(c, n, &p);
		if (r) return r;
		lf->kind = pb::LEAF_RGBA_F32;
		lf->ptr = p;
		lf->w = n->w;
		lf->h = n->h;
		lf->pitch = n->w * 16;
		return PB_OK;
	}

	int leaf_spec(const NodeP &n, pb::Leaf *lf) {
		memset(lf, 0, sizeof *lf);
		keep.push_back(n);
		if (0 == set_basic_leaf(n, lf)) return PB_OK;
		if (n->kind == N_TRANSFORM) {
			const NodeP &child = n->in[0];
			if (0 != set_basic_leaf(child, lf)) {
				int r = as_rgba_leaf(child, lf);
				if (r) return r;
			}
			lf->has_xf = 1;
			lf->xf_w = n->w;
			lf->xf_h = n->h;
			memcpy(lf->m, n->mat, sizeof lf->m);
			if (n->lanczos) return attach_lanczos(c, lf, n->lanczos);
			return PB_OK;
		}
		return as_rgba_leaf(n, lf);
	}

	int layer_spec(const NodeP &n, pb::Layer *ly) {
		memset(ly, 0, sizeof *ly);
		int r;
		if (n->kind == N_DISSOLVE) {
			ly->kind = pb::LAYER_DISSOLVE;
			ly->mix = n->mix;
			if ((r = leaf_spec(n->in[0], &ly->a))) return r;
			return leaf_spec(n->in[1], &ly->b);
		}
		if (n->kind == N_WIPE_MASK) {
			ly->kind = pb::LAYER_WIPE_MASK;
			if ((r = leaf_spec(n->in[0], &ly->a))) return r;
			if ((r = leaf_spec(n->in[1], &ly->b))) return r;
			return leaf_spec(n->in[2], &ly->mask);
		}
		ly->kind = pb::LAYER_DIRECT;
		return leaf_spec(n, &ly->a);
	}

	int compile(const NodeP &root) {
		memset(&d, 0, sizeof d);
		d.out_w = root->w;
		d.out_h = root->h;
		if (root->kind == N_COMBINE) {
			std::vector<NodeP> layers = root->in;
			// combine_N with N > kMaxLayers: fold the bottom layers first
			while ((int)layers.size() > pb::kMaxLayers) {
				auto sub = std::make_shared<Node>();
				sub->kind = N_COMBINE;
				sub->ctx = c;
				sub->w = root->w;
				sub->h = root->h;
				sub->in.assign(layers.begin(), layers.begin() + pb::kMaxLayers);
				layers.erase(layers.begin(), layers.begin() + pb::kMaxLayers);
				layers.insert(layers.begin(), sub);
			}
			d.n_layers = (int)layers.size();
			for (int i = 0; i < d.n_layers; ++i) {
				int r = layer_spec(layers[i], &d.layers[i]);
				if (r) return r;
			}
		} else {
			d.n_layers = 1;
			int r = layer_spec(root, &d.layers[0]);
			if (r) return r;
		}
		return PB_OK;
	}
};


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
		if (f.version == lut->version) {
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
	for (size_t i = 0; i < c->lut_tables.size(); ++i)
		if (c->lut_tables[i].hash == res[0].hash) table = (int)i;
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
	c->lut_fits.push_back({lut->version, table});
	*table_out = table;
	return PB_OK;
}

int lut_table_by_raw(pb_ctx *c, const float *raw) {
	for (size_t i = 0; i < c->lut_tables.size(); ++i)
		if (c->lut_tables[i].raw == raw) return (int)i;
	return -1;
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
	const bool planar_sink = d.sink == pb::SINK_YUV422P10 || d.sink == pb::SINK_YUV422P8 || d.sink == pb::SINK_YUV420P || d.sink == pb::SINK_NV12 ||
	                         d.sink == pb::SINK_RGBA8 || d.sink == pb::SINK_BGRA8;
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
			const bool rgba = lf.kind == pb::LEAF_RGBA8 || lf.kind == pb::LEAF_BGRA8 || lf.kind == pb::LEAF_RGBA_F32;
			if (lf.kind == pb::LEAF_RGBA_F32) any_f32 = true;
			if (!(ycc || rgba) || lf.w < 6 || lf.lz_tx) return 0;
			if (lf.kind != pb::LEAF_V210 || lf.w % 6 != 0) any_planar = true;   // general load path (formats, partial last groups)
			if (rgba) any_rgba = true;
			else rc_ycc[lf.rc] = true;
			if (lf.has_xf) {
				for (float v : lf.m)
					if (!(v == v) || v > 1e30f || v < -1e30f) return 0;
				if (lf.m[1] != 0.0f || lf.m[3] != 0.0f) return 0;   // rotation / shear
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
	d.n_strips = (d.out_w / 6 + d.strip_groups - 1) / d.strip_groups;
	if (d.n_strips > pb::kMaxStrips) return 0;
	std::shared_ptr<const pb_ctx::SampleTab::Opq> opq[3 * pb::kMaxLayers], tab_of[3 * pb::kMaxLayers];
	bool big_rows = false;
	for (int i = 0; i < n_leaves; ++i) {
		pb_ctx::SampleTab *t;
		int fits = 0;
		int r = get_tabs(c, *leaves[i], d.out_w, d.out_h, d.strip_groups, &t, &fits);
		if (r) return r;
		if (!fits) return 0;
		const bool leaf_rgba = leaves[i]->kind == pb::LEAF_RGBA8 || leaves[i]->kind == pb::LEAF_BGRA8 || leaves[i]->kind == pb::LEAF_RGBA_F32;
		if (leaf_rgba && fits != 1) return 0;   // four planes: 32 source groups per row at most
		if (fits == 2 || leaf_rgba) big_rows = true;
		opq[i] = leaf_rgba ? nullptr : t->opq;   // the alpha of an rgba8 leaf is data: never certifiably opaque
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
	bool cull = !(c->flags & PB_CTX_NO_CULL) && !any_f32;   // (an RGBA-f32 frame may hold NaN / inf: NaN * 0 != 0)
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
	const int wt = lut_table_by_raw(c, d.wc.lut);
	if (wt < 0 || !c->lut_tables[wt].unit_range) return 0;
	// (and, being inside the code range, need no saturation: the encoder drops the clamp of convert_ushort_sat_rte)
	const double code_max = (d.sink == pb::SINK_V210 || d.sink == pb::SINK_YUV422P10) ? 1023.0 : 255.0;
	for (int row = 0; row < 3; ++row) {
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
	d.wc.lut_slot = slot_of(wt);
	if (d.wc.lut_slot < 0) all_d8 = false;
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
	const bool plain_tables = all_d8 && n_slots <= 2 && d.sparse_cm && d.wc.lut_slot >= 0 && c->lut_tables[slots[d.wc.lut_slot]].lp.affine == 1;
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
	d.direct_mode = d.n_ops == 1 && d.layers[0].kind == pb::LAYER_DIRECT && !d.layers[0].a.has_xf && d.layers[0].a.kind == pb::LEAF_V210 &&
	                d.sink == pb::SINK_V210 && d.out_w % 48 == 0 && all_d8 && n_slots <= 2 && d.sparse_cm && !any_planar && !big_rows &&
	                d.rc[0].lut_slot >= 0 && c->lut_tables[slots[d.rc[0].lut_slot]].lp.affine != 1 &&
	                d.wc.lut_slot >= 0 && c->lut_tables[slots[d.wc.lut_slot]].lp.affine == 1 && !(c->flags & PB_CTX_NO_DIRECT);
	if (big_rows) {   // 2 x 64 KiB of tables + 20 x 4.5 KiB of rows is what an SM holds
		if (d.n_luts > 2) return 0;
		any_planar = true;
	}
	if (any_rgba && d.n_luts == 0) return 0;   // rgba8 leaves ride on the big-row variants, which exist for shared-memory tables
	d.any_planar = any_planar;
	if (any_planar && !(d.n_luts > 0 && d.sparse_cm)) return 0;   // planar variants exist for the common configuration only
	for (int i = 0; i < d.n_luts; ++i) {
		d.luts[i].d8 = c->lut_tables[slots[i]].d8;
		d.luts[i].lp = c->lut_tables[slots[i]].lp;
	}
	if (d.n_luts) d.wlp = d.luts[d.wc.lut_slot].lp;
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
int launch_compiled(pb_ctx *c, cudaStream_t s, const pb::FusedDesc &d_in, bool march, void *out_rgba) {
	pb::FusedDesc bg_copy;
	const pb::FusedDesc *dp = &d_in;
	if (march && d_in.bg_single) {   // the second phase claims its items from a counter that is never reset (pb_march.cu)
		if (!c->bg_counter) {
			CU(cudaMalloc(&c->bg_counter, sizeof(unsigned int)));
			CU(cudaMemsetAsync(c->bg_counter, 0, sizeof(unsigned int), s));
			c->bg_next_base = 0;
		}
		bg_copy = d_in;
		bg_copy.bg_counter = c->bg_counter;
		bg_copy.bg_base = c->bg_next_base;
		const int n_lines = d_in.out_h;
		const int pairs = (d_in.out_w / 6 + d_in.single_strip_groups - 1) / d_in.single_strip_groups;
		const unsigned total = (unsigned)pairs * (unsigned)((n_lines + d_in.single_lines - 1) / d_in.single_lines);
		const unsigned items1 = (unsigned)n_lines * (unsigned)d_in.n_strips;   // the grid launch_fused_march picks (phase-1 items)
		const unsigned grid = std::max(1u, std::min((unsigned)c->prop.multiProcessorCount, (items1 + pb::kMarchWarps - 1) / pb::kMarchWarps));
		c->bg_next_base += total + grid * pb::kMarchWarps;
		dp = &bg_copy;
	}
	const pb::FusedDesc &d = *dp;
	cudaError_t e = march ? pb::launch_fused_march(s, d, c->prop.multiProcessorCount) : pb::launch_fused(s, d, out_rgba);
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
	if (!out_rgba) {
		int r = prepare_march(c, d);
		if (r < 0) return r;
		march = r == 1;
	}
	int r = launch_compiled(c, s, d, march, out_rgba);
	if (r) return r;
	if (march_out) *march_out = march;
	return PB_OK;
}

void record_launch(pb_ctx *c, const Compiler &cc, void *out_rgba, pb_buf *out_buf, bool march = false);
// further destination planes of the launch just recorded: a replayable chain keeps them alive too
void record_extra_output(pb_ctx *c, pb_buf *b) {
	if (!c->recording || c->recording->items.empty()) return;
	b->refs.fetch_add(1);
	c->recording->items.back().keep.push_back(std::shared_ptr<void>(b, [](void *p) {
		pb_buf *bb = static_cast<pb_buf *>(p);
		std::lock_guard<std::recursive_mutex> lk(bb->ctx->mu);
		buf_release_locked(bb);
	}));
}

// A Writer other than v210 whose input is still an expression: evaluate the layer graph inside the writer (one launch,
// no RGBA-f32 frame).  `outs` are the destination planes (addref'd by a recorded chain through outs[0] only: the
// recorder keeps the expression nodes; planes stay alive because the caller's job holds them until the request ends).
int launch_fused_sink(pb_ctx *c, cudaStream_t s, pb_buf *in, int sink, pb_buf *const *outs, int n_outs, int interlace, const pb::WriteConsts &wc,
                      int W, int H) {
	Compiler cc{c};
	int r = cc.compile(in->expr);
	if (r) return r;
	if (cc.d.out_w != W || cc.d.out_h != H) return fail(PB_ERR_ARG, "writer is %dx%d but input image is %dx%d", W, H, cc.d.out_w, cc.d.out_h);
	cc.d.wc = wc;
	cc.d.interlace = interlace;
	cc.d.sink = sink;
	cc.d.out = outs[0]->dev;
	cc.d.out_u = n_outs > 1 ? outs[1]->dev : nullptr;
	cc.d.out_v = n_outs > 2 ? outs[2]->dev : nullptr;
	bool march = false;
	if ((r = launch_desc(c, s, cc.d, nullptr, &march))) return r;
	c->stats.fused_launches++;
	if (march) c->stats.march_launches++;
	record_launch(c, cc, nullptr, outs[0], march);
	for (int i = 1; i < n_outs; ++i) record_extra_output(c, outs[i]);
	return PB_OK;
}

void record_launch(pb_ctx *c, const Compiler &cc, void *out_rgba, pb_buf *out_buf, bool march) {
	if (!c->recording) return;
	pb_chain::Item it;
	it.d = cc.d;
	it.march = march;
	it.out_rgba = out_rgba;
	for (const auto &k : cc.keep) it.keep.push_back(std::static_pointer_cast<void>(k));
	it.out_buf = out_buf;
	if (out_buf) out_buf->refs.fetch_add(1);
	c->recording->items.push_back(std::move(it));
}

// write node n as RGBA-f32 into HBM (cached on the node)
int materialise_node(pb_ctx *c, const NodeP &n, const void **dev_out) {
	if (n->kind == N_LEAF_RGBA) {
		*dev_out = n->src->dev;
		return PB_OK;
	}
	if (!n->mat_dev) {
		void *p;
		CU(c->pool.dev_get((size_t)n->w * n->h * 16, &p));
		Compiler cc{c};
		int r = cc.compile(n);
		if (r) {
			c->pool.dev_put((size_t)n->w * n->h * 16, p);
			return r;
		}
		cudaError_t e = pb::launch_fused(c->q[PB_QUEUE_PROCESS], cc.d, p);
		if (e != cudaSuccess) {
			c->pool.dev_put((size_t)n->w * n->h * 16, p);
			return fail(PB_ERR_CUDA, "fused materialise launch: %s", cudaGetErrorString(e));
		}
		c->stats.kernel_launches++;
		c->stats.fused_launches++;
		c->stats.materialised++;
		n->mat_dev = p;
		cc.keep.push_back(n);   // a recorded chain must keep the node (and its mat_dev) alive
		record_launch(c, cc, p, nullptr);
	}
	*dev_out = n->mat_dev;
	return PB_OK;
}

// make a deferred buffer real
int materialise_buf(pb_buf *b) {
	if (!b->expr) return PB_OK;
	pb_ctx *c = b->ctx;
	NodeP n = b->expr;
	int r = ensure_dev(b);
	if (r) return r;
	if (n->kind == N_LEAF_RGBA) {
		CU(cudaMemcpyAsync(b->dev, n->src->dev, b->bytes, cudaMemcpyDeviceToDevice, c->q[PB_QUEUE_PROCESS]));
	} else if (n->mat_dev) {
		CU(cudaMemcpyAsync(b->dev, n->mat_dev, b->bytes, cudaMemcpyDeviceToDevice, c->q[PB_QUEUE_PROCESS]));
	} else {
		Compiler cc{c};
		if ((r = cc.compile(n))) return r;
		cudaError_t e = pb::launch_fused(c->q[PB_QUEUE_PROCESS], cc.d, b->dev);
		if (e != cudaSuccess) return fail(PB_ERR_CUDA, "fused materialise launch: %s", cudaGetErrorString(e));
		c->stats.kernel_launches++;
		c->stats.fused_launches++;
		c->stats.materialised++;
		record_launch(c, cc, b->dev, b);
	}
	b->expr.reset();
	return PB_OK;
}

// RGBA input that must be real memory for a stand-alone kernel
int real_input(pb_buf *b, const void **p) {
	int r = materialise_buf(b);
	if (r) return r;
	if ((r = flush_host(b, b->ctx->q[PB_QUEUE_PROCESS]))) return r;
	if (!b->dev) {
		// Never written.  The reference reads whatever the fresh SVM allocation holds (this
		// really happens: Yadif runs with a `next` frame whose ToRGBA job is still queued,
		// yadif.ts:88-113 vs macadamProducer.ts:193-227).  We define it as zeros.
		if ((r = ensure_dev(b))) return r;
		CU(cudaMemsetAsync(b->dev, 0, b->bytes, b->ctx->q[PB_QUEUE_PROCESS]));
	}
	*p = b->dev;
	return PB_OK;
}

int real_output(pb_buf *b, void **p) {
	b->expr.reset();
	b->host_dirty = false;
	b->version = ++b->ctx->version_counter;
	int r = ensure_dev(b);
	if (r) return r;
	*p = b->dev;
	return PB_OK;
}

NodeP new_node(pb_ctx *c, NodeKind k, int w, int h) {
	auto n = std::make_shared<Node>();
	n->kind = k;
	n->ctx = c;
	n->w = w;
	n->h = h;
	return n;
}

void set_deferred(pb_buf *out, NodeP n) {
	pb_ctx *c = out->ctx;
	if (out->dev && !out->dev_external) {   // drop stale storage: the frame lives in the expression now
		c->pool.dev_put(out->bytes, out->dev);
		out->dev = nullptr;
	}
	out->host_dirty = false;
	out->expr = std::move(n);
	c->stats.deferred_nodes++;
}

int check_image(pb_buf *b, int w, int h, const char *what) {
	if (b->bytes < (size_t)w * h * 16) return fail(PB_ERR_ARG, "%s buffer too small for %dx%d RGBA-f32", what, w, h);
	return PB_OK;
}

int run_locked(pb_ctx *c, pb_prog *g, const pb_param *p, int n, cudaStream_t s) {
	const bool defer = (c->flags & PB_CTX_DEFER) != 0;
	const int W = g->w, H = g->h;
	int r;
	bool fused_launch = false;
	cudaError_t e = cudaSuccess;
	switch (g->op) {
		case PB_OP_V210_READ: {
			pb_buf *in, *out, *lut;
			pb::ReadConsts rc;
			if ((r = need_buf(p, n, "input", &in)) || (r = need_buf(p, n, "output", &out))) return r;
			if ((r = make_read_consts(c, p, n, true, &rc, &lut))) return r;
			if (in->bytes < (size_t)v210_pitch_bytes(W) * H) return fail(PB_ERR_ARG, "v210 input buffer too small");
			if ((r = check_image(out, W, H, "output"))) return r;
			if ((r = flush_host(in, s))) return r;
			if (!in->dev) return fail(PB_ERR_STATE, "v210 input has no contents");
			if (defer) {
				NodeP nd = new_node(c, N_LEAF_V210, W, H);
				nd->src = in;
				in->refs.fetch_add(1);
				nd->lut_buf = lut;
				lut->refs.fetch_add(1);
				nd->rc = rc;
				out->w = W;
				out->h = H;
				set_deferred(out, nd);
				return PB_OK;
			}
			void *o;
			if ((r = real_output(out, &o))) return r;
			e = pb::launch_v210_read(s, in->dev, o, W, H, rc);
			break;
		}
		case PB_OP_RGBA8_READ:
		case PB_OP_BGRA8_READ: {
			pb_buf *in, *out, *lut;
			pb::ReadConsts rc;
			if ((r = need_buf(p, n, "input", &in)) || (r = need_buf(p, n, "output", &out))) return r;
			if ((r = make_read_consts(c, p, n, false, &rc, &lut))) return r;
			if (in->bytes < (size_t)W * H * 4) return fail(PB_ERR_ARG, "rgba8 input buffer too small");
			if ((r = check_image(out, W, H, "output"))) return r;
			if ((r = flush_host(in, s))) return r;
			if (!in->dev) return fail(PB_ERR_STATE, "rgba8 input has no contents");
			if (defer) {
				NodeP nd = new_node(c, N_LEAF_PACKED, W, H);
				nd->leaf_kind = g->op == PB_OP_BGRA8_READ ? pb::LEAF_BGRA8 : pb::LEAF_RGBA8;
				nd->src = in;
				in->refs.fetch_add(1);
				nd->lut_buf = lut;
				lut->refs.fetch_add(1);
				nd->rc = rc;
				out->w = W;
				out->h = H;
				set_deferred(out, nd);
				return PB_OK;
			}
			void *o;
			if ((r = real_output(out, &o))) return r;
			e = pb::launch_rgba8_read(s, in->dev, o, W, H, g->op == PB_OP_BGRA8_READ, rc);
			break;
		}
		case PB_OP_V210_WRITE:
		case PB_OP_RGBA8_WRITE:
		case PB_OP_BGRA8_WRITE: {
			pb_buf *in, *out;
			pb::WriteConsts wc;
			double il = 0;
			const bool v210 = g->op == PB_OP_V210_WRITE;
			if ((r = need_buf(p, n, "input", &in)) || (r = need_buf(p, n, "output", &out))) return r;
			if ((r = make_write_consts(c, p, n, v210, &wc))) return r;
			if (find(p, n, "interlace") && (r = need_num(p, n, "interlace", &il))) return r;
			const int interlace = (int)il;
			if (interlace != 0 && interlace != 1 && interlace != 3) return fail(PB_ERR_ARG, "interlace must be 0, 1 or 3");
			const size_t need = v210 ? (size_t)v210_pitch_bytes(W) * H : (size_t)W * H * 4;
			if (out->bytes < need) return fail(PB_ERR_ARG, "packed output buffer too small");
			if (in->w && (in->w != W || in->h != H)) return fail(PB_ERR_ARG, "writer is %dx%d but input image is %dx%d", W, H, in->w, in->h);
			out->expr.reset();
			// a field write must keep the other field's lines: push pending host contents first
			if (interlace != 0 && (r = flush_host(out, s))) return r;
			out->host_dirty = false;
			if ((r = ensure_dev(out))) return r;
			if (v210 && in->expr) {
				Compiler cc{c};
				if ((r = cc.compile(in->expr))) return r;
				cc.d.wc = wc;
				cc.d.interlace = interlace;
				cc.d.out = out->dev;
				cc.d.out_pitch = v210_pitch_bytes(W);
				bool march = false;
				if ((r = launch_desc(c, s, cc.d, nullptr, &march))) return r;
				c->stats.fused_launches++;
				if (march) c->stats.march_launches++;
				record_launch(c, cc, nullptr, out, march);
				fused_launch = true;
				break;
			}
			if (!v210 && in->expr) {   // ScreenConsumer path: the layer graph is evaluated inside the rgba8 / bgra8 writer
				pb_buf *outs[1] = {out};
				if ((r = launch_fused_sink(c, s, in, g->op == PB_OP_BGRA8_WRITE ? pb::SINK_BGRA8 : pb::SINK_RGBA8, outs, 1, interlace, wc, W, H))) return r;
				fused_launch = true;
				break;
			}
			const void *src;
			if ((r = real_input(in, &src))) return r;
			if (v210) e = pb::launch_v210_write(s, src, out->dev, W, H, interlace, wc);
			else e = pb::launch_rgba8_write(s, src, out->dev, W, H, interlace, g->op == PB_OP_BGRA8_WRITE, wc);
			break;
		}
		case PB_OP_YUV422P10_READ:
		case PB_OP_YUV422P8_READ: {
			const int bits = g->op == PB_OP_YUV422P8_READ ? 8 : 10;
			pb_buf *iy, *iu, *iv, *out, *lut;
			pb::ReadConsts rc;
			if ((r = need_buf(p, n, "inputY", &iy)) || (r = need_buf(p, n, "inputU", &iu)) || (r = need_buf(p, n, "inputV", &iv)) ||
			    (r = need_buf(p, n, "output", &out)))
				return r;
			if ((r = make_read_consts(c, p, n, true, &rc, &lut))) return r;
			const size_t luma = (size_t)((W + 7) / 8 * 8) * (bits == 8 ? 1 : 2) * H;
			if (iy->bytes < luma || iu->bytes < luma / 2 || iv->bytes < luma / 2) return fail(PB_ERR_ARG, "yuv422p input plane too small");
			if ((r = check_image(out, W, H, "output"))) return r;
			if ((r = flush_host(iy, s)) || (r = flush_host(iu, s)) || (r = flush_host(iv, s))) return r;
			if (!iy->dev || !iu->dev || !iv->dev) return fail(PB_ERR_STATE, "yuv422p input has no contents");
			if (defer) {
				NodeP nd = new_node(c, N_LEAF_PACKED, W, H);
				nd->leaf_kind = bits == 8 ? pb::LEAF_YUV422P8 : pb::LEAF_YUV422P10;
				nd->src = iy; nd->src_u = iu; nd->src_v = iv;
				iy->refs.fetch_add(1); iu->refs.fetch_add(1); iv->refs.fetch_add(1);
				nd->lut_buf = lut;
				lut->refs.fetch_add(1);
				nd->rc = rc;
				out->w = W;
				out->h = H;
				set_deferred(out, nd);
				return PB_OK;
			}
			void *o;
			if ((r = real_output(out, &o))) return r;
			out->w = W;
			out->h = H;
			e = pb::launch_yuv422p_read(s, bits, iy->dev, iu->dev, iv->dev, o, W, H, rc);
			break;
		}
		case PB_OP_YUV422P10_WRITE:
		case PB_OP_YUV422P8_WRITE: {
			const int bits = g->op == PB_OP_YUV422P8_WRITE ? 8 : 10;
			pb_buf *in, *oy, *ou, *ov;
			pb::WriteConsts wc;
			double il = 0;
			if ((r = need_buf(p, n, "input", &in)) || (r = need_buf(p, n, "outputY", &oy)) || (r = need_buf(p, n, "outputU", &ou)) ||
			    (r = need_buf(p, n, "outputV", &ov)))
				return r;
			if ((r = make_write_consts(c, p, n, true, &wc))) return r;
			if (find(p, n, "interlace") && (r = need_num(p, n, "interlace", &il))) return r;
			const int interlace = (int)il;
			if (interlace != 0 && interlace != 1 && interlace != 3) return fail(PB_ERR_ARG, "interlace must be 0, 1 or 3");
			const size_t luma = (size_t)((W + 7) / 8 * 8) * (bits == 8 ? 1 : 2) * H;
			if (oy->bytes < luma || ou->bytes < luma / 2 || ov->bytes < luma / 2) return fail(PB_ERR_ARG, "yuv422p output plane too small");
			if (in->w && (in->w != W || in->h != H)) return fail(PB_ERR_ARG, "writer is %dx%d but input image is %dx%d", W, H, in->w, in->h);
			pb_buf *outs[3] = {oy, ou, ov};
			for (pb_buf *o : outs) {
				o->expr.reset();
				if (interlace != 0 && (r = flush_host(o, s))) return r;   // a field write keeps the other field's lines
				o->host_dirty = false;
				if ((r = ensure_dev(o))) return r;
				o->version = ++c->version_counter;
			}
			if (in->expr) {   // FFmpegConsumer path (yuv422p8): the layer graph is evaluated inside the planar writer
				if ((r = launch_fused_sink(c, s, in, bits == 8 ? pb::SINK_YUV422P8 : pb::SINK_YUV422P10, outs, 3, interlace, wc, W, H))) return r;
				fused_launch = true;
				break;
			}
			const void *src;
			if ((r = real_input(in, &src))) return r;
			e = pb::launch_yuv422p_write(s, bits, src, oy->dev, ou->dev, ov->dev, W, H, interlace, wc);
			break;
		}
		case PB_OP_YUV420P_READ:
		case PB_OP_NV12_READ: {
			const bool nv12 = g->op == PB_OP_NV12_READ;
			pb_buf *iy, *iu, *iv = nullptr, *out, *lut;
			pb::ReadConsts rc;
			if ((r = need_buf(p, n, "inputY", &iy)) || (r = need_buf(p, n, nv12 ? "inputC" : "inputU", &iu)) ||
			    (!nv12 && (r = need_buf(p, n, "inputV", &iv))) || (r = need_buf(p, n, "output", &out)))
				return r;
			if (H & 1) return fail(PB_ERR_ARG, "4:2:0 packers need an even height, found %d", H);   // the reference launches height / 2 work-groups
			if ((r = make_read_consts(c, p, n, true, &rc, &lut))) return r;
			const size_t luma = (size_t)((W + 7) / 8 * 8) * H;
			if (iy->bytes < luma || iu->bytes < (nv12 ? luma / 2 : luma / 4) || (iv && iv->bytes < luma / 4)) return fail(PB_ERR_ARG, "4:2:0 input plane too small");
			if ((r = check_image(out, W, H, "output"))) return r;
			if ((r = flush_host(iy, s)) || (r = flush_host(iu, s)) || (iv && (r = flush_host(iv, s)))) return r;
			if (!iy->dev || !iu->dev || (iv && !iv->dev)) return fail(PB_ERR_STATE, "4:2:0 input has no contents");
			if (defer) {
				NodeP nd = new_node(c, N_LEAF_PACKED, W, H);
				nd->leaf_kind = nv12 ? pb::LEAF_NV12 : pb::LEAF_YUV420P;
				nd->src = iy; nd->src_u = iu; nd->src_v = iv;
				iy->refs.fetch_add(1); iu->refs.fetch_add(1);
				if (iv) iv->refs.fetch_add(1);
				nd->lut_buf = lut;
				lut->refs.fetch_add(1);
				nd->rc = rc;
				out->w = W;
				out->h = H;
				set_deferred(out, nd);
				return PB_OK;
			}
			void *o;
			if ((r = real_output(out, &o))) return r;
			out->w = W;
			out->h = H;
			e = pb::launch_yuv420_read(s, nv12, iy->dev, iu->dev, iv ? iv->dev : nullptr, o, W, H, rc);
			break;
		}
		case PB_OP_YUV420P_WRITE:
		case PB_OP_NV12_WRITE: {
			const bool nv12 = g->op == PB_OP_NV12_WRITE;
			pb_buf *in, *oy, *ou, *ov = nullptr;
			pb::WriteConsts wc;
			double il = 0;
			if ((r = need_buf(p, n, "input", &in)) || (r = need_buf(p, n, "outputY", &oy)) || (r = need_buf(p, n, nv12 ? "outputC" : "outputU", &ou)) ||
			    (!nv12 && (r = need_buf(p, n, "outputV", &ov))))
				return r;
			if (H & 1) return fail(PB_ERR_ARG, "4:2:0 packers need an even height, found %d", H);
			if ((r = make_write_consts(c, p, n, true, &wc))) return r;
			if (find(p, n, "interlace") && (r = need_num(p, n, "interlace", &il))) return r;
			const int interlace = (int)il;
			if (interlace != 0 && interlace != 1 && interlace != 3) return fail(PB_ERR_ARG, "interlace must be 0, 1 or 3");
			const size_t luma = (size_t)((W + 7) / 8 * 8) * H;
			if (oy->bytes < luma || ou->bytes < (nv12 ? luma / 2 : luma / 4) || (ov && ov->bytes < luma / 4)) return fail(PB_ERR_ARG, "4:2:0 output plane too small");
			if (in->w && (in->w != W || in->h != H)) return fail(PB_ERR_ARG, "writer is %dx%d but input image is %dx%d", W, H, in->w, in->h);
			pb_buf *outs[3] = {oy, ou, ov};
			for (pb_buf *o : outs) {
				if (!o) continue;
				o->expr.reset();
				if (interlace != 0 && (r = flush_host(o, s))) return r;   // a field write keeps the other field's luma lines
				o->host_dirty = false;
				if ((r = ensure_dev(o))) return r;
				o->version = ++c->version_counter;
			}
			if (in->expr) {
				if ((r = launch_fused_sink(c, s, in, nv12 ? pb::SINK_NV12 : pb::SINK_YUV420P, outs, nv12 ? 2 : 3, interlace, wc, W, H))) return r;
				fused_launch = true;
				break;
			}
			const void *src;
			if ((r = real_input(in, &src))) return r;
			e = pb::launch_yuv420_write(s, nv12, src, oy->dev, ou->dev, ov ? ov->dev : nullptr, W, H, interlace, wc);
			break;
		}
		case PB_OP_COMBINE: {
			pb_buf *out, *ins[64];
			int cnt = 0;
			char name[16];
			if ((r = need_buf(p, n, "output", &out))) return r;
			for (; cnt < 64; ++cnt) {
				snprintf(name, sizeof name, "l%dIn", cnt);
				if (!find(p, n, name)) break;
				if ((r = need_buf(p, n, name, &ins[cnt]))) return r;
			}
			if (cnt < 2) return fail(PB_ERR_ARG, "combine needs at least l0In and l1In");
			if ((r = check_image(out, W, H, "output"))) return r;
			if (defer) {
				NodeP nd = new_node(c, N_COMBINE, W, H);
				for (int i = 0; i < cnt; ++i) {
					NodeP x;
					if ((r = input_expr(ins[i], &x))) return r;
					if (x->w != W || x->h != H) return fail(PB_ERR_ARG, "combine layer %d is %dx%d, expected %dx%d", i, x->w, x->h, W, H);
					nd->in.push_back(x);
				}
				out->w = W;
				out->h = H;
				set_deferred(out, nd);
				return PB_OK;
			}
			if (cnt > pb::kMaxLayers) return fail(PB_ERR_ARG, "eager combine supports at most %d layers", pb::kMaxLayers);
			const void *src[pb::kMaxLayers];
			for (int i = 0; i < cnt; ++i)
				if ((r = real_input(ins[i], &src[i]))) return r;
			void *o;
			if ((r = real_output(out, &o))) return r;
			e = pb::launch_combine(s, src, cnt, o, W, H);
			break;
		}
		case PB_OP_DISSOLVE:
		case PB_OP_MIX: {
			pb_buf *in0, *in1, *out;
			double mix;
			if ((r = need_buf(p, n, "input0", &in0)) || (r = need_buf(p, n, "input1", &in1)) || (r = need_buf(p, n, "output", &out)) ||
			    (r = need_num(p, n, "mix", &mix)))
				return r;
			if ((r = check_image(out, W, H, "output"))) return r;
			if (defer) {
				NodeP nd = new_node(c, N_DISSOLVE, W, H);
				NodeP a, b;
				if ((r = input_expr(in0, &a)) || (r = input_expr(in1, &b))) return r;
				if (a->w != W || a->h != H || b->w != W || b->h != H) return fail(PB_ERR_ARG, "dissolve inputs must be %dx%d", W, H);
				nd->in = {a, b};
				nd->mix = (float)mix;
				out->w = W;
				out->h = H;
				set_deferred(out, nd);
				return PB_OK;
			}
			const void *a, *b;
			void *o;
			if ((r = real_input(in0, &a)) || (r = real_input(in1, &b)) || (r = real_output(out, &o))) return r;
			e = pb::launch_dissolve(s, a, b, (float)mix, o, W, H);
			break;
		}
		case PB_OP_WIPE_MASK: {
			pb_buf *in0, *in1, *mask, *out;
			if ((r = need_buf(p, n, "input0", &in0)) || (r = need_buf(p, n, "input1", &in1)) || (r = need_buf(p, n, "maskIn", &mask)) ||
			    (r = need_buf(p, n, "output", &out)))
				return r;
			if ((r = check_image(out, W, H, "output"))) return r;
			if (defer) {
				NodeP nd = new_node(c, N_WIPE_MASK, W, H);
				NodeP a, b, m;
				if ((r = input_expr(in0, &a)) || (r = input_expr(in1, &b)) || (r = input_expr(mask, &m))) return r;
				if (a->w != W || a->h != H || b->w != W || b->h != H || m->w != W || m->h != H)
					return fail(PB_ERR_ARG, "wipe inputs must be %dx%d", W, H);
				nd->in = {a, b, m};
				out->w = W;
				out->h = H;
				set_deferred(out, nd);
				return PB_OK;
			}
			const void *a, *b, *m;
			void *o;
			if ((r = real_input(in0, &a)) || (r = real_input(in1, &b)) || (r = real_input(mask, &m)) || (r = real_output(out, &o))) return r;
			e = pb::launch_wipe_mask(s, a, b, m, o, W, H);
			break;
		}
		case PB_OP_WIPE: {
			pb_buf *in0, *in1, *out;
			double wipe;
			if ((r = need_buf(p, n, "input0", &in0)) || (r = need_buf(p, n, "input1", &in1)) || (r = need_buf(p, n, "output", &out)) ||
			    (r = need_num(p, n, "wipe", &wipe)))
				return r;
			if ((r = check_image(out, W, H, "output"))) return r;
			const void *a, *b;
			void *o;
			if ((r = real_input(in0, &a)) || (r = real_input(in1, &b)) || (r = real_output(out, &o))) return r;
			e = pb::launch_wipe(s, a, b, (float)wipe, o, W, H);
			break;
		}
		case PB_OP_TRANSFORM: {
			pb_buf *in, *out, *mb;
			float m9[9];
			if ((r = need_buf(p, n, "input", &in)) || (r = need_buf(p, n, "output", &out)) || (r = need_buf(p, n, "transformMatrix", &mb)))
				return r;
			if ((r = host_floats(mb, 9, m9, "transformMatrix"))) return r;
			if ((r = check_image(out, W, H, "output"))) return r;
			double lanczos = 0;   // extension (not in the reference): Transform.run({..., filter: 'lanczos3'}) binds lanczos = 3
			if (find(p, n, "lanczos") && (r = need_num(p, n, "lanczos", &lanczos))) return r;
			if (defer || lanczos != 0) {
				NodeP child;
				if ((r = input_expr(in, &child))) return r;
				NodeP nd = new_node(c, N_TRANSFORM, W, H);
				nd->in = {child};
				memcpy(nd->mat, m9, sizeof nd->mat);
				nd->lanczos = (int)lanczos;
				out->w = W;
				out->h = H;
				set_deferred(out, nd);
				if (!defer) {   // eager mode: evaluate the one-node expression now (the generic kernel holds the only Lanczos sampler)
					if ((r = materialise_buf(out))) return r;
					return PB_OK;
				}
				return PB_OK;
			}
			if (in->w <= 0 || in->h <= 0) return fail(PB_ERR_ARG, "transform input was created without imageDims");
			const void *src;
			void *o;
			if ((r = real_input(in, &src)) || (r = real_output(out, &o))) return r;
			e = pb::launch_transform(s, src, in->w, in->h, m9, o, W, H);
			break;
		}
		case PB_OP_RESIZE: {
			pb_buf *in, *out, *fb;
			double scale, ox, oy;
			float flip[4];
			if ((r = need_buf(p, n, "input", &in)) || (r = need_buf(p, n, "output", &out)) || (r = need_buf(p, n, "flip", &fb)) ||
			    (r = need_num(p, n, "scale", &scale)) || (r = need_num(p, n, "offsetX", &ox)) || (r = need_num(p, n, "offsetY", &oy)))
				return r;
			if ((r = host_floats(fb, 4, flip, "flip"))) return r;
			if ((r = check_image(out, W, H, "output"))) return r;
			if (in->w <= 0 || in->h <= 0) return fail(PB_ERR_ARG, "resize input was created without imageDims");
			const void *src;
			void *o;
			if ((r = real_input(in, &src)) || (r = real_output(out, &o))) return r;
			e = pb::launch_resize(s, src, in->w, in->h, (float)scale, (float)ox, (float)oy, flip, o, W, H);
			break;
		}
		case PB_OP_YADIF: {
			pb_buf *prev, *cur, *next, *out;
			double parity, tff, skip;
			if ((r = need_buf(p, n, "prev", &prev)) || (r = need_buf(p, n, "cur", &cur)) || (r = need_buf(p, n, "next", &next)) ||
			    (r = need_buf(p, n, "output", &out)) || (r = need_num(p, n, "parity", &parity)) || (r = need_num(p, n, "tff", &tff)) ||
			    (r = need_num(p, n, "skipSpatial", &skip)))
				return r;
			if ((r = check_image(out, W, H, "output"))) return r;
			const void *a, *b, *d;
			void *o;
			if ((r = real_input(prev, &a)) || (r = real_input(cur, &b)) || (r = real_input(next, &d)) || (r = real_output(out, &o))) return r;
			out->w = W;
			out->h = H;
			e = pb::launch_yadif(s, a, b, d, (int)parity, tff != 0, skip != 0, o, W, H);
			break;
		}
		default:
			return fail(PB_ERR_ARG, "unknown op %d", g->op);
	}
	if (e != cudaSuccess) return fail(PB_ERR_CUDA, "kernel launch (op %d): %s", g->op, cudaGetErrorString(e));
	c->stats.kernel_launches++;
	if (c->recording && !fused_launch)
		c->recording->complete = false;   // a stand-alone kernel ran: the chain cannot reproduce it
	return PB_OK;
}

}  // namespace

// ---- C ABI ------------------------------------------------------------------------------------
extern "C" {

const char *pb_last_error(void) { return g_err.c_str(); }
const char *pb_version(void) { return "phaneron_b200 0.1 (sm_100a)"; }

int pb_ctx_create(int gpu_index, unsigned flags, pb_ctx **out) {
	if (!out) return fail(PB_ERR_ARG, "out is null");
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		return fail(PB_ERR_NO_DEVICE, "no CUDA device (%s); phaneron_b200 has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
	if (gpu_index < 0 || gpu_index >= count) return fail(PB_ERR_ARG, "gpu_index %d out of range (%d devices)", gpu_index, count);
	CU(cudaSetDevice(gpu_index));
	auto *c = new pb_ctx;
	c->dev = gpu_index;
	c->flags = flags;
	c->allow_march = !(flags & PB_CTX_NO_MARCH);
	CU(cudaGetDeviceProperties(&c->prop, gpu_index));
	for (auto &q : c->q) CU(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
	CU(cudaEventCreate(&c->ev0));
	CU(cudaEventCreate(&c->ev1));
	CU(cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming));
	*out = c;
	return PB_OK;
}

int pb_ctx_destroy(pb_ctx *c) {
	if (!c) return PB_OK;
	cudaSetDevice(c->dev);
	cudaDeviceSynchronize();
	c->pool.destroy();
	for (auto &t : c->tabs) cudaFree(t.dev);
	for (auto &t : c->lut_tables) {
		cudaFree(t.raw);
		cudaFree(t.d8);
	}
	for (auto &lo : c->line_ops) cudaFree(lo.dev);
	for (cudaEvent_t e : c->copy_events) cudaEventDestroy(e);
	for (auto &e : c->lanczos_tabs) cudaFree(e.dev);
	for (auto &e : c->line_pairs) cudaFree(e.dev);
	cudaFree(c->bg_counter);
	cudaFree(c->lut_cands_dev);
	cudaFree(c->lut_res_dev);
	cudaFree(c->lut_scratch);
	for (auto &q : c->q) cudaStreamDestroy(q);
	cudaEventDestroy(c->ev0);
	cudaEventDestroy(c->ev1);
	cudaEventDestroy(c->ev_x);
	delete c;
	return PB_OK;
}

int pb_ctx_info(pb_ctx *c, char *buf, size_t n) {
	if (!c || !buf) return fail(PB_ERR_ARG, "null argument");
	snprintf(buf, n,
	         "{\"vendor\":\"NVIDIA Corporation\",\"name\":\"phaneron_b200\",\"version\":\"%s\",\"devices\":[{\"type\":\"GPU\","
	         "\"name\":\"%s\",\"computeCapability\":\"%d.%d\",\"multiProcessorCount\":%d,\"totalGlobalMem\":%zu}]}",
	         pb_version(), c->prop.name, c->prop.major, c->prop.minor, c->prop.multiProcessorCount, (size_t)c->prop.totalGlobalMem);
	return PB_OK;
}

int pb_ctx_stats(pb_ctx *c, pb_stats *out) {
	if (!c || !out) return fail(PB_ERR_ARG, "null argument");
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	*out = c->stats;
	out->dev_bytes_live = c->pool.dev_live;
	out->dev_bytes_pooled = c->pool.dev_pooled;
	out->lut_tables = c->lut_tables.size();
	out->lut_tables_d8 = 0;
	out->lut_tables_poly = 0;
	for (const auto &t : c->lut_tables) {
		out->lut_tables_d8 += t.d8 ? 1 : 0;
		out->lut_tables_poly += (t.d8 && t.lp.affine == 2) ? 1 : 0;
	}
	return PB_OK;
}

int pb_ctx_set_flags(pb_ctx *c, unsigned flags) {
	if (!c) return fail(PB_ERR_ARG, "null context");
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	c->flags = flags;
	c->allow_march = !(flags & PB_CTX_NO_MARCH);
	return PB_OK;
}

int pb_buf_create(pb_ctx *c, size_t bytes, int dir, int svm, int image_w, int image_h, const char *owner, pb_buf **out) {
	if (!c || !out) return fail(PB_ERR_ARG, "null argument");
	if (bytes == 0) return fail(PB_ERR_ARG, "zero-sized buffer");
	if (image_w < 0 || image_h < 0) return fail(PB_ERR_ARG, "negative image dimensions");
	if (image_w && (size_t)image_w * image_h * 16 > bytes) return fail(PB_ERR_ARG, "imageDims %dx%d exceed %zu bytes", image_w, image_h, bytes);
	auto *b = new pb_buf;
	b->ctx = c;
	b->bytes = bytes;
	b->dir = dir;
	b->svm = svm;
	b->w = image_w;
	b->h = image_h;
	if (owner) b->owner = owner;
	*out = b;
	return PB_OK;
}

int pb_buf_wrap(pb_ctx *c, void *dev_ptr, size_t bytes, int image_w, int image_h, pb_buf **out) {
	if (!dev_ptr) return fail(PB_ERR_ARG, "null device pointer");
	int r = pb_buf_create(c, bytes, PB_DIR_READWRITE, PB_SVM_NONE, image_w, image_h, "wrapped", out);
	if (r) return r;
	(*out)->dev = dev_ptr;
	(*out)->dev_external = true;
	return PB_OK;
}

int pb_buf_addref(pb_buf *b) {
	if (!b) return fail(PB_ERR_ARG, "null buffer");
	b->refs.fetch_add(1);
	return PB_OK;
}

int pb_buf_release(pb_buf *b) {
	if (!b) return fail(PB_ERR_ARG, "null buffer");
	pb_ctx *c = b->ctx;
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	cudaSetDevice(c->dev);
	buf_release_locked(b);
	return PB_OK;
}

int pb_buf_refs(pb_buf *b) { return b ? b->refs.load() : 0; }
size_t pb_buf_bytes(pb_buf *b) { return b ? b->bytes : 0; }

void *pb_buf_host_ptr(pb_buf *b) {
	if (!b) return nullptr;
	std::lock_guard<std::recursive_mutex> lk(b->ctx->mu);
	cudaSetDevice(b->ctx->dev);
	if (ensure_host(b)) return nullptr;
	return b->host;
}

void *pb_buf_dev_ptr(pb_buf *b) {
	if (!b) return nullptr;
	std::lock_guard<std::recursive_mutex> lk(b->ctx->mu);
	cudaSetDevice(b->ctx->dev);
	if (materialise_buf(b)) return nullptr;
	if (flush_host(b, b->ctx->q[PB_QUEUE_PROCESS])) return nullptr;
	if (ensure_dev(b)) return nullptr;
	return b->dev;
}

int pb_buf_is_deferred(pb_buf *b) { return (b && b->expr) ? 1 : 0; }

namespace {
cudaEvent_t take_copy_event(pb_ctx *c) {   // call with c->mu held
	if (!c->copy_events.empty()) {
		cudaEvent_t e = c->copy_events.back();
		c->copy_events.pop_back();
		return e;
	}
	cudaEvent_t e = nullptr;
	cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync);
	return e;
}
int wait_copy_event(pb_ctx *c, cudaEvent_t ev) {   // call WITHOUT c->mu
	const cudaError_t err = cudaEventSynchronize(ev);
	{
		std::lock_guard<std::recursive_mutex> lk(c->mu);
		c->copy_events.push_back(ev);
	}
	if (err != cudaSuccess) return fail(PB_ERR_CUDA, "waiting for a copy: %s", cudaGetErrorString(err));
	return PB_OK;
}
}  // namespace

int pb_buf_host_access(pb_buf *b, int mode, int queue, const void *src, size_t src_bytes) {
	if (!b) return fail(PB_ERR_ARG, "null buffer");
	if (queue < 0 || queue > 2) return fail(PB_ERR_ARG, "bad queue %d", queue);
	pb_ctx *c = b->ctx;
	cudaStream_t s;
	cudaEvent_t ev = nullptr;
	static const bool trace = getenv("PB_TRACE") != nullptr;   // host-side timing of frame-sized copies, for tools/e2e_probe.py
	const auto t_in = std::chrono::steady_clock::now();
	auto t_locked = t_in, t_queued = t_in;
	struct Trace {
		const bool on;
		const int mode, queue;
		const size_t bytes;
		const std::chrono::steady_clock::time_point &a, &b, &c;
		~Trace() {
			if (!on || bytes < (1u << 20)) return;
			const auto d = std::chrono::steady_clock::now();
			auto us = [](auto x, auto y) { return (long)std::chrono::duration_cast<std::chrono::microseconds>(y - x).count(); };
			fprintf(stderr, "[pb trace] hostAccess mode %d queue %d %zu B: lock %ld us, enqueue %ld us, wait %ld us\n", mode, queue, bytes, us(a, b), us(b, c), us(c, d));
		}
	} tr{trace, mode, queue, src ? src_bytes : b->bytes, t_in, t_locked, t_queued};
	{
		std::lock_guard<std::recursive_mutex> lk(c->mu);
		t_locked = std::chrono::steady_clock::now();
		cudaSetDevice(c->dev);
		s = c->q[queue];
		int r;
		switch (mode) {
			case PB_ACCESS_WRITEONLY:
				b->expr.reset();
				if ((!src || src_bytes <= 65536) && (r = ensure_host(b))) return r;
				if (src) {
					if (src_bytes > b->bytes) return fail(PB_ERR_ARG, "source (%zu bytes) larger than buffer (%zu)", src_bytes, b->bytes);
					if ((r = ensure_dev(b))) return r;
					if (src_bytes <= 65536) {
						// small constants (matrices, flip values) are also read from the host face
						memcpy(b->host, src, src_bytes);
						CU(cudaMemcpyAsync(b->dev, b->host, src_bytes, cudaMemcpyHostToDevice, s));
					} else {
						// frames: DMA straight from the caller's memory (pinned if it came from pb_host_alloc)
						CU(cudaMemcpyAsync(b->dev, src, src_bytes, cudaMemcpyHostToDevice, s));
					}
					c->stats.h2d_bytes += src_bytes;
					b->host_dirty = false;
					b->version = ++c->version_counter;
				} else {
					b->host_dirty = true;   // host will write through pb_buf_host_ptr(); flushed on next use
					return PB_OK;
				}
				break;
			case PB_ACCESS_READONLY: {
				if ((r = ensure_host(b))) return r;
				if (b->host_dirty && !b->expr) return PB_OK;   // host face is the newest copy
				const bool was_deferred = (bool)b->expr;
				if ((r = materialise_buf(b))) return r;
				if (!b->dev) {   // never written: reads as zeros
					memset(b->host, 0, b->bytes);
					return PB_OK;
				}
				if (was_deferred || s != c->q[PB_QUEUE_PROCESS]) {   // order the copy after the producing kernels
					CU(cudaEventRecord(c->ev_x, c->q[PB_QUEUE_PROCESS]));
					CU(cudaStreamWaitEvent(s, c->ev_x, 0));
				}
				CU(cudaMemcpyAsync(b->host, b->dev, b->bytes, cudaMemcpyDeviceToHost, s));
				c->stats.d2h_bytes += b->bytes;
				break;
			}
			case PB_ACCESS_NONE:
				if (!b->host_dirty) return PB_OK;   // nothing to hand back: do not wait for other producers' copies on this queue
				if ((r = flush_host(b, s))) return r;
				break;
			default:
				return fail(PB_ERR_ARG, "bad access mode %d", mode);
		}
		if (queue != PB_QUEUE_PROCESS) {
			ev = take_copy_event(c);
			if (ev) CU(cudaEventRecord(ev, s));
		}
		t_queued = std::chrono::steady_clock::now();
	}
	if (ev) return wait_copy_event(c, ev);
	CU(cudaStreamSynchronize(s));
	return PB_OK;
}

int pb_buf_upload_async(pb_buf *b, int queue, const void *src, size_t bytes) {
	if (!b || !src) return fail(PB_ERR_ARG, "null argument");
	if (queue < 0 || queue > 2) return fail(PB_ERR_ARG, "bad queue %d", queue);
	if (bytes > b->bytes) return fail(PB_ERR_ARG, "upload larger than buffer");
	pb_ctx *c = b->ctx;
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	cudaSetDevice(c->dev);
	b->expr.reset();
	b->host_dirty = false;
	int r = ensure_dev(b);
	if (r) return r;
	CU(cudaMemcpyAsync(b->dev, src, bytes, cudaMemcpyHostToDevice, c->q[queue]));
	c->stats.h2d_bytes += bytes;
	b->version = ++c->version_counter;
	return PB_OK;
}

int pb_buf_download_async(pb_buf *b, int queue, void *dst, size_t bytes) {
	if (!b || !dst) return fail(PB_ERR_ARG, "null argument");
	if (queue < 0 || queue > 2) return fail(PB_ERR_ARG, "bad queue %d", queue);
	if (bytes > b->bytes) return fail(PB_ERR_ARG, "download larger than buffer");
	pb_ctx *c = b->ctx;
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	cudaSetDevice(c->dev);
	int r = materialise_buf(b);
	if (r) return r;
	if (!b->dev) return fail(PB_ERR_STATE, "buffer has no device contents");
	CU(cudaMemcpyAsync(dst, b->dev, bytes, cudaMemcpyDeviceToHost, c->q[queue]));
	c->stats.d2h_bytes += bytes;
	return PB_OK;
}

void *pb_host_alloc(size_t bytes) {
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes) != cudaSuccess) {
		fail(PB_ERR_CUDA, "cudaMallocHost(%zu) failed", bytes);
		return nullptr;
	}
	return p;
}

void pb_host_free(void *p) {
	if (p) cudaFreeHost(p);
}

int pb_prog_create(pb_ctx *c, int op, int width, int height, pb_prog **out) {
	if (!c || !out) return fail(PB_ERR_ARG, "null argument");
	if (width <= 0 || height <= 0) return fail(PB_ERR_ARG, "bad program dimensions %dx%d", width, height);
	switch (op) {
		case PB_OP_V210_READ: case PB_OP_V210_WRITE: case PB_OP_RGBA8_READ: case PB_OP_RGBA8_WRITE: case PB_OP_BGRA8_READ:
		case PB_OP_BGRA8_WRITE: case PB_OP_COMBINE: case PB_OP_DISSOLVE: case PB_OP_WIPE_MASK: case PB_OP_TRANSFORM:
		case PB_OP_YADIF: case PB_OP_MIX: case PB_OP_WIPE: case PB_OP_RESIZE:
		case PB_OP_YUV422P10_READ: case PB_OP_YUV422P10_WRITE: case PB_OP_YUV422P8_READ: case PB_OP_YUV422P8_WRITE:
		case PB_OP_YUV420P_READ: case PB_OP_YUV420P_WRITE: case PB_OP_NV12_READ: case PB_OP_NV12_WRITE:
			break;
		default:
			return fail(PB_ERR_ARG, "unknown op %d", op);
	}
	if ((op == PB_OP_V210_READ || op == PB_OP_V210_WRITE) && (width % 2)) return fail(PB_ERR_ARG, "v210 width must be even");
	*out = new pb_prog{c, op, width, height};
	return PB_OK;
}

int pb_prog_destroy(pb_prog *g) {
	delete g;
	return PB_OK;
}

int pb_run_program(pb_ctx *c, pb_prog *g, const pb_param *params, int num_params, int queue, pb_timings *t) {
	if (!c || !g || (num_params && !params)) return fail(PB_ERR_ARG, "null argument");
	if (g->ctx != c) return fail(PB_ERR_ARG, "program belongs to another context");
	if (queue < 0 || queue > 2) return fail(PB_ERR_ARG, "bad queue %d", queue);
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	CU(cudaSetDevice(c->dev));
	cudaStream_t s = c->q[queue];
	if (t) {
		memset(t, 0, sizeof *t);
		CU(cudaEventRecord(c->ev0, s));
	}
	const uint64_t before = c->stats.kernel_launches;
	const auto h0 = std::chrono::steady_clock::now();
	int r = run_locked(c, g, params, num_params, s);
	c->stats.run_program_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - h0).count();
	c->stats.run_program_calls += 1;
	if (r) return r;
	if (t && c->stats.kernel_launches != before) {
		CU(cudaEventRecord(c->ev1, s));
		CU(cudaEventSynchronize(c->ev1));
		float ms = 0.f;
		CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
		t->kernelExec = (uint32_t)(ms * 1000.0f + 0.5f);
		t->totalTime = t->kernelExec;
	}
	return PB_OK;
}

int pb_wait_finish(pb_ctx *c, int queue) {
	if (!c) return fail(PB_ERR_ARG, "null context");
	if (queue < 0 || queue > 2) return fail(PB_ERR_ARG, "bad queue %d", queue);
	CU(cudaSetDevice(c->dev));
	if (queue != PB_QUEUE_PROCESS) {   // copy queues: sleep on an event (see pb_ctx::copy_events)
		cudaEvent_t ev;
		{
			std::lock_guard<std::recursive_mutex> lk(c->mu);
			ev = take_copy_event(c);
			if (ev) CU(cudaEventRecord(ev, c->q[queue]));
		}
		if (ev) return wait_copy_event(c, ev);
	}
	CU(cudaStreamSynchronize(c->q[queue]));
	return PB_OK;
}

int pb_queue_wait_queue(pb_ctx *c, int queue, int on_queue) {
	if (!c) return fail(PB_ERR_ARG, "null context");
	if (queue < 0 || queue > 2 || on_queue < 0 || on_queue > 2) return fail(PB_ERR_ARG, "bad queue");
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	CU(cudaSetDevice(c->dev));
	CU(cudaEventRecord(c->ev_x, c->q[on_queue]));
	CU(cudaStreamWaitEvent(c->q[queue], c->ev_x, 0));
	return PB_OK;
}

int pb_chain_begin(pb_ctx *c) {
	if (!c) return fail(PB_ERR_ARG, "null context");
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	if (c->recording) return fail(PB_ERR_STATE, "already recording");
	c->recording = new pb_chain;
	c->recording->ctx = c;
	return PB_OK;
}

int pb_chain_end(pb_ctx *c, pb_chain **out) {
	if (!c || !out) return fail(PB_ERR_ARG, "null argument");
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	if (!c->recording) return fail(PB_ERR_STATE, "not recording");
	*out = c->recording;
	c->recording = nullptr;
	return PB_OK;
}

int pb_chain_info(pb_chain *ch, int *launches, int *complete) {
	if (!ch) return fail(PB_ERR_ARG, "null chain");
	if (launches) *launches = (int)ch->items.size();
	if (complete) *complete = ch->complete ? 1 : 0;
	return PB_OK;
}

int pb_chain_replay(pb_chain *ch, int queue) {
	if (!ch || queue < 0 || queue > 2) return fail(PB_ERR_ARG, "bad argument");
	pb_ctx *c = ch->ctx;
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	CU(cudaSetDevice(c->dev));
	for (const auto &it : ch->items) {
		int r = launch_compiled(c, c->q[queue], it.d, it.march, it.out_rgba);
		if (r) return r;
		c->stats.kernel_launches++;
		c->stats.fused_launches++;
		if (it.march) c->stats.march_launches++;
	}
	return PB_OK;
}

int pb_chain_destroy(pb_chain *ch) {
	if (!ch) return PB_OK;
	pb_ctx *c = ch->ctx;
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	cudaSetDevice(c->dev);
	for (auto &it : ch->items) {
		it.keep.clear();
		if (it.out_buf) buf_release_locked(it.out_buf);
	}
	delete ch;
	return PB_OK;
}

struct pb_event {
	pb_ctx *ctx;
	cudaEvent_t ev;
};

int pb_event_create(pb_ctx *c, pb_event **out) {
	if (!c || !out) return fail(PB_ERR_ARG, "null argument");
	CU(cudaSetDevice(c->dev));
	auto *e = new pb_event{c, nullptr};
	CU(cudaEventCreate(&e->ev));
	*out = e;
	return PB_OK;
}
int pb_event_record(pb_event *e, int queue) {
	if (!e || queue < 0 || queue > 2) return fail(PB_ERR_ARG, "bad argument");
	CU(cudaSetDevice(e->ctx->dev));
	CU(cudaEventRecord(e->ev, e->ctx->q[queue]));
	return PB_OK;
}
int pb_event_sync(pb_event *e) {
	if (!e) return fail(PB_ERR_ARG, "null event");
	CU(cudaEventSynchronize(e->ev));
	return PB_OK;
}
int pb_event_elapsed_ms(pb_event *a, pb_event *b, float *ms) {
	if (!a || !b || !ms) return fail(PB_ERR_ARG, "null argument");
	CU(cudaEventElapsedTime(ms, a->ev, b->ev));
	return PB_OK;
}
int pb_event_destroy(pb_event *e) {
	if (e) {
		cudaEventDestroy(e->ev);
		delete e;
	}
	return PB_OK;
}

void *pb_ctx_stream(pb_ctx *c, int queue) {
	if (!c || queue < 0 || queue > 2) return nullptr;
	return (void *)c->q[queue];
}

}  // extern "C"
