// pb_recorder.cu -- refcounted buffers with a pinned host face and the frame-expression recorder that turns
// phaneron's per-stage jobs (clJobQueue.ts:122-128) into one fused launch per output frame.
//
// Design: in PB_CTX_DEFER mode a job whose output is an RGBA-f32 image
// is not executed; its output buffer gets an expression node that references the input
// expressions (and holds references on the packed leaves).  A packed writer is a sink: it
// flattens the expression into a pb::FusedDesc and launches once.  Anything the fused kernel
// cannot express is materialised bottom-up with the stand-alone kernels and re-enters as an
// RGBA leaf.  Host reads of a deferred frame materialise it on demand.
#include "pb_internal.h"

namespace pbrt {

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
		if (n->kind == N_YADIF) {   // a de-interlaced field (yadifCl.ts:105-167), evaluated where it is sampled
			lf->kind = pb::LEAF_YADIF;
			lf->ptr = n->src->dev;
			lf->ptr_u = n->src_u->dev;
			lf->ptr_v = n->src_v->dev;
			lf->w = n->w;
			lf->h = n->h;
			lf->pitch = n->w * 16;
			lf->yadif = n->yadif;
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
			// A rotated / sheared Transform (the Mixer's DVE rotation) of a packed source: no kernel can share conversions between
			// output pixels there (the generic kernel converts every tap: four conversions per pixel and leaf), so the source is
			// made real once as RGBA-f32 (the direct kernel: 31 us at 2160p) and sampled as a frame -- by the march kernel too,
			// which takes the taps of RGBA-f32 leaves straight from global memory at any affine position (eval_leaf_f32).
			const bool rotated = (n->mat[1] != 0.0f || n->mat[3] != 0.0f) && !n->lanczos && (c->flags & PB_CTX_DEFER);
			const bool packed_child = child->kind == N_LEAF_V210 || child->kind == N_LEAF_PACKED;
			if ((rotated && packed_child) || 0 != set_basic_leaf(child, lf)) {
				int r = as_rgba_leaf(child, lf);
				if (r) return r;
				if (rotated && packed_child) lf->finite_lut = child->rc.lut;   // table values x gamut matrix: finite if the table is
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
	it.pre = std::move(c->pending_pre);
	it.scratch = std::move(c->pending_scratch);
	c->pending_pre.clear();
	c->pending_scratch.clear();
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
		cc.d.sink = pb::SINK_RGBA_F32;
		cc.d.out = p;
		cc.d.out_pitch = n->w * 16;
		bool march = false;
		if ((r = launch_desc(c, c->q[PB_QUEUE_PROCESS], cc.d, p, &march))) {
			c->pool.dev_put((size_t)n->w * n->h * 16, p);
			return r;
		}
		c->stats.kernel_launches++;
		c->stats.fused_launches++;
		if (march) c->stats.march_launches++;
		c->stats.materialised++;
		n->mat_dev = p;
		cc.keep.push_back(n);   // a recorded chain must keep the node (and its mat_dev) alive
		record_launch(c, cc, p, nullptr, march);
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
		cc.d.sink = pb::SINK_RGBA_F32;
		cc.d.out = b->dev;
		cc.d.out_pitch = n->w * 16;
		bool march = false;
		if ((r = launch_desc(c, c->q[PB_QUEUE_PROCESS], cc.d, b->dev, &march))) return r;
		c->stats.kernel_launches++;
		c->stats.fused_launches++;
		if (march) c->stats.march_launches++;
		c->stats.materialised++;
		record_launch(c, cc, b->dev, b, march);
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
			// (while a chain is being recorded a frame that already is real memory goes through the fused path too, as an RGBA
			// leaf: a recorded chain replays fused launches only -- the ROUTE payload frame that FromRGBA packs, bench_route.py)
			if (v210 && (in->expr || (c->recording && in->w > 0 && (c->flags & PB_CTX_DEFER)))) {
				NodeP root;
				if ((r = input_expr(in, &root))) return r;
				Compiler cc{c};
				if ((r = cc.compile(root))) return r;
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
			if ((r = real_input(prev, &a)) || (r = real_input(cur, &b)) || (r = real_input(next, &d))) return r;
			if (c->flags & PB_CTX_DEFER) {
				// The three frames of the window are real RGBA-f32 frames now (each ToRGBA output is made real once, by the direct
				// kernel, and serves six field evaluations).  The field itself is only recorded: it is computed inside the fused
				// launch that consumes it (Mixer Transform -> Combine -> FromRGBA), pixel by pixel where it is sampled, and never
				// exists in HBM (yadif.ts:88-113 would write it and the next stage read it back).
				if (prev->w != W || prev->h != H || cur->w != W || cur->h != H || next->w != W || next->h != H)
					return fail(PB_ERR_ARG, "yadif: the three frames must be %dx%d images", W, H);
				NodeP yn = new_node(c, N_YADIF, W, H);
				yn->src = cur;
				yn->src_u = prev;
				yn->src_v = next;
				cur->refs.fetch_add(1);
				prev->refs.fetch_add(1);
				next->refs.fetch_add(1);
				yn->yadif = ((int)parity & 1) | (tff != 0 ? 2 : 0) | (skip != 0 ? 4 : 0);
				set_deferred(out, yn);
				out->w = W;
				out->h = H;
				c->stats.deferred_nodes++;
				return PB_OK;
			}
			if ((r = real_output(out, &o))) return r;
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

}  // namespace pbrt
