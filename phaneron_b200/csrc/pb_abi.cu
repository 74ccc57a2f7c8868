// pb_abi.cu -- the C ABI (include/phaneron_b200.h): contexts, buffers, programs, queues, events, recorded chains.
#include "pb_internal.h"

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
	c->march_sms = c->prop.multiProcessorCount;
	for (auto &q : c->q) CU(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
	c->pool.queues = c->q;
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
	for (auto &e : c->lanczos_tabs) {
		cudaFree(e.dev);
		cudaFree(e.dstrip);
	}
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
	{   // a fresh content id: two buffers never share one (the gamma-table cache is keyed by it), and 0 is never cached
		std::lock_guard<std::recursive_mutex> lk(c->mu);
		b->version = ++c->version_counter;
	}
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

int pb_buf_trim(pb_buf *b) {
	if (!b) return fail(PB_ERR_ARG, "null buffer");
	pb_ctx *c = b->ctx;
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	b->expr.reset();
	if (b->dev && !b->dev_external) c->pool.dev_put(b->bytes, b->dev);
	b->dev = nullptr;
	b->host_dirty = false;
	return PB_OK;
}
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
	if (launches) {   // fused launches and the first passes that go with them
		*launches = 0;
		for (const auto &it : ch->items) *launches += 1 + (int)it.pre.size();
	}
	if (complete) *complete = ch->complete ? 1 : 0;
	return PB_OK;
}

int pb_chain_replay(pb_chain *ch, int queue) {
	if (!ch || queue < 0 || queue > 2) return fail(PB_ERR_ARG, "bad argument");
	pb_ctx *c = ch->ctx;
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	CU(cudaSetDevice(c->dev));
	for (const auto &it : ch->items) {
		int r = launch_compiled(c, c->q[queue], it.d, it.march, it.out_rgba, it.pre.empty() ? nullptr : &it.pre);
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
		for (auto &sc : it.scratch) c->pool.dev_put(sc.second, sc.first);
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

