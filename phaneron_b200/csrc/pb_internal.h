// pb_internal.h -- what the translation units of the runtime share: contexts, refcounted buffers, programs, recorded
// chains and the frame-expression nodes.  Internal to libphaneron_b200.so; the C ABI is
// include/phaneron_b200.h.
//   pb_recorder.cu     buffers, the deferred frame-expression recorder, run_locked (one job of clJobQueue.ts:122-128)
//   pb_lut_cache.cu    gamma tables: content hash, one-byte forms (pb_lut.cuh), candidate models
//   pb_march_prep.cu   march-kernel preparation: exact sampling tables, occlusion analysis, Lanczos taps, launch
//   pb_abi.cu          the extern "C" entry points
//   pb_route.cu        ROUTE between GPUs over NCCL (pb_comm_*, pb_route_*)
#pragma once
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

namespace pbrt {

extern thread_local std::string g_err;
int fail(int code, const char *fmt, ...);


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
	// A block goes back to its free list while kernels / copies that use it may still be in flight: the owner's last
	// reference often drops right after the launch was ISSUED (deferred frames release their packed sources as soon as the
	// fused launch is queued).  So every pooled block carries two events, recorded when it was returned on the process and the
	// load queue; whoever takes the block next makes all three queues wait for the ones that have not completed yet
	// (device-side waits, no host stall).  Without this an upload on the load queue could overwrite a source that a kernel
	// on the process queue was still reading.
	struct Block {
		void *p;
		cudaEvent_t ev[2];
	};
	std::unordered_map<size_t, std::vector<Block>> dev;
	std::unordered_map<size_t, std::vector<void *>> host;
	std::vector<cudaEvent_t> events;   // spare events
	cudaStream_t *queues = nullptr;    // the context's three queues (load, process, unload)
	size_t dev_pooled = 0, dev_live = 0;
	static constexpr size_t kMaxPooled = size_t(24) << 30;

	cudaEvent_t take_event() {
		if (!events.empty()) {
			cudaEvent_t e = events.back();
			events.pop_back();
			return e;
		}
		cudaEvent_t e = nullptr;
		cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
		return e;
	}
	cudaError_t dev_get(size_t n, void **p) {
		auto &v = dev[n];
		if (!v.empty()) {
			Block b = v.back();
			v.pop_back();
			for (cudaEvent_t e : b.ev) {
				if (!e) continue;
				if (queues && cudaEventQuery(e) != cudaSuccess)
					for (int q = 0; q < 3; ++q) cudaStreamWaitEvent(queues[q], e, 0);
				events.push_back(e);
			}
			*p = b.p;
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
			cudaFree(p);   // (synchronises with outstanding work on the block)
			return;
		}
		Block b{p, {nullptr, nullptr}};
		if (queues) {
			b.ev[0] = take_event();
			b.ev[1] = take_event();
			if (b.ev[0]) cudaEventRecord(b.ev[0], queues[1]);   // PB_QUEUE_PROCESS
			if (b.ev[1]) cudaEventRecord(b.ev[1], queues[0]);   // PB_QUEUE_LOAD
		}
		dev[n].push_back(b);
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
			for (Block &b : kv.second) {
				cudaFree(b.p);
				for (cudaEvent_t e : b.ev)
					if (e) events.push_back(e);
			}
		dev.clear();
		dev_pooled = 0;
	}
	void destroy() {
		trim();
		for (cudaEvent_t e : events) cudaEventDestroy(e);
		events.clear();
		for (auto &kv : host)
			for (void *p : kv.second) cudaFreeHost(p);
		host.clear();
	}
};

}  // namespace pbrt
using namespace pbrt;

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
		float *wx = nullptr, *wy = nullptr, *wxt = nullptr;
		// march kernel: first tap per output column / line on the host, and (built on demand for one strip width) the per-strip
		// source footprints {flags, first group, groups, 0} with the strips / lines outside of which all taps are border texels
		std::vector<int> h_i0, h_j0;
		int strip_groups = 0, fits = 1, s0 = 0, s1 = -1, y0 = 0, y1 = -1;
		int4 *dstrip = nullptr;
		std::shared_ptr<const SampleTab::Opq> opq;   // footprint accounting only (never opaque: the alpha is a weight sum)
	};
	std::vector<LanczosTab> lanczos_tabs;
	// blocking-sync events for waits on the copy queues: a host thread waiting for a frame-sized DMA sleeps instead of
	// spinning, and waits for ITS copy only, not for whatever other producers have queued behind it
	std::vector<cudaEvent_t> copy_events;
	bool allow_march = true;
	// first passes (and their intermediates) of the launch launch_desc has just issued, for record_launch to take over
	std::vector<pb::HPassDesc> pending_pre;
	std::vector<std::pair<void *, size_t>> pending_scratch;
	int march_sms = 0;   // SMs the persistent march kernels occupy; fewer than all while a ROUTE communicator needs SMs of its own (pb_route.cu)
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
	unsigned int *bg_counter = nullptr;   // ring of device counters of the background pass: one per launch, zeroed on its stream (launch_compiled)
	unsigned int bg_next_base = 0;        // next ring slot
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
		std::vector<pb::HPassDesc> pre;            // launches that go before it: first passes of separable Lanczos leaves
		std::vector<std::pair<void *, size_t>> scratch;   // their intermediates (pool blocks, given back with the chain)
	};
	std::vector<Item> items;
	bool complete = true;
};

namespace pbrt {


enum NodeKind { N_LEAF_V210, N_LEAF_RGBA, N_TRANSFORM, N_DISSOLVE, N_WIPE_MASK, N_COMBINE, N_LEAF_PACKED, N_YADIF };

struct Node {
	NodeKind kind;
	pb_ctx *ctx;
	int w = 0, h = 0;            // dimensions of the image this node produces
	std::vector<NodeP> in;
	pb_buf *src = nullptr;       // leaves: referenced source buffer
	pb_buf *src_u = nullptr, *src_v = nullptr;   // N_LEAF_PACKED, planar formats: chroma planes
	int leaf_kind = 0;           // N_LEAF_PACKED: pb::LeafKind (rgba8, bgra8, yuv422p10/8, yuv420p, nv12)
	int yadif = 0;               // N_YADIF (src = cur, src_u = prev, src_v = next: real RGBA-f32 frames): parity | tff << 1 | skipSpatial << 2
	pb_buf *lut_buf = nullptr;   // packed leaves: referenced gamma LUT buffer
	pb::ReadConsts rc{};         // packed leaves
	float mat[6] = {0};          // transform
	int lanczos = 0;             // transform: 0 = the reference's bilinear sampler, else Lanczos lobes
	float mix = 0.f;             // dissolve
	void *mat_dev = nullptr;     // RGBA-f32 copy if this node had to be materialised
	~Node();
};

inline int v210_pitch_bytes(int w) { return ((w + 47) / 48) * 128; }

// pb_recorder.cu
void buf_release_locked(pb_buf *b);
void buf_free(pb_buf *b);
int ensure_dev(pb_buf *b);
int ensure_host(pb_buf *b);
int flush_host(pb_buf *b, cudaStream_t s);
const pb_param *find(const pb_param *p, int n, const char *name);
int need_buf(const pb_param *p, int n, const char *name, pb_buf **out);
int need_num(const pb_param *p, int n, const char *name, double *out);
int host_floats(pb_buf *b, int count, float *out, const char *what);
int input_expr(pb_buf *b, NodeP *out);
int materialise_node(pb_ctx *c, const NodeP &n, const void **dev_out);
int materialise_buf(pb_buf *b);
int run_locked(pb_ctx *c, pb_prog *g, const pb_param *p, int n, cudaStream_t s);
// pb_lut_cache.cu
int lut_table_of(pb_ctx *c, pb_buf *lut, int *table_out);
int lut_table_by_raw(pb_ctx *c, const float *raw);
// pb_march_prep.cu
int attach_lanczos(pb_ctx *c, pb::Leaf *lf, int lobes);
int prepare_march(pb_ctx *c, pb::FusedDesc &d);
int launch_compiled(pb_ctx *c, cudaStream_t s, const pb::FusedDesc &d_in, bool march, void *out_rgba, const std::vector<pb::HPassDesc> *pre = nullptr);
int launch_desc(pb_ctx *c, cudaStream_t s, pb::FusedDesc &d, void *out_rgba, bool *march_out);

}  // namespace pbrt
