// pb_route.cu -- ROUTE between channels that live on different GPUs.
//
// In the reference a ROUTE producer forks another channel's pipes inside one process and one OpenCL device: the routed
// "frame" is a reference to that channel's combined RGBA-f32 OpenCLBuffer (routeProducer.ts:63-73, channel.ts:289-300,
// routeSource.ts:26-31).  Here channels shard one per GPU, so a ROUTE whose source channel lives on
// another GPU is the path's single exchange step: the source side sends the frame point-to-point, the destination side
// receives it into an image buffer that enters its layer stack like any other source.  No reduction, no collective.
//
//   * one process per GPU (the layout bench.py and the tests use): NCCL point-to-point over NVLink / NVSwitch,
//     ncclSend / ncclRecv on a side stream of the context, grouped per frame period (pb_route_begin .. pb_route_end);
//   * one process driving several GPUs (a Node.js host with one context per device): pb_route_copy_peer, a peer-to-peer
//     cudaMemcpyPeerAsync between two contexts' buffers.
//
// Ordering is by CUDA events only, the host never blocks in the steady state: the side stream waits for the process queue
// before it reads a frame to send or overwrites a landing buffer, and pb_route_wait makes a queue wait for the exchange.
// Sent frames and landing buffers are reference-held until the exchange has completed.
//
// libnccl.so.2 is loaded on first use (dlopen): the library itself does not depend on NCCL, and a process that already
// carries a NCCL (torch.distributed) shares that copy.
#include <dlfcn.h>

#include "pb_internal.h"

namespace {

// the few NCCL declarations used (nccl.h, NCCL 2.x ABI)
typedef struct ncclComm *ncclComm_t;
struct ncclUniqueId_ {
	char internal[128];
};
typedef int ncclResult_t;       // ncclSuccess == 0
constexpr int kNcclUint8 = 1;   // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1

struct Nccl {
	void *handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	std::string why;
};

Nccl *nccl() {
	static Nccl n;
	static std::once_flag once;
	std::call_once(once, [] {
		const char *names[] = {getenv("PB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
		for (const char *nm : names) {
			if (!nm || !*nm) continue;
			n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
			if (n.handle) break;
		}
		if (!n.handle) {
			n.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "dlopen failed");
			return;
		}
		bool ok = true;
		auto sym = [&](const char *s) {
			void *p = dlsym(n.handle, s);
			if (!p) {
				ok = false;
				n.why = std::string("NCCL symbol missing: ") + s;
			}
			return p;
		};
		n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
		n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
		n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
		n.Send = (decltype(n.Send))sym("ncclSend");
		n.Recv = (decltype(n.Recv))sym("ncclRecv");
		n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
		n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
		n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
		if (!ok) {
			dlclose(n.handle);
			n.handle = nullptr;
		}
	});
	return &n;
}

#define NC(call)                                                                                                    \
	do {                                                                                                            \
		ncclResult_t r__ = (call);                                                                                  \
		if (r__ != 0) return fail(PB_ERR_CUDA, "%s: %s", #call, nccl()->GetErrorString ? nccl()->GetErrorString(r__) : "NCCL error"); \
	} while (0)

}  // namespace

struct pb_comm {
	pb_ctx *ctx = nullptr;
	ncclComm_t comm = nullptr;
	int rank = 0, world = 1;
	cudaStream_t rs = nullptr;           // the side stream the exchanges run on
	cudaEvent_t ev_ready = nullptr;      // process queue -> side stream
	// The last kRing exchanges, newest at `head`: the event recorded at their end on the side stream and the buffers they hold.
	// Keeping several lets exchange n run while frame n + 1 is composed from what exchange n - 1 delivered (pb_route_wait_age).
	static constexpr int kRing = 4;
	struct Gen {
		cudaEvent_t done = nullptr;      // side stream -> whoever consumes the received frames
		std::vector<pb_buf *> held;
		bool pending = false;            // ended, buffers still held
	} gen[kRing];
	int head = 0;
	bool open = false;                   // between begin and end
	uint64_t bytes_sent = 0, bytes_received = 0;
};

namespace {

void release_gen(pb_comm *m, pb_comm::Gen &g) {   // with ctx->mu held
	for (pb_buf *b : g.held) buf_release_locked(b);
	g.held.clear();
	g.pending = false;
}

// the side stream must not touch `b` before everything queued so far on the process (and load) queue has run
int order_after_queues(pb_comm *m) {
	for (int q : {PB_QUEUE_PROCESS, PB_QUEUE_LOAD}) {
		CU(cudaEventRecord(m->ev_ready, m->ctx->q[q]));
		CU(cudaStreamWaitEvent(m->rs, m->ev_ready, 0));
	}
	return PB_OK;
}

}  // namespace

extern "C" {

int pb_comm_unique_id(void *out128) {
	if (!out128) return fail(PB_ERR_ARG, "null argument");
	Nccl *n = nccl();
	if (!n->handle) return fail(PB_ERR_STATE, "%s", n->why.c_str());
	ncclUniqueId_ id;
	NC(n->GetUniqueId(&id));
	memcpy(out128, id.internal, sizeof id.internal);
	return PB_OK;
}

int pb_comm_init(pb_ctx *c, int rank, int world, const void *id128, pb_comm **out) {
	if (!c || !id128 || !out) return fail(PB_ERR_ARG, "null argument");
	if (world < 1 || rank < 0 || rank >= world) return fail(PB_ERR_ARG, "rank %d of %d", rank, world);
	// The fused kernels are persistent, one CTA per SM with nearly all of its registers: NCCL's copy kernel would only get an
	// SM when a fused launch ends, and the exchange would serialise with the frames it is meant to overlap.  So while a
	// communicator spans more than this GPU, the march kernels can leave PB_ROUTE_SMS SMs to NCCL, which is then told to use that
	// many channels.  Measured (profiles/r02_route_sms.txt, 2 GPUs): NCCL's point-to-point kernel moves ~10 GB/s per CTA, so
	// a reservation small enough not to hurt the frames (8-16 SMs) starves the exchange (433 / 248 us per frame period against
	// 152 us with no reservation); the default is therefore 0 and the knob stays for hosts with other trade-offs.
	int route_sms = 0;
	if (world > 1) {
		const char *e = getenv("PB_ROUTE_SMS");
		route_sms = e ? atoi(e) : 0;
		route_sms = std::max(0, std::min(route_sms, c->prop.multiProcessorCount / 2));
		if (route_sms > 0) {
			char v[16];
			snprintf(v, sizeof v, "%d", route_sms);
			setenv("NCCL_MIN_CTAS", v, 0);
			setenv("NCCL_MAX_CTAS", v, 0);
		}
	}
	Nccl *n = nccl();
	if (!n->handle) return fail(PB_ERR_STATE, "%s", n->why.c_str());
	CU(cudaSetDevice(c->dev));
	auto *m = new pb_comm;
	m->ctx = c;
	m->rank = rank;
	m->world = world;
	ncclUniqueId_ id;
	memcpy(id.internal, id128, sizeof id.internal);
	ncclResult_t r = n->CommInitRank(&m->comm, world, id, rank);
	if (r != 0) {
		delete m;
		return fail(PB_ERR_CUDA, "ncclCommInitRank: %s", n->GetErrorString(r));
	}
	{   // the exchange stream outranks the queues: its copy CTAs take the SMs that free up first
		int lo = 0, hi = 0;
		CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
		CU(cudaStreamCreateWithPriority(&m->rs, cudaStreamNonBlocking, hi));
	}
	CU(cudaEventCreateWithFlags(&m->ev_ready, cudaEventDisableTiming));
	for (auto &g : m->gen) CU(cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming));
	{
		std::lock_guard<std::recursive_mutex> lk(c->mu);
		c->march_sms = c->prop.multiProcessorCount - route_sms;
	}
	*out = m;
	return PB_OK;
}

int pb_comm_info(pb_comm *m, int *rank, int *world, uint64_t *bytes_sent, uint64_t *bytes_received) {
	if (!m) return fail(PB_ERR_ARG, "null comm");
	if (rank) *rank = m->rank;
	if (world) *world = m->world;
	if (bytes_sent) *bytes_sent = m->bytes_sent;
	if (bytes_received) *bytes_received = m->bytes_received;
	return PB_OK;
}

int pb_comm_destroy(pb_comm *m) {
	if (!m) return PB_OK;
	cudaSetDevice(m->ctx->dev);
	cudaStreamSynchronize(m->rs);
	{
		std::lock_guard<std::recursive_mutex> lk(m->ctx->mu);
		for (auto &g : m->gen) release_gen(m, g);
	}
	{
		std::lock_guard<std::recursive_mutex> lk(m->ctx->mu);
		m->ctx->march_sms = m->ctx->prop.multiProcessorCount;
	}
	if (m->comm && nccl()->CommDestroy) nccl()->CommDestroy(m->comm);
	cudaStreamDestroy(m->rs);
	cudaEventDestroy(m->ev_ready);
	for (auto &g : m->gen) cudaEventDestroy(g.done);
	delete m;
	return PB_OK;
}

int pb_route_begin(pb_comm *m) {
	if (!m) return fail(PB_ERR_ARG, "null comm");
	if (m->open) return fail(PB_ERR_STATE, "pb_route_begin: an exchange is already open");
	CU(cudaSetDevice(m->ctx->dev));
	{   // earlier exchanges that have completed give their buffers back; the slot about to be reused has to (normally it
		// finished several frame periods ago, so this does not block)
		std::lock_guard<std::recursive_mutex> lk(m->ctx->mu);
		const int next = (m->head + 1) % pb_comm::kRing;
		for (int k = 0; k < pb_comm::kRing; ++k) {
			pb_comm::Gen &g = m->gen[k];
			if (!g.pending) continue;
			if (k == next) CU(cudaEventSynchronize(g.done));
			else if (cudaEventQuery(g.done) != cudaSuccess) continue;
			release_gen(m, g);
		}
		m->head = next;
	}
	NC(nccl()->GroupStart());
	m->open = true;
	return PB_OK;
}

int pb_route_send(pb_comm *m, pb_buf *frame, int peer) {
	if (!m || !frame) return fail(PB_ERR_ARG, "null argument");
	if (!m->open) return fail(PB_ERR_STATE, "pb_route_send outside pb_route_begin / pb_route_end");
	if (peer < 0 || peer >= m->world) return fail(PB_ERR_ARG, "peer %d of %d", peer, m->world);
	if (frame->ctx != m->ctx) return fail(PB_ERR_ARG, "frame belongs to another context");
	std::lock_guard<std::recursive_mutex> lk(m->ctx->mu);
	CU(cudaSetDevice(m->ctx->dev));
	int r;
	if (frame->expr && (r = materialise_buf(frame))) return r;   // the channel frame as RGBA-f32: one fused launch on the process queue
	if ((r = flush_host(frame, m->ctx->q[PB_QUEUE_PROCESS]))) return r;
	if (!frame->dev) return fail(PB_ERR_STATE, "routed frame '%s' has no contents", frame->owner.c_str());
	if ((r = order_after_queues(m))) return r;
	NC(nccl()->Send(frame->dev, frame->bytes, kNcclUint8, peer, m->comm, m->rs));
	frame->refs.fetch_add(1);
	m->gen[m->head].held.push_back(frame);
	m->bytes_sent += frame->bytes;
	return PB_OK;
}

int pb_route_recv(pb_comm *m, pb_buf *landing, int peer) {
	if (!m || !landing) return fail(PB_ERR_ARG, "null argument");
	if (!m->open) return fail(PB_ERR_STATE, "pb_route_recv outside pb_route_begin / pb_route_end");
	if (peer < 0 || peer >= m->world) return fail(PB_ERR_ARG, "peer %d of %d", peer, m->world);
	if (landing->ctx != m->ctx) return fail(PB_ERR_ARG, "landing buffer belongs to another context");
	std::lock_guard<std::recursive_mutex> lk(m->ctx->mu);
	CU(cudaSetDevice(m->ctx->dev));
	landing->expr.reset();   // whatever the buffer stood for, it now holds the received frame
	landing->host_dirty = false;
	int r;
	if ((r = ensure_dev(landing))) return r;
	if ((r = order_after_queues(m))) return r;   // earlier readers of the landing buffer finish before it is overwritten
	NC(nccl()->Recv(landing->dev, landing->bytes, kNcclUint8, peer, m->comm, m->rs));
	landing->version = ++m->ctx->version_counter;
	landing->refs.fetch_add(1);
	m->gen[m->head].held.push_back(landing);
	m->bytes_received += landing->bytes;
	return PB_OK;
}

int pb_route_end(pb_comm *m) {
	if (!m) return fail(PB_ERR_ARG, "null comm");
	if (!m->open) return fail(PB_ERR_STATE, "pb_route_end without pb_route_begin");
	CU(cudaSetDevice(m->ctx->dev));
	m->open = false;
	NC(nccl()->GroupEnd());
	CU(cudaEventRecord(m->gen[m->head].done, m->rs));
	m->gen[m->head].pending = true;
	return PB_OK;
}

int pb_route_wait_age(pb_comm *m, int queue, int age) {
	if (!m) return fail(PB_ERR_ARG, "null comm");
	if (queue < 0 || queue > 2) return fail(PB_ERR_ARG, "bad queue %d", queue);
	if (age < 0 || age >= pb_comm::kRing - 1) return fail(PB_ERR_ARG, "age %d: the last %d exchanges can be waited for", age, pb_comm::kRing - 1);
	if (m->open) return fail(PB_ERR_STATE, "pb_route_wait inside an open exchange");
	CU(cudaSetDevice(m->ctx->dev));
	// (an exchange that has been released already has completed; its event still says so)
	pb_comm::Gen &g = m->gen[(m->head + pb_comm::kRing - age) % pb_comm::kRing];
	CU(cudaStreamWaitEvent(m->ctx->q[queue], g.done, 0));
	return PB_OK;
}

int pb_route_wait(pb_comm *m, int queue) { return pb_route_wait_age(m, queue, 0); }

int pb_route_sync(pb_comm *m) {
	if (!m) return fail(PB_ERR_ARG, "null comm");
	if (m->open) return fail(PB_ERR_STATE, "pb_route_sync inside an open exchange");
	CU(cudaSetDevice(m->ctx->dev));
	CU(cudaStreamSynchronize(m->rs));
	std::lock_guard<std::recursive_mutex> lk(m->ctx->mu);
	for (auto &g : m->gen) release_gen(m, g);
	return PB_OK;
}

// one process, two contexts (a Node.js host with a context per GPU): the routed frame crosses with a peer-to-peer copy on
// the destination's load queue, ordered after the source's process queue; consumers order themselves with
// pb_queue_wait_queue(dst_ctx, PB_QUEUE_PROCESS, PB_QUEUE_LOAD) or pb_wait_finish(dst_ctx, PB_QUEUE_LOAD)
int pb_route_copy_peer(pb_buf *src, pb_buf *dst) {
	if (!src || !dst) return fail(PB_ERR_ARG, "null argument");
	if (src->bytes > dst->bytes) return fail(PB_ERR_ARG, "routed frame holds %zu bytes, destination %zu", src->bytes, dst->bytes);
	pb_ctx *cs = src->ctx, *cd = dst->ctx;
	cudaEvent_t ev = nullptr;
	{
		std::lock_guard<std::recursive_mutex> lk(cs->mu);
		CU(cudaSetDevice(cs->dev));
		int r;
		if (src->expr && (r = materialise_buf(src))) return r;
		if ((r = flush_host(src, cs->q[PB_QUEUE_PROCESS]))) return r;
		if (!src->dev) return fail(PB_ERR_STATE, "routed frame '%s' has no contents", src->owner.c_str());
		CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
		CU(cudaEventRecord(ev, cs->q[PB_QUEUE_PROCESS]));
	}
	std::lock_guard<std::recursive_mutex> lk(cd->mu);
	CU(cudaSetDevice(cd->dev));
	dst->expr.reset();
	dst->host_dirty = false;
	int r = ensure_dev(dst);
	if (r) return r;
	cudaStream_t s = cd->q[PB_QUEUE_LOAD];
	CU(cudaStreamWaitEvent(s, ev, 0));
	CU(cudaEventRecord(cd->ev_x, cd->q[PB_QUEUE_PROCESS]));   // earlier readers of dst on the destination's process queue
	CU(cudaStreamWaitEvent(s, cd->ev_x, 0));
	CU(cudaMemcpyPeerAsync(dst->dev, cd->dev, src->dev, cs->dev, src->bytes, s));
	dst->version = ++cd->version_counter;
	// the source frame must outlive the copy: the caller keeps its reference until the destination's load queue has been
	// waited on (as with any hostAccess copy); the event is destroyed once recorded work has been consumed
	CU(cudaEventDestroy(ev));
	return PB_OK;
}

}  // extern "C"
