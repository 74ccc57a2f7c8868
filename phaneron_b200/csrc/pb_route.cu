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

// stream memory operations of the driver API (the runtime hands out the entry points)
typedef int (*StreamValue32Fn)(cudaStream_t, unsigned long long, uint32_t, unsigned int);
struct DriverOps {
	StreamValue32Fn wait32 = nullptr, write32 = nullptr;
};
DriverOps *driver_ops() {
	static DriverOps d;
	static std::once_flag once;
	std::call_once(once, [] {
		void *w = nullptr, *r = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &w, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) d.wait32 = (StreamValue32Fn)w;
		if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &r, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) d.write32 = (StreamValue32Fn)r;
	});
	return &d;
}
constexpr unsigned kWaitGeq = 0x0;   // CU_STREAM_WAIT_VALUE_GEQ

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

	// ---- copy-engine transport (pb_route_attach) ------------------------------------------------------------------------------
	// NCCL's point-to-point kernel needs SMs, and the fused kernels are persistent: the exchange would wait for the end of a
	// launch and then run in its place (profiles/r02_route_sms.txt).  Between GPUs of one node the frame can instead be PUSHED
	// by the copy engines straight into the receiver's landing buffer (CUDA IPC mapping), with two 32-bit sequence numbers in
	// device memory for flow control, written and waited for by the streams themselves (cuStreamWriteValue32 /
	// cuStreamWaitValue32): no SM, no host round trip.
	//   flags[0] "ready":  written by the sender into the RECEIVER's flags after its copy: exchanges delivered so far
	//   flags[1] "credit": written by the receiver into the SENDER's flags when the frames that read an exchange are done
	struct Ce {
		bool on = false;
		int peer_in = -1, peer_out = -1, slots = 0;
		pb_buf *landing[4] = {};          // this rank's landing buffers, in the order they are filled
		void *remote_landing[4] = {};     // peer_out's landing buffers, mapped here
		uint32_t *flags = nullptr;        // this rank's flags (written by the peers)
		uint32_t *flags_of_in = nullptr;  // peer_in's flags, mapped: credit goes there
		uint32_t *flags_of_out = nullptr; // peer_out's flags, mapped: ready goes there
		uint32_t sent = 0, received = 0;  // exchanges so far (1-based sequence numbers)
		uint32_t credit_due = 0;          // sequence number of the last exchange a queue was made to wait for
		uint32_t credit_given = 0;
	} ce;
	uint32_t gen_recv_seq[kRing] = {};   // per exchange of the ring: sequence number of its copy-engine receive (0: none)
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
	if (m->ce.flags) {
		std::lock_guard<std::recursive_mutex> lk(m->ctx->mu);
		for (int i = 0; i < 4; ++i) {
			if (m->ce.remote_landing[i]) cudaIpcCloseMemHandle(m->ce.remote_landing[i]);
			if (m->ce.landing[i]) buf_release_locked(m->ce.landing[i]);
		}
		if (m->ce.flags_of_in && m->ce.flags_of_in != m->ce.flags_of_out) cudaIpcCloseMemHandle(m->ce.flags_of_in);
		if (m->ce.flags_of_out) cudaIpcCloseMemHandle(m->ce.flags_of_out);
		cudaFree(m->ce.flags);
	}
	if (m->comm && nccl()->CommDestroy) nccl()->CommDestroy(m->comm);
	cudaStreamDestroy(m->rs);
	cudaEventDestroy(m->ev_ready);
	for (auto &g : m->gen) cudaEventDestroy(g.done);
	delete m;
	return PB_OK;
}

// Collective over the communicator, once: every rank names the landing buffers it will receive into (in the order it will
// pass them to pb_route_recv, round robin), the rank it receives from and the rank it sends to.  The ranks exchange CUDA IPC
// handles (over NCCL, a few hundred bytes) and map each other's buffers.  From then on pb_route_send to `peer_out` and
// pb_route_recv from `peer_in` use the copy engines; anything else (and every failure to map) stays on NCCL.
int pb_route_attach(pb_comm *m, pb_buf **landing, int n, int peer_in, int peer_out) {
	if (!m || !landing) return fail(PB_ERR_ARG, "null argument");
	if (n < 2 || n > 4) return fail(PB_ERR_ARG, "pb_route_attach: 2..4 landing buffers, found %d", n);
	if (peer_in < 0 || peer_in >= m->world || peer_out < 0 || peer_out >= m->world) return fail(PB_ERR_ARG, "bad peer");
	if (m->open) return fail(PB_ERR_STATE, "pb_route_attach inside an open exchange");
	if (m->world < 2 || peer_in == m->rank || peer_out == m->rank || getenv("PB_ROUTE_NCCL_ONLY")) return PB_OK;   // nothing to map: NCCL path
	DriverOps *ops = driver_ops();
	if (!ops->wait32 || !ops->write32) return PB_OK;
	pb_ctx *c = m->ctx;
	std::lock_guard<std::recursive_mutex> lk(c->mu);
	CU(cudaSetDevice(c->dev));
	pb_comm::Ce &ce = m->ce;
	CU(cudaMalloc(&ce.flags, 64));
	CU(cudaMemsetAsync(ce.flags, 0, 64, m->rs));
	struct Packet {
		cudaIpcMemHandle_t flags, landing[4];
		int slots, rank;
		unsigned long long bytes;
	} mine{}, from_in{}, from_out{};
	CU(cudaIpcGetMemHandle(&mine.flags, ce.flags));
	for (int i = 0; i < n; ++i) {
		pb_buf *b = landing[i];
		if (!b || b->ctx != c) return fail(PB_ERR_ARG, "landing buffer %d belongs to another context", i);
		int r = ensure_dev(b);
		if (r) return r;
		b->expr.reset();
		CU(cudaIpcGetMemHandle(&mine.landing[i], b->dev));
		b->refs.fetch_add(1);
		ce.landing[i] = b;
	}
	mine.slots = n;
	mine.rank = m->rank;
	mine.bytes = landing[0]->bytes;
	Packet *dev = nullptr;
	CU(cudaMalloc(&dev, 3 * sizeof(Packet)));
	CU(cudaMemcpyAsync(dev, &mine, sizeof mine, cudaMemcpyHostToDevice, m->rs));
	// my handles go to both neighbours (peer_in writes my landing buffers and my "ready"; peer_out writes my "credit")
	NC(nccl()->GroupStart());
	NC(nccl()->Send(dev, sizeof(Packet), kNcclUint8, peer_in, m->comm, m->rs));
	NC(nccl()->Send(dev, sizeof(Packet), kNcclUint8, peer_out, m->comm, m->rs));
	NC(nccl()->Recv(dev + 1, sizeof(Packet), kNcclUint8, peer_out, m->comm, m->rs));   // (same order on every rank: from the rank I send to ...
	NC(nccl()->Recv(dev + 2, sizeof(Packet), kNcclUint8, peer_in, m->comm, m->rs));    //  ... then from the rank I receive from)
	NC(nccl()->GroupEnd());
	CU(cudaMemcpyAsync(&from_out, dev + 1, sizeof(Packet), cudaMemcpyDeviceToHost, m->rs));
	CU(cudaMemcpyAsync(&from_in, dev + 2, sizeof(Packet), cudaMemcpyDeviceToHost, m->rs));
	CU(cudaStreamSynchronize(m->rs));
	cudaFree(dev);
	bool ok = from_out.rank == peer_out && from_in.rank == peer_in && from_out.slots >= 2 && from_out.slots <= 4 && from_out.bytes == mine.bytes;
	void *p = nullptr;
	if (ok && cudaIpcOpenMemHandle(&p, from_out.flags, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) ce.flags_of_out = (uint32_t *)p;
	else ok = false;
	if (ok && peer_in == peer_out) ce.flags_of_in = ce.flags_of_out;   // (two ranks: one neighbour, one mapping)
	else if (ok && cudaIpcOpenMemHandle(&p, from_in.flags, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) ce.flags_of_in = (uint32_t *)p;
	else ok = false;
	for (int i = 0; ok && i < from_out.slots; ++i) {
		if (cudaIpcOpenMemHandle(&p, from_out.landing[i], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) ce.remote_landing[i] = p;
		else ok = false;
	}
	cudaGetLastError();   // (a failed mapping is not an error of the library: the NCCL path stays)
	ce.peer_in = peer_in;
	ce.peer_out = peer_out;
	ce.slots = n;
	ce.on = ok && from_out.slots == n;
	return PB_OK;
}

int pb_route_transport(pb_comm *m) { return (m && m->ce.on) ? 1 : 0; }

int pb_route_begin(pb_comm *m) {
	if (!m) return fail(PB_ERR_ARG, "null comm");
	if (m->open) return fail(PB_ERR_STATE, "pb_route_begin: an exchange is already open");
	CU(cudaSetDevice(m->ctx->dev));
	if (m->ce.on && m->ce.credit_due > m->ce.credit_given) {
		// the frames that read the exchange(s) waited for since the last begin are queued by now (the contract of pb_route_wait):
		// when they have run, the sender may overwrite those landing buffers
		if (driver_ops()->write32(m->ctx->q[PB_QUEUE_PROCESS], (unsigned long long)(uintptr_t)(m->ce.flags_of_in + 1), m->ce.credit_due, 0) != 0)
			return fail(PB_ERR_CUDA, "cuStreamWriteValue32 (credit) failed");
		m->ce.credit_given = m->ce.credit_due;
	}
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
		m->gen_recv_seq[next] = 0;
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
	if (m->ce.on && peer == m->ce.peer_out && frame->bytes == m->ce.landing[0]->bytes) {
		pb_comm::Ce &ce = m->ce;
		const uint32_t seq = ++ce.sent;   // 1-based
		const int slot = (int)((seq - 1) % (uint32_t)ce.slots);
		DriverOps *ops = driver_ops();
		// the receiver must be done with what this slot held: exchange seq - slots
		if (seq > (uint32_t)ce.slots && ops->wait32(m->rs, (unsigned long long)(uintptr_t)(ce.flags + 1), seq - (uint32_t)ce.slots, kWaitGeq) != 0)
			return fail(PB_ERR_CUDA, "cuStreamWaitValue32 (credit) failed");
		CU(cudaMemcpyAsync(ce.remote_landing[slot], frame->dev, frame->bytes, cudaMemcpyDeviceToDevice, m->rs));   // copy engines, over NVLink
		if (ops->write32(m->rs, (unsigned long long)(uintptr_t)(ce.flags_of_out + 0), seq, 0) != 0) return fail(PB_ERR_CUDA, "cuStreamWriteValue32 (ready) failed");
	} else {
		NC(nccl()->Send(frame->dev, frame->bytes, kNcclUint8, peer, m->comm, m->rs));
	}
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
	if (m->ce.on && peer == m->ce.peer_in) {
		// the sender's copy engine fills the buffer; nothing runs here.  The credit protocol keeps the sender off the buffer until
		// its earlier readers are done; pb_route_wait makes the consumer wait for the "ready" sequence number.
		pb_comm::Ce &ce = m->ce;
		if (landing != ce.landing[ce.received % (uint32_t)ce.slots])
			return fail(PB_ERR_STATE, "pb_route_recv: attached landing buffers are filled in the order they were attached");
		m->gen_recv_seq[m->head] = ++ce.received;
	} else {
		if ((r = order_after_queues(m))) return r;   // earlier readers of the landing buffer finish before it is overwritten
		NC(nccl()->Recv(landing->dev, landing->bytes, kNcclUint8, peer, m->comm, m->rs));
	}
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
	const int gi = (m->head + pb_comm::kRing - age) % pb_comm::kRing;
	pb_comm::Gen &g = m->gen[gi];
	CU(cudaStreamWaitEvent(m->ctx->q[queue], g.done, 0));
	if (m->ce.on && m->gen_recv_seq[gi]) {   // the frame pushed by the peer's copy engine: wait for its sequence number, on the device
		if (driver_ops()->wait32(m->ctx->q[queue], (unsigned long long)(uintptr_t)(m->ce.flags + 0), m->gen_recv_seq[gi], kWaitGeq) != 0)
			return fail(PB_ERR_CUDA, "cuStreamWaitValue32 (ready) failed");
		m->ce.credit_due = std::max(m->ce.credit_due, m->gen_recv_seq[gi]);
	}
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
