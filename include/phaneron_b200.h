/*
 * phaneron_b200.h -- C ABI of libphaneron_b200.so, the B200 (sm_100a) replacement
 * for the `nodencl` OpenCL addon that the .ts files of Streampunk/phaneron's src/process and
 * src/clJobQueue.ts bind to.
 *
 * Every entry point cites the nodencl call (by its call sites in the reference,
 * paths under /root/reference/) that it replaces.  Plain pointers and sizes only;
 * no C++/torch types.  Thread-safe per context (one mutex per pb_ctx); every call
 * returns PB_OK or a negative pb_status, with the message in pb_last_error()
 * (thread-local).  Intended callers: the N-API shim (napi/phaneron_napi.cc, one
 * async work item per call, mirroring nodencl's Promises) and Python ctypes
 * (phaneron_b200/_lib.py).
 *
 * There is NO CPU fallback: pb_ctx_create fails when no CUDA device is present.
 */
#ifndef PHANERON_B200_H
#define PHANERON_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_ctx pb_ctx;   /* nodencl clContext           (index.ts:94-102)  */
typedef struct pb_buf pb_buf;   /* nodencl OpenCLBuffer        (io.ts:61-77)      */
typedef struct pb_prog pb_prog; /* nodencl OpenCLProgram       (packer.ts:97-103) */

typedef enum pb_status {
	PB_OK = 0,
	PB_ERR_CUDA = -1,     /* a CUDA runtime call failed */
	PB_ERR_ARG = -2,      /* bad argument / missing kernel parameter */
	PB_ERR_NO_DEVICE = -3,
	PB_ERR_STATE = -4     /* e.g. running a program on a released buffer */
} pb_status;

/* clContext.queue.{load,process,unload} (clJobQueue.ts:126,131; transform.ts:100) */
typedef enum pb_queue { PB_QUEUE_LOAD = 0, PB_QUEUE_PROCESS = 1, PB_QUEUE_UNLOAD = 2 } pb_queue;
/* createBuffer(numBytes, 'readonly'|'writeonly'|'readwrite', 'none'|'coarse'|'fine', ...) */
typedef enum pb_dir { PB_DIR_READONLY = 0, PB_DIR_WRITEONLY = 1, PB_DIR_READWRITE = 2 } pb_dir;
typedef enum pb_svm { PB_SVM_NONE = 0, PB_SVM_COARSE = 1, PB_SVM_FINE = 2 } pb_svm;
/* buffer.hostAccess('none'|'readonly'|'writeonly', queue?, src?) */
typedef enum pb_access { PB_ACCESS_NONE = 0, PB_ACCESS_READONLY = 1, PB_ACCESS_WRITEONLY = 2 } pb_access;

/* What createProgram's OpenCL source string + entry name selected in the reference.
   ('read'/'write' are reused by all packers, so identity comes from the op.) */
typedef enum pb_op {
	PB_OP_V210_READ = 1,    /* v210.ts:25-111    */
	PB_OP_V210_WRITE = 2,   /* v210.ts:113-195   */
	PB_OP_RGBA8_READ = 3,   /* rgba8.ts:25-67    */
	PB_OP_RGBA8_WRITE = 4,  /* rgba8.ts:69-103   */
	PB_OP_BGRA8_READ = 5,   /* bgra8.ts:25-67    */
	PB_OP_BGRA8_WRITE = 6,  /* bgra8.ts:69-103   */
	PB_OP_COMBINE = 10,     /* combine.ts:24-68  (N from the lKIn params bound) */
	PB_OP_DISSOLVE = 11,    /* transition.ts:60-65 */
	PB_OP_WIPE_MASK = 12,   /* transition.ts:66-73 */
	PB_OP_TRANSFORM = 13,   /* transform.ts:36-59 */
	PB_OP_YADIF = 14,       /* yadifCl.ts:105-167 */
	PB_OP_MIX = 15,         /* mix.ts:30-45 */
	PB_OP_WIPE = 16,        /* wipe.ts:30-47 */
	PB_OP_RESIZE = 17,      /* resize.ts:35-59 */
	PB_OP_YUV422P10_READ = 20,   /* yuv422p10.ts:25-124  (params inputY, inputU, inputV, output, ...) */
	PB_OP_YUV422P10_WRITE = 21,  /* yuv422p10.ts:126-219 (params input, outputY, outputU, outputV, ...) */
	PB_OP_YUV422P8_READ = 22,    /* yuv422p8.ts:25-124  */
	PB_OP_YUV422P8_WRITE = 23,   /* yuv422p8.ts:126-219 */
	PB_OP_YUV420P_READ = 24,     /* yuv420p.ts:25-140  (params inputY, inputU, inputV, output, ...) */
	PB_OP_YUV420P_WRITE = 25,    /* yuv420p.ts:142-238 (params input, outputY, outputU, outputV, ...) */
	PB_OP_NV12_READ = 26,        /* nv12.ts:24-132  (params inputY, inputC, output, ...) */
	PB_OP_NV12_WRITE = 27        /* nv12.ts:134-240 (params input, outputY, outputC, ...) */
} pb_op;

/* One kernel argument, bound BY KERNEL PARAMETER NAME as nodencl's runProgram
   does (clJobQueue.ts:126; names from each getKernelParams, e.g. v210.ts:297-309). */
typedef enum pb_param_kind { PB_PARAM_BUF = 0, PB_PARAM_NUM = 1 } pb_param_kind;
typedef struct pb_param {
	const char *name;
	int kind;     /* pb_param_kind */
	pb_buf *buf;  /* PB_PARAM_BUF */
	double num;   /* PB_PARAM_NUM: numbers and booleans */
} pb_param;

/* nodencl RunTimings, microseconds (clJobQueue.ts:183-190) */
typedef struct pb_timings {
	uint32_t dataToKernel;
	uint32_t kernelExec;
	uint32_t totalTime;
} pb_timings;

typedef struct pb_stats {
	uint64_t kernel_launches;   /* every CUDA kernel this context launched */
	uint64_t fused_launches;    /* of which: fused chain launches */
	uint64_t march_launches;    /* of which: the march kernel (pb_march.cu) */
	uint64_t deferred_nodes;    /* jobs recorded into the frame-expression DAG */
	uint64_t materialised;      /* deferred RGBA frames that had to be written to HBM */
	uint64_t h2d_bytes, d2h_bytes;
	uint64_t dev_bytes_live, dev_bytes_pooled;
	uint64_t lut_tables;        /* distinct gamma tables seen (by content) */
	uint64_t lut_tables_d8;     /* of which: held in the lossless one-byte form the march kernel keeps in shared memory */
	uint64_t march_src_bytes;   /* PB_CTX_FOOTPRINT: distinct packed source bytes read by the last march launch */
	uint64_t lut_tables_poly;   /* of lut_tables_d8: decoded without MUFU (polynomial power segment) */
	uint64_t run_program_ns;    /* host time spent inside pb_run_program (recording, flattening, launch issue), nanoseconds */
	uint64_t run_program_calls;
} pb_stats;

const char *pb_last_error(void);
const char *pb_version(void);

/* new clContext({platformIndex, deviceIndex, overlapping}) + initialise()
   (index.ts:94-102).  flags: bit0 = defer RGBA intermediates and fuse at sinks
   (default behaviour of the product); 0 = eager, one launch per job. */
#define PB_CTX_DEFER 1u
/* bit1 = never pick the march kernel (always the generic fused kernel); for A/B tests */
#define PB_CTX_NO_MARCH 2u
/* bit2 = march kernel gathers from the raw 256 KiB gamma tables in global memory instead of the
   one-byte shared-memory form (A/B tests; faster only on very coherent pictures) */
#define PB_CTX_RAW_LUT 4u
/* bit3 = no occlusion culling in the march kernel.  By default, ops of layers that lie under a layer whose alpha is
   exactly 1.0f over a whole strip line are skipped: combine.ts:49-59 multiplies them by 1 - 1 = 0, so the output
   bytes are identical either way (tests/test_gpu_chain.py); the flag exists for A/B measurements. */
#define PB_CTX_NO_CULL 8u
/* bit4 = account, per march launch, the distinct packed source bytes the launch reads (pb_stats.march_src_bytes);
   costs a host pass over the op masks, so it is off by default (bench.py uses it for the roofline figure) */
#define PB_CTX_FOOTPRINT 16u
/* bit5 = do not use the dedicated 1:1 v210 -> v210 kernel (k_march_direct); A/B tests against the general march kernel */
#define PB_CTX_NO_DIRECT 32u
int pb_ctx_create(int gpu_index, unsigned flags, pb_ctx **out);
int pb_ctx_destroy(pb_ctx *ctx);
/* getPlatformInfo() (index.ts:103-107): JSON text into buf */
int pb_ctx_info(pb_ctx *ctx, char *buf, size_t buf_len);
int pb_ctx_stats(pb_ctx *ctx, pb_stats *out);
int pb_ctx_set_flags(pb_ctx *ctx, unsigned flags);

/* createBuffer(numBytes, dir, svm, imageDims?, owner?) (io.ts:61-77; mixer.ts:196-207) */
int pb_buf_create(pb_ctx *ctx, size_t bytes, int dir, int svm, int image_w, int image_h,
                  const char *owner, pb_buf **out);
/* wrap device memory owned by the caller (ROUTE frames received over NCCL) */
int pb_buf_wrap(pb_ctx *ctx, void *dev_ptr, size_t bytes, int image_w, int image_h, pb_buf **out);
int pb_buf_addref(pb_buf *buf);   /* OpenCLBuffer.addRef()  */
int pb_buf_release(pb_buf *buf);  /* OpenCLBuffer.release() */
int pb_buf_refs(pb_buf *buf);
/* give the device memory (and any deferred expression) of a buffer back to the pool while keeping the handle and its
   pinned host face: what the N-API wrapper calls when the last USER reference has gone and only the node Buffer's own
   reference keeps the bytes visible to JavaScript until the garbage collector runs */
int pb_buf_trim(pb_buf *buf);
size_t pb_buf_bytes(pb_buf *buf);
/* host-addressable storage of the Buffer subclass (pinned); valid until release */
void *pb_buf_host_ptr(pb_buf *buf);
/* device pointer; materialises a deferred frame.  NULL on error. */
void *pb_buf_dev_ptr(pb_buf *buf);
/* 1 if the frame currently exists only as a deferred expression (no HBM copy) */
int pb_buf_is_deferred(pb_buf *buf);
/* hostAccess(mode, queue, src): WRITEONLY+src = H2D copy of src; WRITEONLY w/o src = map
   for host writes (flushed to the device before the next use); READONLY = D2H into the
   host storage (materialising if deferred); NONE = hand back to the device.  Blocks
   until the copy on `queue` is complete, like the resolved Promise. */
int pb_buf_host_access(pb_buf *buf, int mode, int queue, const void *src, size_t src_bytes);
/* async variants for callers that pipeline (bench e2e, ROUTE): enqueue only */
int pb_buf_upload_async(pb_buf *buf, int queue, const void *pinned_src, size_t bytes);
int pb_buf_download_async(pb_buf *buf, int queue, void *pinned_dst, size_t bytes);
/* pinned host allocations for callers' staging rings */
void *pb_host_alloc(size_t bytes);
void pb_host_free(void *p);

/* createProgram(source, {name, globalWorkItems, workItemsPerGroup}) (packer.ts:97-103,
   imageProcess.ts:68-74): the op enum stands in for the OpenCL source. */
int pb_prog_create(pb_ctx *ctx, int op, int width, int height, pb_prog **out);
int pb_prog_destroy(pb_prog *prog);

/* runProgram(program, params, queue) (clJobQueue.ts:122-128).  In PB_CTX_DEFER mode a
   job whose output is an RGBA frame is only RECORDED (timings zero); a packed sink
   (any *_WRITE op) compiles the recorded expression into one fused launch. */
int pb_run_program(pb_ctx *ctx, pb_prog *prog, const pb_param *params, int num_params, int queue,
                   pb_timings *timings);
/* waitFinish(queue) (clJobQueue.ts:131) */
int pb_wait_finish(pb_ctx *ctx, int queue);
/* make `queue` wait (device-side) for everything enqueued so far on `on_queue` */
int pb_queue_wait_queue(pb_ctx *ctx, int queue, int on_queue);
/* the CUDA stream behind a queue id (for event timing by the caller) */
void *pb_ctx_stream(pb_ctx *ctx, int queue);

/* CUDA-event timing on a queue's stream (the device-side replacement for the hrtime
   bookkeeping in clJobQueue.ts:159-215); used by bench.py */
typedef struct pb_event pb_event;
int pb_event_create(pb_ctx *ctx, pb_event **out);
int pb_event_record(pb_event *ev, int queue);
int pb_event_sync(pb_event *ev);
int pb_event_elapsed_ms(pb_event *start, pb_event *stop, float *ms);
int pb_event_destroy(pb_event *ev);

/* Record / replay of fused launches.  phaneron re-issues an identical job list every
   frame (same programs, same buffer roles); a host that recycles its buffers can record
   the launches of one frame and replay them without re-walking the job queue.  bench.py
   uses this to time the device path without Python in the loop.  Buffers referenced by
   a chain stay alive until pb_chain_destroy. */
typedef struct pb_chain pb_chain;
int pb_chain_begin(pb_ctx *ctx);
int pb_chain_end(pb_ctx *ctx, pb_chain **out);
/* launches in the chain; *complete = 0 if a non-replayable (stand-alone) kernel ran while recording */
int pb_chain_info(pb_chain *chain, int *launches, int *complete);
int pb_chain_replay(pb_chain *chain, int queue);
int pb_chain_destroy(pb_chain *chain);

/* colourMaths.ts exports, so hosts without colourMaths.ts (Python, C++) agree with the
   TypeScript bit for bit.  Return 1 if colspec was known, 0 if it fell back to '709'. */
int pb_gamma2linear_lut(const char *colspec, float *out65536);   /* colourMaths.ts:130-149 */
int pb_linear2gamma_lut(const char *colspec, float *out65536);   /* colourMaths.ts:151-169 */
int pb_ycbcr2rgb_matrix(const char *colspec, int num_bits, int luma_black, int luma_white,
                        int chr_range, float *out12);              /* colourMaths.ts:276-332 */
int pb_rgb2ycbcr_matrix(const char *colspec, int num_bits, int luma_black, int luma_white,
                        int chr_range, float *out12);              /* colourMaths.ts:334-390 */
int pb_rgb2rgb_matrix(const char *src, const char *dst, float *out9); /* colourMaths.ts:392-394 */
/* Transform.getKernelParams matrix build (transform.ts:119-171) */
int pb_transform_matrix(int width, int height, int flip_h, int flip_v, double anchor_x,
                        double anchor_y, double scale_x, double scale_y, double offset_x,
                        double offset_y, double rotate_turns, float *out9);

/* ---- ROUTE between channels on different GPUs ------------------------------------------------------------------
   The reference's ROUTE producer forks another channel's pipes: the routed frame is a REFERENCE to that channel's
   combined RGBA-f32 OpenCLBuffer (routeProducer.ts:63-73, channel.ts:289-300, routeSource.ts:26-31), all inside one
   process and one device.  With channels sharded one per GPU the frame has to cross: these calls are that hand-off
   (csrc/pb_route.cu).  One process per GPU: NCCL point-to-point (ncclSend / ncclRecv over NVLink) on a side stream of the
   context, one group per frame period; the host never blocks in the steady state.  libnccl.so.2 is loaded on first use.

     rank 0: pb_comm_unique_id(id); the host carries the 128 bytes to the other ranks (its control plane)
     every rank: pb_comm_init(ctx, rank, world, id, &comm)
     every frame period: pb_route_begin; pb_route_send(frame, peer) / pb_route_recv(landing, peer) ...; pb_route_end;
                         pb_route_wait(comm, PB_QUEUE_PROCESS) where a received frame is first consumed
   Sent frames (materialised on demand if still deferred) and landing buffers stay referenced until the exchange has
   completed; the next pb_route_begin (or pb_route_sync) gives them back. */
typedef struct pb_comm pb_comm;
#define PB_COMM_ID_BYTES 128
int pb_comm_unique_id(void *out128);
int pb_comm_init(pb_ctx *ctx, int rank, int world, const void *id128, pb_comm **out);
int pb_comm_info(pb_comm *comm, int *rank, int *world, uint64_t *bytes_sent, uint64_t *bytes_received);
int pb_comm_destroy(pb_comm *comm);
int pb_route_begin(pb_comm *comm);
int pb_route_send(pb_comm *comm, pb_buf *frame, int peer);
int pb_route_recv(pb_comm *comm, pb_buf *landing, int peer);
int pb_route_end(pb_comm *comm);
/* Copy-engine transport between the GPUs of one node (optional; collective over the communicator, once).  Every rank names
   the 2..4 landing buffers it will receive into -- in the order it will pass them to pb_route_recv, round robin --, the rank
   it receives from and the rank it sends to.  The ranks exchange CUDA IPC handles and map each other's buffers; from then on
   pb_route_send(frame, peer_out) pushes the frame into the receiver's landing buffer with the copy engines (no SM: NCCL's
   copy kernel would wait for the persistent fused kernels) and two sequence numbers in device memory, written and waited for
   by the streams themselves, do the flow control.  Contract: the launches that read a received buffer are queued between the
   pb_route_wait that made a queue wait for it and the next pb_route_begin.  Where mapping is impossible the calls stay on
   NCCL; pb_route_transport says which it is (1 = copy engines). */
int pb_route_attach(pb_comm *comm, pb_buf **landing, int n_landing, int peer_in, int peer_out);
int pb_route_transport(pb_comm *comm);
/* make `queue` wait, on the device, for the exchange last ended */
int pb_route_wait(pb_comm *comm, int queue);
/* ... or for the one `age` exchanges before it (age 0 = the last one; up to 2): a host that composes frame n + 1 from what
   exchange n - 1 delivered lets exchange n overlap that frame's kernels */
int pb_route_wait_age(pb_comm *comm, int queue, int age);
/* block the host until the exchange last ended has completed and release the buffers it held */
int pb_route_sync(pb_comm *comm);
/* one process, one context per GPU (a Node.js host): peer-to-peer copy of a routed frame into a buffer of another
   context, on the destination's load queue, ordered after the source's process queue */
int pb_route_copy_peer(pb_buf *src, pb_buf *dst);

#ifdef __cplusplus
}
#endif
#endif
