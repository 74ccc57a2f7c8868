// phaneron_napi.cc -- the N-API addon that stands where `nodencl` stands under Streampunk/phaneron
// (package.json:22), over the C ABI of libphaneron_b200.so (include/phaneron_b200.h).
//
// It exports what the 41 importing files of the reference use, with nodencl's shapes:
//
//   new clContext({platformIndex, deviceIndex, overlapping})     index.ts:94-102
//     .initialise(): Promise<void>
//     .getPlatformInfo(): {vendor, devices: [{type, ...}]}        index.ts:103-107
//     .queue.{load, process, unload}: number                      clJobQueue.ts:126,131
//     .createBuffer(numBytes, dir, svm, imageDims?, owner?): Promise<OpenCLBuffer>      io.ts:61-77, mixer.ts:196-207
//     .createProgram(source, {name, globalWorkItems, workItemsPerGroup?, op?}): Promise<OpenCLProgram>
//                                                                 packer.ts:97-103, imageProcess.ts:68-74
//     .runProgram(program, params, queue): Promise<RunTimings>    clJobQueue.ts:122-128
//     .waitFinish(queue?): Promise<void>                          clJobQueue.ts:131
//     .logBuffers(), .close()
//   OpenCLBuffer extends Buffer (the bytes are the pinned host face of the pb_buf):
//     .hostAccess(mode, queue?, src?): Promise<void>              io.ts:89-94,172; loadSave.ts:76,87,98
//     .addRef(), .release(), .timestamp, .loadstamp, .creationTime, .numBytes, .owner
//   ROUTE between GPUs (not in nodencl: the reference has one device):
//     clContext.uniqueId(): Buffer(128) ; ctx.createComm(rank, world, id): RouteComm
//     comm.begin() / send(buf, peer) / recv(buf, peer) / end() / wait(queue?, age?) / sync() / close()
//     clContext.routeCopyPeer(src, dst)   (one process, one context per GPU)
//
// Every call that can take time is an AsyncWorker on the libuv pool, exactly as nodencl runs its calls; errors are
// rejected Promises carrying pb_last_error() fetched ON the worker thread (the message is thread-local).
//
// createProgram: nodencl JIT-compiles the OpenCL source string.  Here programs are precompiled, so the program is
// selected by name: options.op when present (`${packImpl.name}_${programName}`, the one-line edit of ts/packer.ts.patch --
// every packer reuses the entry names 'read' / 'write'), else options.name, which is unique for every image process
// ('combine_N', 'transition_dissolve', 'transition_wipe', 'transform', 'yadif', 'mixer', 'wipe', 'resize').
//
// Build (on a machine with Node.js):  node-gyp / cmake-js with node-addon-api, linking libphaneron_b200.so; see napi/binding.gyp.
// In the build image there is no Node.js: `g++ -std=c++17 -fsyntax-only -Inapi/stub -Iinclude napi/phaneron_napi.cc`
// checks this file against the stub of the node-addon-api surface it uses (tests/test_host.py).
#include <napi.h>

#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "phaneron_b200.h"

namespace {

// ---- one async work item per native call ---------------------------------------------------------------------------------
class PbWorker : public Napi::AsyncWorker {
 public:
	using Fn = std::function<int()>;
	using Done = std::function<Napi::Value(Napi::Env)>;
	PbWorker(Napi::Env env, Fn fn, Done done) : Napi::AsyncWorker(env), deferred_(Napi::Promise::Deferred::New(env)), fn_(std::move(fn)), done_(std::move(done)) {}
	Napi::Promise Promise() const { return deferred_.Promise(); }
	void Execute() override {
		if (fn_() != PB_OK) SetError(pb_last_error());   // same thread as the failing call
	}
	void OnOK() override { deferred_.Resolve(done_ ? done_(Env()) : Env().Undefined()); }
	void OnError(const Napi::Error &e) override { deferred_.Reject(e.Value()); }

 private:
	Napi::Promise::Deferred deferred_;
	Fn fn_;
	Done done_;
};

Napi::Promise run_async(Napi::Env env, PbWorker::Fn fn, PbWorker::Done done = nullptr) {
	auto *w = new PbWorker(env, std::move(fn), std::move(done));
	Napi::Promise p = w->Promise();
	w->Queue();
	return p;
}

Napi::Error napi_error(Napi::Env env) { return Napi::Error::New(env, pb_last_error()); }

int enum_of(Napi::Env env, const Napi::Value &v, const char *what, const char *const *names, int n) {
	if (!v.IsString()) throw Napi::TypeError::New(env, std::string(what) + " must be a string");
	const std::string s = v.As<Napi::String>().Utf8Value();
	for (int i = 0; i < n; ++i)
		if (s == names[i]) return i;
	std::string all;
	for (int i = 0; i < n; ++i) all += std::string(i ? "|" : "") + names[i];
	throw Napi::Error::New(env, std::string(what) + " must be one of " + all + ", found '" + s + "'");
}
const char *const kDir[] = {"readonly", "writeonly", "readwrite"};   // pb_dir
const char *const kSvm[] = {"none", "coarse", "fine"};               // pb_svm
const char *const kAccess[] = {"none", "readonly", "writeonly"};     // pb_access

// program names -> pb_op
struct OpName {
	const char *name;
	int op;
};
const OpName kOps[] = {
	{"v210_read", PB_OP_V210_READ},           {"v210_write", PB_OP_V210_WRITE},
	{"rgba8_read", PB_OP_RGBA8_READ},         {"rgba8_write", PB_OP_RGBA8_WRITE},
	{"bgra8_read", PB_OP_BGRA8_READ},         {"bgra8_write", PB_OP_BGRA8_WRITE},
	{"yuv422p10_read", PB_OP_YUV422P10_READ}, {"yuv422p10_write", PB_OP_YUV422P10_WRITE},
	{"yuv422p10le_read", PB_OP_YUV422P10_READ}, {"yuv422p10le_write", PB_OP_YUV422P10_WRITE},   // the PackImpl names itself 'yuv422p10le' (yuv422p10.ts:297)
	{"yuv422p8_read", PB_OP_YUV422P8_READ},   {"yuv422p8_write", PB_OP_YUV422P8_WRITE},
	{"yuv420p_read", PB_OP_YUV420P_READ},     {"yuv420p_write", PB_OP_YUV420P_WRITE},
	{"nv12_read", PB_OP_NV12_READ},           {"nv12_write", PB_OP_NV12_WRITE},
	{"transition_dissolve", PB_OP_DISSOLVE},  {"transition_wipe", PB_OP_WIPE_MASK},
	{"transform", PB_OP_TRANSFORM},           {"yadif", PB_OP_YADIF},
	{"mixer", PB_OP_MIX},                     {"wipe", PB_OP_WIPE},
	{"resize", PB_OP_RESIZE},
};
int op_of(const std::string &name) {
	if (name.rfind("combine_", 0) == 0) return PB_OP_COMBINE;   // combine_N: N comes from the lKIn parameters bound (combine.ts:24-68)
	for (const OpName &o : kOps)
		if (name == o.name) return o.op;
	return 0;
}

pb_buf *buf_of(const Napi::Value &v) {
	if (!v.IsObject()) return nullptr;
	Napi::Object o = v.As<Napi::Object>();
	if (!o.Has("_pb")) return nullptr;
	Napi::Value h = o.Get("_pb");
	return h.IsExternal() ? h.As<Napi::External<pb_buf>>().Data() : nullptr;
}

// ---- OpenCLBuffer ------------------------------------------------------------------------------------------------------------
// A node Buffer over the pinned host face of a pb_buf, with nodencl's extra members.  The wrapper holds one reference of
// its own (dropped when the Buffer is collected), so the bytes stay valid as long as JavaScript can see them; when the last
// USER reference goes (release() brings the count down to the wrapper's), device memory and any deferred expression are
// given back at once (pb_buf_trim) instead of waiting for the garbage collector.
Napi::Value make_buffer(Napi::Env env, pb_buf *buf, size_t bytes, const std::string &owner) {
	void *host = pb_buf_host_ptr(buf);
	if (!host) throw napi_error(env);
	pb_buf_addref(buf);   // the wrapper's reference
	Napi::Buffer<uint8_t> b = Napi::Buffer<uint8_t>::New(env, static_cast<uint8_t *>(host), bytes, [buf](Napi::Env, uint8_t *) { pb_buf_release(buf); });
	b.Set("_pb", Napi::External<pb_buf>::New(env, buf));
	b.Set("numBytes", Napi::Number::New(env, (double)bytes));
	b.Set("owner", Napi::String::New(env, owner));
	b.Set("timestamp", Napi::Number::New(env, 0));
	b.Set("loadstamp", Napi::Number::New(env, 0));
	b.Set("creationTime", Napi::Number::New(env, 0));
	b.Set("addRef", Napi::Function::New(env, [buf](const Napi::CallbackInfo &) { pb_buf_addref(buf); }));
	b.Set("release", Napi::Function::New(env, [buf](const Napi::CallbackInfo &i) {
		if (pb_buf_refs(buf) <= 1) throw Napi::Error::New(i.Env(), "OpenCLBuffer released more often than referenced");
		pb_buf_release(buf);
		if (pb_buf_refs(buf) == 1) pb_buf_trim(buf);   // only the wrapper is left: hand device memory back now
	}));
	b.Set("refs", Napi::Function::New(env, [buf](const Napi::CallbackInfo &i) -> Napi::Value { return Napi::Number::New(i.Env(), pb_buf_refs(buf) - 1); }));
	// hostAccess(mode, queue?, src?): writeonly + src = H2D copy of src; writeonly = map for host writes; readonly = D2H
	// (materialises a deferred frame); none = hand back to the device.  Resolves when the copy on `queue` has completed.
	b.Set("hostAccess", Napi::Function::New(env, [buf](const Napi::CallbackInfo &i) -> Napi::Value {
		Napi::Env e = i.Env();
		const int mode = i.Length() > 0 && !i[0].IsUndefined() ? enum_of(e, i[0], "hostAccess mode", kAccess, 3) : PB_ACCESS_NONE;
		const int queue = i.Length() > 1 && i[1].IsNumber() ? i[1].As<Napi::Number>().Int32Value() : PB_QUEUE_LOAD;
		const void *src = nullptr;
		size_t n = 0;
		std::shared_ptr<Napi::ObjectReference> keep;   // the source Buffer must outlive the copy
		if (i.Length() > 2 && i[2].IsBuffer()) {
			Napi::Buffer<uint8_t> s = i[2].As<Napi::Buffer<uint8_t>>();
			src = s.Data();
			n = s.Length();
			keep = std::make_shared<Napi::ObjectReference>(Napi::Persistent(i[2].As<Napi::Object>()));
		}
		return run_async(e, [=] { return pb_buf_host_access(buf, mode, queue, src, n); }, [keep](Napi::Env env2) { return env2.Undefined(); });
	}));
	return b;
}

// ---- RouteComm -----------------------------------------------------------------------------------------------------------------
class RouteComm : public Napi::ObjectWrap<RouteComm> {
 public:
	static Napi::FunctionReference ctor;
	static void Init(Napi::Env env) {
		Napi::Function f = DefineClass(env, "RouteComm",
		                               {InstanceMethod("begin", &RouteComm::Begin), InstanceMethod("send", &RouteComm::Send), InstanceMethod("recv", &RouteComm::Recv),
		                                InstanceMethod("end", &RouteComm::End), InstanceMethod("wait", &RouteComm::Wait), InstanceMethod("sync", &RouteComm::Sync),
		                                InstanceMethod("info", &RouteComm::Info), InstanceMethod("close", &RouteComm::Close)});
		ctor = Napi::Persistent(f);
	}
	explicit RouteComm(const Napi::CallbackInfo &info) : Napi::ObjectWrap<RouteComm>(info) {
		if (info.Length() > 0 && info[0].IsExternal()) comm_ = info[0].As<Napi::External<pb_comm>>().Data();
	}
	~RouteComm() { pb_comm_destroy(comm_); }

 private:
	pb_comm *need(Napi::Env env) {
		if (!comm_) throw Napi::Error::New(env, "RouteComm is closed");
		return comm_;
	}
	void check(Napi::Env env, int rc) {
		if (rc != PB_OK) throw napi_error(env);
	}
	Napi::Value Begin(const Napi::CallbackInfo &i) { check(i.Env(), pb_route_begin(need(i.Env()))); return i.Env().Undefined(); }
	Napi::Value Send(const Napi::CallbackInfo &i) {
		pb_buf *b = buf_of(i[0]);
		if (!b) throw Napi::TypeError::New(i.Env(), "send(buffer, peer): buffer must be an OpenCLBuffer");
		check(i.Env(), pb_route_send(need(i.Env()), b, i[1].As<Napi::Number>().Int32Value()));
		return i.Env().Undefined();
	}
	Napi::Value Recv(const Napi::CallbackInfo &i) {
		pb_buf *b = buf_of(i[0]);
		if (!b) throw Napi::TypeError::New(i.Env(), "recv(buffer, peer): buffer must be an OpenCLBuffer");
		check(i.Env(), pb_route_recv(need(i.Env()), b, i[1].As<Napi::Number>().Int32Value()));
		return i.Env().Undefined();
	}
	Napi::Value End(const Napi::CallbackInfo &i) { check(i.Env(), pb_route_end(need(i.Env()))); return i.Env().Undefined(); }
	Napi::Value Wait(const Napi::CallbackInfo &i) {
		const int queue = i.Length() > 0 && i[0].IsNumber() ? i[0].As<Napi::Number>().Int32Value() : PB_QUEUE_PROCESS;
		const int age = i.Length() > 1 && i[1].IsNumber() ? i[1].As<Napi::Number>().Int32Value() : 0;
		check(i.Env(), pb_route_wait_age(need(i.Env()), queue, age));
		return i.Env().Undefined();
	}
	Napi::Value Sync(const Napi::CallbackInfo &i) {
		pb_comm *c = need(i.Env());
		return run_async(i.Env(), [c] { return pb_route_sync(c); });
	}
	Napi::Value Info(const Napi::CallbackInfo &i) {
		int rank = 0, world = 0;
		uint64_t s = 0, r = 0;
		check(i.Env(), pb_comm_info(need(i.Env()), &rank, &world, &s, &r));
		Napi::Object o = Napi::Object::New(i.Env());
		o.Set("rank", Napi::Number::New(i.Env(), rank));
		o.Set("world", Napi::Number::New(i.Env(), world));
		o.Set("bytesSent", Napi::Number::New(i.Env(), (double)s));
		o.Set("bytesReceived", Napi::Number::New(i.Env(), (double)r));
		return o;
	}
	Napi::Value Close(const Napi::CallbackInfo &i) {
		pb_comm_destroy(comm_);
		comm_ = nullptr;
		return i.Env().Undefined();
	}
	pb_comm *comm_ = nullptr;
};
Napi::FunctionReference RouteComm::ctor;

// ---- clContext -----------------------------------------------------------------------------------------------------------------
class Context : public Napi::ObjectWrap<Context> {
 public:
	static Napi::Function Init(Napi::Env env) {
		return DefineClass(env, "clContext",
		                   {InstanceMethod("initialise", &Context::Initialise), InstanceMethod("getPlatformInfo", &Context::GetPlatformInfo),
		                    InstanceMethod("createBuffer", &Context::CreateBuffer), InstanceMethod("createProgram", &Context::CreateProgram),
		                    InstanceMethod("runProgram", &Context::RunProgram), InstanceMethod("waitFinish", &Context::WaitFinish),
		                    InstanceMethod("logBuffers", &Context::LogBuffers), InstanceMethod("stats", &Context::Stats),
		                    InstanceMethod("setFlags", &Context::SetFlags), InstanceMethod("createComm", &Context::CreateComm),
		                    InstanceMethod("close", &Context::Close), StaticMethod("uniqueId", &Context::UniqueId),
		                    StaticMethod("routeCopyPeer", &Context::RouteCopyPeer)});
	}

	explicit Context(const Napi::CallbackInfo &info) : Napi::ObjectWrap<Context>(info) {
		Napi::Env env = info.Env();
		if (info.Length() > 0 && info[0].IsObject()) {
			Napi::Object o = info[0].As<Napi::Object>();
			if (o.Has("platformIndex") && o.Get("platformIndex").IsNumber()) platform_ = o.Get("platformIndex").As<Napi::Number>().Int32Value();
			if (o.Has("deviceIndex") && o.Get("deviceIndex").IsNumber()) device_ = o.Get("deviceIndex").As<Napi::Number>().Int32Value();
			if (o.Has("deferred") && o.Get("deferred").IsBoolean() && !o.Get("deferred").As<Napi::Boolean>().Value()) flags_ &= ~PB_CTX_DEFER;
		}
		Napi::Object q = Napi::Object::New(env);   // clContext.queue.{load,process,unload} (clJobQueue.ts:126,131)
		q.Set("load", Napi::Number::New(env, PB_QUEUE_LOAD));
		q.Set("process", Napi::Number::New(env, PB_QUEUE_PROCESS));
		q.Set("unload", Napi::Number::New(env, PB_QUEUE_UNLOAD));
		info.This().As<Napi::Object>().Set("queue", q);
	}
	~Context() { pb_ctx_destroy(ctx_); }

 private:
	pb_ctx *need(Napi::Env env) {
		if (!ctx_) throw Napi::Error::New(env, "clContext is not initialised");
		return ctx_;
	}

	// initialise(): Promise<void>  (index.ts:102)
	Napi::Value Initialise(const Napi::CallbackInfo &info) {
		Napi::Env env = info.Env();
		if (platform_ != 0) throw Napi::Error::New(env, "phaneron_b200 exposes one platform (CUDA); platformIndex must be 0");
		return run_async(env, [this] { return pb_ctx_create(device_, flags_, &ctx_); });
	}

	// getPlatformInfo(): {vendor, devices:[{type}]}  (index.ts:103-107)
	Napi::Value GetPlatformInfo(const Napi::CallbackInfo &info) {
		Napi::Env env = info.Env();
		char text[1024];
		if (pb_ctx_info(need(env), text, sizeof text) != PB_OK) throw napi_error(env);
		Napi::Object json = env.Global().Get("JSON").As<Napi::Object>();
		Napi::Value dev = json.Get("parse").As<Napi::Function>().Call(json, {Napi::String::New(env, text)});
		// pb_ctx_info describes THE device of this context; nodencl lists all devices of the platform and the caller indexes by
		// deviceIndex: pad the array so that devices[deviceIndex] is this device
		Napi::Object o = dev.As<Napi::Object>();
		Napi::Array devices = o.Get("devices").As<Napi::Array>();
		Napi::Value mine = devices.Get(0u);
		Napi::Array padded = Napi::Array::New(env, (size_t)device_ + 1);
		for (int i = 0; i <= device_; ++i) padded.Set((uint32_t)i, mine);
		o.Set("devices", padded);
		return o;
	}

	// createBuffer(numBytes, dir, svm, imageDims?, owner?): Promise<OpenCLBuffer>
	Napi::Value CreateBuffer(const Napi::CallbackInfo &info) {
		Napi::Env env = info.Env();
		pb_ctx *ctx = need(env);
		if (info.Length() < 3 || !info[0].IsNumber()) throw Napi::TypeError::New(env, "createBuffer(numBytes, bufDir, bufType, imageDims?, owner?)");
		const size_t bytes = (size_t)info[0].As<Napi::Number>().Int64Value();
		const int dir = enum_of(env, info[1], "buffer direction", kDir, 3);
		const int svm = enum_of(env, info[2], "buffer type", kSvm, 3);
		int w = 0, h = 0;
		if (info.Length() > 3 && info[3].IsObject()) {
			Napi::Object d = info[3].As<Napi::Object>();
			w = d.Get("width").As<Napi::Number>().Int32Value();
			h = d.Get("height").As<Napi::Number>().Int32Value();
		}
		const std::string owner = info.Length() > 4 && info[4].IsString() ? info[4].As<Napi::String>().Utf8Value() : "";
		auto out = std::make_shared<pb_buf *>(nullptr);
		// the pinned host allocation (cudaMallocHost on a pool miss) happens on the worker; wrapping it is main-thread work
		return run_async(
		    env,
		    [=] {
			    int r = pb_buf_create(ctx, bytes, dir, svm, w, h, owner.c_str(), out.get());
			    if (r == PB_OK && !pb_buf_host_ptr(*out)) r = PB_ERR_CUDA;
			    return r;
		    },
		    [=](Napi::Env e) { return make_buffer(e, *out, bytes, owner); });   // references: the user's (pb_buf_create) + the wrapper's
	}

	// createProgram(source, {name, globalWorkItems, workItemsPerGroup?, op?}): Promise<OpenCLProgram>
	Napi::Value CreateProgram(const Napi::CallbackInfo &info) {
		Napi::Env env = info.Env();
		pb_ctx *ctx = need(env);
		if (info.Length() < 2 || !info[1].IsObject()) throw Napi::TypeError::New(env, "createProgram(source, options)");
		Napi::Object opts = info[1].As<Napi::Object>();
		std::string name = opts.Has("name") && opts.Get("name").IsString() ? opts.Get("name").As<Napi::String>().Utf8Value() : "";
		std::string key = opts.Has("op") && opts.Get("op").IsString() ? opts.Get("op").As<Napi::String>().Utf8Value() : name;
		const int op = op_of(key);
		if (!op)
			throw Napi::Error::New(env, "createProgram: no precompiled program for '" + key +
			                                "' (packers pass options.op = `${packImpl.name}_${programName}`, ts/packer.ts.patch)");
		// image dimensions: packers give them as options.width / options.height (ts/packer.ts.patch); image processes through
		// globalWorkItems = Uint32Array [width, height] (imageProcess.ts:48-50)
		int w = 0, h = 0;
		if (opts.Has("width") && opts.Get("width").IsNumber()) w = opts.Get("width").As<Napi::Number>().Int32Value();
		if (opts.Has("height") && opts.Get("height").IsNumber()) h = opts.Get("height").As<Napi::Number>().Int32Value();
		if ((!w || !h) && opts.Has("globalWorkItems") && opts.Get("globalWorkItems").IsTypedArray()) {
			Napi::Uint32Array g = opts.Get("globalWorkItems").As<Napi::Uint32Array>();
			if (g.ElementLength() >= 2) {
				w = (int)g[0];
				h = (int)g[1];
			}
		}
		if (w <= 0 || h <= 0) throw Napi::Error::New(env, "createProgram: image width / height missing for '" + key + "'");
		auto out = std::make_shared<pb_prog *>(nullptr);
		return run_async(env, [=] { return pb_prog_create(ctx, op, w, h, out.get()); },
		                 [=](Napi::Env e) {
			                 Napi::Object p = Napi::Object::New(e);
			                 pb_prog *g = *out;
			                 p.Set("_pb", Napi::External<pb_prog>::New(e, g, [](Napi::Env, pb_prog *q) { pb_prog_destroy(q); }));
			                 p.Set("name", Napi::String::New(e, name));
			                 p.Set("op", Napi::String::New(e, key));
			                 p.Set("width", Napi::Number::New(e, w));
			                 p.Set("height", Napi::Number::New(e, h));
			                 return p;
		                 });
	}

	// runProgram(program, {argName: OpenCLBuffer | number | boolean}, queue): Promise<RunTimings>  -- arguments bound BY NAME
	Napi::Value RunProgram(const Napi::CallbackInfo &info) {
		Napi::Env env = info.Env();
		pb_ctx *ctx = need(env);
		if (info.Length() < 2 || !info[0].IsObject() || !info[1].IsObject()) throw Napi::TypeError::New(env, "runProgram(program, params, queue)");
		Napi::Object prog = info[0].As<Napi::Object>();
		if (!prog.Has("_pb") || !prog.Get("_pb").IsExternal()) throw Napi::TypeError::New(env, "runProgram: not an OpenCLProgram");
		pb_prog *g = prog.Get("_pb").As<Napi::External<pb_prog>>().Data();
		Napi::Object params = info[1].As<Napi::Object>();
		const int queue = info.Length() > 2 && info[2].IsNumber() ? info[2].As<Napi::Number>().Int32Value() : PB_QUEUE_PROCESS;

		struct Call {
			std::vector<std::string> names;
			std::vector<pb_param> params;
			std::vector<pb_buf *> held;   // every buffer argument stays referenced until the call has run
			pb_timings t{};
		};
		auto call = std::make_shared<Call>();
		Napi::Array keys = params.GetPropertyNames();
		const uint32_t n = keys.Length();
		call->names.reserve(n);
		for (uint32_t k = 0; k < n; ++k) {
			const std::string name = keys.Get(k).As<Napi::String>().Utf8Value();
			Napi::Value v = params.Get(name);
			pb_param p{};
			if (v.IsNumber()) {
				p.kind = PB_PARAM_NUM;
				p.num = v.As<Napi::Number>().DoubleValue();
			} else if (v.IsBoolean()) {
				p.kind = PB_PARAM_NUM;
				p.num = v.As<Napi::Boolean>().Value() ? 1.0 : 0.0;
			} else if (pb_buf *b = buf_of(v)) {
				p.kind = PB_PARAM_BUF;
				p.buf = b;
				pb_buf_addref(b);
				call->held.push_back(b);
			} else if (v.IsUndefined() || v.IsNull()) {
				continue;   // optional parameters left out by the caller
			} else {
				for (pb_buf *b : call->held) pb_buf_release(b);
				throw Napi::TypeError::New(env, "kernel parameter '" + name + "' must be an OpenCLBuffer, a number or a boolean");
			}
			call->names.push_back(name);
			call->params.push_back(p);
		}
		for (size_t k = 0; k < call->params.size(); ++k) call->params[k].name = call->names[k].c_str();
		return run_async(
		    env,
		    [=] {
			    const int r = pb_run_program(ctx, g, call->params.data(), (int)call->params.size(), queue, &call->t);
			    for (pb_buf *b : call->held) pb_buf_release(b);
			    return r;
		    },
		    [call](Napi::Env e) {
			    Napi::Object t = Napi::Object::New(e);   // RunTimings, microseconds (clJobQueue.ts:183-190)
			    t.Set("dataToKernel", Napi::Number::New(e, call->t.dataToKernel));
			    t.Set("kernelExec", Napi::Number::New(e, call->t.kernelExec));
			    t.Set("totalTime", Napi::Number::New(e, call->t.totalTime));
			    return t;
		    });
	}

	// waitFinish(queue?): Promise<void>
	Napi::Value WaitFinish(const Napi::CallbackInfo &info) {
		pb_ctx *ctx = need(info.Env());
		const int queue = info.Length() > 0 && info[0].IsNumber() ? info[0].As<Napi::Number>().Int32Value() : PB_QUEUE_PROCESS;
		return run_async(info.Env(), [=] { return pb_wait_finish(ctx, queue); });
	}

	Napi::Value LogBuffers(const Napi::CallbackInfo &info) { return Stats(info); }

	Napi::Value Stats(const Napi::CallbackInfo &info) {
		Napi::Env env = info.Env();
		pb_stats s{};
		if (pb_ctx_stats(need(env), &s) != PB_OK) throw napi_error(env);
		Napi::Object o = Napi::Object::New(env);
		o.Set("kernelLaunches", Napi::Number::New(env, (double)s.kernel_launches));
		o.Set("fusedLaunches", Napi::Number::New(env, (double)s.fused_launches));
		o.Set("marchLaunches", Napi::Number::New(env, (double)s.march_launches));
		o.Set("deferredNodes", Napi::Number::New(env, (double)s.deferred_nodes));
		o.Set("materialised", Napi::Number::New(env, (double)s.materialised));
		o.Set("h2dBytes", Napi::Number::New(env, (double)s.h2d_bytes));
		o.Set("d2hBytes", Napi::Number::New(env, (double)s.d2h_bytes));
		o.Set("devBytesLive", Napi::Number::New(env, (double)s.dev_bytes_live));
		o.Set("devBytesPooled", Napi::Number::New(env, (double)s.dev_bytes_pooled));
		return o;
	}

	Napi::Value SetFlags(const Napi::CallbackInfo &info) {
		Napi::Env env = info.Env();
		flags_ = info[0].As<Napi::Number>().Uint32Value();
		if (ctx_ && pb_ctx_set_flags(ctx_, flags_) != PB_OK) throw napi_error(env);
		return env.Undefined();
	}

	// static uniqueId(): Buffer(128) -- rank 0 makes it, the host's control plane carries it to the other ranks
	static Napi::Value UniqueId(const Napi::CallbackInfo &info) {
		Napi::Env env = info.Env();
		uint8_t id[PB_COMM_ID_BYTES];
		if (pb_comm_unique_id(id) != PB_OK) throw napi_error(env);
		return Napi::Buffer<uint8_t>::Copy(env, id, sizeof id);
	}

	// createComm(rank, world, id): Promise<RouteComm>
	Napi::Value CreateComm(const Napi::CallbackInfo &info) {
		Napi::Env env = info.Env();
		pb_ctx *ctx = need(env);
		if (info.Length() < 3 || !info[2].IsBuffer()) throw Napi::TypeError::New(env, "createComm(rank, world, uniqueId)");
		const int rank = info[0].As<Napi::Number>().Int32Value(), world = info[1].As<Napi::Number>().Int32Value();
		Napi::Buffer<uint8_t> idb = info[2].As<Napi::Buffer<uint8_t>>();
		if (idb.Length() != PB_COMM_ID_BYTES) throw Napi::Error::New(env, "createComm: uniqueId must hold 128 bytes");
		auto id = std::make_shared<std::vector<uint8_t>>(idb.Data(), idb.Data() + idb.Length());
		auto out = std::make_shared<pb_comm *>(nullptr);
		return run_async(env, [=] { return pb_comm_init(ctx, rank, world, id->data(), out.get()); },
		                 [out](Napi::Env e) { return RouteComm::ctor.New({Napi::External<pb_comm>::New(e, *out)}); });
	}

	// static routeCopyPeer(src, dst): Promise<void> -- one process, one context per GPU
	static Napi::Value RouteCopyPeer(const Napi::CallbackInfo &info) {
		Napi::Env env = info.Env();
		pb_buf *src = buf_of(info[0]), *dst = buf_of(info[1]);
		if (!src || !dst) throw Napi::TypeError::New(env, "routeCopyPeer(src, dst): both must be OpenCLBuffers");
		pb_buf_addref(src);
		pb_buf_addref(dst);
		return run_async(env, [=] {
			const int r = pb_route_copy_peer(src, dst);
			pb_buf_release(src);
			pb_buf_release(dst);
			return r;
		});
	}

	Napi::Value Close(const Napi::CallbackInfo &info) {
		pb_ctx_destroy(ctx_);
		ctx_ = nullptr;
		return info.Env().Undefined();
	}

	pb_ctx *ctx_ = nullptr;
	int platform_ = 0, device_ = 0;
	unsigned flags_ = PB_CTX_DEFER;
};

// colourMaths.ts stays in TypeScript (bit-identical by construction); these are for hosts that want the library's copies
Napi::Value GammaToLinearLut(const Napi::CallbackInfo &info) {
	Napi::Env env = info.Env();
	const std::string spec = info[0].As<Napi::String>().Utf8Value();
	Napi::Float32Array out = Napi::Float32Array::New(env, 65536);
	pb_gamma2linear_lut(spec.c_str(), out.Data());
	return out;
}
Napi::Value LinearToGammaLut(const Napi::CallbackInfo &info) {
	Napi::Env env = info.Env();
	const std::string spec = info[0].As<Napi::String>().Utf8Value();
	Napi::Float32Array out = Napi::Float32Array::New(env, 65536);
	pb_linear2gamma_lut(spec.c_str(), out.Data());
	return out;
}

Napi::Object InitAll(Napi::Env env, Napi::Object exports) {
	RouteComm::Init(env);
	exports.Set("clContext", Context::Init(env));
	exports.Set("gamma2linearLUT", Napi::Function::New(env, GammaToLinearLut));
	exports.Set("linear2gammaLUT", Napi::Function::New(env, LinearToGammaLut));
	exports.Set("version", Napi::String::New(env, pb_version()));
	return exports;
}

}  // namespace

NODE_API_MODULE(phaneron_b200, InitAll)
