// napi/stub/napi.h -- NOT node-addon-api.  A declaration-only stand-in for the part of node-addon-api's <napi.h> that
// napi/phaneron_napi.cc uses, so that the shim can be syntax- and type-checked in an image without Node.js:
//     g++ -std=c++17 -fsyntax-only -Inapi/stub -Iinclude napi/phaneron_napi.cc
// Signatures follow node-addon-api 3.x-7.x (Napi::ObjectWrap, AsyncWorker, Promise::Deferred, Buffer<T>, External<T>,
// TypedArrayOf<T> ...).  Nothing here has a real definition: it cannot be linked, only compiled against.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <string>
#include <vector>

namespace Napi {

class Value;
class Object;
class String;
class Number;
class Boolean;
class Array;
class Function;
class Error;
class Promise;
class CallbackInfo;
class ObjectReference;
class FunctionReference;
template <typename T> class Buffer;
template <typename T> class External;
template <typename T> class TypedArrayOf;

class Env {
 public:
	Value Undefined() const;
	Value Null() const;
	Object Global() const;
};

class Value {
 public:
	Value();
	bool IsUndefined() const;
	bool IsNull() const;
	bool IsBoolean() const;
	bool IsNumber() const;
	bool IsString() const;
	bool IsObject() const;
	bool IsArray() const;
	bool IsFunction() const;
	bool IsBuffer() const;
	bool IsExternal() const;
	bool IsTypedArray() const;
	template <typename T> T As() const;
	Napi::Env Env() const;
};

class Boolean : public Value {
 public:
	static Boolean New(Napi::Env env, bool v);
	bool Value() const;
};

class Number : public Value {
 public:
	static Number New(Napi::Env env, double v);
	int32_t Int32Value() const;
	uint32_t Uint32Value() const;
	int64_t Int64Value() const;
	double DoubleValue() const;
};

class String : public Value {
 public:
	static String New(Napi::Env env, const std::string &v);
	static String New(Napi::Env env, const char *v);
	std::string Utf8Value() const;
};

class Object : public Value {
 public:
	static Object New(Napi::Env env);
	bool Has(const char *key) const;
	bool Has(const std::string &key) const;
	Value Get(const char *key) const;
	Value Get(const std::string &key) const;
	Value Get(uint32_t index) const;
	template <typename V> void Set(const char *key, const V &v);
	template <typename V> void Set(const std::string &key, const V &v);
	template <typename V> void Set(uint32_t index, const V &v);
	Array GetPropertyNames() const;
};

class Array : public Object {
 public:
	static Array New(Napi::Env env, size_t length);
	uint32_t Length() const;
};

class Function : public Object {
 public:
	using VoidCallback = std::function<void(const CallbackInfo &)>;
	using Callback = std::function<Value(const CallbackInfo &)>;
	template <typename Callable> static Function New(Napi::Env, Callable, const char * = nullptr) { return Function(); }   // (templates over local lambdas need a body)
	Value Call(Value recv, const std::initializer_list<Value> &args) const;
	Object New(const std::initializer_list<Value> &args) const;
};

template <typename T>
class External : public Value {
 public:
	static External New(Napi::Env env, T *data);
	template <typename Finalizer> static External New(Napi::Env, T *, Finalizer) { return External(); }
	T *Data() const;
};

template <typename T>
class Buffer : public Object {
 public:
	template <typename Finalizer> static Buffer<T> New(Napi::Env, T *, size_t, Finalizer) { return Buffer<T>(); }
	static Buffer<T> Copy(Napi::Env env, const T *data, size_t length);
	T *Data() const;
	size_t Length() const;
};

template <typename T>
class TypedArrayOf : public Object {
 public:
	static TypedArrayOf New(Napi::Env env, size_t elementLength);
	size_t ElementLength() const;
	T *Data() const;
	T &operator[](size_t i);
};
using Uint32Array = TypedArrayOf<uint32_t>;
using Float32Array = TypedArrayOf<float>;

class Error {
 public:
	static Error New(Napi::Env env, const std::string &msg);
	static Error New(Napi::Env env, const char *msg);
	Napi::Value Value() const;
	const std::string &Message() const;
};
class TypeError : public Error {
 public:
	static TypeError New(Napi::Env env, const std::string &msg);
	static TypeError New(Napi::Env env, const char *msg);
};

class Promise : public Object {
 public:
	class Deferred {
	 public:
		static Deferred New(Napi::Env env);
		Napi::Promise Promise() const;
		void Resolve(Napi::Value v) const;
		void Reject(Napi::Value v) const;
	};
};

class ObjectReference {
 public:
	Object Value() const;
};
class FunctionReference {
 public:
	Object New(const std::initializer_list<Napi::Value> &args) const;
};
ObjectReference Persistent(Object o);
FunctionReference Persistent(Function f);

class CallbackInfo {
 public:
	Napi::Env Env() const;
	size_t Length() const;
	const Value operator[](size_t i) const;
	Value This() const;
};

class AsyncWorker {
 public:
	void Queue();
	Napi::Env Env() const;
	virtual ~AsyncWorker();

 protected:
	explicit AsyncWorker(Napi::Env env);
	virtual void Execute() = 0;
	virtual void OnOK();
	virtual void OnError(const Error &e);
	void SetError(const std::string &msg);
};

template <typename T>
class ObjectWrap {
 public:
	explicit ObjectWrap(const CallbackInfo &info);
	virtual ~ObjectWrap();
	struct PropertyDescriptor {};
	using InstanceMethodCallback = Value (T::*)(const CallbackInfo &);
	using StaticMethodCallback = Value (*)(const CallbackInfo &);
	static PropertyDescriptor InstanceMethod(const char *name, InstanceMethodCallback cb);
	static PropertyDescriptor StaticMethod(const char *name, StaticMethodCallback cb);
	static Function DefineClass(Napi::Env env, const char *name, const std::initializer_list<PropertyDescriptor> &props);
	static T *Unwrap(Object o);
};

}  // namespace Napi

#define NODE_API_MODULE(modname, regfunc) Napi::Object modname##_register(Napi::Env env, Napi::Object exports) { return regfunc(env, exports); }
