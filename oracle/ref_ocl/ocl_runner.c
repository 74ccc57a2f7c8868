/*
 * ocl_runner.c -- TEST INFRASTRUCTURE: runs OpenCL C kernel source on the NVIDIA OpenCL driver of
 * the GPU box, so that the reference's OWN kernel strings (src/process/*.ts of Streampunk/phaneron,
 * extracted into the git-ignored oracle/_ref/kernels/ by extract_kernels.py) execute on the same
 * B200 as our CUDA path.  This is the "oracle/_ref" leg of the parity story:
 * the reference's executor is Node.js + nodencl + an OpenCL device; node and nodencl are absent,
 * but the kernels are plain OpenCL C and the driver is there.
 *
 * The image has no OpenCL headers and no usable ICD registry, so this file declares the handful
 * of types it needs and calls the driver through the Khronos ICD dispatch table: the vendor
 * library exports clGetExtensionFunctionAddress; "clIcdGetPlatformIDsKHR" yields the platform,
 * whose first word points at the table of entry points in the fixed order of CL/cl_icd.h.  Slot 65
 * (clGetExtensionFunctionAddress) is checked against the library's exported symbol before anything
 * else in the table is trusted.
 *
 * Nothing in phaneron_b200/ links or loads this.
 */
#include <dlfcn.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef cl_ulong cl_bitfield;
typedef void *cl_handle;
typedef struct {
	cl_uint order, type;
} cl_image_format;

enum {
	D_GetPlatformInfo = 1, D_GetDeviceIDs = 2, D_GetDeviceInfo = 3, D_CreateContext = 4, D_CreateCommandQueue = 9, D_CreateBuffer = 14,
	D_CreateImage2D = 15, D_ReleaseMemObject = 18, D_CreateProgramWithSource = 26, D_ReleaseProgram = 29, D_BuildProgram = 30,
	D_GetProgramInfo = 32, D_GetProgramBuildInfo = 33, D_CreateKernel = 34, D_ReleaseKernel = 37, D_SetKernelArg = 38, D_Finish = 47, D_EnqueueReadBuffer = 48,
	D_EnqueueWriteBuffer = 49, D_EnqueueReadImage = 51, D_EnqueueWriteImage = 52, D_EnqueueNDRangeKernel = 59,
	D_GetExtensionFunctionAddress = 65
};
#define CL_DEVICE_TYPE_GPU (1u << 2)
#define CL_DEVICE_NAME 0x102B
#define CL_MEM_READ_WRITE 1u
#define CL_RGBA 0x10B5
#define CL_FLOAT 0x10DE
#define CL_PROGRAM_BUILD_LOG 0x1183
#define CL_PROGRAM_BINARY_SIZES 0x1165
#define CL_PROGRAM_BINARIES 0x1166

typedef cl_int (*fn_IcdGetPlatformIDs)(cl_uint, cl_handle *, cl_uint *);
typedef void *(*fn_GetExt)(const char *);
typedef cl_int (*fn_GetDeviceIDs)(cl_handle, cl_bitfield, cl_uint, cl_handle *, cl_uint *);
typedef cl_int (*fn_GetDeviceInfo)(cl_handle, cl_uint, size_t, void *, size_t *);
typedef cl_handle (*fn_CreateContext)(const intptr_t *, cl_uint, const cl_handle *, void *, void *, cl_int *);
typedef cl_handle (*fn_CreateCommandQueue)(cl_handle, cl_handle, cl_bitfield, cl_int *);
typedef cl_handle (*fn_CreateBuffer)(cl_handle, cl_bitfield, size_t, void *, cl_int *);
typedef cl_handle (*fn_CreateImage2D)(cl_handle, cl_bitfield, const cl_image_format *, size_t, size_t, size_t, void *, cl_int *);
typedef cl_int (*fn_Release)(cl_handle);
typedef cl_handle (*fn_CreateProgramWithSource)(cl_handle, cl_uint, const char **, const size_t *, cl_int *);
typedef cl_int (*fn_BuildProgram)(cl_handle, cl_uint, const cl_handle *, const char *, void *, void *);
typedef cl_int (*fn_GetProgramBuildInfo)(cl_handle, cl_handle, cl_uint, size_t, void *, size_t *);
typedef cl_handle (*fn_CreateKernel)(cl_handle, const char *, cl_int *);
typedef cl_int (*fn_SetKernelArg)(cl_handle, cl_uint, size_t, const void *);
typedef cl_int (*fn_Finish)(cl_handle);
typedef cl_int (*fn_RWBuffer)(cl_handle, cl_handle, cl_uint, size_t, size_t, void *, cl_uint, const void *, void *);
typedef cl_int (*fn_RWImage)(cl_handle, cl_handle, cl_uint, const size_t *, const size_t *, size_t, size_t, void *, cl_uint, const void *, void *);
typedef cl_int (*fn_NDRange)(cl_handle, cl_handle, cl_uint, const size_t *, const size_t *, const size_t *, cl_uint, const void *, void *);

static void **g_tab;
static cl_handle g_platform, g_device, g_ctx, g_queue;
static char g_log[16384];
static char g_name[256];

#define MAX_OBJ 256
static cl_handle g_mem[MAX_OBJ], g_kern[MAX_OBJ], g_prog[MAX_OBJ];
static int g_nmem, g_nkern;

static int fail(const char *what, cl_int e) {
	snprintf(g_log, sizeof g_log, "%s failed (%d)", what, (int)e);
	return -1;
}

const char *ocl_log(void) { return g_log; }
const char *ocl_device_name(void) { return g_name; }

int ocl_init(void) {
	if (g_queue) return 0;
	const char *names[] = {"libnvidia-opencl.so.1", "/usr/lib/libnvidia-opencl.so.1", "/usr/local/nvidia/lib/libnvidia-opencl.so.1",
	                       "/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so.1"};
	void *lib = NULL;
	for (size_t i = 0; i < sizeof names / sizeof *names && !lib; ++i) lib = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
	if (!lib) {
		snprintf(g_log, sizeof g_log, "no NVIDIA OpenCL driver library: %s", dlerror());
		return -1;
	}
	fn_GetExt get_ext = (fn_GetExt)dlsym(lib, "clGetExtensionFunctionAddress");
	void *exported_info = dlsym(lib, "clGetPlatformInfo");
	if (!get_ext) return fail("dlsym(clGetExtensionFunctionAddress)", 0);
	fn_IcdGetPlatformIDs icd = (fn_IcdGetPlatformIDs)get_ext("clIcdGetPlatformIDsKHR");
	if (!icd) return fail("clIcdGetPlatformIDsKHR lookup", 0);
	cl_uint n = 0;
	cl_int e = icd(1, &g_platform, &n);
	if (e || n < 1) return fail("clIcdGetPlatformIDsKHR", e);
	g_tab = *(void ***)g_platform;
	(void)exported_info;   /* the exported clGetPlatformInfo is a thunk, not the table entry: only slot 65 can be anchored */
	if (g_tab[D_GetExtensionFunctionAddress] != (void *)get_ext) {
		snprintf(g_log, sizeof g_log, "ICD dispatch table does not have the expected layout (clGetExtensionFunctionAddress is not slot %d)",
		         D_GetExtensionFunctionAddress);
		return -1;
	}
	e = ((fn_GetDeviceIDs)g_tab[D_GetDeviceIDs])(g_platform, CL_DEVICE_TYPE_GPU, 1, &g_device, &n);
	if (e || n < 1) return fail("clGetDeviceIDs", e);
	((fn_GetDeviceInfo)g_tab[D_GetDeviceInfo])(g_device, CL_DEVICE_NAME, sizeof g_name, g_name, NULL);
	g_ctx = ((fn_CreateContext)g_tab[D_CreateContext])(NULL, 1, &g_device, NULL, NULL, &e);
	if (e) return fail("clCreateContext", e);
	g_queue = ((fn_CreateCommandQueue)g_tab[D_CreateCommandQueue])(g_ctx, g_device, 0, &e);
	if (e) return fail("clCreateCommandQueue", e);
	return 0;
}

/* build `src` with `options`, return a kernel id for entry point `name` */
int ocl_kernel(const char *src, const char *name, const char *options) {
	if (g_nkern >= MAX_OBJ) return fail("too many kernels", 0);
	cl_int e;
	cl_handle p = ((fn_CreateProgramWithSource)g_tab[D_CreateProgramWithSource])(g_ctx, 1, &src, NULL, &e);
	if (e) return fail("clCreateProgramWithSource", e);
	e = ((fn_BuildProgram)g_tab[D_BuildProgram])(p, 1, &g_device, options ? options : "", NULL, NULL);
	if (e) {
		size_t used = (size_t)snprintf(g_log, sizeof g_log, "clBuildProgram failed (%d): ", (int)e);
		((fn_GetProgramBuildInfo)g_tab[D_GetProgramBuildInfo])(p, g_device, CL_PROGRAM_BUILD_LOG, sizeof g_log - used - 1, g_log + used, NULL);
		return -1;
	}
	cl_handle k = ((fn_CreateKernel)g_tab[D_CreateKernel])(p, name, &e);
	if (e) return fail("clCreateKernel", e);
	g_prog[g_nkern] = p;
	g_kern[g_nkern] = k;
	return g_nkern++;
}

/* what the driver compiled the kernel's program to (NVIDIA returns PTX text); returns the size */
typedef cl_int (*fn_GetProgramInfo)(cl_handle, cl_uint, size_t, void *, size_t *);
long ocl_program_binary(int kern, char *dst, size_t cap) {
	size_t size = 0;
	cl_int e = ((fn_GetProgramInfo)g_tab[D_GetProgramInfo])(g_prog[kern], CL_PROGRAM_BINARY_SIZES, sizeof size, &size, NULL);
	if (e) return fail("clGetProgramInfo(sizes)", e);
	if (size + 1 > cap) return (long)size;
	unsigned char *ptrs[1] = {(unsigned char *)dst};
	e = ((fn_GetProgramInfo)g_tab[D_GetProgramInfo])(g_prog[kern], CL_PROGRAM_BINARIES, sizeof ptrs, ptrs, NULL);
	if (e) return fail("clGetProgramInfo(binaries)", e);
	dst[size] = 0;
	return (long)size;
}

static int new_mem(cl_handle m) {
	for (int i = 0; i < g_nmem; ++i)
		if (!g_mem[i]) {
			g_mem[i] = m;
			return i;
		}
	if (g_nmem >= MAX_OBJ) return fail("too many memory objects", 0);
	g_mem[g_nmem] = m;
	return g_nmem++;
}

int ocl_buffer(size_t bytes, const void *init) {
	cl_int e;
	cl_handle m = ((fn_CreateBuffer)g_tab[D_CreateBuffer])(g_ctx, CL_MEM_READ_WRITE, bytes, NULL, &e);
	if (e) return fail("clCreateBuffer", e);
	if (init) {
		e = ((fn_RWBuffer)g_tab[D_EnqueueWriteBuffer])(g_queue, m, 1, 0, bytes, (void *)init, 0, NULL, NULL);
		if (e) return fail("clEnqueueWriteBuffer", e);
	}
	return new_mem(m);
}

/* RGBA float32 image2d_t, as nodencl makes of a buffer created with imageDims */
int ocl_image(int w, int h, const float *init) {
	cl_int e;
	cl_image_format f = {CL_RGBA, CL_FLOAT};
	cl_handle m = ((fn_CreateImage2D)g_tab[D_CreateImage2D])(g_ctx, CL_MEM_READ_WRITE, &f, (size_t)w, (size_t)h, 0, NULL, &e);
	if (e) return fail("clCreateImage2D", e);
	if (init) {
		size_t o[3] = {0, 0, 0}, r[3] = {(size_t)w, (size_t)h, 1};
		e = ((fn_RWImage)g_tab[D_EnqueueWriteImage])(g_queue, m, 1, o, r, 0, 0, (void *)init, 0, NULL, NULL);
		if (e) return fail("clEnqueueWriteImage", e);
	}
	return new_mem(m);
}

int ocl_read_buffer(int mem, void *dst, size_t bytes) {
	cl_int e = ((fn_RWBuffer)g_tab[D_EnqueueReadBuffer])(g_queue, g_mem[mem], 1, 0, bytes, dst, 0, NULL, NULL);
	return e ? fail("clEnqueueReadBuffer", e) : 0;
}

int ocl_write_buffer(int mem, const void *src, size_t bytes) {
	cl_int e = ((fn_RWBuffer)g_tab[D_EnqueueWriteBuffer])(g_queue, g_mem[mem], 1, 0, bytes, (void *)src, 0, NULL, NULL);
	return e ? fail("clEnqueueWriteBuffer", e) : 0;
}

int ocl_read_image(int mem, float *dst, int w, int h) {
	size_t o[3] = {0, 0, 0}, r[3] = {(size_t)w, (size_t)h, 1};
	cl_int e = ((fn_RWImage)g_tab[D_EnqueueReadImage])(g_queue, g_mem[mem], 1, o, r, 0, 0, dst, 0, NULL, NULL);
	return e ? fail("clEnqueueReadImage", e) : 0;
}

int ocl_release(int mem) {
	if (mem < 0 || mem >= g_nmem || !g_mem[mem]) return 0;
	((fn_Release)g_tab[D_ReleaseMemObject])(g_mem[mem]);
	g_mem[mem] = NULL;
	return 0;
}

int ocl_arg_mem(int kern, int idx, int mem) {
	cl_int e = ((fn_SetKernelArg)g_tab[D_SetKernelArg])(g_kern[kern], (cl_uint)idx, sizeof(cl_handle), &g_mem[mem]);
	return e ? fail("clSetKernelArg(mem)", e) : 0;
}
int ocl_arg_u32(int kern, int idx, uint32_t v) {
	cl_int e = ((fn_SetKernelArg)g_tab[D_SetKernelArg])(g_kern[kern], (cl_uint)idx, 4, &v);
	return e ? fail("clSetKernelArg(u32)", e) : 0;
}
int ocl_arg_f32(int kern, int idx, float v) {
	cl_int e = ((fn_SetKernelArg)g_tab[D_SetKernelArg])(g_kern[kern], (cl_uint)idx, 4, &v);
	return e ? fail("clSetKernelArg(f32)", e) : 0;
}

/* enqueue only (for timing the reference's launch sequence back to back); ocl_finish() waits */
int ocl_enqueue(int kern, int dims, size_t g0, size_t g1, size_t local0) {
	size_t g[2] = {g0, g1}, l[2] = {local0, 1};
	cl_int e = ((fn_NDRange)g_tab[D_EnqueueNDRangeKernel])(g_queue, g_kern[kern], (cl_uint)dims, NULL, g, local0 ? l : NULL, 0, NULL, NULL);
	return e ? fail("clEnqueueNDRangeKernel", e) : 0;
}
int ocl_finish(void) {
	cl_int e = ((fn_Finish)g_tab[D_Finish])(g_queue);
	return e ? fail("clFinish", e) : 0;
}

/* globalWorkItems / workItemsPerGroup as nodencl's createProgram takes them (local0 = 0: let the driver choose) */
int ocl_run(int kern, int dims, size_t g0, size_t g1, size_t local0) {
	size_t g[2] = {g0, g1}, l[2] = {local0, 1};
	cl_int e = ((fn_NDRange)g_tab[D_EnqueueNDRangeKernel])(g_queue, g_kern[kern], (cl_uint)dims, NULL, g, local0 ? l : NULL, 0, NULL, NULL);
	if (e) return fail("clEnqueueNDRangeKernel", e);
	e = ((fn_Finish)g_tab[D_Finish])(g_queue);
	return e ? fail("clFinish", e) : 0;
}
