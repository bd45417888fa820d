// Minimal single-process CUDA execution-model emulator — TEST INFRASTRUCTURE ONLY.
//
// Lets tests/ compile the *unchanged* kernel sources of laps_b200/csrc with g++ and run them on
// the CPU of the build container (which has no GPU), so that index math, shared-memory staging,
// barriers and launch geometry are validated before a GPU call is spent.  One CTA runs at a time;
// its threads are ucontext fibers that yield at __syncthreads().  Nothing in the product package
// includes or links this file; the shipped library is the nvcc build and fails loudly without it.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define LAPS_EMU 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define __constant__ static

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3e { unsigned x, y, z; };
struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
struct alignas(16) double4e { double x, y, z, w; };

namespace emu {
extern uint3e g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char* g_dyn_smem;
void yield_barrier();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
extern double g_shfl_scratch[2048];
extern unsigned long long g_shfl_scratch_u[2048];
// "device" memory is POSIX shared memory so that another emulator PROCESS (a second rank of a
// gloo test) can map it through the cudaIpc* stand-ins below.
void* shm_alloc(size_t n);
void shm_free(void* p);
int shm_export(void* p, char name_out[64]);
void* shm_import(const char name[64]);
void shm_unmap(void* p);
void spin_pause();
unsigned long long now_ns();
}  // namespace emu

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

static inline void __syncthreads() { emu::yield_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() {}
static inline void __threadfence_system() {}

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __drcp_rn(double a) { return 1.0 / a; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }

// Block-wide lock-step shuffle emulation: every thread of the CTA must execute the call.
static inline double __shfl_xor_sync(unsigned, double v, int lanemask) {
  unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  emu::g_shfl_scratch[tid] = v;
  emu::yield_barrier();
  unsigned src = (tid & ~31u) | ((tid ^ (unsigned)lanemask) & 31u);
  unsigned nthr = blockDim.x * blockDim.y * blockDim.z;
  double r = src < nthr ? emu::g_shfl_scratch[src] : v;
  emu::yield_barrier();
  return r;
}
static inline double __shfl_down_sync(unsigned, double v, int delta) {
  unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  emu::g_shfl_scratch[tid] = v;
  emu::yield_barrier();
  unsigned lane = tid & 31u;
  unsigned src = tid + delta;
  unsigned nthr = blockDim.x * blockDim.y * blockDim.z;
  double r = (lane + delta < 32 && src < nthr) ? emu::g_shfl_scratch[src] : v;
  emu::yield_barrier();
  return r;
}

static inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
  unsigned long long o = *p; if (v < o) *p = v; return o;
}
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  unsigned long long o = *p; if (v > o) *p = v; return o;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }

static inline double fma(double a, double b, double c, int) { return std::fma(a, b, c); }
static inline double fmin_(double a, double b) { return a < b ? a : b; }

// ---------------------------------------------------------------- runtime API stand-ins
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorPeerAccessAlreadyEnabled = 704 };
struct cudaIpcMemHandle_t { char reserved[64]; };
#define cudaIpcMemLazyEnablePeerAccess 1
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize, cudaFuncAttributePreferredSharedMemoryCarveout };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = emu::shm_alloc(n ? n : 1); return *p ? 0 : 2; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { emu::shm_free(p); return 0; }
template <class T> static inline cudaError_t cudaMallocAsync(T** p, size_t n, cudaStream_t) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { if (p) emu::shm_free(p); return 0; }
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { return emu::shm_export(p, h->reserved); }
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { *p = emu::shm_import(h.reserved); return *p ? 0 : 2; }
static inline cudaError_t cudaIpcCloseMemHandle(void* p) { emu::shm_unmap(p); return 0; }
static inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return 0; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? 0 : 2; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
#define cudaEventDisableTiming 2
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
#define cudaStreamNonBlocking 1

#define LAPS_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
