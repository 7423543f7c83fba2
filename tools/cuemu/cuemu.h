// cuemu.h — functional CPU emulation of the CUDA execution model, for TESTS ONLY.
//
// Purpose: this container has no GPU and a round's GPU minutes are finite.  The kernels of oak_b200/csrc are
// compiled a second time with g++ against this header (tools/cuemu/build_emu.py) into liboak_b200_emu.so, which
// has the same C ABI, so the GPU parity tests can exercise the *kernel sources themselves* — block/warp
// decomposition, shared-memory indexing, shuffles, barriers, mma fragment layouts — on the CPU before they
// are spent GPU time on.  It is a checker like compute-sanitizer, not a product path: nothing under oak_b200/
// builds or loads it, oak_b200/_lib.py refuses an emulated library unless OAK_B200_TEST_EMU=1 is set by a test,
// and no performance number can come from it.
//
// Model: one thread block at a time; every CUDA thread is a ucontext fiber with its own stack; a fiber runs
// until it reaches a synchronisation point (__syncthreads, warp collectives) and then yields to the next one.
// `__shared__` variables are function-local statics (valid because blocks run one after the other).  The fiber
// order can be permuted per block (CUEMU_SHUFFLE=seed) to expose missing barriers.  Streams are synchronous,
// device memory is host memory.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <type_traits>

#define OAK_CUEMU 1

// ---------------------------------------------------------------------------------------------------------
// language keywords
// ---------------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 { double x, y; };
struct alignas(8) float2 { float x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

namespace cuemu {
struct ThreadCtx {
  uint3 tid;
  int linear;
};
const ThreadCtx &cur();
const uint3 &block_idx();
const dim3 &block_dim();
const dim3 &grid_dim();
void *dyn_smem();
void sync_block();
void sync_warp(unsigned mask);
uint64_t shfl(unsigned mask, uint64_t v, int src_lane_of_me);   // src lane absolute in the warp
unsigned ballot(unsigned mask, bool pred);
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
long launches();
}  // namespace cuemu

#define threadIdx (cuemu::cur().tid)
#define blockIdx (cuemu::block_idx())
#define blockDim (cuemu::block_dim())
#define gridDim (cuemu::grid_dim())
#define warpSize 32

// ---------------------------------------------------------------------------------------------------------
// synchronisation and warp collectives
// ---------------------------------------------------------------------------------------------------------
static inline void __syncthreads() { cuemu::sync_block(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { cuemu::sync_warp(mask); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

namespace cuemu {
template <class T> inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle of > 8 bytes");
  uint64_t b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T> inline T from_bits(uint64_t b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}
inline int lane_id() { return cur().linear & 31; }
}  // namespace cuemu

template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  const int lane = cuemu::lane_id();
  const int base = lane & ~(width - 1);
  return cuemu::from_bits<T>(cuemu::shfl(mask, cuemu::to_bits(v), base + (src & (width - 1))));
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  const int lane = cuemu::lane_id();
  int src = lane ^ lanemask;
  if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
  return cuemu::from_bits<T>(cuemu::shfl(mask, cuemu::to_bits(v), src));
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = cuemu::lane_id();
  int src = lane + (int)delta;
  if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
  return cuemu::from_bits<T>(cuemu::shfl(mask, cuemu::to_bits(v), src));
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  const int lane = cuemu::lane_id();
  int src = lane - (int)delta;
  if (src < (lane & ~(width - 1))) src = lane;
  return cuemu::from_bits<T>(cuemu::shfl(mask, cuemu::to_bits(v), src));
}
static inline unsigned __ballot_sync(unsigned mask, int pred) { return cuemu::ballot(mask, pred != 0); }
static inline int __any_sync(unsigned mask, int pred) { return cuemu::ballot(mask, pred != 0) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return cuemu::ballot(mask, pred == 0) == 0; }
static inline unsigned __activemask() { return 0xffffffffu; }

// fp64 tensor-core instruction mma.sync.aligned.m8n8k4.row.col.f64 (PTX ISA, "Matrix fragments for
// mma.m8n8k4 with .f64"): D(8x8) = A(8x4) B(4x8) + C, lane = 4 g + t:
//   a = A[g][t] ; b = B[t][g] ; c0,c1 / d0,d1 = C/D[g][2t], [g][2t+1]
// every product-sum is one fused multiply-add in k order, as the hardware chain does.
static inline void cuemu_dmma_m8n8k4(double &d0, double &d1, double a, double b, double c0, double c1) {
  const int lane = cuemu::lane_id(), g = lane >> 2, t = lane & 3;
  double acc0 = c0, acc1 = c1;
  for (int k = 0; k < 4; k++) {
    const double ak = __shfl_sync(0xffffffffu, a, 4 * g + k);        // A[g][k]
    const double b0 = __shfl_sync(0xffffffffu, b, 4 * (2 * t) + k);  // B[k][2t]   held by lane 4*(2t)+k
    const double b1 = __shfl_sync(0xffffffffu, b, 4 * (2 * t + 1) + k);
    acc0 = fma(ak, b0, acc0);
    acc1 = fma(ak, b1, acc1);
  }
  d0 = acc0;
  d1 = acc1;
}

// ---------------------------------------------------------------------------------------------------------
// integer / floating-point intrinsics
// ---------------------------------------------------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __drcp_rn(double a) { return 1. / a; }
static inline double rsqrt(double x) { return 1. / sqrt(x); }
static inline float rsqrtf(float x) { return 1.f / sqrtf(x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline double __longlong_as_double(long long v) { return cuemu::from_bits<double>((uint64_t)v); }
static inline long long __double_as_longlong(double v) { return (long long)cuemu::to_bits(v); }
static inline int __double2hiint(double v) { return (int)(cuemu::to_bits(v) >> 32); }
static inline int __double2loint(double v) { return (int)(cuemu::to_bits(v) & 0xffffffffu); }
static inline double __hiloint2double(int hi, int lo) {
  return cuemu::from_bits<double>(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo);
}
static inline int __float_as_int(float v) { return cuemu::from_bits<int>(cuemu::to_bits(v)); }
static inline float __int_as_float(int v) { return cuemu::from_bits<float>(cuemu::to_bits(v)); }
static inline unsigned __float_as_uint(float v) { return cuemu::from_bits<unsigned>(cuemu::to_bits(v)); }
static inline float __uint_as_float(unsigned v) { return cuemu::from_bits<float>(cuemu::to_bits(v)); }
template <class T> static inline T __ldg(const T *p) { return *p; }

// CUDA declares min/max for every arithmetic combination in the global namespace
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
static inline std::common_type_t<A, B> min(A a, B b) { return b < a ? b : a; }
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
static inline std::common_type_t<A, B> max(A a, B b) { return a < b ? b : a; }

// atomics: fibers never run concurrently, plain read-modify-write is atomic
template <class T, class U> static inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> static inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U> static inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T> static inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

// ---------------------------------------------------------------------------------------------------------
// runtime API subset (synchronous; device memory = host memory)
// ---------------------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
struct cuemu_stream_t { int id; };
struct cuemu_event_t { double t; };
typedef cuemu_stream_t *cudaStream_t;
typedef cuemu_event_t *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; char name[64]; };
struct cudaIpcMemHandle_t { char reserved[64]; };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
  memset(p, 0, sizeof *p);
  p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
  snprintf(p->name, sizeof p->name, "cuemu (CPU emulation, tests only)");
  return cudaSuccess;
}
static inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -5; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t bytes) {
  void *q = nullptr;
  if (posix_memalign(&q, 256, bytes ? bytes : 256) != 0) { *p = nullptr; return cudaErrorMemoryAllocation; }
  memset(q, 0xFF, bytes ? bytes : 256);  // uninitialised device memory is not zero: NaN pattern, -1 as an integer
  *p = q;
  return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)p, bytes); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind) {
  for (size_t r = 0; r < h; r++) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t = nullptr) {
  return cudaMemcpy2D(d, dp, s, sp, w, h, k);
}
static inline cudaError_t cudaMemset(void *p, int v, size_t n) { if (n) memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new cuemu_stream_t{1}; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = new cuemu_stream_t{1}; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new cuemu_stream_t{1}; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
double cuemu_now_ms();
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cuemu_event_t{0.}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new cuemu_event_t{0.}; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = cuemu_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof *h); memcpy(h, &p, sizeof p); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, &h, sizeof *p); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
// page-locked host memory: plain heap memory here (the pointer-attribute query says "unregistered")
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) { a->type = cudaMemoryTypeUnregistered; a->device = 0; a->devicePointer = nullptr; a->hostPointer = const_cast<void *>(p); return cudaSuccess; }
constexpr unsigned cudaHostRegisterPortable = 1, cudaHostAllocPortable = 1;
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned) { *p = malloc(bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
