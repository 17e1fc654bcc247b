// tools/simt/include/simt.h -- a small SIMT emulator: compiles the repo's CUDA sources with g++ and runs every CUDA
// thread as a fiber on the host, so that kernel LOGIC can be checked against the oracle on a machine without a GPU.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in zpaq_b200/ (the product) includes, links or loads this; the library built from
// it (tools/simt/_build/libzpq_simt.so) is only ever loaded by tools/simt/run.py.  It is a pre-flight check, not a
// parity claim: memory-model races, occupancy and timing are outside what it can see, and the `-m gpu` tests on a real
// B200 stay the gate.
//
// What is emulated: __global__ launches (grid of CTAs, one OS thread per CTA in flight, one fiber per CUDA thread),
// dynamic shared memory, __syncthreads, named barriers (bar.sync id, count), warp collectives (__shfl*_sync,
// __ballot_sync, __all_sync, __any_sync, __reduce_add_sync, __syncwarp) with exited lanes counted as arrived, spin
// waits on volatile shared words (the translator puts a yield into empty spin bodies), global atomics, and the slice of
// the runtime API the C-ABI layer uses (synchronous streams, malloc-backed device memory poisoned with 0xA5).
#ifndef ZPQ_SIMT_H
#define ZPQ_SIMT_H
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <functional>

// ---- language keywords ----
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local

// ---- vector types ----
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// ---- the running thread ----
struct SimtThreadView { uint3 tid, bid; dim3 bdim, gdim; };
extern "C" const SimtThreadView* simt_view();
#define threadIdx (simt_view()->tid)
#define blockIdx (simt_view()->bid)
#define blockDim (simt_view()->bdim)
#define gridDim (simt_view()->gdim)

extern "C" {
unsigned char* simt_dyn_smem();
// every lane deposits one 64-bit value; returns the warp's value array (valid until the lane's next collective)
// and the set of lanes that took part (mask minus exited lanes)
const uint64_t* simt_warp_exchange(unsigned mask, uint64_t v, unsigned* took_part);
void simt_syncthreads();
void simt_named_barrier(int id, int count);
void simt_named_arrive(int id, int count);
void simt_spin();
int simt_lane();
void simt_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}

// ---- warp collectives ----
template <typename T> static inline T simt_unpack(uint64_t v) { T r; memcpy(&r, &v, sizeof(T)); return r; }
template <typename T> static inline uint64_t simt_pack(T v) { static_assert(sizeof(T) <= 8, "shuffle payload"); uint64_t r = 0; memcpy(&r, &v, sizeof(T)); return r; }

template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  unsigned tp; const uint64_t* a = simt_warp_exchange(mask, simt_pack(v), &tp);
  const int lane = simt_lane();
  const int s = (lane & ~(width - 1)) | (src & (width - 1));
  return ((tp >> s) & 1u) ? simt_unpack<T>(a[s]) : v;
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  unsigned tp; const uint64_t* a = simt_warp_exchange(mask, simt_pack(v), &tp);
  const int lane = simt_lane();
  const int s = lane - (int)delta;
  if (s < (lane & ~(width - 1)) || !((tp >> s) & 1u)) return v;
  return simt_unpack<T>(a[s]);
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  unsigned tp; const uint64_t* a = simt_warp_exchange(mask, simt_pack(v), &tp);
  const int lane = simt_lane();
  const int s = lane + (int)delta;
  if (s > (lane | (width - 1)) || !((tp >> s) & 1u)) return v;
  return simt_unpack<T>(a[s]);
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
  unsigned tp; const uint64_t* a = simt_warp_exchange(mask, simt_pack(v), &tp);
  const int s = simt_lane() ^ x;
  (void)width;
  return ((tp >> s) & 1u) ? simt_unpack<T>(a[s]) : v;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  unsigned tp; const uint64_t* a = simt_warp_exchange(mask, pred ? 1u : 0u, &tp);
  unsigned r = 0;
  for (int l = 0; l < 32; ++l) if (((tp >> l) & 1u) && a[l]) r |= 1u << l;
  return r;
}
static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
  unsigned tp; const uint64_t* a = simt_warp_exchange(mask, v, &tp);
  unsigned r = 0;
  for (int l = 0; l < 32; ++l) if (((tp >> l) & 1u) && (unsigned)a[l] == v) r |= 1u << l;
  return r;
}
static inline int __all_sync(unsigned mask, int pred) {
  unsigned tp; const uint64_t* a = simt_warp_exchange(mask, pred ? 1u : 0u, &tp);
  for (int l = 0; l < 32; ++l) if (((tp >> l) & 1u) && !a[l]) return 0;
  return 1;
}
static inline int __any_sync(unsigned mask, int pred) {
  unsigned tp; const uint64_t* a = simt_warp_exchange(mask, pred ? 1u : 0u, &tp);
  for (int l = 0; l < 32; ++l) if (((tp >> l) & 1u) && a[l]) return 1;
  return 0;
}
static inline int __reduce_add_sync(unsigned mask, int v) {
  unsigned tp; const uint64_t* a = simt_warp_exchange(mask, simt_pack(v), &tp);
  unsigned r = 0;
  for (int l = 0; l < 32; ++l) if ((tp >> l) & 1u) r += (unsigned)simt_unpack<int>(a[l]);
  return (int)r;
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
  unsigned tp; const uint64_t* a = simt_warp_exchange(mask, simt_pack(v), &tp);
  unsigned r = 0;
  for (int l = 0; l < 32; ++l) if ((tp >> l) & 1u) r += simt_unpack<unsigned>(a[l]);
  return r;
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { unsigned tp; (void)simt_warp_exchange(mask, 0, &tp); }
static inline void __syncthreads() { simt_syncthreads(); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- scalar intrinsics ----
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { const uint64_t v = (uint64_t)hi << 32 | lo; return (unsigned)((v << (s & 31)) >> 32); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { const uint64_t v = (uint64_t)hi << 32 | lo; return (unsigned)(v >> (s & 31)); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  const uint64_t v = (uint64_t)b << 32 | a; unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    const unsigned sel = (s >> (4 * i)) & 15u; unsigned byte = (unsigned)(v >> ((sel & 7u) * 8)) & 255u;
    if (sel & 8u) byte = (byte & 128u) ? 255u : 0u;
    r |= byte << (8 * i);
  }
  return r;
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
long long clock64();

static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

// CUDA's global min/max overload set (mixed signedness promotes to unsigned, as in device code)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
static inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }

// ---- runtime API (synchronous; "device" memory is host memory) ----
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct SimtStream* cudaStream_t;
typedef struct SimtEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1 };
enum { cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
  char name[256];
  size_t totalGlobalMem, sharedMemPerBlockOptin, sharedMemPerMultiprocessor;
  int multiProcessorCount, major, minor, l2CacheSize, clockRate;
};
extern "C" {
cudaError_t cudaMalloc(void** p, size_t n);
cudaError_t cudaFree(void* p);
cudaError_t cudaMemGetInfo(size_t* fr, size_t* tot);
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st);
cudaError_t cudaMemset(void* d, int v, size_t n);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* st, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t st);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaGetLastError();
cudaError_t cudaPeekAtLastError();
const char* cudaGetErrorString(cudaError_t e);
}
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
#endif
