// zpq_kcommon.cuh -- device helpers shared by the fast decode kernel
// (zpq_kernels2.cu) and the skewed-pipeline encode kernel (zpq_kernels3.cu).
#ifndef ZPQ_KCOMMON_CUH
#define ZPQ_KCOMMON_CUH
#include <cuda_runtime.h>
#include <stdint.h>
#include "zpq_dev.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

#define FULL 0xffffffffu
#define ST_OK 0
#define ST_VM (-3)
#define ST_FULL (-4)
#define ST_CORRUPT (-5)
#define ST_EOF (-6)
#define ST_EOS (-11)

// shared-memory layout of the CTA-wide tables
#define S2_STRETCH 0                       // i16[16384]  stretcht[16384..32767]
#define S2_SQUASH (S2_STRETCH + 32768)     // u16[1344]   squasht[1376..2719]
#define S2_DT (S2_SQUASH + 2688)           // u16[1024]   dt[i]/2
#define S2_NS (S2_DT + 2048)               // u16[256]    next0 | next1<<8
#define S2_DT2K (S2_NS + 512)              // u16[256]
#define S2_STATIC (S2_DT2K + 512)          // 38528

struct Tabs {
  const short* stretch_hi;
  const u16* squash;
  const u16* dt;
  const u16* ns;
  const u16* dt2k;
};

__device__ __forceinline__ int t_stretch(const Tabs& S, u32 x) {
  return x >= 16384u ? (int)S.stretch_hi[x - 16384u] : -(int)S.stretch_hi[16383u - x];
}
// squash(x), x in [-2048, 2047]: 0 below -672, 32767 from 672 up (libzpaq.cpp:1739-1741)
__device__ __forceinline__ int t_squash(const Tabs& S, int x) {
  const int i = x + 672;
  const int v = (int)S.squash[min(max(i, 0), 1343)];
  return i < 0 ? 0 : (i > 1343 ? 32767 : v);
}
__device__ __forceinline__ int clamp2k(int x) { return min(max(x, -2048), 2047); }
__device__ __forceinline__ int clamp512k(int x) { return min(max(x, -(1 << 19)), (1 << 19) - 1); }

__device__ __forceinline__ u32 row_get(const uint4& r, int k) {
  u32 w = (k & 8) ? ((k & 4) ? r.w : r.z) : ((k & 4) ? r.y : r.x);
  return (w >> ((k & 3) * 8)) & 255u;
}
__device__ __forceinline__ void row_set(uint4& r, int k, u32 v) {
  const u32 sh = (k & 3) * 8;
  const u32 keep = ~(255u << sh), val = v << sh;
  if (k & 8) { if (k & 4) r.w = (r.w & keep) | val; else r.z = (r.z & keep) | val; }
  else       { if (k & 4) r.y = (r.y & keep) | val; else r.x = (r.x & keep) | val; }
}

// Predictor::find (libzpaq.cpp:2072-2088)
__device__ __forceinline__ void d_find(u8* ht, u32 ht_mask, int shift, u32 cxt, uint4& row, u32& off) {
  const u32 chk = (cxt >> shift) & 255u;
  const u32 h0 = (cxt * 16u) & ht_mask, h1 = h0 ^ 16u, h2 = h0 ^ 32u;
  const uint4 r0 = *reinterpret_cast<const uint4*>(ht + h0);
  const uint4 r1 = *reinterpret_cast<const uint4*>(ht + h1);
  const uint4 r2 = *reinterpret_cast<const uint4*>(ht + h2);
  if ((r0.x & 255u) == chk) { row = r0; off = h0; return; }
  if ((r1.x & 255u) == chk) { row = r1; off = h1; return; }
  if ((r2.x & 255u) == chk) { row = r2; off = h2; return; }
  const u32 p0 = (r0.x >> 8) & 255u, p1 = (r1.x >> 8) & 255u, p2 = (r2.x >> 8) & 255u;
  if (p0 <= p1 && p0 <= p2) off = h0;
  else if (p1 < p2) off = h1;
  else off = h2;
  row = make_uint4(chk, 0, 0, 0);
}


// Stage the CTA-wide tables into shared memory (call with all threads, then __syncthreads()).
__device__ __forceinline__ void stage_tables(unsigned char* smem, const ZTables* T) {
  const uint4* src = reinterpret_cast<const uint4*>(T->stretch_hi);
  uint4* dst = reinterpret_cast<uint4*>(smem + S2_STRETCH);
  for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) dst[i] = src[i];
  u16* sq = reinterpret_cast<u16*>(smem + S2_SQUASH);
  for (int i = threadIdx.x; i < 1344; i += blockDim.x) sq[i] = T->squash[1376 + i];
  u16* dt = reinterpret_cast<u16*>(smem + S2_DT);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) dt[i] = (u16)(T->dt[i] >> 1);
  u16* ns = reinterpret_cast<u16*>(smem + S2_NS);
  u16* d2 = reinterpret_cast<u16*>(smem + S2_DT2K);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) { ns[i] = (u16)(T->ns[i] & 0xffffu); d2[i] = T->dt2k[i]; }
}
__device__ __forceinline__ Tabs make_tabs(const unsigned char* smem) {
  Tabs S;
  S.stretch_hi = reinterpret_cast<const short*>(smem + S2_STRETCH);
  S.squash = reinterpret_cast<const u16*>(smem + S2_SQUASH);
  S.dt = reinterpret_cast<const u16*>(smem + S2_DT);
  S.ns = reinterpret_cast<const u16*>(smem + S2_NS);
  S.dt2k = reinterpret_cast<const u16*>(smem + S2_DT2K);
  return S;
}
#endif
