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
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
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


// ---- eight int16 predictions in one 16-byte vector; flattened context recipes (zpq_hcomp.cpp) ----
__device__ __forceinline__ int vget(const uint4& v, int bit) {
  const u32 w = (bit & 4) ? ((bit & 2) ? v.w : v.z) : ((bit & 2) ? v.y : v.x);
  return (int)(short)(w >> ((bit & 1) * 16));
}
__device__ __forceinline__ void vset(uint4& v, int bit, int p) {
  const u32 sh = (bit & 1) * 16;
  const u32 keep = ~(0xffffu << sh), val = ((u32)p & 0xffffu) << sh;
  if (bit & 4) { if (bit & 2) v.w = (v.w & keep) | val; else v.z = (v.z & keep) | val; }
  else         { if (bit & 2) v.y = (v.y & keep) | val; else v.x = (v.x & keep) | val; }
}

// Flattened recipe (zpq_hcomp.cpp: build_wave_plan / FlatEval).  st[] = private segment states.
// The op words are fetched four at a time (16-byte loads from the read-only path) so that the
// interpreter loop does not pay one L1 round trip per op.
struct FlatState { u32 acc, prev_new, prev_old; int k; bool skip; u32 w_pending; };

// Byte sources older than 8 bytes: `in` is either the input stream (encode: in[bpos-off]) or, when
// ring_mask != 0, the HCOMP M ring (decode: m[(C+off) & mask], exactly what the bytecode reads).
__device__ __forceinline__ void flat_op(FlatState& F, u32 w, u32 wnext, u32 st[ZPQ_FLAT_SEGS], u32 cur, u64 hs, int bpos,
                                        const u8* in, u32 dist_last, u32 vc, u32 ring_mask = 0) {
  if (F.skip) { F.skip = false; return; }          // second word of a ZR_WORD op
  const u32 op = w & 15u, src = (w >> 4) & 15u, arg = (w >> 8) & 255u, off = w >> 16;
  if (op == ZR_END) return;
  if (op == ZR_SEG) {
    if (F.k >= 0) {
      const u32 old = F.k == 0 ? st[0] : (F.k == 1 ? st[1] : (F.k == 2 ? st[2] : st[3]));
      if (F.k == 0) st[0] = F.acc; else if (F.k == 1) st[1] = F.acc; else if (F.k == 2) st[2] = F.acc; else st[3] = F.acc;
      F.prev_old = old; F.prev_new = F.acc;
    }
    ++F.k;
    const u32 own = F.k == 0 ? st[0] : (F.k == 1 ? st[1] : (F.k == 2 ? st[2] : st[3]));
    F.acc = src == ZB_ZERO ? 0u : (src == ZB_OWN ? own : F.prev_new);
    return;
  }
  u32 x = 0;
  if (src == ZS_BYTE) {
    if (off < 8) x = (u32)(hs >> (off * 8)) & arg;
    else if (ring_mask) x = (u32)in[(vc + off) & ring_mask] & arg;
    else x = ((int)off <= bpos) ? ((u32)in[bpos - (int)off] & arg) : 0u;
  }
  else if (src == ZS_POSAND) x = vc & arg;
  else if (src == ZS_POSMOD) x = arg ? vc % arg : 0u;
  else if (src == ZS_DIST) x = min(dist_last - vc, 255u);
  if (op == ZR_FOLD) F.acc = (F.acc + x + 512u) * 773u;
  else if (op == ZR_SHL8ADD) F.acc = (F.acc << 8) + x;
  else if (op == ZR_SHR) F.acc >>= arg;
  else if (op == ZR_SHL) F.acc <<= arg;
  else if (op == ZR_ROLL) F.acc = F.acc * off + cur + 1u;
  else if (op == ZR_WORD) {
    const u32 v2 = off & 255u, v3 = off >> 8, mult = wnext & 255u, first = (wnext >> 8) & 1u;
    if ((((cur & arg) - v2) & 255u) < v3) F.acc = F.acc * mult + cur + 1u;
    else F.acc = first ? 0u : F.prev_old;
    F.skip = true;
  }
}

__device__ __forceinline__ u32 flat_eval(const u32* __restrict__ ops, u32 len, u32 st[ZPQ_FLAT_SEGS], u32 cur, u64 hs,
                                         int bpos, const u8* in, u32 dist_last, u32 ring_mask = 0) {
  const u32 vc = (u32)(-(bpos + 1));
  FlatState F; F.acc = 0; F.prev_new = 0; F.prev_old = 0; F.k = -1; F.skip = false; F.w_pending = 0;
  const uint4* __restrict__ o4 = reinterpret_cast<const uint4*>(ops);     // rows are 160 bytes: 16-byte aligned
  uint4 q = __ldg(o4);
  for (u32 t = 0; t < len; t += 4) {
    const uint4 qn = (t + 4 < len) ? __ldg(o4 + (t >> 2) + 1) : make_uint4(0, 0, 0, 0);
    flat_op(F, q.x, q.y, st, cur, hs, bpos, in, dist_last, vc, ring_mask);
    flat_op(F, q.y, q.z, st, cur, hs, bpos, in, dist_last, vc, ring_mask);
    flat_op(F, q.z, q.w, st, cur, hs, bpos, in, dist_last, vc, ring_mask);
    flat_op(F, q.w, qn.x, st, cur, hs, bpos, in, dist_last, vc, ring_mask);
    q = qn;
  }
  if (F.k == 0) st[0] = F.acc; else if (F.k == 1) st[1] = F.acc; else if (F.k == 2) st[2] = F.acc; else if (F.k == 3) st[3] = F.acc;
  return F.acc;
}


// The same interpreter with one op per iteration and no unrolling: a quarter of the code of flat_eval (which is inlined
// at every call site of every role of the wavefront encoder; the instruction caches of an SM hold all resident roles).
__device__ __forceinline__ u32 flat_eval_compact(const u32* __restrict__ ops, u32 len, u32 st[ZPQ_FLAT_SEGS], u32 cur, u64 hs,
                                                 int bpos, const u8* in, u32 dist_last, u32 ring_mask = 0) {
  const u32 vc = (u32)(-(bpos + 1));
  FlatState F; F.acc = 0; F.prev_new = 0; F.prev_old = 0; F.k = -1; F.skip = false; F.w_pending = 0;
  u32 w = len ? __ldg(ops) : 0u;
#pragma unroll 1
  for (u32 t = 0; t < len; ++t) {
    const u32 wn = (t + 1 < len) ? __ldg(ops + t + 1) : 0u;
    flat_op(F, w, wn, st, cur, hs, bpos, in, dist_last, vc, ring_mask);
    w = wn;
  }
  if (F.k == 0) st[0] = F.acc; else if (F.k == 1) st[1] = F.acc; else if (F.k == 2) st[2] = F.acc; else if (F.k == 3) st[3] = F.acc;
  return F.acc;
}

// Recipe of a wavefront-encoder thread: the order-n byte hash (ZWave.fold_n = n + 1) straight from the history register,
// any other recipe through the flat interpreter.  hs = the last eight bytes, newest in the low byte.
template <bool COMPACT = false>
__device__ __forceinline__ u32 wave_hash(u32 fold_n1, const u32* __restrict__ ops, u32 len, u32 st[ZPQ_FLAT_SEGS], u32 cur, u64 hs,
                                         int bpos, const u8* in, u32 dist_last) {
  if (fold_n1) {
    u32 acc = 0;
    for (u32 k = 0; k + 1 < fold_n1; ++k) acc = (acc + ((u32)(hs >> (8u * k)) & 255u) + 512u) * 773u;
    st[0] = acc;
    return acc;
  }
  if (COMPACT) return flat_eval_compact(ops, len, st, cur, hs, bpos, in, dist_last);
  return flat_eval(ops, len, st, cur, hs, bpos, in, dist_last);
}

// Slot-scheduled hashes (ZHashPlan, zpq_dev.h): the recipes of all lanes as ONE sequence of slots with a warp-uniform op
// kind, so the loop below has no per-lane control flow apart from the rare byte sources older than 8 bytes and the
// position-modulo source.  `mine_mask` = lanes of this warp that own a component (slots without any of them are skipped),
// `vc` = the HCOMP byte counter (0 - bytes seen), `ring` = the M ring the bytecode reads old bytes from.
// Returns the lane's new hash (h_old for a lane that owns nothing).
__device__ __forceinline__ u32 hplan_eval(const ZHashPlan* __restrict__ P, u32 mine_mask, int lane, u32 last_seg, u32 fold_n1,
                                          u32 st[ZPQ_FLAT_SEGS], u32 cur, u64 hs, u32 vc, const u8* ring, u32 ring_mask,
                                          u32 dist_last, u32 h_old) {
  // lanes whose recipe is the order-n byte hash (fold_n1 = n + 1): one predicated 8-step loop for all of them
  u32 accf = 0;
  if (__ldg(&P->fold_lanes) & mine_mask) {               // warp-uniform
#pragma unroll
    for (u32 k = 0; k < 8; ++k) {
      const u32 nx = (accf + ((u32)(hs >> (8u * k)) & 255u) + 512u) * 773u;
      if (k + 1u < fold_n1) accf = nx;
    }
  }
  u32 acc = 0, prev_new = 0, prev_old = 0;
  const int ns = (int)P->nslots;
  const uint2* __restrict__ sd = reinterpret_cast<const uint2*>(P->slot);
  const u32* __restrict__ wp = &P->w[0][lane];
  uint2 dn = __ldg(sd);
  u32 wn = __ldg(wp);
  for (int s = 0; s < ns; ++s) {
    const u32 lm = dn.x, kind = dn.y, w = wn;
    if (s + 1 < ns) { dn = __ldg(sd + s + 1); wn = __ldg(wp + (s + 1) * 32); }     // next slot requested one slot early
    if (!(lm & mine_mask)) continue;                      // warp-uniform
    const bool on = (lm >> lane) & 1u;
    const u32 src = (w >> 4) & 15u, arg = (w >> 8) & 255u, off = w >> 16;
    if (kind == ZH_FOLD_BYTE) {
      const u32 x = (u32)(hs >> ((off & 7u) * 8u)) & arg;
      if (on) acc = (acc + x + 512u) * 773u;
    }
    else if (kind == ZH_SEG) {
      const u32 k = arg & 3u;
      if (on) {
        if (k > 0u) {
          const u32 old = k == 1u ? st[0] : (k == 2u ? st[1] : st[2]);
          if (k == 1u) st[0] = acc; else if (k == 2u) st[1] = acc; else st[2] = acc;
          prev_old = old; prev_new = acc;
        }
        const u32 own = k == 0u ? st[0] : (k == 1u ? st[1] : (k == 2u ? st[2] : st[3]));
        acc = src == ZB_ZERO ? 0u : (src == ZB_OWN ? own : prev_new);
      }
    }
    else if (kind == ZH_SHL8_BYTE) {
      const u32 x = (u32)(hs >> ((off & 7u) * 8u)) & arg;
      if (on) acc = (acc << 8) + x;
    }
    else if (kind == ZH_FOLD_OTHER || kind == ZH_SHL8_OTHER) {
      if (on) {
        u32 x = 0;
        if (src == ZS_BYTE) x = (u32)ring[(vc + off) & ring_mask] & arg;      // older than eight bytes: from the M ring
        else if (src == ZS_POSAND) x = vc & arg;
        else if (src == ZS_POSMOD) x = arg ? vc % arg : 0u;
        else if (src == ZS_DIST) x = min(dist_last - vc, 255u);
        acc = kind == ZH_FOLD_OTHER ? (acc + x + 512u) * 773u : (acc << 8) + x;
      }
    }
    else if (kind == ZH_SHR) { if (on) acc >>= arg; }
    else if (kind == ZH_SHL) { if (on) acc <<= arg; }
    else if (kind == ZH_ROLL) { if (on) acc = acc * off + cur + 1u; }
    else if (kind == ZH_WORD) {
      const u32 w2 = __ldg(&P->w2[s][lane]);
      const u32 v2 = off & 255u, v3 = off >> 8, mult = w2 & 255u, first = (w2 >> 8) & 1u;
      if (on) {
        if ((((cur & arg) - v2) & 255u) < v3) acc = acc * mult + cur + 1u;
        else acc = first ? 0u : prev_old;
      }
    }
  }
  if (!((mine_mask >> lane) & 1u)) return h_old;
  if (fold_n1) { st[0] = accf; return accf; }
  if (last_seg == 0u) st[0] = acc; else if (last_seg == 1u) st[1] = acc; else if (last_seg == 2u) st[2] = acc; else st[3] = acc;
  return acc;
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
