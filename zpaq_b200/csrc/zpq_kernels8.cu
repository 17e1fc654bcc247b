// zpq_kernels8.cu -- DECODE with lanes = BLOCKS: one CTA decodes up to eight blocks of one model in lock step, one warp
// per component (round 2), sm_100a.
//
// Why another mapping.  With lane = component (zpq_kernels6/7.cu) a block is two dependent instruction streams of ~500
// warp instructions per bit, and most of them are scalar work that all 32 lanes repeat (profiles/r02e_dec7.summary.txt:
// 990 instructions per block-bit at 9 cycles each, 14 warps and 1.7 of 4 issue slots per SM).  A block cannot get more
// warps there (three per block are 672 threads at 80 registers).  Here the transposition is the other way round:
//
//   * a CTA owns up to 8 blocks of ONE model and decodes bit t of all of them in the same step (the blocks of a batch
//     have the same structure: an end-of-stream flag and eight predicted bits per byte), LANE = BLOCK;
//   * every ICM / ISSE component gets its own warp (its code is the per-lane routine of zpq_kernels7.cu: two-outcome
//     speculation while the bit is being decoded, select + publish after it), CM / MATCH / CONST share one warp
//     (lane = slot * 8 + block), one warp walks the ISSE chains of the eight blocks serially in registers, one warp computes
//     the mixers' dot products (lane = block * 4 + quarter of the inputs), one warp runs MIX2 / SSE / AVG and the eight
//     arithmetic decoders;
//   * so every scalar instruction serves eight blocks, each warp's per-bit stream is short (the longest is a hash warp's
//     ~300), and ~20 warps per SM run side by side: the critical path is  y -> select (hash warps) -> chain -> mixers -> tail;
//   * predictions travel through shared memory (P[parity][block][component], int16; ISSE weight pairs XW[block][k]) and
//     four CTA-wide named barriers per bit: H (hash / misc published), C (chain published), M (mixers published), Y (bits
//     decoded).  Blocks that reach an end of segment hold their state for that step (valid = 0); finished blocks idle.
//
// Same operations on the same private tables in the same order as Predictor::predict0 / update0 (libzpaq.cpp:1854-2066):
// bit-exact; the `-m gpu` tests run this tier against the oracle and the golden archives.
//
// STATUS: opt-in (ZPQ_DEC_KERNEL=8), NOT the production decoder.  Measured on a B200 (profiles/r02i_*, r02j_*; 1 029 blocks
// x 32 KiB, method 5): 21.5 MB/s in its first form, 15.3 MB/s after the chain / mixer / tail roles were rewritten for
// registers and hoisted loads -- against 49.7 MB/s of zpq_kernels7.cu.  The estimate behind this design assumed that the
// short scalar roles would be latency-bound and would profit from instruction-level parallelism; the captures show that a
// single warp retires one instruction per ~8-9 cycles here whatever the dependencies (address arithmetic, predicates,
// descriptor decoding dominate), that the generic chain / mixer / tail roles execute 450-1 100 instructions per step, and
// that they sit one after the other on the critical path (hash select -> chain -> mixers -> tail: ~1 500 instructions, 24 of
// 33 stall cycles per instruction are barrier waits of the 16 hash warps).  Beating the two-warp decoder with this mapping
// needs those three roles at ~60-100 instructions per step, i.e. code specialised at compile time for one model structure.
#include <cuda_runtime.h>
#include <stdint.h>
#include "zpq_dev.h"
#include "zpq_kernels.h"
#include "zpq_kcommon.cuh"

// shared memory: stretch (32 KB) | next-state pairs (512 B) | the hash rows of the current nibble (16 B per block and hash
// component, lane stride 21 rows: conflict-free) ; squash and dt are read through the read-only path (full tables, no clamping)
#define S8_STRETCH 0
#define S8_NS (S8_STRETCH + 32768)
#define T_ROW_STRIDE 21                       // rows per block (20 hash components + 1: lanes fall on different banks)
#define S8_ROWS (S8_NS + 512)
#define S8_STATIC (S8_ROWS + T_NB * T_ROW_STRIDE * 16)
#define T_NB 8                               // blocks per CTA (lanes of a component warp)
#define T_P_BYTES (2 * T_NB * 32 * 2)        // short P[2][8][32]
#define T_NI 16
#define T_XW_BYTES (T_NB * T_NI * 8)         // int2 XW[8][16]
#define T_YV_BYTES (T_NB * 4 + 32)           // u32 YV[8] + meta
#define T_DESC_BYTES ((ZPQ_FAST_MIX + ZPQ_T_SCALAR) * 32)    // StepDesc of the mixers and the scalar steps
#define S8_XCHG (T_P_BYTES + T_XW_BYTES + T_YV_BYTES + T_DESC_BYTES)
#define BAR_H 1
#define BAR_C 2
#define BAR_M 3
#define BAR_Y 4
#define BAR_G 5
#define T_SYNC(id, n) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory")
#define T_ARRIVE(id, n) asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory")

struct ZTArgs {
  ZKernelArgs A;
  const uint32_t* group_first;     // first block (index into A.blocks) of each group
  const uint32_t* group_count;     // blocks in the group (<= nbk)
  int ngroups, nbk, pw_bytes, nwarps;
};

struct Dec8 {
  u32 low, high, curr;
  const u8* in; u64 in_pos, in_end;
  u8* out; u64 out_pos, out_cap;
  int status;
};

__device__ __forceinline__ int dec8_bit(Dec8& C, u32 p16) {
  if (C.curr < C.low || C.curr > C.high) { C.status = ST_CORRUPT; return 0; }
  const u32 mid = C.low + (u32)(((u64)(C.high - C.low) * p16) >> 16);
  int y;
  if (C.curr <= mid) { y = 1; C.high = mid; } else { y = 0; C.low = mid + 1; }
  while ((C.high ^ C.low) < 0x1000000u) {
    C.high = C.high << 8 | 255u;
    C.low = C.low << 8;
    C.low += (C.low == 0);
    if (C.in_pos >= C.in_end) { C.status = ST_EOF; return y; }
    C.curr = C.curr << 8 | (u32)C.in[C.in_pos++];
  }
  return y;
}

__device__ __forceinline__ void find8(const u8* ht, u32 ht_mask, int shift, u32 cxt, uint4& row, u32& off) {
  const u32 chk = (cxt >> shift) & 255u;
  const u32 h0 = (cxt * 16u) & ht_mask, h1 = h0 ^ 16u, h2 = h0 ^ 32u;
  const uint4 r0 = *reinterpret_cast<const uint4*>(ht + h0);
  const uint4 r1 = *reinterpret_cast<const uint4*>(ht + h1);
  const uint4 r2 = *reinterpret_cast<const uint4*>(ht + h2);
  const bool m0 = (r0.x & 255u) == chk, m1 = (r1.x & 255u) == chk, m2 = (r2.x & 255u) == chk;
  const u32 p0 = (r0.x >> 8) & 255u, p1 = (r1.x >> 8) & 255u, p2 = (r2.x >> 8) & 255u;
  const u32 repl = (p0 <= p1 && p0 <= p2) ? h0 : (p1 < p2 ? h1 : h2);
  if (m0) { row = r0; off = h0; }
  else if (m1) { row = r1; off = h1; }
  else if (m2) { row = r2; off = h2; }
  else { row = make_uint4(chk, 0, 0, 0); off = repl; }
}

// What every role follows of its blocks' byte streams (Predictor::update's tail, libzpaq.cpp:2054-2066)
struct Track { int c8, hmap4, ph; u64 hist; u32 vm_c; };
__device__ __forceinline__ void track_init(Track& K) { K.c8 = 1; K.hmap4 = 1; K.ph = 0; K.hist = 0; K.vm_c = 0; }
// returns true when the bit completed a byte (`byte` = its value)
__device__ __forceinline__ bool track_bit(Track& K, int y, u32& byte) {
  const int c8n = K.c8 * 2 + y;
  if (K.ph == 7) {
    byte = (u32)(c8n - 256);
    K.vm_c -= 1; K.hist = (K.hist << 8) | (u64)byte;
    K.c8 = 1; K.hmap4 = 1; K.ph = 0;
    return true;
  }
  K.c8 = c8n;
  if (K.ph == 3) K.hmap4 = (K.hmap4 & 0xf) << 5 | y << 4 | 1;
  else K.hmap4 = (K.hmap4 & 0x1f0) | (((K.hmap4 & 0xf) * 2 + y) & 0xf);
  ++K.ph;
  return false;
}

// context hash of one component for one block after a byte (the flattened recipe of the wavefront encoder, old bytes
// from the block's M ring as the HCOMP program reads them)
struct HRecipe { const u32* ops; u32 len, fold1, dist_byte; };
__device__ __forceinline__ u32 comp_hash(const HRecipe& R, u32 st[ZPQ_FLAT_SEGS], u32& dist_last, u32 byte, const Track& K,
                                         const u8* mring, u32 mmask) {
  if (byte == R.dist_byte) dist_last = K.vm_c;
  if (R.fold1) {
    u32 acc = 0;
    for (u32 k = 0; k + 1 < R.fold1; ++k) acc = (acc + ((u32)(K.hist >> (8u * k)) & 255u) + 512u) * 773u;
    st[0] = acc;
    return acc;
  }
  return flat_eval_compact(R.ops, R.len, st, byte, K.hist, (int)(0u - K.vm_c) - 1, mring, dist_last, mmask);
}

__device__ __forceinline__ int sq8(const u16* __restrict__ gsq, int x) { return (int)__ldg(gsq + (x + 2048)); }     // x in [-2048, 2047]

struct T8 {                       // per-thread view of the CTA's shared memory and the group
  Tabs S;
  u32* sm32;
  short* P;                       // [2][T_NB][32]
  int2* XW;                       // [T_NB][T_NI]
  u8* rows;                       // [T_NB][T_ROW_STRIDE][16]
  const u16* gsq;                 // squasht[4096] (global, read-only path)
  volatile u32* YV;               // [T_NB] : y | valid << 1 | done << 2 ;  YV[T_NB] = all blocks done
  const int* gdt;
  int nthreads;                   // all threads of the CTA: barriers Y (bits decoded) and G (group switch)
  int nsub;                       // barrier H: every warp but mix and tail; barrier C: every warp but misc and tail (both nthreads - 64)
};
#define PIDX(cur, b, c) (((cur) * T_NB + (b)) * 32 + (c))

// ------------------------------------------------------------------------------------------------------------------
// role: one ICM / ISSE component, lane = block
__device__ __forceinline__ void role_hash(const T8& X, const ZKernelArgs& A, const ZModel* __restrict__ ZM, const ZBlock* blks, int nbk,
                                          int c, int hk, int pw_words, u32 wbase0) {
  const int lane = threadIdx.x & 31;
  const bool act = lane < nbk;
  const ZComp cd = ZM->comp[c];
  const bool isse = cd.type == ZT_ISSE;
  const u32 a1 = cd.a1, ht_mask = cd.ht_mask;
  const int jin = (int)cd.a2;                                   // ISSE input component
  const int kx = (int)ZM->tplan.isse_ord[c];
  const u32 toff = ZM->fast.smem_off[c] / 4u;
  u32* const tab = X.sm32 + wbase0 + (u32)(act ? lane : 0) * (u32)pw_words + toff;      // ICM: u32[256]; ISSE: int2[256]
  u8* const rowp = X.rows + ((act ? lane : 0) * T_ROW_STRIDE + hk) * 16;                // this block's row of the current nibble
  const ZBlock B = blks[act ? lane : 0];
  u8* const arena = A.arena + B.arena_off;
  u8* const ht = arena + (u64)cd.ht_off16 * 16;
  const u8* const mring = arena + (u64)ZM->m_off16 * 16;
  const u32 mmask = ZM->m_mask;
  HRecipe R; R.ops = ZM->wave.flat[c]; R.len = ZM->wave.flat_len[c]; R.fold1 = ZM->wave.fold_n[c]; R.dist_byte = ZM->pipe.dist_byte[c];
  // tables of all blocks of the group: initial values (Predictor::init, libzpaq.cpp:1783-1790, 1831-1838)
  for (int b = 0; b < nbk; ++b) {
    const u32 tb = wbase0 + (u32)b * (u32)pw_words + toff;
    if (isse) for (int i = lane; i < 512; i += 32) X.sm32[tb + i] = A.tables->isse_init[i];
    else for (int i = lane; i < 256; i += 32) X.sm32[tb + i] = A.tables->icm_init[i];
  }
  __syncwarp();
  Track K; track_init(K);
  u32 st[ZPQ_FLAT_SEGS] = {0, 0, 0, 0}; u32 dist_last = 0, h = 0;
  u32 row_off = 0, hs = 1, bh = 0, v0 = 0, v1 = 0; int p_own = 0;
#define LOAD_CUR8()                                                                                  \
  do {                                                                                               \
    uint4 row_;                                                                                      \
    find8(ht, ht_mask, (int)a1 + 2, h + 16u * (u32)K.c8, row_, row_off);                             \
    *reinterpret_cast<uint4*>(rowp) = row_;                                                          \
    hs = 1u; bh = (row_.x >> 8) & 255u;                                                              \
    if (isse) { const int2 e_ = reinterpret_cast<const int2*>(tab)[bh]; v0 = (u32)e_.x; v1 = (u32)e_.y; } \
    else { v0 = tab[bh]; p_own = t_stretch(X.S, v0 >> 8); }                                         \
  } while (0)
  if (act) LOAD_CUR8();
  int cur = 0;
  for (;;) {
    // ---- publish the current bit's contribution ----
    if (act) {
      if (isse) X.XW[lane * T_NI + kx] = make_int2((int)v0, (int)v1);
      else X.P[PIDX(cur, lane, c)] = (short)p_own;
    }
    T_ARRIVE(BAR_H, X.nsub);
    // ---- candidates of the next bit (both outcomes), independent of this bit's predictions ----
    const bool in_nib = (K.ph & 3) != 3;
    u32 nsw = 0, bc0 = 0, bc1 = 0, vc0a = 0, vc0b = 0, vc1a = 0, vc1b = 0; int pc0 = 0, pc1 = 0;
    if (act) {
      nsw = (u32)X.S.ns[bh];
      if (in_nib) {
        const u32 two = (u32)*reinterpret_cast<const u16*>(rowp + hs * 2u);      // slots 2 hs and 2 hs + 1 are neighbours
        bc0 = two & 255u; bc1 = two >> 8;
        if (isse) {
          const int2 e0 = reinterpret_cast<const int2*>(tab)[bc0], e1 = reinterpret_cast<const int2*>(tab)[bc1];
          vc0a = (u32)e0.x; vc0b = (u32)e0.y; vc1a = (u32)e1.x; vc1b = (u32)e1.y;
        } else {
          vc0a = tab[bc0]; vc1a = tab[bc1];
          pc0 = t_stretch(X.S, vc0a >> 8); pc1 = t_stretch(X.S, vc1a >> 8);
        }
      }
      else if (K.ph == 3) {
        prefetch_l1(ht + (((h + 16u * (u32)(K.c8 * 2)) * 16u) & ht_mask & ~63u));
        prefetch_l1(ht + (((h + 16u * (u32)(K.c8 * 2 + 1)) * 16u) & ht_mask & ~63u));
      }
      if (K.ph == 2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) prefetch_l2(ht + (((h + 16u * (u32)(K.c8 * 4 + k)) * 16u) & ht_mask & ~63u));
      }
    }
    T_SYNC(BAR_C, X.nsub);                                        // the chain warp has published every ISSE output
    // ---- both trained values of the entry in use ----
    u32 tr0a = 0, tr0b = 0, tr1a = 0, tr1b = 0;
    if (act) {
      if (isse) {
        const int xj = (int)X.P[PIDX(cur, lane, jin)], p = (int)X.P[PIDX(cur, lane, c)];
        const int sq = sq8(X.gsq, p);
        const int e0 = -sq, e1 = 32767 - sq;
        tr0a = (u32)clamp512k((int)v0 + ((e0 * xj + (1 << 12)) >> 13));
        tr1a = (u32)clamp512k((int)v0 + ((e1 * xj + (1 << 12)) >> 13));
        tr0b = (u32)clamp512k((int)v1 + ((e0 + 16) >> 5));
        tr1b = (u32)clamp512k((int)v1 + ((e1 + 16) >> 5));
      } else {
        tr0a = v0 + (u32)(((int)(0 - (int)(v0 >> 8))) >> 2);
        tr1a = v0 + (u32)(((int)(32767 - (int)(v0 >> 8))) >> 2);
      }
    }
    T_SYNC(BAR_Y, X.nthreads);
    if (X.YV[T_NB]) break;
    const u32 yv = act ? X.YV[lane] : 0u;
    if (yv & 2u) {
      const int y = (int)(yv & 1u);
      const u32 tra = y ? tr1a : tr0a, trb = y ? tr1b : tr0b;
      const u32 nx = (nsw >> (8 * y)) & 255u;
      const u32 bh_old = bh, hs_old = hs;
      if (in_nib) {
        const u32 bn = y ? bc1 : bc0;
        const bool same = bn == bh_old;
        v0 = same ? tra : (y ? vc1a : vc0a);
        v1 = same ? trb : (y ? vc1b : vc0b);
        if (!isse) { p_own = y ? pc1 : pc0; if (same) p_own = t_stretch(X.S, tra >> 8); }
        bh = bn; hs = hs * 2u + (u32)y;
      }
      if (isse) reinterpret_cast<int2*>(tab)[bh_old] = make_int2((int)tra, (int)trb);
      else tab[bh_old] = tra;
      rowp[hs_old] = (u8)nx;
      if (!in_nib) *reinterpret_cast<uint4*>(ht + row_off) = *reinterpret_cast<const uint4*>(rowp);
      u32 byte = 0;
      const int ph_old = K.ph;
      if (track_bit(K, y, byte)) { h = comp_hash(R, st, dist_last, byte, K, mring, mmask); LOAD_CUR8(); }
      else if (ph_old == 3) LOAD_CUR8();
    }
    cur ^= 1;
  }
#undef LOAD_CUR8
}

// ------------------------------------------------------------------------------------------------------------------
// role: CM / MATCH / CONST components, lane = slot * 8 + block (up to four components)
__device__ __forceinline__ void role_misc(const T8& X, const ZKernelArgs& A, const ZModel* __restrict__ ZM, const ZBlock* blks, int nbk) {
  const int lane = threadIdx.x & 31;
  const int slot = lane >> 3, b = lane & 7;
  const int nm = (int)ZM->tplan.nmisc;
  const bool act = slot < nm && b < nbk;
  const int c = act ? (int)ZM->tplan.mcomp[slot] : 0;
  const ZComp cd = ZM->comp[c];
  const u32 type = act ? cd.type : (u32)ZT_NONE;
  const ZBlock B = blks[b < nbk ? b : 0];
  u8* const arena = A.arena + B.arena_off;
  u32* const gtab = reinterpret_cast<u32*>(arena + (u64)cd.cm_off16 * 16);
  u8* const ht = arena + (u64)cd.ht_off16 * 16;
  const u32 cm_mask = cd.cm_mask, ht_mask = cd.ht_mask, limit = cd.limit;
  const u8* const mring = arena + (u64)ZM->m_off16 * 16;
  const u32 mmask = ZM->m_mask;
  HRecipe R; R.ops = ZM->wave.flat[c]; R.len = ZM->wave.flat_len[c]; R.fold1 = ZM->wave.fold_n[c]; R.dist_byte = ZM->pipe.dist_byte[c];
  Track K; track_init(K);
  u32 st[ZPQ_FLAT_SEGS] = {0, 0, 0, 0}; u32 dist_last = 0, h = 0;
  int p_own = (type == ZT_CONS) ? ((int)cd.a1 - 128) * 4 : 0;
  u32 cxa = 0, v0 = 0;
  u32 ma = 0, mb = 0, mpos = 0, mbit = 0, mpred = 0, mdt = 0; int mexp = 0;
  if (type == ZT_CM) { cxa = (h ^ 1u) & cm_mask; v0 = gtab[cxa]; p_own = t_stretch(X.S, v0 >> 17); }
  int cur = 0;
  for (;;) {
    if (act) X.P[PIDX(cur, b, c)] = (short)p_own;
    T_ARRIVE(BAR_H, X.nsub);
    u32 cmc0 = 0, cmc1 = 0, cmx0 = 0, cmx1 = 0, cmt0 = 0, cmt1 = 0; int cmp0 = 0, cmp1 = 0, mp_ok = 0;
    if (type == ZT_CM) {
      const u32 count = v0 & 0x3ffu;
      const u32 dtv = (u32)__ldg(X.gdt + count), inc = count < limit ? 1u : 0u;
      cmt0 = v0 + (((u32)(0 - (int)(v0 >> 17)) * dtv) & 0xfffffc00u) + inc;
      cmt1 = v0 + (((u32)(32767 - (int)(v0 >> 17)) * dtv) & 0xfffffc00u) + inc;
      if (K.ph != 7) {
        u32 hm0;
        if (K.ph == 3) hm0 = (u32)((K.hmap4 & 0xf) << 5 | 1);
        else hm0 = (u32)((K.hmap4 & 0x1f0) | (((K.hmap4 & 0xf) * 2) & 0xf));
        const u32 hm1 = hm0 | (K.ph == 3 ? 16u : 1u);
        cmx0 = (h ^ hm0) & cm_mask; cmx1 = (h ^ hm1) & cm_mask;
        cmc0 = gtab[cmx0]; cmc1 = gtab[cmx1];
        cmp0 = t_stretch(X.S, cmc0 >> 17); cmp1 = t_stretch(X.S, cmc1 >> 17);
      }
    }
    else if (type == ZT_MATCH) {
      if (ma != 0 && mbit < 7) {
        const u32 mcn = (mpred >> (6 - mbit)) & 1u;
        mp_ok = t_stretch(X.S, (u32)((int)mdt * (1 - 2 * (int)mcn)) & 32767u);
      }
      if (K.ph == 0) prefetch_l1(gtab + (h & cm_mask));
    }
    T_SYNC(BAR_Y, X.nthreads);
    if (X.YV[T_NB]) break;
    const u32 yv = act ? X.YV[b] : 0u;
    if (yv & 2u) {
      const int y = (int)(yv & 1u);
      const int c8n = K.c8 * 2 + y;
      if (type == ZT_CM) {
        gtab[cxa] = y ? cmt1 : cmt0;
        if (K.ph != 7) {
          const u32 cxn = y ? cmx1 : cmx0;
          const bool same = cxn == cxa;
          const u32 tr = y ? cmt1 : cmt0;
          v0 = same ? tr : (y ? cmc1 : cmc0);
          p_own = y ? cmp1 : cmp0;
          if (same) p_own = t_stretch(X.S, tr >> 17);
          cxa = cxn;
        }
      }
      else if (type == ZT_MATCH) {
        // MATCH update (libzpaq.cpp:1990-2008); h is still the hash as of the previous byte here
        if (ma != 0 && mexp != y) ma = 0;
        if (++mbit == 8) {
          mbit = 0;
          ht[mpos] = (u8)(c8n & 255);
          mpos = (mpos + 1) & ht_mask;
          u32* idx = gtab + (h & cm_mask);
          if (ma == 0) {
            mb = mpos - *idx;
            if (mb & ht_mask)
              while (ma < 255 && ht[(mpos - ma - 1) & ht_mask] == ht[(mpos - ma - mb - 1) & ht_mask]) ++ma;
          }
          else ma += (ma < 255);
          *idx = mpos;
          mpred = ht[(mpos - mb) & ht_mask];
          mdt = ma ? 2048u / ma : 0u;
          if (ma != 0) { mexp = (int)((mpred >> 7) & 1u); p_own = t_stretch(X.S, (u32)((int)mdt * (1 - 2 * mexp)) & 32767u); }
          else p_own = 0;
        } else {
          if (ma != 0) { mexp = (int)((mpred >> (7 - mbit)) & 1u); p_own = mp_ok; }
          else p_own = 0;
        }
      }
      u32 byte = 0;
      if (track_bit(K, y, byte)) {
        h = comp_hash(R, st, dist_last, byte, K, mring, mmask);
        if (type == ZT_CM) { cxa = (h ^ 1u) & cm_mask; v0 = gtab[cxa]; p_own = t_stretch(X.S, v0 >> 17); }
      }
    }
    cur ^= 1;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// role: every ISSE of the eight blocks, in component order, lane = block.  The weight pairs of all ISSEs and the inputs of
// the chain roots are loaded first (independent loads, back to back); the chains are then multiply-shift-clamp in
// registers, three dependent instructions per link.  (The plan guarantees that an ISSE's input is an ICM or the ISSE
// right before it.)
__device__ __forceinline__ void role_chain(const T8& X, const ZModel* __restrict__ ZM, int nbk) {
  const int lane = threadIdx.x & 31;
  const bool act = lane < nbk;
  const int ni = (int)ZM->tplan.nisse;
  u32 cj[T_NI];                                   // component | input << 8 | root << 16, warp-uniform
#pragma unroll
  for (int k = 0; k < T_NI; ++k) {
    cj[k] = 0;
    if (k < ni) {
      const u32 c = ZM->tplan.icomp[k], j = ZM->tplan.ijn[k];
      const u32 root = (k == 0 || j != (u32)ZM->tplan.icomp[k > 0 ? k - 1 : 0]) ? 1u : 0u;
      cj[k] = c | j << 8 | root << 16;
    }
  }
  int cur = 0;
  for (;;) {
    T_SYNC(BAR_H, X.nsub);
    if (act) {
      const short* Pc = X.P + PIDX(cur, lane, 0);
      const int2* xw = X.XW + lane * T_NI;
      int2 w[T_NI]; int pin[T_NI];
#pragma unroll
      for (int k = 0; k < T_NI; ++k) if (k < ni) { w[k] = xw[k]; pin[k] = (cj[k] >> 16) ? (int)Pc[(cj[k] >> 8) & 255u] : 0; }
      int p = 0;
      short* Pw = X.P + PIDX(cur, lane, 0);
#pragma unroll
      for (int k = 0; k < T_NI; ++k)
        if (k < ni) {
          const int pj = (cj[k] >> 16) ? pin[k] : p;
          p = clamp2k((w[k].x * pj + w[k].y * 64) >> 16);
          Pw[cj[k] & 255u] = (short)p;
        }
    }
    T_ARRIVE(BAR_C, X.nsub);
    T_SYNC(BAR_Y, X.nthreads);
    if (X.YV[T_NB]) break;
    cur ^= 1;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Static description of a mixer / scalar step, kept in shared memory by the warp that uses it (eight words each): the
// per-lane registers are for the state that changes (hash, row, entry), not for constants every lane would repeat.
struct StepDesc { u32 w0, aux, a5, cmm, limit, cm_off16, rlen, rfold_dist; };
// w0: kind | comp << 8 | j << 16 | k << 24 (mixers: owner in comp, j0 in j, m in k);  rfold_dist: fold_n | dist_byte << 8
__device__ __forceinline__ void fill_desc(StepDesc* D, const ZModel* __restrict__ ZM, int c, u32 kind) {
  const ZComp cd = ZM->comp[c];
  u32 j, k, aux;
  if (kind == ZT_MIX) { j = cd.a2; k = cd.a3; aux = cd.a4; }
  else if (kind == ZT_AVG) { j = cd.a1; k = cd.a2; aux = cd.a3; }
  else { j = cd.a2; k = cd.a3; aux = cd.a4; }
  D->w0 = kind | (u32)c << 8 | j << 16 | k << 24;
  D->aux = aux; D->a5 = kind == ZT_SSE ? 0xffffffffu : (u32)cd.a5; D->cmm = cd.cm_mask; D->limit = cd.limit; D->cm_off16 = cd.cm_off16;
  D->rlen = ZM->wave.flat_len[c]; D->rfold_dist = (u32)ZM->wave.fold_n[c] | (u32)ZM->pipe.dist_byte[c] << 8;
}
__device__ __forceinline__ u32 desc_hash(const StepDesc& D, const ZModel* __restrict__ ZM, u32 st[ZPQ_FLAT_SEGS], u32& dist_last, u32 byte,
                                         const Track& K, const u8* mring, u32 mmask) {
  HRecipe R; R.ops = ZM->wave.flat[(D.w0 >> 8) & 255u]; R.len = D.rlen; R.fold1 = D.rfold_dist & 255u; R.dist_byte = D.rfold_dist >> 8;
  return comp_hash(R, st, dist_last, byte, K, mring, mmask);
}

// role: the MIX components, lane = block * 4 + quarter (the quarter's inputs are j = quarter, quarter + 4, ...)
__device__ __forceinline__ void role_mix(const T8& X, const ZKernelArgs& A, const ZModel* __restrict__ ZM, const ZBlock* blks, int nbk,
                                         StepDesc* MD) {
  const int lane = threadIdx.x & 31;
  const int b = lane >> 2, part = lane & 3;
  const bool act = b < nbk;
  const int nmix = (int)ZM->fast.nmix;
  const ZBlock B = blks[act ? b : 0];
  u8* const arena = A.arena + B.arena_off;
  const u8* const mring = arena + (u64)ZM->m_off16 * 16;
  const u32 mmask = ZM->m_mask;
  if (lane < nmix) fill_desc(&MD[lane], ZM, (int)ZM->fast.mix_owner[lane], ZT_MIX);
  __syncwarp();
  u32 h[ZPQ_FAST_MIX], base[ZPQ_FAST_MIX], dist_last[ZPQ_FAST_MIX]; int pm[ZPQ_FAST_MIX];
  u32 st[ZPQ_FAST_MIX][ZPQ_FLAT_SEGS];
#pragma unroll
  for (int q = 0; q < ZPQ_FAST_MIX; ++q) { h[q] = 0; base[q] = 0; dist_last[q] = 0; pm[q] = 0; st[q][0] = st[q][1] = st[q][2] = st[q][3] = 0; }
  Track K; track_init(K);
  int cur = 0;
  for (;;) {
    // rows of this bit; the rows of both candidates of the next bit towards L1
#pragma unroll
    for (int q = 0; q < ZPQ_FAST_MIX; ++q)
      if (q < nmix) {
        const StepDesc D = MD[q];
        const u32 m = D.w0 >> 24;
        int* const tab = reinterpret_cast<int*>(arena + (u64)D.cm_off16 * 16);
        base[q] = ((h[q] + ((u32)K.c8 & D.a5)) & D.cmm) * m;
        if (act && K.ph != 7) {
          prefetch_l1(tab + ((h[q] + ((u32)(K.c8 * 2) & D.a5)) & D.cmm) * m + part * 8);
          prefetch_l1(tab + ((h[q] + ((u32)(K.c8 * 2 + 1) & D.a5)) & D.cmm) * m + part * 8);
        }
      }
    // weights of every mixer of this bit: independent loads, issued before the predictions are waited for
    int wv[ZPQ_FAST_MIX][8];
#pragma unroll
    for (int q = 0; q < ZPQ_FAST_MIX; ++q)
      if (q < nmix) {
        const int m = (int)(MD[q].w0 >> 24);
        const int* const tab = reinterpret_cast<const int*>(arena + (u64)MD[q].cm_off16 * 16) + base[q];
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int j = part + 4 * i; wv[q][i] = (act && j < m) ? tab[j] : 0; }
      }
    T_SYNC(BAR_C, X.nsub);
#pragma unroll
    for (int q = 0; q < ZPQ_FAST_MIX; ++q)
      if (q < nmix) {
        const StepDesc D = MD[q];
        const int m = (int)(D.w0 >> 24), j0 = (int)((D.w0 >> 16) & 255u), owner = (int)((D.w0 >> 8) & 255u);
        const short* Pc = X.P + PIDX(cur, b, j0);
        int xv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int j = part + 4 * i; xv[i] = (act && j < m) ? (int)Pc[j] : 0; }
        int sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += (wv[q][i] >> 8) * xv[i];
        sum += __shfl_xor_sync(FULL, sum, 1);
        sum += __shfl_xor_sync(FULL, sum, 2);
        pm[q] = clamp2k(sum >> 8);
        if (act && part == 0) X.P[PIDX(cur, b, owner)] = (short)pm[q];
        __syncwarp();                                             // a later mixer may take this one as an input
      }
    T_ARRIVE(BAR_M, 64);
    T_SYNC(BAR_Y, X.nthreads);
    if (X.YV[T_NB]) break;
    const u32 yv = act ? X.YV[b] : 0u;
    if (yv & 2u) {
      const int y = (int)(yv & 1u);
#pragma unroll
      for (int q = 0; q < ZPQ_FAST_MIX; ++q)
        if (q < nmix) {
          const StepDesc D = MD[q];
          const int m = (int)(D.w0 >> 24), j0 = (int)((D.w0 >> 16) & 255u);
          int* const tab = reinterpret_cast<int*>(arena + (u64)D.cm_off16 * 16) + base[q];
          const int err = ((y * 32767 - sq8(X.gsq, pm[q])) * (int)D.aux) >> 4;
          const short* Pc = X.P + PIDX(cur, b, j0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int j = part + 4 * i;
            if (j < m) tab[j] = clamp512k(wv[q][i] + ((err * (int)Pc[j] + (1 << 12)) >> 13));
          }
        }
      u32 byte = 0;
      if (track_bit(K, y, byte)) {
#pragma unroll
        for (int q = 0; q < ZPQ_FAST_MIX; ++q)
          if (q < nmix) h[q] = desc_hash(MD[q], ZM, st[q], dist_last[q], byte, K, mring, mmask);
      }
    }
    cur ^= 1;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// role: MIX2 / SSE / AVG steps and the arithmetic decoders, lane = block
#define T_MAXS ZPQ_T_SCALAR
__device__ __forceinline__ void role_tail(const T8& X, const ZKernelArgs& A, const ZModel* __restrict__ ZM, const ZBlock* blks, int nbk,
                                          int first_blk, StepDesc* SD) {
  const int lane = threadIdx.x & 31;
  const bool act = lane < nbk;
  const int n = (int)ZM->n;
  const ZBlock B = blks[act ? lane : 0];
  u8* const arena = A.arena + B.arena_off;
  u8* const mring = arena + (u64)ZM->m_off16 * 16;
  const u32 mmask = ZM->m_mask;
  const int ns = (int)ZM->tplan.nscalar;
  if (lane < ns) fill_desc(&SD[lane], ZM, (int)ZM->tplan.scomp[lane], ZM->comp[ZM->tplan.scomp[lane]].type);
  __syncwarp();
  u32 h[T_MAXS], cxt[T_MAXS], v0[T_MAXS], dist_last[T_MAXS]; int xd[T_MAXS];
  u32 st[T_MAXS][ZPQ_FLAT_SEGS];
#pragma unroll
  for (int s = 0; s < T_MAXS; ++s) { h[s] = 0; cxt[s] = 0; v0[s] = 0; dist_last[s] = 0; xd[s] = 0; st[s][0] = st[s][1] = st[s][2] = st[s][3] = 0; }
  Track K; track_init(K);
  Dec8 C;
  C.low = 1; C.high = 0xffffffffu; C.curr = 0; C.status = ST_OK;
  C.in = A.in + B.in_off; C.in_pos = 0; C.in_end = 0;
  C.out = A.out + B.out_off; C.out_pos = 0; C.out_cap = B.out_cap;
  int status = ST_OK;
  bool done = !act;
  u32 sg = 0; u64 out_start = 0;
  bool at_seg_start = true;                  // the next thing to do is to open segment `sg`
  int cur = 0;
  for (;;) {
    // ---- this bit's table entries of the scalar steps (and the candidates of the next bit towards L1) ----
#pragma unroll
    for (int s = 0; s < T_MAXS; ++s)
      if (s < ns && act) {
        const StepDesc D = SD[s];
        const u32 kind = D.w0 & 255u;
        u32* const gt = reinterpret_cast<u32*>(arena + (u64)D.cm_off16 * 16);
        if (kind == ZT_MIX2) {
          cxt[s] = (h[s] + ((u32)K.c8 & D.a5)) & D.cmm;
          v0[s] = (u32)reinterpret_cast<const u16*>(gt)[cxt[s]];
        }
        else if (kind == ZT_SSE) {
          cxt[s] = ((h[s] + (u32)K.c8) * 32u) & D.cmm;
          if (K.ph != 7) {
            prefetch_l1(gt + (((h[s] + (u32)(K.c8 * 2)) * 32u) & D.cmm));
            prefetch_l1(gt + (((h[s] + (u32)(K.c8 * 2 + 1)) * 32u) & D.cmm));
          }
        }
      }
    // ---- open a segment / end-of-stream flag of a byte (Decoder::decompress, libzpaq.cpp:2127-2137), per block ----
    bool valid = false;
    if (!done) {
      if (at_seg_start) {
        if (sg >= B.nseg) done = true;
        else {
          const u64 seg_n = A.seg_len[B.seg_first + sg];
          out_start = C.out_pos;
          C.in_end = C.in_pos + seg_n;
          C.curr = 0;
          for (int i = 0; i < 4; ++i) {
            if (C.in_pos >= C.in_end) { status = ST_EOF; break; }
            C.curr = C.curr << 8 | (u32)C.in[C.in_pos++];
          }
          at_seg_start = false;
          if (status != ST_OK) done = true;
        }
      }
      if (!done && K.ph == 0) {
        const int eos = dec8_bit(C, 0);
        if (C.status != ST_OK) { status = C.status; done = true; }
        else if (eos) {
          if (C.curr != 0) { status = ST_EOS; done = true; }
          else {
            if (C.in_pos != C.in_end) { status = ST_CORRUPT; done = true; }      // segment must end with its 4 zeros
            A.seg_out[B.seg_first + sg] = C.out_pos - out_start;
            ++sg; at_seg_start = true;
            if (sg >= B.nseg) done = true;
          }
        }
        else valid = true;
      }
      else if (!done) valid = true;
      if (done && status != ST_OK && sg < B.nseg) A.seg_out[B.seg_first + sg] = C.out_pos - out_start;
    }
    T_SYNC(BAR_M, 64);
    int y = 0;
    int res[T_MAXS];
#pragma unroll
    for (int s = 0; s < T_MAXS; ++s) res[s] = 0;
    if (valid) {
      // ---- the scalar steps: every input is loaded from P first (independent loads); an input that is an earlier scalar
      // step's output is taken from its register instead, so the steps are a register chain
      const short* Pc = X.P + PIDX(cur, lane, 0);
      int pjl[T_MAXS], pkl[T_MAXS];
#pragma unroll
      for (int s = 0; s < T_MAXS; ++s) { pjl[s] = 0; pkl[s] = 0; if (s < ns) { pjl[s] = (int)Pc[(SD[s].w0 >> 16) & 255u]; pkl[s] = (int)Pc[SD[s].w0 >> 24]; } }
      int pfin = (int)Pc[n - 1];
#pragma unroll
      for (int s = 0; s < T_MAXS; ++s)
        if (s < ns) {
          const StepDesc D = SD[s];
          const u32 kind = D.w0 & 255u;
          const u32 comp = (D.w0 >> 8) & 255u, sj = (D.w0 >> 16) & 255u, sk = D.w0 >> 24;
          int pj = pjl[s], pk = pkl[s];
#pragma unroll
          for (int e = 0; e < T_MAXS; ++e) if (e < s) { const u32 ce = (SD[e].w0 >> 8) & 255u; if (ce == sj) pj = res[e]; if (ce == sk) pk = res[e]; }
          int r;
          if (kind == ZT_SSE) {
            const u32* const gt = reinterpret_cast<const u32*>(arena + (u64)D.cm_off16 * 16);
            int pq = min(max(pj + 992, 0), 1983);
            const int wt = pq & 63;
            pq >>= 6;
            const u32 q0 = gt[cxt[s] + (u32)pq], q1 = gt[cxt[s] + (u32)pq + 1u];
            r = t_stretch(X.S, (u32)(((int)(q0 >> 10) * (64 - wt) + (int)(q1 >> 10) * wt) >> 13));
            cxt[s] += (u32)pq + (u32)(wt >> 5); v0[s] = (wt >> 5) ? q1 : q0;
          } else {
            xd[s] = pj - pk;
            if (kind == ZT_MIX2) { const int w = (int)v0[s]; r = (w * pj + (65536 - w) * pk) >> 16; }
            else r = (pj * (int)D.aux + pk * (256 - (int)D.aux)) >> 8;                 // AVG
          }
          res[s] = r;
          if ((int)comp == n - 1) pfin = r;
        }
      y = dec8_bit(C, (u32)sq8(X.gsq, pfin) * 2u + 1u);
      if (C.status != ST_OK) { status = C.status; done = true; valid = false; A.seg_out[B.seg_first + sg] = C.out_pos - out_start; }
    }
    if (lane < T_NB) X.YV[lane] = (u32)y | (valid ? 2u : 0u) | (done ? 4u : 0u);
    const bool all = __all_sync(FULL, done);
    if (lane == 0) X.YV[T_NB] = all ? 1u : 0u;
    __syncwarp();
    T_ARRIVE(BAR_Y, X.nthreads);
    if (all) break;
    if (valid) {
      // ---- train (libzpaq.cpp:1954-2052) ----
#pragma unroll
      for (int s = 0; s < T_MAXS; ++s)
        if (s < ns) {
          const StepDesc D = SD[s];
          const u32 kind = D.w0 & 255u;
          u32* const gt = reinterpret_cast<u32*>(arena + (u64)D.cm_off16 * 16);
          if (kind == ZT_MIX2) {
            const int err = ((y * 32767 - sq8(X.gsq, res[s])) * (int)D.aux) >> 5;
            int w = (int)v0[s];
            w += (err * xd[s] + (1 << 12)) >> 13;
            w = min(max(w, 0), 65535);
            reinterpret_cast<u16*>(gt)[cxt[s]] = (u16)w;
          }
          else if (kind == ZT_SSE) {
            const u32 vv = v0[s];
            const u32 count = vv & 0x3ffu;
            const int err = y * 32767 - (int)(vv >> 17);
            const u32 prod = (u32)err * (u32)__ldg(X.gdt + count);
            gt[cxt[s]] = vv + (prod & 0xfffffc00u) + (count < D.limit ? 1u : 0u);
          }
        }
      u32 byte = 0;
      if (track_bit(K, y, byte)) {
        mring[K.vm_c & mmask] = (u8)byte;                 // the M ring of the HCOMP machine: one writer per block
        if (C.out_pos < C.out_cap) C.out[C.out_pos] = (u8)byte;
        else { status = ST_FULL; done = true; A.seg_out[B.seg_first + sg] = C.out_pos + 1 - out_start; }
        ++C.out_pos;
#pragma unroll
        for (int s = 0; s < T_MAXS; ++s)
          if (s < ns) h[s] = desc_hash(SD[s], ZM, st[s], dist_last[s], byte, K, mring, mmask);
      }
    }
    cur ^= 1;
  }
  if (act) { A.status[first_blk + lane] = status; A.out_len[first_blk + lane] = C.out_pos; }
}

// ------------------------------------------------------------------------------------------------------------------
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) zpq_dec_kernel8(const ZTArgs W) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ZKernelArgs& A = W.A;
  {
    const ZTables* T = A.tables;
    const uint4* src = reinterpret_cast<const uint4*>(T->stretch_hi);
    uint4* dst = reinterpret_cast<uint4*>(smem + S8_STRETCH);
    for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) dst[i] = src[i];
    u16* ns = reinterpret_cast<u16*>(smem + S8_NS);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) ns[i] = (u16)(T->ns[i] & 0xffffu);
  }
  __syncthreads();
  T8 X;
  X.S.stretch_hi = reinterpret_cast<const short*>(smem + S8_STRETCH);
  X.S.squash = nullptr;
  X.rows = smem + S8_ROWS;
  X.gsq = A.tables->squash;
  X.S.ns = reinterpret_cast<const u16*>(smem + S8_NS);
  X.S.dt = nullptr; X.S.dt2k = nullptr;
  X.sm32 = reinterpret_cast<u32*>(smem);
  unsigned char* xchg = smem + S8_STATIC + (size_t)W.nbk * W.pw_bytes;
  X.P = reinterpret_cast<short*>(xchg);
  X.XW = reinterpret_cast<int2*>(xchg + T_P_BYTES);
  X.YV = reinterpret_cast<volatile u32*>(xchg + T_P_BYTES + T_XW_BYTES);
  X.gdt = A.tables->dt;
  X.nthreads = (int)blockDim.x;
  X.nsub = (int)blockDim.x - 64;
  volatile u32* grp = X.YV + T_NB + 1;
  StepDesc* descs = reinterpret_cast<StepDesc*>(xchg + T_P_BYTES + T_XW_BYTES + T_YV_BYTES);
  const int warp = threadIdx.x >> 5;
  const u32 wbase0 = (u32)S8_STATIC / 4u;
  const int pw_words = W.pw_bytes / 4;

  for (;;) {
    if (threadIdx.x == 0) { *grp = atomicAdd(A.counter, 1u); X.YV[T_NB] = 0u; }
    T_SYNC(BAR_G, X.nthreads);
    const int g = (int)*grp;
    if (g >= W.ngroups) break;
    const int first = (int)W.group_first[g], nbk = (int)W.group_count[g];
    const ZBlock* blks = A.blocks + first;
    const ZModel* __restrict__ ZM = A.models + blks[0].model;
    const int nh = (int)ZM->tplan.nh;
    if (warp < nh) role_hash(X, A, ZM, blks, nbk, (int)ZM->tplan.hcomp[warp], warp, pw_words, wbase0);
    else if (warp == nh) role_misc(X, A, ZM, blks, nbk);
    else if (warp == nh + 1) role_chain(X, ZM, nbk);
    else if (warp == nh + 2) role_mix(X, A, ZM, blks, nbk, descs);
    else if (warp == nh + 3) role_tail(X, A, ZM, blks, nbk, first, descs + ZPQ_FAST_MIX);
    else {
      // spare warps of a CTA launched for a model with more hash components: keep the barriers' thread count
      for (;;) {
        T_ARRIVE(BAR_H, X.nsub);
        T_SYNC(BAR_C, X.nsub);
        T_SYNC(BAR_Y, X.nthreads);
        if (X.YV[T_NB]) break;
      }
    }
    T_SYNC(BAR_G, X.nthreads);
  }
}

extern "C" int zpq_dec8_static_smem(void) { return S8_STATIC; }
extern "C" int zpq_dec8_xchg_bytes(void) { return S8_XCHG; }

extern "C" cudaError_t zpq_launch_dec8(const ZKernelArgs* a, const uint32_t* d_group_first, const uint32_t* d_group_count, int ngroups,
                                       int nbk, int pw_bytes, int nwarps, int sm_count, cudaStream_t st) {
  if (ngroups <= 0) return cudaSuccess;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(zpq_dec_kernel8<640>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZPQ_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(zpq_dec_kernel8<768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZPQ_SMEM_MAX);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  ZTArgs W;
  W.A = *a; W.group_first = d_group_first; W.group_count = d_group_count; W.ngroups = ngroups; W.nbk = nbk; W.pw_bytes = pw_bytes; W.nwarps = nwarps;
  const size_t smem = (size_t)S8_STATIC + (size_t)nbk * pw_bytes + S8_XCHG;
  const int ctas = ngroups < sm_count ? ngroups : sm_count;
  // up to 20 warps (16 hash components, the level-5 chain): 96 registers per thread; up to 24 warps: 80
  if (nwarps <= 20) zpq_dec_kernel8<640><<<ctas, nwarps * 32, smem, st>>>(W);
  else zpq_dec_kernel8<768><<<ctas, nwarps * 32, smem, st>>>(W);
  return cudaGetLastError();
}
