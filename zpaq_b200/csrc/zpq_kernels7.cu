// zpq_kernels7.cu -- DECODE, two warps per archive block, critical-first ordering (round 2), sm_100a.
//
// Decoding is one serial chain per block: bit t+1 cannot be predicted before bit t is decoded.  The round-1 two-warp
// decoder (zpq_kernels6.cu) split a bit's work over two warps but kept the reference's order inside each of them
// (predict, wait, train, advance), so the warps mostly waited for each other: its ncu source page shows 17 % + 9 % of
// all samples on the two spin loops and ~1 190 warp instructions per block-bit of which ~280 are polls
// (profiles/r02a_dec6_*).  This kernel keeps the mapping (lane = component; warp A = every ICM / ISSE / CM / MATCH /
// CONST, warp B = every MIX / MIX2 / SSE / AVG and the arithmetic decoder) and reorders the work around the only chain
// that cannot be shortened, y(t) -> ISSE chain(t+1) -> mixers -> SSE -> decoder -> y(t+1):
//
//   * TWO-OUTCOME SPECULATION in warp A.  While warp B mixes and decodes bit t, warp A prepares bit t+1 for y = 0 AND
//     y = 1: both candidate bit-history slots of its hash row, both table entries (and their stretch() values), both
//     trained values of the entry in use (so that "next bit reads the entry just trained" is a select, not arithmetic),
//     both next states.  These are reads plus register arithmetic on private data, so the bytes are unchanged.  When y
//     arrives, A selects, walks the ISSE chain and publishes; the table stores and the row bookkeeping follow after.
//   * UNIFORM SCALARS in warp B.  A mixer's redux.sync result is warp-uniform, so MIX2 / SSE / AVG / the decoder are
//     evaluated by every lane from uniform registers (inputs that are mixer outputs are picked with selects, not
//     shuffles); only the SSE interpolation needs two shuffles.  Training, the next bit's rows and contexts come after y
//     has been sent.
//   * HARDWARE NAMED BARRIERS instead of polled mailboxes: per block P ("predictions published": A arrives, B syncs) and
//     Y ("bit decoded": B arrives, A syncs).  bar.arrive does not block the producer, a waiting warp issues nothing.
//     Control messages (end of segment, abort) ride on the same Y hand-shake: B owes exactly one Y per P.
//   * second-nibble hash rows are pulled into L1/L2 for all candidates two bits early; the MATCH index (addressed by
//     the PREVIOUS byte's hash, libzpaq.cpp:1996) at the start of its byte.
// Bit-exactness: every table entry sees the same operations in the same order as in Predictor::predict0/update0
// (libzpaq.cpp:1854-2066); speculation only reads.  Tests run this tier against the oracle and the golden archives.
#include <cuda_runtime.h>
#include <stdint.h>
#include "zpq_dev.h"
#include "zpq_kernels.h"
#include "zpq_kcommon.cuh"

// shared memory: stretch (32 KB) | squash (2 688 B) | ns (512 B) ; dt is read through the read-only path (training only)
#define S7_STRETCH 0
#define S7_SQUASH (S7_STRETCH + 32768)
#define S7_NS (S7_SQUASH + 2688)
#define S7_STATIC (S7_NS + 512)
#define S7_CTRL 96                          // per block: short pp[32] | msg | block index | pad
#define BAR_SYNC64(id) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(64) : "memory")
#define BAR_ARRIVE64(id) asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(64) : "memory")

struct Dec7 {
  u32 low, high, curr;
  const u8* in; u64 in_pos, in_end;
  u8* out; u64 out_pos, out_cap;
  int status;
};

__device__ __forceinline__ int dec7_bit(Dec7& C, u32 p16) {
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

// step e of a signature runs in warp B (MIX / MIX2 / SSE / AVG)
__host__ __device__ constexpr bool sig7_is_b(unsigned long long sig, int e) {
  return ((sig >> (3 * e)) & 7ull) != 0ull && ((sig >> (3 * e)) & 7ull) != (unsigned long long)ZK_ISSE;
}
__host__ __device__ constexpr int sig7_count(unsigned long long sig, int upto, unsigned kind) {
  int c = 0;
  for (int s = 0; s < upto; ++s) if (((sig >> (3 * s)) & 7ull) == kind) ++c;
  return c;
}

// Predictor::find (libzpaq.cpp:2072-2088) with the three probes as one group
__device__ __forceinline__ void find7(const u8* ht, u32 ht_mask, int shift, u32 cxt, uint4& row, u32& off) {
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

// hplan_eval (zpq_kcommon.cuh) over the slots of ONE role only: `list` / `nl` = ZHashPlan.role_slots / role_n of the warp's
// role.  Every listed slot has at least one lane of this warp, so there is nothing to skip.
__device__ __forceinline__ u32 hplan_eval7(const ZHashPlan* __restrict__ P, const u8* __restrict__ list, int nl, u32 fold_mask, u32 mine_mask,
                                           int lane, u32 last_seg, u32 fold_n1, u32 st[ZPQ_FLAT_SEGS], u32 cur, u64 hs, u32 vc,
                                           const u8* ring, u32 ring_mask, u32 dist_last, u32 h_old) {
  u32 accf = 0;
  if (fold_mask) {                                         // warp-uniform
#pragma unroll
    for (u32 k = 0; k < 8; ++k) {
      const u32 nx = (accf + ((u32)(hs >> (8u * k)) & 255u) + 512u) * 773u;
      if (k + 1u < fold_n1) accf = nx;
    }
  }
  u32 acc = 0, prev_new = 0, prev_old = 0;
  const uint2* __restrict__ sd = reinterpret_cast<const uint2*>(P->slot);
  int sn = nl > 0 ? (int)__ldg(list) : 0;
  uint2 dn = __ldg(sd + sn);
  u32 wn = __ldg(&P->w[sn][lane]);
  for (int k = 0; k < nl; ++k) {
    const int s = sn;
    const u32 lm = dn.x, kind = dn.y, w = wn;
    if (k + 1 < nl) { sn = (int)__ldg(list + k + 1); dn = __ldg(sd + sn); wn = __ldg(&P->w[sn][lane]); }
    const bool on = (lm >> lane) & 1u;
    const u32 src = (w >> 4) & 15u, arg = (w >> 8) & 255u, off = w >> 16;
    if (kind == ZH_FOLD_BYTE) {
      const u32 x = (u32)(hs >> ((off & 7u) * 8u)) & arg;
      if (on) acc = (acc + x + 512u) * 773u;
    }
    else if (kind == ZH_SEG) {
      const u32 kk = arg & 3u;
      if (on) {
        if (kk > 0u) {
          const u32 old = kk == 1u ? st[0] : (kk == 2u ? st[1] : st[2]);
          if (kk == 1u) st[0] = acc; else if (kk == 2u) st[1] = acc; else st[2] = acc;
          prev_old = old; prev_new = acc;
        }
        const u32 own = kk == 0u ? st[0] : (kk == 1u ? st[1] : (kk == 2u ? st[2] : st[3]));
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
        if (src == ZS_BYTE) x = (u32)ring[(vc + off) & ring_mask] & arg;
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

template <unsigned long long SIG>
__global__ void __launch_bounds__(448, 1) zpq_dec_kernel7(const ZKernelArgs A, const int pw_bytes, const int nteams) {
  extern __shared__ __align__(16) unsigned char smem[];
  {
    const ZTables* T = A.tables;
    const uint4* src = reinterpret_cast<const uint4*>(T->stretch_hi);
    uint4* dst = reinterpret_cast<uint4*>(smem + S7_STRETCH);
    for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) dst[i] = src[i];
    u16* sq = reinterpret_cast<u16*>(smem + S7_SQUASH);
    for (int i = threadIdx.x; i < 1344; i += blockDim.x) sq[i] = T->squash[1376 + i];
    u16* ns = reinterpret_cast<u16*>(smem + S7_NS);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) ns[i] = (u16)(T->ns[i] & 0xffffu);
  }
  __syncthreads();
  Tabs S;
  S.stretch_hi = reinterpret_cast<const short*>(smem + S7_STRETCH);
  S.squash = reinterpret_cast<const u16*>(smem + S7_SQUASH);
  S.ns = reinterpret_cast<const u16*>(smem + S7_NS);
  S.dt = nullptr; S.dt2k = nullptr;
  const int* __restrict__ gdt = A.tables->dt;
  u32* sm32 = reinterpret_cast<u32*>(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp >> 1, role = warp & 1;
  const int barP = 1 + 2 * team, barY = 2 + 2 * team;
  const u32 wbase = (u32)(S7_STATIC + team * pw_bytes) / 4u;
  unsigned char* ctrlb = smem + S7_STATIC + (size_t)nteams * pw_bytes + (size_t)team * S7_CTRL;
  volatile short* pp = reinterpret_cast<volatile short*>(ctrlb);
  volatile u32* ymsg = reinterpret_cast<volatile u32*>(ctrlb + 64);     // kind << 1 | y ; kind 0 bit, 1 end of segment, 2 abort
  volatile u32* blkw = ymsg + 1;

  for (;;) {
    if (role == 0 && lane == 0) *blkw = atomicAdd(A.counter, 1u);
    BAR_SYNC64(barP);
    const int blk = (int)*blkw;
    if (blk >= A.nblk) break;

    const ZBlock B = A.blocks[blk];
    const ZModel* __restrict__ ZM = A.models + B.model;
    u8* const arena = A.arena + B.arena_off;
    const int n = (int)ZM->n;
    // ---- lane <-> component (same index in both warps; a lane only works in the warp that owns it) ----
    u32 type = ZT_NONE, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, cm_mask = 0, ht_mask = 0, limit = 0, level = 0;
    u32* gtab = nullptr; u8* ht = nullptr;
    if (lane < n) {
      const ZComp cd = ZM->comp[lane];
      // warp A: the hash components and MATCH; warp B: everything else (CM / CONST moved to B in round 2: A was the longer stream)
      const bool inA = cd.type == ZT_ICM || cd.type == ZT_ISSE || cd.type == ZT_MATCH;
      if ((role == 0) == inA) {
        type = cd.type; a1 = cd.a1; a2 = cd.a2; a3 = cd.a3; a4 = cd.a4; a5 = cd.a5;
        cm_mask = cd.cm_mask; ht_mask = cd.ht_mask; limit = cd.limit; level = ZM->fast.level[lane];
        gtab = reinterpret_cast<u32*>(arena + (u64)cd.cm_off16 * 16);
        ht = arena + (u64)cd.ht_off16 * 16;
      }
    }
    const bool mine = type != ZT_NONE;
    const ZHashPlan* __restrict__ HP = &ZM->hplan;
    const u32 mine_mask = __ballot_sync(FULL, mine);
    const u32 hp_last = (lane < n) ? (u32)HP->last_seg[lane] : 0u;
    const u32 hp_fold = (lane < n) ? (u32)HP->fold_n[lane] : 0u;
    const u8* __restrict__ hp_list = HP->role_slots[role];
    const int hp_nl = (int)HP->role_n[role];
    const u32 hp_foldm = HP->role_fold[role];
    const u32 tabw = wbase + ((lane < n) ? ZM->fast.smem_off[lane] : 0u) / 4u;
    const u32 dist_byte = mine ? ZM->pipe.dist_byte[lane] : 0u;
    u32 st[ZPQ_FLAT_SEGS] = {0, 0, 0, 0};
    u32 dist_last = 0, vm_c = 0; u64 hist = 0;
    u8* const mring = arena + (u64)ZM->m_off16 * 16;
    const u32 mmask = ZM->m_mask;
    u32 h = 0;
    int c8 = 1, hmap4 = 1;

#define BYTE_END7(byteval)                                                                           \
    do {                                                                                             \
      const u32 bv_ = (u32)(byteval);                                                                \
      vm_c -= 1; mring[vm_c & mmask] = (u8)bv_;          /* both warps write the same value */       \
      hist = (hist << 8) | (u64)bv_;                                                                 \
      if (mine && bv_ == dist_byte) dist_last = vm_c;                                                \
      h = hplan_eval7(HP, hp_list, hp_nl, hp_foldm, mine_mask, lane, hp_last, hp_fold, st, bv_, hist, vm_c, mring, mmask, dist_last, h);   \
      hmap4 = 1; c8 = 1;                                                                             \
    } while (0)

    if (role == 0) {
      // =============================== warp A: ICM / ISSE / CM / MATCH / CONST ===============================
      __syncwarp();
      for (int src = 0; src < n; ++src) {
        const u32 ty = __shfl_sync(FULL, type, src);
        const u32 tw = __shfl_sync(FULL, tabw, src);
        if (ty == ZT_ICM) for (int i = lane; i < 256; i += 32) sm32[tw + i] = A.tables->icm_init[i];
        else if (ty == ZT_ISSE) for (int i = lane; i < 512; i += 32) sm32[tw + i] = A.tables->isse_init[i];
      }
      __syncwarp();
      const bool hashl = type == ZT_ICM || type == ZT_ISSE;
      const bool isse = type == ZT_ISSE;
      const u32 tw2 = isse ? 2u : 1u, tw1 = isse ? 1u : 0u;      // words per table entry; offset of the second word
      constexpr int NLEV = sig7_count(SIG, ZPQ_FAST_STEPS, ZK_ISSE);      // ISSE levels of the signature (= fast.nlev)
      const int mylev = isse ? (int)level : 0;
      const int jsrc = (int)a2 & 31;
      uint4 row = make_uint4(0, 0, 0, 0); u32 row_off = 0;
      u32 hs = 1;                                                // slot inside the row = hmap4 & 15
      u32 bh = 0, v0 = 0, v1 = 0;                                // state and table entry of the CURRENT bit (hash lanes)
      u32 cxa = 0;                                               // CM: entry index of the current bit
      int p_own = (type == ZT_CONS) ? ((int)a1 - 128) * 4 : 0;   // prediction of a component that needs no input
      u32 ma = 0, mb = 0, mpos = 0, mbit = 0, mpred = 0, mdt = 0; int mexp = 0;      // MATCH
      int ph = 0;                                                // bit position inside the byte, 0..7
      bool abort = false;

      // state of the first bit of a nibble (or of the block): row lookup, entry, own prediction
#define LOAD_CUR7()                                                                                   \
      do {                                                                                            \
        if (hashl) {                                                                                  \
          find7(ht, ht_mask, (int)a1 + 2, h + 16u * (u32)c8, row, row_off);                           \
          hs = 1u;                                                                                    \
          bh = (row.x >> 8) & 255u;                                                                   \
          v0 = sm32[tabw + bh * tw2]; v1 = sm32[tabw + bh * tw2 + tw1];                               \
          if (!isse) p_own = t_stretch(S, v0 >> 8);                                                   \
        }                                                                                             \
      } while (0)
      LOAD_CUR7();
      if (type == ZT_CM) { cxa = (h ^ 1u) & cm_mask; v0 = gtab[cxa]; p_own = t_stretch(S, v0 >> 17); }

      for (u32 sg = 0; sg < B.nseg && !abort; ++sg) {
        for (;;) {
          // ---------------- predict (critical path): ISSE chain, publish ----------------
          int p = p_own, xj = 0;
          {
            // the chain, one level per step of the signature: shuffle, multiply-add, shift, clamp, two selects
            const int c64 = (int)v1 * 64;
#pragma unroll
            for (int lev = 1; lev <= NLEV; ++lev) {
              const int pj = __shfl_sync(FULL, p, jsrc);
              const int q = clamp2k(((int)v0 * pj + c64) >> 16);
              const bool on = mylev == lev;
              xj = on ? pj : xj; p = on ? q : p;
            }
          }
          if (mine) pp[lane] = (short)p;
          BAR_ARRIVE64(barP);
          // ---------------- while warp B mixes and decodes: everything bit t+1 can need, for y = 0 and y = 1 ----------------
          u32 tr0a = 0, tr0b = 0, tr1a = 0, tr1b = 0, nsw = 0;        // trained entry / next states of the entry in use
          u32 bc0 = 0, bc1 = 0, vc0a = 0, vc0b = 0, vc1a = 0, vc1b = 0; int pc0 = 0, pc1 = 0;     // candidates of the next bit
          const bool in_nib = (ph & 3) != 3;                           // next bit uses the same row
          if (hashl) {
            nsw = (u32)S.ns[bh];
            if (isse) {
              const int sq = t_squash(S, p);
              const int e0 = -sq, e1 = 32767 - sq;
              tr0a = (u32)clamp512k((int)v0 + ((e0 * xj + (1 << 12)) >> 13));
              tr1a = (u32)clamp512k((int)v0 + ((e1 * xj + (1 << 12)) >> 13));
              tr0b = (u32)clamp512k((int)v1 + ((e0 + 16) >> 5));
              tr1b = (u32)clamp512k((int)v1 + ((e1 + 16) >> 5));
            } else {
              tr0a = v0 + (u32)(((int)(0 - (int)(v0 >> 8))) >> 2);
              tr1a = v0 + (u32)(((int)(32767 - (int)(v0 >> 8))) >> 2);
            }
            if (in_nib) {
              bc0 = row_get(row, (int)(hs * 2u)); bc1 = row_get(row, (int)(hs * 2u + 1u));
              vc0a = sm32[tabw + bc0 * tw2]; vc0b = sm32[tabw + bc0 * tw2 + tw1];
              vc1a = sm32[tabw + bc1 * tw2]; vc1b = sm32[tabw + bc1 * tw2 + tw1];
              if (!isse) { pc0 = t_stretch(S, vc0a >> 8); pc1 = t_stretch(S, vc1a >> 8); }
            }
            else if (ph == 3) {
              // second nibble: both candidate row groups towards L1 (all four were pulled into L2 one bit earlier)
              prefetch_l1(ht + (((h + 16u * (u32)(c8 * 2)) * 16u) & ht_mask & ~63u));
              prefetch_l1(ht + (((h + 16u * (u32)(c8 * 2 + 1)) * 16u) & ht_mask & ~63u));
            }
            if (ph == 2) {     // c8 = 1bb: the second nibble's context is c8 * 4 + k
#pragma unroll
              for (int k = 0; k < 4; ++k) prefetch_l2(ht + (((h + 16u * (u32)(c8 * 4 + k)) * 16u) & ht_mask & ~63u));
            }
          }
          u32 cmc0 = 0, cmc1 = 0, cmx0 = 0, cmx1 = 0, cmt0 = 0, cmt1 = 0; int cmp0 = 0, cmp1 = 0;
          int mp_ok = 0;
          if (type == ZT_CM) {
            const u32 count = v0 & 0x3ffu;
            const u32 dtv = (u32)__ldg(gdt + count), inc = count < limit ? 1u : 0u;
            cmt0 = v0 + (((u32)(0 - (int)(v0 >> 17)) * dtv) & 0xfffffc00u) + inc;
            cmt1 = v0 + (((u32)(32767 - (int)(v0 >> 17)) * dtv) & 0xfffffc00u) + inc;
            if (ph != 7) {
              u32 hm0;
              if (ph == 3) hm0 = (u32)((hmap4 & 0xf) << 5 | 1);
              else hm0 = (u32)((hmap4 & 0x1f0) | (((hmap4 & 0xf) * 2) & 0xf));
              const u32 hm1 = hm0 | (ph == 3 ? 16u : 1u);
              cmx0 = (h ^ hm0) & cm_mask; cmx1 = (h ^ hm1) & cm_mask;
              cmc0 = gtab[cmx0]; cmc1 = gtab[cmx1];
              cmp0 = t_stretch(S, cmc0 >> 17); cmp1 = t_stretch(S, cmc1 >> 17);
            }
          }
          else if (type == ZT_MATCH) {
            // prediction of the next bit of this byte if the match survives this one
            if (ma != 0 && mbit < 7) {
              const u32 mcn = (mpred >> (6 - mbit)) & 1u;
              mp_ok = t_stretch(S, (u32)((int)mdt * (1 - 2 * (int)mcn)) & 32767u);
            }
            if (ph == 0) prefetch_l1(gtab + (h & cm_mask));          // index slot read at the end of this byte (old hash, gotcha 6)
          }
          // ---------------- the decoded bit ----------------
          BAR_SYNC64(barY);
          const u32 msg = *ymsg;
          const u32 kind = msg >> 1;
          if (kind == 2u) { abort = true; break; }
          if (kind == 1u) break;                               // end of segment: the prediction above was not used
          const int y = (int)(msg & 1u);
          const int c8n = c8 * 2 + y;
          // ---------------- select the next bit's state (critical), then store what bit t trained ----------------
          if (hashl) {
            const u32 tra = y ? tr1a : tr0a, trb = y ? tr1b : tr0b;
            const u32 nx = (nsw >> (8 * y)) & 255u;
            const u32 bh_old = bh, hs_old = hs;
            if (in_nib) {
              const u32 bn = y ? bc1 : bc0;
              const bool same = bn == bh_old;
              v0 = same ? tra : (y ? vc1a : vc0a);
              v1 = same ? trb : (y ? vc1b : vc0b);
              if (!isse) { p_own = y ? pc1 : pc0; if (same) p_own = t_stretch(S, tra >> 8); }
              bh = bn; hs = hs * 2u + (u32)y;
            }
            sm32[tabw + bh_old * tw2] = tra;
            if (isse) sm32[tabw + bh_old * tw2 + 1u] = trb;
            row_set(row, (int)hs_old, nx);
            if (!in_nib) *reinterpret_cast<uint4*>(ht + row_off) = row;
          }
          else if (type == ZT_CM) {
            gtab[cxa] = y ? cmt1 : cmt0;
            if (ph != 7) {
              const u32 cxn = y ? cmx1 : cmx0;
              const bool same = cxn == cxa;
              const u32 tr = y ? cmt1 : cmt0;
              v0 = same ? tr : (y ? cmc1 : cmc0);
              p_own = y ? cmp1 : cmp0;
              if (same) p_own = t_stretch(S, tr >> 17);
              cxa = cxn;
            }
          }
          else if (type == ZT_MATCH) {
            // MATCH update (libzpaq.cpp:1990-2008); h is still the hash as of the previous byte here (gotcha 6)
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
              if (ma != 0) { mexp = (int)((mpred >> 7) & 1u); p_own = t_stretch(S, (u32)((int)mdt * (1 - 2 * mexp)) & 32767u); }
              else p_own = 0;
            } else {
              if (ma != 0) { mexp = (int)((mpred >> (7 - mbit)) & 1u); p_own = mp_ok; }
              else p_own = 0;
            }
          }
          // ---------------- advance c8 / hmap4; nibble and byte boundaries ----------------
          if (ph == 7) {
            BYTE_END7(c8n - 256);
            ph = 0;
            LOAD_CUR7();
            if (type == ZT_CM) { cxa = (h ^ 1u) & cm_mask; v0 = gtab[cxa]; p_own = t_stretch(S, v0 >> 17); }
          } else {
            c8 = c8n;
            if (ph == 3) { hmap4 = (hmap4 & 0xf) << 5 | y << 4 | 1; LOAD_CUR7(); }
            else hmap4 = (hmap4 & 0x1f0) | (((hmap4 & 0xf) * 2 + y) & 0xf);
            ++ph;
          }
        }
      }
    } else {
      // ================== warp B: MIX, MIX2, SSE, AVG and the arithmetic decoder ==================
      constexpr int NMIXC = SIG ? sig7_count(SIG, ZPQ_FAST_STEPS, ZK_MIX) : ZPQ_FAST_MIX;
      constexpr int NSSEC = SIG ? sig7_count(SIG, ZPQ_FAST_STEPS, ZK_SSE) : ZPQ_FAST_SSE;
      const int nsteps = (int)ZM->fast.nsteps, nmix = SIG ? NMIXC : (int)ZM->fast.nmix, nsse = SIG ? NSSEC : (int)ZM->fast.nsse;
      u32 sd[ZPQ_FAST_STEPS];
#pragma unroll
      for (int s = 0; s < ZPQ_FAST_STEPS; ++s) sd[s] = ZM->fast.sd[s];
      // the last component is produced by the last step of this warp (else it is one of warp A's)
      const u32 sd_last = nsteps > 0 ? ZM->fast.sd[nsteps - 1] : 0u;
      const bool fin_in_b = nsteps > 0 && (sd_last & 7u) != ZK_ISSE && (int)((sd_last >> 3) & 31u) == n - 1;
      int* mix_tab[ZPQ_FAST_MIX]; u32 mix_owner[ZPQ_FAST_MIX], mix_m[ZPQ_FAST_MIX], mix_rate[ZPQ_FAST_MIX], mixbase[ZPQ_FAST_MIX];
#pragma unroll
      for (int q = 0; q < ZPQ_FAST_MIX; ++q) {
        mix_tab[q] = reinterpret_cast<int*>(arena + (u64)ZM->fast.mix_off16[q] * 16);
        mix_owner[q] = ZM->fast.mix_owner[q]; mix_m[q] = ZM->fast.mix_m[q]; mix_rate[q] = ZM->fast.mix_rate[q]; mixbase[q] = 0;
      }
      const u32* sse_tab[ZPQ_FAST_SSE]; u32 sse_owner[ZPQ_FAST_SSE];
#pragma unroll
      for (int r = 0; r < ZPQ_FAST_SSE; ++r) {
        sse_tab[r] = reinterpret_cast<const u32*>(arena + (u64)ZM->fast.sse_off16[r] * 16);
        sse_owner[r] = ZM->fast.sse_owner[r];
      }
      int mixw0 = 0, mixw1 = 0, mixw2 = 0, mixx0 = 0, mixx1 = 0, mixx2 = 0; u32 sserow0 = 0, sserow1 = 0;
      u32 cxt = 0, v0 = 0;
      // MIX: ((h + (c8 & mask)) & (size-1)) * m   MIX2: (h + (c8 & mask)) & (size-1)   SSE: ((h + c8) * 32) & (32*size-1)
      const bool ctx_on = type == ZT_MIX || type == ZT_MIX2 || type == ZT_SSE;
      const u32 ctx_m1 = type == ZT_SSE ? 0xffffffffu : a5;
      const u32 ctx_m2 = type == ZT_SSE ? cm_mask >> 5 : cm_mask;
      const u32 ctx_mul = type == ZT_MIX ? a3 : (type == ZT_SSE ? 32u : 1u);
      // candidate rows of the next bit: valid when every MIX has a full c8 mask and >= 256 rows and every SSE has
      // >= 256 rows, so that the two candidate rows differ from the row trained in this bit
      int mc00 = 0, mc01 = 0, mc10 = 0, mc11 = 0, mc20 = 0, mc21 = 0; u32 sc00 = 0, sc01 = 0, sc10 = 0, sc11 = 0;
      bool candv = false; int ylast = 0;
      const bool cand_ok = __all_sync(FULL, (type != ZT_MIX || ((a5 & 255u) == 255u && cm_mask >= 255u)) &&
                                            (type != ZT_SSE || cm_mask >= 256u * 32u - 1u));
      // CM / CONST lanes of this warp: plain predict-after-update (warp B has the slack after it has sent the bit)
      int p_ownB = (type == ZT_CONS) ? ((int)a1 - 128) * 4 : 0;
      u32 cxaB = 0, cmv = 0;
      if (type == ZT_CM) { cxaB = (h ^ 1u) & cm_mask; cmv = gtab[cxaB]; p_ownB = t_stretch(S, cmv >> 17); }
      int status = ST_OK;
      Dec7 C;
      C.low = 1; C.high = 0xffffffffu; C.curr = 0; C.status = ST_OK;
      C.in = A.in + B.in_off; C.in_pos = 0; C.in_end = 0;
      C.out = A.out + B.out_off; C.out_pos = 0; C.out_cap = B.out_cap;

#define SEND7(kind, y) do { __syncwarp(); if (lane == 0) *ymsg = (u32)(kind) << 1 | (u32)(y); __syncwarp(); BAR_ARRIVE64(barY); } while (0)
// a control message answers the prediction warp A has published (or is about to publish) for a bit that will not be decoded
#define REPLY_CTRL7(kind) do { BAR_SYNC64(barP); SEND7(kind, 0); } while (0)

      for (u32 sg = 0; sg < B.nseg && status == ST_OK; ++sg) {
        const u64 seg_n = A.seg_len[B.seg_first + sg];
        const u64 out_start = C.out_pos;
        C.in_end = C.in_pos + seg_n;
        C.curr = 0;
        for (int i = 0; i < 4; ++i) {
          if (C.in_pos >= C.in_end) { status = ST_EOF; break; }
          C.curr = C.curr << 8 | (u32)C.in[C.in_pos++];
        }
        if (status != ST_OK) { REPLY_CTRL7(2); break; }
        for (;;) {
          const int eos = dec7_bit(C, 0);
          if (C.status != ST_OK) { status = C.status; REPLY_CTRL7(2); break; }
          if (eos) {
            if (C.curr != 0) { status = ST_EOS; REPLY_CTRL7(2); }
            else REPLY_CTRL7(1);
            break;
          }
          int cbyte = 0;
          bool failed = false;
          for (int bit = 7; bit >= 0; --bit) {
            // ---- this bit's rows and contexts (while warp A walks its chain) ----
            u32 cxn0 = 0, cxn1 = 0;
            if (ctx_on) {
              cxt = ((h + ((u32)c8 & ctx_m1)) & ctx_m2) * ctx_mul;
              cxn0 = ((h + ((u32)(c8 * 2) & ctx_m1)) & ctx_m2) * ctx_mul;
              cxn1 = ((h + ((u32)(c8 * 2 + 1) & ctx_m1)) & ctx_m2) * ctx_mul;
              if (type == ZT_MIX2) v0 = reinterpret_cast<const u16*>(gtab)[cxt];
            }
            const bool use_cand = candv;
            candv = cand_ok && c8 < 128;
#pragma unroll
            for (int q = 0; q < ZPQ_FAST_MIX; ++q)
              if (q < NMIXC && q < nmix) {
                const u32 base = __shfl_sync(FULL, cxt, (int)mix_owner[q]);
                mixbase[q] = base;
                int wq;
                if (use_cand) wq = q == 0 ? (ylast ? mc01 : mc00) : (q == 1 ? (ylast ? mc11 : mc10) : (ylast ? mc21 : mc20));
                else wq = lane < (int)mix_m[q] ? mix_tab[q][base + lane] : 0;
                if (q == 0) mixw0 = wq; else if (q == 1) mixw1 = wq; else mixw2 = wq;
                if (candv) {
                  const u32 nb0 = __shfl_sync(FULL, cxn0, (int)mix_owner[q]), nb1 = __shfl_sync(FULL, cxn1, (int)mix_owner[q]);
                  const int m_ = (int)mix_m[q];
                  const int n0 = lane < m_ ? mix_tab[q][nb0 + lane] : 0;
                  const int n1 = lane < m_ ? mix_tab[q][nb1 + lane] : 0;
                  if (q == 0) { mc00 = n0; mc01 = n1; } else if (q == 1) { mc10 = n0; mc11 = n1; } else { mc20 = n0; mc21 = n1; }
                }
              }
#pragma unroll
            for (int r = 0; r < ZPQ_FAST_SSE; ++r)
              if (r < NSSEC && r < nsse) {
                const u32 rb = __shfl_sync(FULL, cxt, (int)sse_owner[r]);
                u32 rowv;
                if (use_cand) rowv = r == 0 ? (ylast ? sc01 : sc00) : (ylast ? sc11 : sc10);
                else rowv = sse_tab[r][rb + lane];
                if (r == 0) sserow0 = rowv; else sserow1 = rowv;
                if (candv) {
                  const u32 nb0 = __shfl_sync(FULL, cxn0, (int)sse_owner[r]), nb1 = __shfl_sync(FULL, cxn1, (int)sse_owner[r]);
                  const u32 n0 = sse_tab[r][nb0 + lane], n1 = sse_tab[r][nb1 + lane];
                  if (r == 0) { sc00 = n0; sc01 = n1; } else { sc10 = n0; sc11 = n1; }
                }
              }
            // weights of the MIX2 steps as uniform values (their owner lanes loaded them above)
            int w2u[ZPQ_FAST_STEPS];
#pragma unroll
            for (int s = 0; s < ZPQ_FAST_STEPS; ++s) {
              w2u[s] = 0;
              const u32 ksig = (u32)((SIG >> (3 * s)) & 7ull);
              if (SIG ? (ksig == ZK_MIX2) : (s < nsteps && (sd[s] & 7u) == ZK_MIX2)) w2u[s] = (int)__shfl_sync(FULL, v0, (int)((sd[s] >> 3) & 31u));
            }
            // ---- predictions of warp A's components ----
            BAR_SYNC64(barP);
            int pv = (int)pp[lane];                              // pv = prediction of component `lane` (warp A's lanes for now)
            if (type == ZT_CM || type == ZT_CONS) pv = p_ownB;
            int pu[ZPQ_FAST_STEPS];                              // outputs of the steps, warp-uniform
            int xjs[ZPQ_FAST_STEPS], xks[ZPQ_FAST_STEPS];        // inputs of MIX2 steps (training), warp-uniform
            u32 ssek0 = 0, ssek1 = 0, ssev0 = 0, ssev1 = 0;     // SSE: bucket index / value to train, per SSE
            int pfin = 0;
#pragma unroll
            for (int s = 0; s < ZPQ_FAST_STEPS; ++s) {
              pu[s] = 0; xjs[s] = 0; xks[s] = 0;
              const u32 ksig = (u32)((SIG >> (3 * s)) & 7ull);
              if (SIG ? (ksig != 0u && ksig != ZK_ISSE) : (s < nsteps && (sd[s] & 7u) != ZK_ISSE)) {
                const u32 d = sd[s];
                const u32 kind = SIG ? ksig : (d & 7u);
                const int owner = (int)((d >> 3) & 31u), sj = (int)((d >> 8) & 31u), sk = (int)((d >> 13) & 63u);
                int res;
                if (kind == ZK_MIX) {
                  const u32 q = SIG ? (u32)sig7_count(SIG, s, ZK_MIX) : d >> 27;
                  const int x = sj == 0 ? pv : __shfl_sync(FULL, pv, (sj + lane) & 31);      // sj is warp-uniform
                  const int w = q == 0 ? mixw0 : (q == 1 ? mixw1 : mixw2);
                  if (q == 0) mixx0 = x; else if (q == 1) mixx1 = x; else mixx2 = x;
                  const int sum = __reduce_add_sync(FULL, lane < sk ? (w >> 8) * x : 0);
                  res = clamp2k(sum >> 8);
                } else {
                  // scalar steps: inputs from the uniform outputs of earlier steps, else from warp A's predictions
                  const int pj = __shfl_sync(FULL, pv, sj);                      // owner lanes hold the outputs of earlier steps
                  if (kind == ZK_SSE) {
                    int pq = min(max(pj + 992, 0), 1983);
                    const int wt = pq & 63;
                    pq >>= 6;
                    const bool first = (SIG ? (u32)sig7_count(SIG, s, ZK_SSE) : d >> 27) == 0;
                    const u32 rowv = first ? sserow0 : sserow1;
                    const u32 q0 = __shfl_sync(FULL, rowv, pq), q1 = __shfl_sync(FULL, rowv, pq + 1);
                    res = t_stretch(S, (u32)(((int)(q0 >> 10) * (64 - wt) + (int)(q1 >> 10) * wt) >> 13));
                    const u32 kk = (u32)pq + (u32)(wt >> 5), vv = (wt >> 5) ? q1 : q0;
                    if (first) { ssek0 = kk; ssev0 = vv; } else { ssek1 = kk; ssev1 = vv; }
                  } else {
                    const int pk = __shfl_sync(FULL, pv, sk & 31);
                    xjs[s] = pj; xks[s] = pk;
                    if (kind == ZK_MIX2) { const int w = w2u[s]; res = (w * pj + (65536 - w) * pk) >> 16; }
                    else { const int wt = (int)((d >> 19) & 255u); res = (pj * wt + pk * (256 - wt)) >> 8; }       // ZK_AVG
                  }
                }
                pu[s] = res;
                if (lane == owner) pv = res;
                if (owner == n - 1) pfin = res;
              }
            }
            if (!fin_in_b) pfin = __shfl_sync(FULL, pv, n - 1);
            const int y = dec7_bit(C, (u32)t_squash(S, pfin) * 2u + 1u);
            if (C.status != ST_OK) { status = C.status; SEND7(2, 0); failed = true; break; }
            SEND7(0, y);
            ylast = y;
            // ---- train own components (libzpaq.cpp:1954-2052) ----
            // squash of every step output at once: an owner lane holds its component's prediction in pv
            const int sqv = t_squash(S, pv);
#pragma unroll
            for (int s = 0; s < ZPQ_FAST_STEPS; ++s) {
              const u32 ksig = (u32)((SIG >> (3 * s)) & 7ull);
              if (SIG ? (ksig == ZK_MIX2 || ksig == ZK_SSE) : (s < nsteps && ((sd[s] & 7u) == ZK_MIX2 || (sd[s] & 7u) == ZK_SSE))) {
                const u32 d = sd[s];
                const u32 kind = SIG ? ksig : (d & 7u);
                const int owner = (int)((d >> 3) & 31u);
                if (kind == ZK_MIX2) {
                  if (lane == owner) {
                    const int err = ((y * 32767 - sqv) * (int)a4) >> 5;
                    int w = (int)v0;
                    w += (err * (xjs[s] - xks[s]) + (1 << 12)) >> 13;
                    w = min(max(w, 0), 65535);
                    reinterpret_cast<u16*>(gtab)[cxt] = (u16)w;
                  }
                } else {
                  const bool first = (SIG ? (u32)sig7_count(SIG, s, ZK_SSE) : d >> 27) == 0;
                  if (lane == owner) {
                    const u32 vv = first ? ssev0 : ssev1, kk = first ? ssek0 : ssek1;
                    const u32 count = vv & 0x3ffu;
                    const int err = y * 32767 - (int)(vv >> 17);
                    const u32 prod = (u32)err * (u32)__ldg(gdt + count);
                    gtab[cxt + kk] = vv + (prod & 0xfffffc00u) + (count < limit ? 1u : 0u);
                  }
                }
              }
            }
#pragma unroll
            for (int q = 0; q < ZPQ_FAST_MIX; ++q)
              if (q < NMIXC && q < nmix) {
                const int err = ((y * 32767 - __shfl_sync(FULL, sqv, (int)mix_owner[q])) * (int)mix_rate[q]) >> 4;       // warp-uniform
                const int w = q == 0 ? mixw0 : (q == 1 ? mixw1 : mixw2);
                const int x = q == 0 ? mixx0 : (q == 1 ? mixx1 : mixx2);
                if (lane < (int)mix_m[q]) mix_tab[q][mixbase[q] + lane] = clamp512k(w + ((err * x + (1 << 12)) >> 13));
              }
            if (type == ZT_CM) {                                  // train(), libzpaq.h:1151-1157
              const u32 count = cmv & 0x3ffu;
              const int err = y * 32767 - (int)(cmv >> 17);
              gtab[cxaB] = cmv + (((u32)err * (u32)__ldg(gdt + count)) & 0xfffffc00u) + (count < limit ? 1u : 0u);
            }
            const int c8n = c8 * 2 + y;
            if (c8n >= 256) { cbyte = c8n - 256; BYTE_END7(cbyte); }
            else {
              c8 = c8n;
              if (c8n >= 16 && c8n < 32) hmap4 = (hmap4 & 0xf) << 5 | y << 4 | 1;
              else hmap4 = (hmap4 & 0x1f0) | (((hmap4 & 0xf) * 2 + y) & 0xf);
            }
            if (type == ZT_CM) { cxaB = (h ^ (u32)hmap4) & cm_mask; cmv = gtab[cxaB]; p_ownB = t_stretch(S, cmv >> 17); }
          }
          if (failed) break;
          if (C.out_pos < C.out_cap) { if (lane == 0) C.out[C.out_pos] = (u8)cbyte; }
          else { status = ST_FULL; }
          ++C.out_pos;
          if (status != ST_OK) { REPLY_CTRL7(2); break; }
        }
        if (status == ST_OK && C.in_pos != C.in_end) status = ST_CORRUPT;     // segment must end with its 4 zeros
        if (lane == 0) A.seg_out[B.seg_first + sg] = C.out_pos - out_start;
      }
      if (lane == 0) { A.status[blk] = status; A.out_len[blk] = C.out_pos; }
    }
    BAR_SYNC64(barP);
  }
}

// ---------------------------------------------------------------------------
static const unsigned long long kSig7L4 = 1ull | 1ull << 3 | 1ull << 6 | 1ull << 9 | 1ull << 12 | 2ull << 15;
static const unsigned long long kSig7L5 = 1ull | 1ull << 3 | 1ull << 6 | 1ull << 9 | 1ull << 12 | 1ull << 15 | 1ull << 18 |
                                          2ull << 21 | 2ull << 24 | 3ull << 27 | 4ull << 30 | 3ull << 33;
// BASELINE configs[3]: the 17-component chain (8 ISSE levels, one 16-input MIX)
static const unsigned long long kSig7C17 = 1ull | 1ull << 3 | 1ull << 6 | 1ull << 9 | 1ull << 12 | 1ull << 15 | 1ull << 18 | 1ull << 21 | 2ull << 24;

// the model behind a BWT ("ci1": an order-0 ICM and one order-1 ISSE link, libzpaq.cpp:7623, 7644): one chain level, no mixer
static const unsigned long long kSig7Bwt = 1ull;

template <unsigned long long SIG>
static cudaError_t launch7(const ZKernelArgs* a, int ctas, int threads, size_t smem, int pw_bytes, int teams, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(zpq_dec_kernel7<SIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZPQ_SMEM_MAX);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  zpq_dec_kernel7<SIG><<<ctas, threads, smem, st>>>(*a, pw_bytes, teams);
  return cudaGetLastError();
}

extern "C" int zpq_dec7_static_smem(void) { return S7_STATIC; }
extern "C" int zpq_dec7_ctrl_bytes(void) { return S7_CTRL; }

extern "C" cudaError_t zpq_launch_dec7(const ZKernelArgs* a, int nblk, int sm_count, int teams, int pw_bytes,
                                       unsigned long long sig, cudaStream_t st) {
  if (nblk <= 0) return cudaSuccess;
  if (teams > 7) teams = 7;                                       // 14 named barriers
  const size_t smem = (size_t)S7_STATIC + (size_t)teams * (pw_bytes + S7_CTRL);
  int ctas = (nblk + teams - 1) / teams;
  if (ctas > sm_count) ctas = sm_count;
  const int th = teams * 64;
  if (sig == kSig7L5) return launch7<kSig7L5>(a, ctas, th, smem, pw_bytes, teams, st);
  if (sig == kSig7L4) return launch7<kSig7L4>(a, ctas, th, smem, pw_bytes, teams, st);
  if (sig == kSig7C17) return launch7<kSig7C17>(a, ctas, th, smem, pw_bytes, teams, st);
  if (sig == kSig7Bwt) return launch7<kSig7Bwt>(a, ctas, th, smem, pw_bytes, teams, st);
  return cudaErrorInvalidValue;        // zpq_dec7_takes(sig) is false: the caller uses zpq_kernels6.cu
}

// Step signatures this kernel is instantiated for: the two makeConfig families, the 17-component chain and the BWT model.  Any other
// step list is decoded by the round-1 two-warp kernel, whose generic instantiation reads the list at run time (a fully
// run-time step list here exceeds ptxas' predicate budget: every step kind of every step becomes a live predicate).
extern "C" int zpq_dec7_takes(unsigned long long sig) { return sig == kSig7L5 || sig == kSig7L4 || sig == kSig7C17 || sig == kSig7Bwt; }
