// zpq_kernels6.cu -- DECODE with two cooperating warps per archive block, sm_100a.
//
// Decoding is sequential per bit (bit t+1 needs bit t), and with one warp per block a single
// scheduler issues the whole ~700-instruction bit loop in order (profiles/r01_decode_*).  The work of
// one bit splits into two streams that only meet twice:
//
//   warp A  every ICM / ISSE component: hash rows, shared-memory cm tables, the ISSE chain (shuffles
//           inside the warp).  Publishes its predictions in a 64-byte shared array, then waits for y.
//   warp B  CM, MATCH, CONST, AVG, MIX2, SSE, every MIX and the arithmetic decoder.  Looks up its own
//           tables and fetches MIX / SSE rows while A walks the chain, then mixes, decodes y, publishes it.
//
// After y both warps train their own components at the same time, and at byte boundaries both
// evaluate the (flattened, lane-private) context recipes of their own lanes.  Lane = component index in
// both warps (a lane is idle in the warp that does not own its component), so the code of
// zpq_kernels2.cu carries over.  Hand-offs are two words in shared memory:
//     flagA = seq            "predictions of bit seq are in pp[]"
//     flagB = seq<<3|kind<<1|y   kind 0: decoded bit y, 1: end of segment, 2: abort (error in B)
// Bit-exactness: same operations on the same private tables in the same order as the one-warp kernel.
#include <cuda_runtime.h>
#include <stdint.h>
#include "zpq_dev.h"
#include "zpq_kernels.h"
#include "zpq_kcommon.cuh"

#define S6_STATIC (S2_DT2K)                 // the dt2k table is not staged: MATCH computes 2048/len directly
#define S6_CTRL 80                          // per block: short pp[32], flagA, flagB, block index, pad
#define PAIR_BAR(team) asm volatile("bar.sync %0, %1;" ::"r"((team) + 1), "r"(64) : "memory")
#define SEQ_MASK 0x1fffffffu
#define SPIN_CAP (1u << 27)                 // polls of a hand-shake word before the block is given up (seconds, not microseconds)
#define ST_HUNG (-9)                        // reported as ZPQ_E_CUDA: an internal hand-shake was lost

struct Dec6 {
  u32 low, high, curr;
  const u8* in; u64 in_pos, in_end;
  u8* out; u64 out_pos, out_cap;
  int status;
};

__device__ __forceinline__ int dec6_bit(Dec6& C, u32 p16) {
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

// steps of kind `kind` among the first `upto` steps of a signature: the weight / row register slot of a step
__host__ __device__ constexpr int sig_count(unsigned long long sig, int upto, unsigned kind) {
  int c = 0;
  for (int s = 0; s < upto; ++s) if (((sig >> (3 * s)) & 7ull) == kind) ++c;
  return c;
}

// V bit 0: CM / CONST components live in warp A, which requests both candidate entries of the next bit
//          while it waits for the decoder;
// V bit 1: warp B requests both candidate MIX / SSE rows of the next bit at the top of the current bit;
// V bit 2: context hashes from the slot-scheduled plan (ZHashPlan) instead of lane-by-lane interpretation;
// V bit 4: warp B keeps the predictions in registers through the steps (shuffles instead of shared-memory round trips);
// V bit 3: the MATCH component lives in warp A too (its byte-boundary search runs in A's slack, off warp B's critical path).
template <unsigned long long SIG, int V>
__global__ void __launch_bounds__(448, 1) zpq_dec_kernel6(const ZKernelArgs A, const int pw_bytes, const int nteams) {
  extern __shared__ __align__(16) unsigned char smem[];
  {
    const ZTables* T = A.tables;
    const uint4* src = reinterpret_cast<const uint4*>(T->stretch_hi);
    uint4* dst = reinterpret_cast<uint4*>(smem + S2_STRETCH);
    for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) dst[i] = src[i];
    u16* sq = reinterpret_cast<u16*>(smem + S2_SQUASH);
    for (int i = threadIdx.x; i < 1344; i += blockDim.x) sq[i] = T->squash[1376 + i];
    u16* dt = reinterpret_cast<u16*>(smem + S2_DT);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) dt[i] = (u16)(T->dt[i] >> 1);
    u16* ns = reinterpret_cast<u16*>(smem + S2_NS);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) ns[i] = (u16)(T->ns[i] & 0xffffu);
  }
  __syncthreads();
  Tabs S = make_tabs(smem);
  S.dt2k = nullptr;
  u32* sm32 = reinterpret_cast<u32*>(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp >> 1, role = warp & 1;
  const u32 wbase = (u32)(S6_STATIC + team * pw_bytes) / 4u;
  unsigned char* ctrlb = smem + S6_STATIC + (size_t)nteams * pw_bytes + (size_t)team * S6_CTRL;
  volatile short* pp = reinterpret_cast<volatile short*>(ctrlb);
  volatile u32* flagA = reinterpret_cast<volatile u32*>(ctrlb + 64);
  volatile u32* flagB = flagA + 1;
  volatile u32* blkw = flagA + 2;

  for (;;) {
    if (role == 0 && lane == 0) { *flagA = 0; *flagB = 0; *blkw = atomicAdd(A.counter, 1u); }
    PAIR_BAR(team);
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
      const bool hashc = cd.type == ZT_ICM || cd.type == ZT_ISSE || ((V & 1) && (cd.type == ZT_CM || cd.type == ZT_CONS)) ||
                         ((V & 8) && cd.type == ZT_MATCH);
      if ((role == 0) == hashc) {
        type = cd.type; a1 = cd.a1; a2 = cd.a2; a3 = cd.a3; a4 = cd.a4; a5 = cd.a5;
        cm_mask = cd.cm_mask; ht_mask = cd.ht_mask; limit = cd.limit; level = ZM->fast.level[lane];
        gtab = reinterpret_cast<u32*>(arena + (u64)cd.cm_off16 * 16);
        ht = arena + (u64)cd.ht_off16 * 16;
      }
    }
    const bool mine = type != ZT_NONE;
    // context hashes: V bit 2 = slot-scheduled plan (zpq_hcomp.cpp build_hash_plan; the host only picks this
    // instantiation when every model of the launch has one), else the flat recipes
    const ZHashPlan* __restrict__ HP = &ZM->hplan;
    const u32 mine_mask = __ballot_sync(FULL, mine);
    const u32 hp_last = (lane < n) ? (u32)HP->last_seg[lane] : 0u;
    const u32 hp_fold = (lane < n) ? (u32)HP->fold_n[lane] : 0u;
    const u32 tabw = wbase + ((lane < n) ? ZM->fast.smem_off[lane] : 0u) / 4u;
    if (role == 0) {
      if (type == ZT_ICM || type == ZT_ISSE) { /* tables initialised below by the whole warp */ }
      __syncwarp();
      for (int src = 0; src < n; ++src) {
        const u32 ty = __shfl_sync(FULL, type, src);
        const u32 tw = __shfl_sync(FULL, tabw, src);
        if (ty == ZT_ICM) for (int i = lane; i < 256; i += 32) sm32[tw + i] = A.tables->icm_init[i];
        else if (ty == ZT_ISSE) for (int i = lane; i < 512; i += 32) sm32[tw + i] = A.tables->isse_init[i];
      }
      __syncwarp();
    }
    // context recipes: flattened, lane-private (zpq_hcomp.cpp build_wave_plan)
    const u32* const fops = ZM->wave.flat[lane < n ? lane : 0];
    const u32 flen = mine ? ZM->wave.flat_len[lane] : 0u;
    const u32 dist_byte = mine ? ZM->pipe.dist_byte[lane] : 0u;
    u32 st[ZPQ_FLAT_SEGS] = {0, 0, 0, 0};
    u32 dist_last = 0, vm_c = 0; u64 hist = 0;
    u8* const mring = arena + (u64)ZM->m_off16 * 16;
    const u32 mmask = ZM->m_mask;

    u32 h = 0;
    int p = (type == ZT_CONS) ? ((int)a1 - 128) * 4 : 0;
    int c8 = 1, hmap4 = 1;
    u32 seq = 1;
    int status = ST_OK;

#define BYTE_END(byteval)                                                                            \
    do {                                                                                             \
      const u32 bv_ = (u32)(byteval);                                                                \
      vm_c -= 1; mring[vm_c & mmask] = (u8)bv_;          /* both warps write the same value */       \
      hist = (hist << 8) | (u64)bv_;                                                                 \
      if (mine && bv_ == dist_byte) dist_last = vm_c;                                                \
      if (V & 4) h = hplan_eval(HP, mine_mask, lane, hp_last, hp_fold, st, bv_, hist, vm_c, mring, mmask, dist_last, h);   \
      else if (mine) h = flat_eval(fops, flen, st, bv_, hist, (int)(0u - vm_c) - 1, mring, dist_last, mmask);     \
      hmap4 = 1; c8 = 1;                                                                             \
    } while (0)

    if (role == 0) {
      // =================================== warp A: ICM / ISSE ===================================
      uint4 row = make_uint4(0, 0, 0, 0); u32 row_off = 0;
      u32 bh = 0, v0 = 0, v1 = 0; int xj = 0;
      u32 cxa = 0, cand0 = 0, cand1 = 0; bool cand = false; int ylast = 0;
      u32 ma = 0, mb = 0, mpos = 0, mbit = 0, mpred = 0, mc = 0, mdt = 0;      // MATCH state (V bit 3)
      const bool hashl = type == ZT_ICM || type == ZT_ISSE;
      const int nlev = (int)ZM->fast.nlev;
      bool abort = false, hungA = false;
      for (u32 sg = 0; sg < B.nseg && !abort; ++sg) {
        for (;;) {
          // ---- predict ----
          const bool nib = (c8 == 1) || ((c8 & 0xf0) == 16);
          if ((V & 1) && type == ZT_CM) {
            cxa = (h ^ (u32)hmap4) & cm_mask;
            v0 = cand ? (ylast ? cand1 : cand0) : gtab[cxa];
            p = t_stretch(S, v0 >> 17);
          }
          if ((V & 8) && type == ZT_MATCH) {
            if (ma == 0) p = 0;
            else {
              mc = (mpred >> (7 - mbit)) & 1u;
              p = t_stretch(S, (u32)((int)mdt * (1 - 2 * (int)mc)) & 32767u);        // mdt = dt2k[ma] = 2048 / ma
            }
          }
          if ((V & 1) ? hashl : (mine && type != ZT_MATCH)) {
            if (nib) d_find(ht, ht_mask, (int)a1 + 2, h + 16u * (u32)c8, row, row_off);
            bh = row_get(row, hmap4 & 15);
            if (type == ZT_ISSE) {
              const uint2 w = *reinterpret_cast<const uint2*>(sm32 + tabw + bh * 2);
              v0 = w.x; v1 = w.y;
            } else {
              v0 = sm32[tabw + bh];
              p = t_stretch(S, v0 >> 8);
            }
          }
#pragma unroll 1
          for (int lev = 1; lev <= nlev; ++lev) {
            const int pj = __shfl_sync(FULL, p, (int)a2 & 31);
            if (type == ZT_ISSE && level == (u32)lev) {
              xj = pj;
              p = clamp2k(((int)v0 * pj + (int)v1 * 64) >> 16);
            }
          }
          if (mine) pp[lane] = (short)p;
          __threadfence_block();
          __syncwarp();
          if (lane == 0) *flagA = seq;
          if ((V & 1) && type == ZT_CM) {
            // both entries the next bit can use (hmap4 values of one byte are distinct and < 512)
            cand = c8 < 128 && cm_mask >= 511u;
            if (cand) {
              u32 hm0;
              if (c8 >= 8 && c8 < 16) hm0 = (u32)((hmap4 & 0xf) << 5 | 1);
              else hm0 = (u32)((hmap4 & 0x1f0) | (((hmap4 & 0xf) * 2) & 0xf));
              const u32 hm1 = hm0 | ((c8 >= 8 && c8 < 16) ? 16u : 1u);
              cand0 = gtab[(h ^ hm0) & cm_mask];
              cand1 = gtab[(h ^ hm1) & cm_mask];
            }
          }
          // ---- wait for the decoder ----
          u32 msg;
          // (spin cap: a lost hand-shake becomes an aborted block with status ST_HUNG instead of a hung GPU)
          for (u32 spin = 0; ((msg = *flagB) >> 3) != (seq & SEQ_MASK);) { if (++spin > SPIN_CAP) { msg = 2u << 1; hungA = true; break; } }
          __threadfence_block();
          seq = (seq + 1) & SEQ_MASK;
          const u32 kind = (msg >> 1) & 3u;
          if (kind == 2u) { abort = true; break; }
          if (kind == 1u) break;                              // end of segment (the prediction above was not used)
          const int y = (int)(msg & 1u);
          // ---- update ----
          if ((V & 1) && type == ZT_CM) {
            const u32 count = v0 & 0x3ffu;
            const int err = y * 32767 - (int)(v0 >> 17);
            const u32 prod = (u32)err * ((u32)S.dt[count] * 2u);
            gtab[cxa] = v0 + (prod & 0xfffffc00u) + (count < limit ? 1u : 0u);
            ylast = y;
          }
          if ((V & 8) && type == ZT_MATCH) {
            // MATCH update (libzpaq.cpp:1990-2008); h is still the hash as of the previous byte here (gotcha 6)
            if ((int)mc != y) ma = 0;
            if (++mbit == 8) {
              mbit = 0;
              ht[mpos] = (u8)((c8 * 2 + y) & 255);
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
            }
          }
          if ((V & 1) ? hashl : (mine && type != ZT_MATCH)) {
            const u32 nx = ((u32)S.ns[bh] >> (8 * y)) & 255u;
            row_set(row, hmap4 & 15, nx);
            if (type == ZT_ISSE) {
              const int err = y * 32767 - t_squash(S, p);
              const int w0 = clamp512k((int)v0 + ((err * xj + (1 << 12)) >> 13));
              const int w1 = clamp512k((int)v1 + ((err + 16) >> 5));
              *reinterpret_cast<uint2*>(sm32 + tabw + bh * 2) = make_uint2((u32)w0, (u32)w1);
            } else {
              sm32[tabw + bh] = v0 + (u32)(((int)(y * 32767 - (int)(v0 >> 8))) >> 2);
            }
          }
          const int c8n = c8 * 2 + y;
          if (((V & 1) ? hashl : (mine && type != ZT_MATCH)) && (c8n >= 256 || (c8n >= 16 && c8n < 32))) *reinterpret_cast<uint4*>(ht + row_off) = row;
          if (c8n >= 256) BYTE_END(c8n - 256);
          else {
            c8 = c8n;
            if (c8n >= 16 && c8n < 32) hmap4 = (hmap4 & 0xf) << 5 | y << 4 | 1;
            else hmap4 = (hmap4 & 0x1f0) | (((hmap4 & 0xf) * 2 + y) & 0xf);
          }
        }
      }
      if (hungA && lane == 0) { A.status[blk] = ST_HUNG; A.out_len[blk] = 0; }     // warp B never answered
    } else {
      // ================ warp B: CM, MATCH, CONST, AVG, MIX2, SSE, MIX, arithmetic decoder ================
      constexpr int NMIXC = SIG ? sig_count(SIG, ZPQ_FAST_STEPS, ZK_MIX) : ZPQ_FAST_MIX;     // known from the signature
      constexpr int NSSEC = SIG ? sig_count(SIG, ZPQ_FAST_STEPS, ZK_SSE) : ZPQ_FAST_SSE;
      const int nsteps = (int)ZM->fast.nsteps, nmix = SIG ? NMIXC : (int)ZM->fast.nmix, nsse = SIG ? NSSEC : (int)ZM->fast.nsse;
      u32 sd[ZPQ_FAST_STEPS];
#pragma unroll
      for (int s = 0; s < ZPQ_FAST_STEPS; ++s) sd[s] = ZM->fast.sd[s];
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
      u32 cxt = 0, v0 = 0; int xj = 0, xk = 0;
      u32 ma = 0, mb = 0, mpos = 0, mbit = 0, mpred = 0, mc = 0, mdt = 0;
      // MIX: ((h + (c8 & mask)) & (size-1)) * m   MIX2: (h + (c8 & mask)) & (size-1)   SSE: ((h + c8) * 32) & (32*size-1)
      const bool ctx_on = type == ZT_MIX || type == ZT_MIX2 || type == ZT_SSE;
      const u32 ctx_m1 = type == ZT_SSE ? 0xffffffffu : a5;
      const u32 ctx_m2 = type == ZT_SSE ? cm_mask >> 5 : cm_mask;
      const u32 ctx_mul = type == ZT_MIX ? a3 : (type == ZT_SSE ? 32u : 1u);
      // candidate rows of the next bit (V bit 1): valid when every MIX has a full c8 mask and >= 256 rows and
      // every SSE has >= 256 rows, so that the two candidate rows differ from the row trained in this bit
      int mc00 = 0, mc01 = 0, mc10 = 0, mc11 = 0, mc20 = 0, mc21 = 0; u32 sc00 = 0, sc01 = 0, sc10 = 0, sc11 = 0;
      bool candv = false; int ylast = 0;
      const bool cand_ok = (V & 2) && __all_sync(FULL, (type != ZT_MIX || ((a5 & 255u) == 255u && cm_mask >= 255u)) &&
                                                       (type != ZT_SSE || cm_mask >= 256u * 32u - 1u));

      Dec6 C;
      C.low = 1; C.high = 0xffffffffu; C.curr = 0; C.status = ST_OK;
      C.in = A.in + B.in_off; C.in_pos = 0; C.in_end = 0;
      C.out = A.out + B.out_off; C.out_pos = 0; C.out_cap = B.out_cap;

#define SEND(kind, y) do { __threadfence_block(); __syncwarp(); if (lane == 0) *flagB = (seq & SEQ_MASK) << 3 | (u32)(kind) << 1 | (u32)(y); seq = (seq + 1) & SEQ_MASK; } while (0)
// Control messages (end of segment, abort) leave the per-bit hand-shake: flagB is a one-word mailbox, so before it is
// overwritten warp A must have consumed the previous message -- which it has once it published the prediction it is about
// to wait on (flagA == seq).
#define SEND_CTRL(kind) do { for (u32 spin_ = 0; *flagA != seq;) { if (++spin_ > SPIN_CAP) break; } SEND(kind, 0); } while (0)

      for (u32 sg = 0; sg < B.nseg && status == ST_OK; ++sg) {
        const u64 seg_n = A.seg_len[B.seg_first + sg];
        const u64 out_start = C.out_pos;
        C.in_end = C.in_pos + seg_n;
        C.curr = 0;
        for (int i = 0; i < 4; ++i) {
          if (C.in_pos >= C.in_end) { status = ST_EOF; break; }
          C.curr = C.curr << 8 | (u32)C.in[C.in_pos++];
        }
        if (status != ST_OK) { SEND_CTRL(2); break; }
        for (;;) {
          const int eos = dec6_bit(C, 0);
          if (C.status != ST_OK) { status = C.status; SEND_CTRL(2); break; }
          if (eos) {
            if (C.curr != 0) { status = ST_EOS; SEND_CTRL(2); }
            else SEND_CTRL(1);
            break;
          }
          int cbyte = 0;
          bool failed = false;
          for (int bit = 7; bit >= 0; --bit) {
            // ---- own table lookups and early row fetches (overlap warp A's ISSE chain) ----
            if (!(V & 1) && type == ZT_CM) {
              cxt = (h ^ (u32)hmap4) & cm_mask;
              v0 = gtab[cxt];
              p = t_stretch(S, v0 >> 17);
            }
            else if (type == ZT_MATCH) {
              if (ma == 0) p = 0;
              else {
                mc = (mpred >> (7 - mbit)) & 1u;
                p = t_stretch(S, (u32)((int)mdt * (1 - 2 * (int)mc)) & 32767u);      // mdt = dt2k[ma] = 2048 / ma
              }
            }
            // row / entry of the components indexed by h + c8: one formula with per-lane constants (set per block).
            // cxn0 / cxn1 = the row of the NEXT bit when this bit decodes to 0 / 1
            u32 cxn0 = 0, cxn1 = 0;
            if (ctx_on) {
              cxt = ((h + ((u32)c8 & ctx_m1)) & ctx_m2) * ctx_mul;
              if (V & 2) {
                cxn0 = ((h + ((u32)(c8 * 2) & ctx_m1)) & ctx_m2) * ctx_mul;
                cxn1 = ((h + ((u32)(c8 * 2 + 1) & ctx_m1)) & ctx_m2) * ctx_mul;
              }
              if (type == ZT_MIX2) v0 = reinterpret_cast<const u16*>(gtab)[cxt];
            }
            const bool use_cand = (V & 2) && candv;
            candv = (V & 2) && cand_ok && c8 < 128;
#pragma unroll
            for (int q = 0; q < ZPQ_FAST_MIX; ++q)
              if (q < NMIXC && q < nmix) {
                const u32 base = __shfl_sync(FULL, cxt, (int)mix_owner[q]);
                mixbase[q] = base;
                int wq;
                if (use_cand) wq = q == 0 ? (ylast ? mc01 : mc00) : (q == 1 ? (ylast ? mc11 : mc10) : (ylast ? mc21 : mc20));
                else wq = lane < (int)mix_m[q] ? mix_tab[q][base + lane] : 0;
                if (q == 0) mixw0 = wq; else if (q == 1) mixw1 = wq; else mixw2 = wq;
                if ((V & 2) && candv) {
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
                if ((V & 2) && candv) {
                  const u32 nb0 = __shfl_sync(FULL, cxn0, (int)sse_owner[r]), nb1 = __shfl_sync(FULL, cxn1, (int)sse_owner[r]);
                  const u32 n0 = sse_tab[r][nb0 + lane], n1 = sse_tab[r][nb1 + lane];
                  if (r == 0) { sc00 = n0; sc01 = n1; } else { sc10 = n0; sc11 = n1; }
                }
              }
            if (type == ZT_CM || type == ZT_MATCH || type == ZT_CONS) pp[lane] = (short)p;
            // ---- predictions of the hash components ----
            bool hungB = false;
            for (u32 spin = 0; *flagA != seq;) { if (++spin > SPIN_CAP) { hungB = true; break; } }
            if (hungB) { status = ST_HUNG; SEND(2, 0); failed = true; break; }      // warp A never published: give the block up
            __threadfence_block();
            __syncwarp();
            // V bit 4: the predictions stay in registers from here on (pv = prediction of component `lane`): a step reads
            // its inputs with shuffles and its result replaces pv in the owner lane, so the steps need no shared-memory
            // round trip (store, __syncwarp, load) between them
            int pv = (V & 16) ? (int)pp[lane] : 0;
            int pm0 = 0, pm1 = 0, pm2 = 0;                    // mixer outputs (warp-uniform), for the training below
            // ---- remaining steps ----
#pragma unroll
            for (int s = 0; s < ZPQ_FAST_STEPS; ++s) {
              const u32 ksig = (u32)((SIG >> (3 * s)) & 7ull);
              if (SIG ? (ksig != 0u) : (s < nsteps)) {
                const u32 d = sd[s];
                const u32 kind = SIG ? ksig : (d & 7u);
                const int owner = (int)((d >> 3) & 31u), sj = (int)((d >> 8) & 31u), sk = (int)((d >> 13) & 63u);
                if (kind == ZK_ISSE) { /* done by warp A */ }
                else if (kind == ZK_MIX) {
                  const u32 q = SIG ? (u32)sig_count(SIG, s, ZK_MIX) : d >> 27;
                  int x;
                  if (V & 16) x = sj == 0 ? pv : __shfl_sync(FULL, pv, (sj + lane) & 31);      // sj is warp-uniform
                  else x = (int)pp[(sj + lane) & 31];
                  const int w = q == 0 ? mixw0 : (q == 1 ? mixw1 : mixw2);
                  if (q == 0) mixx0 = x; else if (q == 1) mixx1 = x; else mixx2 = x;
                  const int sum = __reduce_add_sync(FULL, lane < sk ? (w >> 8) * x : 0);
                  if (V & 16) {
                    const int pu = clamp2k(sum >> 8);              // the reduction is warp-uniform: every lane has the output
                    if (q == 0) pm0 = pu; else if (q == 1) pm1 = pu; else pm2 = pu;
                    if (lane == owner) { p = pu; pv = pu; }
                  } else {
                    if (lane == owner) { p = clamp2k(sum >> 8); pp[lane] = (short)p; }
                    __syncwarp();
                  }
                }
                else if (kind == ZK_MIX2) {
                  const int pj = (V & 16) ? __shfl_sync(FULL, pv, sj) : (int)pp[sj];
                  const int pk = (V & 16) ? __shfl_sync(FULL, pv, sk & 31) : (int)pp[sk & 31];
                  if (lane == owner) {
                    xj = pj; xk = pk;
                    const int w = (int)v0;
                    p = (w * pj + (65536 - w) * pk) >> 16;
                    if (V & 16) pv = p; else pp[lane] = (short)p;
                  }
                  if (!(V & 16)) __syncwarp();
                }
                else if (kind == ZK_SSE) {
                  const int pj = (V & 16) ? __shfl_sync(FULL, pv, sj) : (int)pp[sj];
                  int pq = min(max(pj + 992, 0), 1983);
                  const int wt = pq & 63;
                  pq >>= 6;
                  const u32 rowv = (SIG ? (u32)sig_count(SIG, s, ZK_SSE) : d >> 27) == 0 ? sserow0 : sserow1;
                  const u32 q0 = __shfl_sync(FULL, rowv, pq), q1 = __shfl_sync(FULL, rowv, pq + 1);
                  if (lane == owner) {
                    p = t_stretch(S, (u32)(((int)(q0 >> 10) * (64 - wt) + (int)(q1 >> 10) * wt) >> 13));
                    if (wt >> 5) { cxt += (u32)pq + 1u; v0 = q1; } else { cxt += (u32)pq; v0 = q0; }
                    if (V & 16) pv = p; else pp[lane] = (short)p;
                  }
                  if (!(V & 16)) __syncwarp();
                }
                else {   // ZK_AVG
                  const int pj = (V & 16) ? __shfl_sync(FULL, pv, sj) : (int)pp[sj];
                  const int pk = (V & 16) ? __shfl_sync(FULL, pv, sk & 31) : (int)pp[sk & 31];
                  const int wt = (int)((d >> 19) & 255u);
                  if (lane == owner) { p = (pj * wt + pk * (256 - wt)) >> 8; if (V & 16) pv = p; else pp[lane] = (short)p; }
                  if (!(V & 16)) __syncwarp();
                }
              }
            }
            const int pfin = (V & 16) ? __shfl_sync(FULL, pv, n - 1) : (int)pp[n - 1];
            const int y = dec6_bit(C, (u32)t_squash(S, pfin) * 2u + 1u);
            if (C.status != ST_OK) { status = C.status; SEND(2, 0); failed = true; break; }
            SEND(0, y);
            ylast = y;
            // ---- update own components ----
            if (type == ZT_CM || type == ZT_SSE) {
              const u32 count = v0 & 0x3ffu;
              const int err = y * 32767 - (int)(v0 >> 17);
              const u32 prod = (u32)err * ((u32)S.dt[count] * 2u);
              gtab[cxt] = v0 + (prod & 0xfffffc00u) + (count < limit ? 1u : 0u);
            }
            else if (type == ZT_MIX2) {
              const int err = ((y * 32767 - t_squash(S, p)) * (int)a4) >> 5;
              int w = (int)v0;
              w += (err * (xj - xk) + (1 << 12)) >> 13;
              w = min(max(w, 0), 65535);
              reinterpret_cast<u16*>(gtab)[cxt] = (u16)w;
            }
            else if (type == ZT_MATCH) {
              if ((int)mc != y) ma = 0;
              if (++mbit == 8) {
                mbit = 0;
                ht[mpos] = (u8)((c8 * 2 + y) & 255);
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
                mdt = ma ? 2048u / ma : 0u;                 // once per byte instead of once per bit
              }
            }
#pragma unroll
            for (int q = 0; q < ZPQ_FAST_MIX; ++q)
              if (q < NMIXC && q < nmix) {
                int err;
                if (V & 16) err = ((y * 32767 - t_squash(S, q == 0 ? pm0 : (q == 1 ? pm1 : pm2))) * (int)mix_rate[q]) >> 4;   // uniform
                else {
                  const int e0 = ((y * 32767 - t_squash(S, p)) * (int)mix_rate[q]) >> 4;
                  err = __shfl_sync(FULL, e0, (int)mix_owner[q]);
                }
                const int w = q == 0 ? mixw0 : (q == 1 ? mixw1 : mixw2);
                const int x = q == 0 ? mixx0 : (q == 1 ? mixx1 : mixx2);
                if (lane < (int)mix_m[q]) mix_tab[q][mixbase[q] + lane] = clamp512k(w + ((err * x + (1 << 12)) >> 13));
              }
            const int c8n = c8 * 2 + y;
            if (c8n >= 256) { cbyte = c8n - 256; BYTE_END(cbyte); }
            else {
              c8 = c8n;
              if (c8n >= 16 && c8n < 32) hmap4 = (hmap4 & 0xf) << 5 | y << 4 | 1;
              else hmap4 = (hmap4 & 0x1f0) | (((hmap4 & 0xf) * 2 + y) & 0xf);
            }
          }
          if (failed) break;
          if (C.out_pos < C.out_cap) { if (lane == 0) C.out[C.out_pos] = (u8)cbyte; }
          else { status = ST_FULL; }
          ++C.out_pos;
          if (status != ST_OK) {
            // warp A is about to publish the first bit of the next byte: consume it and stop the block
            SEND_CTRL(2);
            break;
          }
        }
        if (status == ST_OK && C.in_pos != C.in_end) status = ST_CORRUPT;     // segment must end with its 4 zeros
        if (lane == 0) A.seg_out[B.seg_first + sg] = C.out_pos - out_start;
      }
      if (lane == 0) { A.status[blk] = status; A.out_len[blk] = C.out_pos; }
    }
    PAIR_BAR(team);
  }
}

// ---------------------------------------------------------------------------
static const unsigned long long kSig6L4 = 1ull | 1ull << 3 | 1ull << 6 | 1ull << 9 | 1ull << 12 | 2ull << 15;
static const unsigned long long kSig6L5 = 1ull | 1ull << 3 | 1ull << 6 | 1ull << 9 | 1ull << 12 | 1ull << 15 | 1ull << 18 |
                                          2ull << 21 | 2ull << 24 | 3ull << 27 | 4ull << 30 | 3ull << 33;

// BASELINE configs[3]: the 17-component chain (8 ISSE levels, one 16-input MIX)
static const unsigned long long kSig6C17 = 1ull | 1ull << 3 | 1ull << 6 | 1ull << 9 | 1ull << 12 | 1ull << 15 | 1ull << 18 | 1ull << 21 | 2ull << 24;

template <unsigned long long SIG, int V>
static cudaError_t launch6(const ZKernelArgs* a, int ctas, int threads, size_t smem, int pw_bytes, int teams, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(zpq_dec_kernel6<SIG, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZPQ_SMEM_MAX);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  zpq_dec_kernel6<SIG, V><<<ctas, threads, smem, st>>>(*a, pw_bytes, teams);
  return cudaGetLastError();
}

extern "C" int zpq_dec6_static_smem(void) { return S6_STATIC; }
extern "C" int zpq_dec6_ctrl_bytes(void) { return S6_CTRL; }

template <int V>
static cudaError_t launch6v(const ZKernelArgs* a, int ctas, int th, size_t smem, int pw_bytes, int teams, unsigned long long sig, cudaStream_t st) {
  if (sig == kSig6L5) return launch6<kSig6L5, V>(a, ctas, th, smem, pw_bytes, teams, st);
  if (sig == kSig6L4) return launch6<kSig6L4, V>(a, ctas, th, smem, pw_bytes, teams, st);
  if (sig == kSig6C17) return launch6<kSig6C17, V>(a, ctas, th, smem, pw_bytes, teams, st);
  return launch6<0ull, V>(a, ctas, th, smem, pw_bytes, teams, st);
}

extern "C" cudaError_t zpq_launch_dec6(const ZKernelArgs* a, int nblk, int sm_count, int teams, int pw_bytes,
                                       unsigned long long sig, int variant, cudaStream_t st) {
  if (nblk <= 0) return cudaSuccess;
  const size_t smem = (size_t)S6_STATIC + (size_t)teams * (pw_bytes + S6_CTRL);
  int ctas = (nblk + teams - 1) / teams;
  if (ctas > sm_count) ctas = sm_count;
  const int th = teams * 64;
  switch (variant & 31) {
    case 1: return launch6v<1>(a, ctas, th, smem, pw_bytes, teams, sig, st);
    case 2: return launch6v<2>(a, ctas, th, smem, pw_bytes, teams, sig, st);
    case 3: return launch6v<3>(a, ctas, th, smem, pw_bytes, teams, sig, st);
    case 7: return launch6v<7>(a, ctas, th, smem, pw_bytes, teams, sig, st);
    case 15: return launch6v<15>(a, ctas, th, smem, pw_bytes, teams, sig, st);
    case 31: return launch6v<31>(a, ctas, th, smem, pw_bytes, teams, sig, st);
    default: return launch6v<0>(a, ctas, th, smem, pw_bytes, teams, sig, st);
  }
}
