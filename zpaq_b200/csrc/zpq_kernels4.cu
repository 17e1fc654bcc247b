// zpq_kernels4.cu -- ENCODE with three warps per archive block ("team" kernel), sm_100a.
//
// zpq_kernels3.cu showed that encoding is a feed-forward pipeline over components, but with one warp
// per block a single scheduler still issues the whole ~900-instruction bit loop in order (IPC 0.11,
// profiles/r01_v3_*).  Because pipeline stages only exchange predictions that are at least one BYTE
// old, they need no lock-step: this kernel gives every block a team of three warps that run the same
// byte-skewed schedule independently, each with ONE code path:
//
//   role 0  all ICM / ISSE components   (lane = component; hash rows, shared-memory cm tables, recipes)
//   role 1  all MIX components          (lane = input; coalesced weight row, redux.sync dot product;
//                                        the next bit's row is fetched while the current bit trains)
//   role 2  CM, MATCH, CONST, AVG, MIX2, SSE (lane = component; SSE rows fetched by the whole warp)
//
// Predictions travel through per-component rings in the block's arena (32 byte-slots x 8 x int16,
// one 16-byte vector per byte).  A component at pipeline depth d codes byte K-d in byte-iteration K
// (same plan as ZPipe), so everything it reads was written at least one iteration earlier; roles only
// synchronise through three "iterations completed" counters in shared memory:
//   role r may run iteration K   when  done[p] >= K        for every role p it reads from, and
//                                      done[c] + 16 >= K    for every role c that reads it (ring reuse).
// The arithmetic coder rides with the role that owns the last component.  Bit-exactness argument: as in
// zpq_kernels3.cu (private tables, exact p[j] of the same bit); tests run this tier against the oracle.
#include <cuda_runtime.h>
#include <stdint.h>
#include "zpq_dev.h"
#include "zpq_kernels.h"
#include "zpq_kcommon.cuh"

#define TEAM_BAR(team) asm volatile("bar.sync %0, %1;" ::"r"((team) + 1), "r"(96) : "memory")

struct Enc4 {
  u32 low, high;
  u8* out; u64 out_pos, out_cap;
  int status;
};

__device__ __forceinline__ void enc4_bit(Enc4& C, int lane, int y, u32 p16) {
  const u32 mid = C.low + (u32)(((u64)(C.high - C.low) * p16) >> 16);
  if (y) C.high = mid; else C.low = mid + 1;
  while ((C.high ^ C.low) < 0x1000000u) {
    if (C.out_pos < C.out_cap) { if (lane == 0) C.out[C.out_pos] = (u8)(C.high >> 24); }
    else C.status = ST_FULL;
    ++C.out_pos;
    C.high = C.high << 8 | 255u;
    C.low = C.low << 8;
    C.low += (C.low == 0);
  }
}

// element `bit` (0..7) of a vector of eight int16 predictions
__device__ __forceinline__ int vec_get(const uint4& v, int bit) {
  const u32 w = (bit & 4) ? ((bit & 2) ? v.w : v.z) : ((bit & 2) ? v.y : v.x);
  return (int)(short)(w >> ((bit & 1) * 16));
}
__device__ __forceinline__ void vec_set(uint4& v, int bit, int p) {
  const u32 sh = (bit & 1) * 16;
  const u32 keep = ~(0xffffu << sh), val = ((u32)p & 0xffffu) << sh;
  if (bit & 4) { if (bit & 2) v.w = (v.w & keep) | val; else v.z = (v.z & keep) | val; }
  else         { if (bit & 2) v.y = (v.y & keep) | val; else v.x = (v.x & keep) | val; }
}

// One component's context hash for the NEXT byte from its recipe (zpq_hcomp.cpp); lane-private.
// `seed` = new hash of the lane's seed component (one stage ahead), `wprev` = hash of the previous
// word-model lane before its last update, `hs` = last 8 input bytes with `cur` as byte 0.
__device__ __forceinline__ u32 recipe_eval(const u32* rop, u32 rlen, u32 rstart, u32 own, u32 seed, u32 wprev,
                                           u32 cur, u64 hs, int bpos, const u8* in, u32 dist_last) {
  const u32 vc = (u32)(-(bpos + 1));            // the VM's C register after this byte
  u32 acc = rstart == ZB_ZERO ? 0u : (rstart == ZB_OWN ? own : seed);
  bool skip = false;
#pragma unroll
  for (int t = 0; t < ZPQ_FAST_RCP; ++t) {
    if (t < (int)rlen && !skip) {
      const u32 w = rop[t];
      const u32 op = w & 15u, src = (w >> 4) & 15u, arg = (w >> 8) & 255u, off = w >> 16;
      u32 x = 0;
      if (src == ZS_BYTE) {
        if (off < 8) x = (u32)(hs >> (off * 8)) & arg;
        else x = ((int)off <= bpos) ? ((u32)in[bpos - (int)off] & arg) : 0u;
      }
      else if (src == ZS_POSAND) x = vc & arg;
      else if (src == ZS_POSMOD) x = arg ? vc % arg : 0u;
      else if (src == ZS_DIST) x = min(dist_last - vc, 255u);
      if (op == ZR_FOLD) acc = (acc + x + 512u) * 773u;
      else if (op == ZR_SHL8ADD) acc = (acc << 8) + x;
      else if (op == ZR_SHR) acc >>= arg;
      else if (op == ZR_SHL) acc <<= arg;
      else if (op == ZR_ROLL) acc = acc * off + cur + 1u;
      else if (op == ZR_WORD) {
        const u32 w2 = rop[t + 1 < ZPQ_FAST_RCP ? t + 1 : t];
        const u32 v2 = off & 255u, v3 = off >> 8, mult = w2 & 255u, first = (w2 >> 8) & 1u;
        if ((((cur & arg) - v2) & 255u) < v3) acc = acc * mult + cur + 1u;
        else acc = first ? 0u : wprev;
        skip = true;
      }
    }
    else skip = false;
  }
  return acc;
}

struct Seg {          // segment bookkeeping of the coder
  u32 seg, nseg, first; u64 seg_end, seg_start_out;
};

__device__ __forceinline__ void coder_flush_segments(Enc4& C, Seg& G, const ZKernelArgs& A, int lane, u64 bytes_done) {
  // Compressor::endSegment: EOS after the last byte of a segment (libzpaq.cpp:2950-2952)
  while (G.seg < G.nseg && G.seg_end == bytes_done) {
    enc4_bit(C, lane, 1, 0);
    if (lane == 0) A.seg_out[G.first + G.seg] = C.out_pos - G.seg_start_out;
    G.seg_start_out = C.out_pos; ++G.seg;
    if (G.seg < G.nseg) G.seg_end += A.seg_len[G.first + G.seg];
  }
}


// Everything a role needs to know about its block (uniform across the team's three warps).
struct TeamCtx {
  const ZKernelArgs* A; const ZModel* ZM; u8* arena; uint4* ringv; const u8* in;
  volatile u32* ctrl; const u32* sm32c; u32* sm32;
  int n, N, total_iters, dcod, lane, role;
  u32 pmask, cmask, wbase;
  bool has_coder;
};

#define WAIT_TURN(K)                                                                         \
    do {                                                                                     \
      const u32 kk_ = (u32)(K);                                                              \
      for (int r_ = 0; r_ < ZPQ_ROLES; ++r_) {                                               \
        if ((pmask >> r_) & 1u) while (ctrl[r_] < kk_) { }                                   \
        if ((cmask >> r_) & 1u) while (ctrl[r_] + 16u < kk_) { }                             \
      }                                                                                      \
      __threadfence_block();                                                                 \
    } while (0)
#define END_TURN(K)                                                                          \
    do { __threadfence_block(); __syncwarp(); if (lane == 0) ctrl[role] = (u32)(K) + 1u; } while (0)

#define TEAM_UNPACK                                                                          \
  const ZKernelArgs& A = *T.A; const ZModel* __restrict__ ZM = T.ZM; u8* const arena = T.arena;  \
  uint4* const ringv = T.ringv; const u8* const in = T.in; volatile u32* const ctrl = T.ctrl;    \
  u32* const sm32 = T.sm32; const int n = T.n, N = T.N, total_iters = T.total_iters, dcod = T.dcod; \
  const int lane = T.lane, role = T.role; const u32 pmask = T.pmask, cmask = T.cmask, wbase = T.wbase; \
  const bool has_coder = T.has_coder; (void)arena; (void)sm32; (void)wbase; (void)n;

// roles 0 (ICM / ISSE) and 2 (CM, MATCH, CONST, AVG, MIX2, SSE): lane = component
template <int ROLE>
__device__ __forceinline__ void lane_role(const TeamCtx& T, const Tabs& S, Enc4& C, Seg& G) {
  TEAM_UNPACK
    // =====================================================================================
      // roles 0 and 2: lane = component
      // =====================================================================================
      const int ci = ZM->team.comp_at[ROLE][lane];
      const bool have = ci != 0xff;
      u32 type = ZT_NONE, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, cm_mask = 0, ht_mask = 0, limit = 0;
      u32* gtab = nullptr; u8* ht = nullptr;
      int depth = 0;
      u32 tabw = 0;
      u32 rop[ZPQ_FAST_RCP];
#pragma unroll
      for (int t = 0; t < ZPQ_FAST_RCP; ++t) rop[t] = 0;
      u32 rstart = 0, rlen = 0, dist_byte = 0;
      int seed_lane = 0, wprev_lane = 0;
      if (have) {
        const ZComp cd = ZM->comp[ci];
        type = cd.type; a1 = cd.a1; a2 = cd.a2; a3 = cd.a3; a4 = cd.a4; a5 = cd.a5;
        cm_mask = cd.cm_mask; ht_mask = cd.ht_mask; limit = cd.limit;
        gtab = reinterpret_cast<u32*>(arena + (u64)cd.cm_off16 * 16);
        ht = arena + (u64)cd.ht_off16 * 16;
        depth = ZM->pipe.depth[ci];
        tabw = wbase + ZM->fast.smem_off[ci] / 4u;
#pragma unroll
        for (int t = 0; t < ZPQ_FAST_RCP; ++t) rop[t] = ZM->rcp.ops[t][ci];
        rstart = ZM->rcp.start[ci]; rlen = ZM->rcp.len[ci];
        dist_byte = ZM->pipe.dist_byte[ci];
        // seed / word-shift sources are components of the same role (host-checked): find their lanes
        const int sb = ZM->rcp.base[ci];
        for (int l = 0; l < 32; ++l) {
          const int cl = ZM->team.comp_at[ROLE][l];
          if (cl == sb) seed_lane = l;
          if (cl == ci - 1) wprev_lane = l;
        }
      }
      const bool is_hash = (type == ZT_ICM || type == ZT_ISSE);
      const int lane_last = __ffs(__ballot_sync(FULL, have && ci == n - 1)) - 1;    // lane of the last component (coder)
      // SSE components of this role: owners and tables (uniform)
      int sse_lane0 = -1, sse_lane1 = -1;
      const u32* sse_tab0 = nullptr; const u32* sse_tab1 = nullptr;
      if (ROLE == 2) {
        if (ZM->team.nsse > 0) {
          const int c0 = (int)ZM->team.sse_comp[0];
          sse_lane0 = __ffs(__ballot_sync(FULL, have && ci == c0)) - 1;
          sse_tab0 = reinterpret_cast<const u32*>(arena + (u64)ZM->comp[c0].cm_off16 * 16);
        }
        if (ZM->team.nsse > 1) {
          const int c1 = (int)ZM->team.sse_comp[1];
          sse_lane1 = __ffs(__ballot_sync(FULL, have && ci == c1)) - 1;
          sse_tab1 = reinterpret_cast<const u32*>(arena + (u64)ZM->comp[c1].cm_off16 * 16);
        }
      }

      int p = (type == ZT_CONS) ? ((int)a1 - 128) * 4 : 0;
      u32 h = 0, h_next = 0, h_prev = 0, dist_last = 0;
      u64 hist = 0;
      uint4 row = make_uint4(0, 0, 0, 0); u32 row_off = 0;
      uint4 pf0 = row, pf1 = row, pf2 = row; u32 pf_h0 = 0; bool pf_valid = false;
      u32 bh = 0, v0 = 0, v1 = 0, cxt = 0; int xj = 0, xk = 0;
      u32 ma = 0, mb = 0, mpos = 0, mpred = 0, mc = 0;

      for (int K = 0; K < total_iters; ++K) {
        WAIT_TURN(K);
        const int bpos = K - depth;
        const bool active = have && bpos >= 0 && bpos < N;
        const u32 cur = active ? (u32)in[bpos] : 0u;
        const u32 slot = (u32)bpos & 31u;
        // predictions of the input components for this byte (written >= 1 iteration ago)
        uint4 vin_j = make_uint4(0, 0, 0, 0), vin_k = vin_j;
        if (active) {
          if (type == ZT_ISSE || type == ZT_SSE) vin_j = ringv[a2 * 32u + slot];
          else if (type == ZT_MIX2) { vin_j = ringv[a2 * 32u + slot]; vin_k = ringv[a3 * 32u + slot]; }
          else if (type == ZT_AVG) { vin_j = ringv[a1 * 32u + slot]; vin_k = ringv[a2 * 32u + slot]; }
        }
        uint4 vout = make_uint4(0, 0, 0, 0);
        const int cpos = K - dcod;
        const bool cactive = has_coder && cpos >= 0 && cpos < N;
        const u32 ccur = __shfl_sync(FULL, cur, lane_last < 0 ? 0 : lane_last);
        if (cactive) enc4_bit(C, lane, 0, 0);                  // per-byte EOS flag, p = 0 (libzpaq.cpp:2426)

#pragma unroll 1
        for (int bit = 0; bit < 8; ++bit) {
          const u32 c8 = (0x100u | cur) >> (8 - bit);
          const u32 hmap4 = bit < 4 ? c8 : ((((0x10u | (cur >> 4)) << 4)) | ((0x10u | (cur & 15u)) >> (8 - bit)));
          const int y = (int)((cur >> (7 - bit)) & 1u);
          const int hslot = (int)(hmap4 & 15u);

          if ((bit & 3) == 0) {
            if (bit == 4) {
              // context hashes for the NEXT byte
              const u32 hn_old = h_next, hp_old = h_prev;
              const u32 seed = __shfl_sync(FULL, hn_old, seed_lane);
              const u32 wprev = __shfl_sync(FULL, hp_old, wprev_lane);
              if (active) {
                const u64 hs = (hist << 8) | (u64)cur;
                if (cur == dist_byte) dist_last = (u32)(-(bpos + 1));
                h_next = recipe_eval(rop, rlen, rstart, hn_old, seed, wprev, cur, hs, bpos, in, dist_last);
                h_prev = hn_old; hist = hs;
              }
            }
            if (is_hash && active) {
              // Predictor::find (libzpaq.cpp:2072-2088) on rows requested one nibble ago
              const u32 cx = h + 16u * c8;
              const u32 chk = (cx >> (a1 + 2)) & 255u;
              const u32 h0 = (cx * 16u) & ht_mask;
              uint4 r0, r1, r2;
              if (pf_valid && pf_h0 == h0) { r0 = pf0; r1 = pf1; r2 = pf2; }
              else {
                r0 = *reinterpret_cast<const uint4*>(ht + h0);
                r1 = *reinterpret_cast<const uint4*>(ht + (h0 ^ 16u));
                r2 = *reinterpret_cast<const uint4*>(ht + (h0 ^ 32u));
              }
              if ((r0.x & 255u) == chk) { row = r0; row_off = h0; }
              else if ((r1.x & 255u) == chk) { row = r1; row_off = h0 ^ 16u; }
              else if ((r2.x & 255u) == chk) { row = r2; row_off = h0 ^ 32u; }
              else {
                const u32 q0 = (r0.x >> 8) & 255u, q1 = (r1.x >> 8) & 255u, q2 = (r2.x >> 8) & 255u;
                if (q0 <= q1 && q0 <= q2) row_off = h0;
                else if (q1 < q2) row_off = h0 ^ 16u;
                else row_off = h0 ^ 32u;
                row = make_uint4(chk, 0, 0, 0);
              }
              const u32 ncx = bit == 0 ? (h + 16u * (0x10u | (cur >> 4))) : (h_next + 16u);
              const u32 nh0 = (ncx * 16u) & ht_mask;
              pf_valid = ((nh0 ^ row_off) & ~63u) != 0u && (bit == 0 || bpos + 1 < N);
              if (pf_valid) {
                pf_h0 = nh0;
                pf0 = *reinterpret_cast<const uint4*>(ht + nh0);
                pf1 = *reinterpret_cast<const uint4*>(ht + (nh0 ^ 16u));
                pf2 = *reinterpret_cast<const uint4*>(ht + (nh0 ^ 32u));
              }
            }
          }

          // ---- predict ----
          if constexpr (ROLE == 0) {
            if (active) {
              bh = row_get(row, hslot);
              if (type == ZT_ISSE) {
                const uint2 w = *reinterpret_cast<const uint2*>(sm32 + tabw + bh * 2);
                v0 = w.x; v1 = w.y; xj = vec_get(vin_j, bit);
                p = clamp2k(((int)v0 * xj + (int)v1 * 64) >> 16);
              } else {
                v0 = sm32[tabw + bh];
                p = t_stretch(S, v0 >> 8);
              }
            }
          } else {
            if (active) {
              if (type == ZT_CM) {
                cxt = (h ^ hmap4) & cm_mask;
                v0 = gtab[cxt];
                p = t_stretch(S, v0 >> 17);
              }
              else if (type == ZT_MATCH) {
                if (ma == 0) p = 0;
                else {
                  mc = (mpred >> (7 - bit)) & 1u;
                  p = t_stretch(S, (u32)((int)S.dt2k[ma] * (1 - 2 * (int)mc)) & 32767u);
                }
              }
              else if (type == ZT_MIX2) {
                cxt = (h + (c8 & a5)) & cm_mask;
                const int w = (int)reinterpret_cast<const u16*>(gtab)[cxt];
                v0 = (u32)w; xj = vec_get(vin_j, bit); xk = vec_get(vin_k, bit);
                p = (w * xj + (65536 - w) * xk) >> 16;
              }
              else if (type == ZT_AVG) p = (vec_get(vin_j, bit) * (int)a3 + vec_get(vin_k, bit) * (256 - (int)a3)) >> 8;
              else if (type == ZT_SSE) { cxt = ((h + c8) * 32u) & cm_mask; xj = vec_get(vin_j, bit); }
            }
            // SSE: the warp fetches the owner's 32-entry row, the two buckets come back by shuffle
#pragma unroll
            for (int r = 0; r < ZPQ_FAST_SSE; ++r) {
              const int ol = r == 0 ? sse_lane0 : sse_lane1;
              if (ol >= 0) {
                const bool oact = __shfl_sync(FULL, (int)active, ol) != 0;
                const u32 rb = __shfl_sync(FULL, cxt, ol);
                const int pj = __shfl_sync(FULL, xj, ol);
                const u32* tab = r == 0 ? sse_tab0 : sse_tab1;
                u32 rowv = 0;
                if (oact) rowv = tab[rb + lane];
                int pq = min(max(pj + 992, 0), 1983);
                const int wt = pq & 63;
                pq >>= 6;
                const u32 q0 = __shfl_sync(FULL, rowv, pq), q1 = __shfl_sync(FULL, rowv, pq + 1);
                if (lane == ol && active) {
                  p = t_stretch(S, (u32)(((int)(q0 >> 10) * (64 - wt) + (int)(q1 >> 10) * wt) >> 13));
                  if (wt >> 5) { cxt += (u32)pq + 1u; v0 = q1; } else { cxt += (u32)pq; v0 = q0; }
                }
              }
            }
          }

          // ---- coder ----
          if (has_coder) {
            const int pfin = __shfl_sync(FULL, p, lane_last < 0 ? 0 : lane_last);
            if (cactive) enc4_bit(C, lane, (int)((ccur >> (7 - bit)) & 1u), (u32)t_squash(S, pfin) * 2u + 1u);
          }

          // ---- update (libzpaq.cpp:1954-2052) ----
          if (active) {
            if (is_hash) {
              const u32 nx = ((u32)S.ns[bh] >> (8 * y)) & 255u;
              row_set(row, hslot, nx);
              if (type == ZT_ISSE) {
                const int err = y * 32767 - t_squash(S, p);
                const int w0 = clamp512k((int)v0 + ((err * xj + (1 << 12)) >> 13));
                const int w1 = clamp512k((int)v1 + ((err + 16) >> 5));
                *reinterpret_cast<uint2*>(sm32 + tabw + bh * 2) = make_uint2((u32)w0, (u32)w1);
              } else {
                sm32[tabw + bh] = v0 + (u32)(((int)(y * 32767 - (int)(v0 >> 8))) >> 2);
              }
              if ((bit & 3) == 3) *reinterpret_cast<uint4*>(ht + row_off) = row;
            }
            else if (type == ZT_CM || type == ZT_SSE) {
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
              if (bit == 7) {
                ht[mpos] = (u8)cur;
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
              }
            }
            vec_set(vout, bit, p);
          }
        }  // bit

        if (active) { ringv[(u32)ci * 32u + slot] = vout; h = h_next; }
        if (cactive) coder_flush_segments(C, G, A, lane, (u64)(cpos + 1));
        END_TURN(K);
      }

}

// role 1: MIX components; lane l = input l of the mixer being evaluated
__device__ __forceinline__ void mix_role(const TeamCtx& T, const Tabs& S, Enc4& C, Seg& G) {
  TEAM_UNPACK
      // =====================================================================================
      // role 1: MIX components; lane l = input l of the mixer being evaluated
      // =====================================================================================
      const int nmix = (int)ZM->team.nmix;
      // per-mixer uniform state (up to ZPQ_FAST_MIX mixers, processed one after the other each bit)
      u32 mh[ZPQ_FAST_MIX], mhn[ZPQ_FAST_MIX]; u64 mhist[ZPQ_FAST_MIX];
      uint4 mvin[ZPQ_FAST_MIX], mvout[ZPQ_FAST_MIX];
      int mw[ZPQ_FAST_MIX]; u32 mwbase[ZPQ_FAST_MIX]; bool mwvalid[ZPQ_FAST_MIX];
#pragma unroll
      for (int q = 0; q < ZPQ_FAST_MIX; ++q) {
        mh[q] = mhn[q] = 0; mhist[q] = 0; mvin[q] = mvout[q] = make_uint4(0, 0, 0, 0); mw[q] = 0; mwbase[q] = 0; mwvalid[q] = false;
      }
      int pcod = 0;
      int mcomp[ZPQ_FAST_MIX], mm[ZPQ_FAST_MIX], mj0[ZPQ_FAST_MIX], mrate[ZPQ_FAST_MIX], mdepth[ZPQ_FAST_MIX];
      u32 mmask5[ZPQ_FAST_MIX], mcmmask[ZPQ_FAST_MIX], mrlen[ZPQ_FAST_MIX]; int* mtab[ZPQ_FAST_MIX];
#pragma unroll
      for (int q = 0; q < ZPQ_FAST_MIX; ++q) {
        mcomp[q] = 0; mm[q] = 0; mj0[q] = 0; mrate[q] = 0; mdepth[q] = 0; mmask5[q] = 0; mcmmask[q] = 0; mrlen[q] = 0; mtab[q] = nullptr;
        if (q < nmix) {
          const int cq = (int)ZM->team.mix_comp[q];
          const ZComp cd = ZM->comp[cq];
          mcomp[q] = cq; mm[q] = cd.a3; mj0[q] = cd.a2; mrate[q] = cd.a4; mmask5[q] = cd.a5; mcmmask[q] = cd.cm_mask;
          mdepth[q] = ZM->pipe.depth[cq]; mrlen[q] = ZM->rcp.len[cq];
          mtab[q] = reinterpret_cast<int*>(arena + (u64)cd.cm_off16 * 16);
        }
      }

      for (int K = 0; K < total_iters; ++K) {
        WAIT_TURN(K);
        const int cpos = K - dcod;
        const bool cactive = has_coder && cpos >= 0 && cpos < N;
        const u32 ccur = cactive ? (u32)in[cpos] : 0u;
        if (cactive) enc4_bit(C, lane, 0, 0);
        u32 mcur[ZPQ_FAST_MIX]; bool mact[ZPQ_FAST_MIX]; int mbpos[ZPQ_FAST_MIX];
#pragma unroll
        for (int q = 0; q < ZPQ_FAST_MIX; ++q) {
          mcur[q] = 0; mact[q] = false; mbpos[q] = 0;
          if (q < nmix) {
            mbpos[q] = K - mdepth[q];
            mact[q] = mbpos[q] >= 0 && mbpos[q] < N;
            if (mact[q]) {
              mcur[q] = (u32)in[mbpos[q]];
              if (lane < mm[q]) mvin[q] = ringv[((u32)mj0[q] + (u32)lane) * 32u + ((u32)mbpos[q] & 31u)];
            }
          }
        }

#pragma unroll 1
        for (int bit = 0; bit < 8; ++bit) {
#pragma unroll
          for (int q = 0; q < ZPQ_FAST_MIX; ++q) {
            if (q < nmix && mact[q]) {
              const int cq = mcomp[q];
              int* const tab = mtab[q];
              const u32 cur = mcur[q];
              const int m = mm[q];
              const u32 a5 = mmask5[q], cmm = mcmmask[q];
              if (bit == 4) {
                // the mixer's own context for the next byte (a pure fold, evaluated uniformly)
                u32 rop[ZPQ_FAST_RCP];
#pragma unroll
                for (int t = 0; t < ZPQ_FAST_RCP; ++t) rop[t] = ZM->rcp.ops[t][cq];
                const u64 hs = (mhist[q] << 8) | (u64)cur;
                mhn[q] = recipe_eval(rop, mrlen[q], ZB_ZERO, 0u, 0u, 0u, cur, hs, mbpos[q], in, 0u);
                mhist[q] = hs;
              }
              const u32 c8 = (0x100u | cur) >> (8 - bit);
              const int y = (int)((cur >> (7 - bit)) & 1u);
              const u32 base = ((mh[q] + (c8 & a5)) & cmm) * (u32)m;
              int w;
              if (mwvalid[q] && mwbase[q] == base) w = mw[q];                     // fetched (or trained) during the previous bit
              else w = lane < m ? tab[base + lane] : 0;
              // request the next bit's row now: its index only depends on input data
              u32 nbase; bool nvalid;
              if (bit < 7) {
                const u32 nc8 = (0x100u | cur) >> (7 - bit);
                nbase = ((mh[q] + (nc8 & a5)) & cmm) * (u32)m; nvalid = true;
              } else { nbase = ((mhn[q] + (1u & a5)) & cmm) * (u32)m; nvalid = mbpos[q] + 1 < N; }
              int wn = 0;
              if (nvalid && nbase != base && lane < m) wn = tab[nbase + lane];
              const int x = vec_get(mvin[q], bit);
              const int sum = __reduce_add_sync(FULL, lane < m ? (w >> 8) * x : 0);
              const int p = clamp2k(sum >> 8);
              const int err = ((y * 32767 - t_squash(S, p)) * mrate[q]) >> 4;
              const int wu = clamp512k(w + ((err * x + (1 << 12)) >> 13));
              if (lane < m) tab[base + lane] = wu;
              mw[q] = (nbase == base) ? wu : wn; mwbase[q] = nbase; mwvalid[q] = nvalid;
              vec_set(mvout[q], bit, p);
              if (cq == n - 1) pcod = p;
            }
          }
          if (cactive) enc4_bit(C, lane, (int)((ccur >> (7 - bit)) & 1u), (u32)t_squash(S, pcod) * 2u + 1u);
        }
#pragma unroll
        for (int q = 0; q < ZPQ_FAST_MIX; ++q) {
          if (q < nmix && mact[q]) {
            if (lane == 0) ringv[(u32)mcomp[q] * 32u + ((u32)mbpos[q] & 31u)] = mvout[q];
            mh[q] = mhn[q];
          }
        }
        if (cactive) coder_flush_segments(C, G, A, lane, (u64)(cpos + 1));
        END_TURN(K);
      }
    
}
__global__ void __launch_bounds__(672, 1) zpq_enc_kernel4(const ZKernelArgs A, const int pw_bytes, const int nteams) {
  extern __shared__ __align__(16) unsigned char smem[];
  stage_tables(smem, A.tables);
  __syncthreads();
  const Tabs S = make_tabs(smem);
  u32* sm32 = reinterpret_cast<u32*>(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp / ZPQ_ROLES, role = warp - team * ZPQ_ROLES;
  const u32 wbase = (u32)(S2_STATIC + team * pw_bytes) / 4u;
  volatile u32* ctrl = reinterpret_cast<volatile u32*>(smem + S2_STATIC + (size_t)nteams * pw_bytes) + team * 4;
  // ctrl[0..2] = byte-iterations completed by role 0..2, ctrl[3] = block index of the team

  for (;;) {
    if (role == 0 && lane == 0) { ctrl[0] = 0; ctrl[1] = 0; ctrl[2] = 0; ctrl[3] = atomicAdd(A.counter, 1u); }
    TEAM_BAR(team);
    const int blk = (int)ctrl[3];
    if (blk >= A.nblk) break;

    const ZBlock B = A.blocks[blk];
    const ZModel* __restrict__ ZM = A.models + B.model;
    u8* const arena = A.arena + B.arena_off;
    const int n = (int)ZM->n;
    const int dmax = (int)ZM->pipe.dmax;
    uint4* const ringv = reinterpret_cast<uint4*>(arena + (u64)ZM->team.ring_off16 * 16);   // [comp][32 slots]
    u64 N = 0;
    for (u32 sg = 0; sg < B.nseg; ++sg) N += A.seg_len[B.seg_first + sg];
    const u8* const in = A.in + B.in_off;
    const u32 pmask = ZM->team.pmask[role], cmask = ZM->team.cmask[role];
    const bool has_coder = (int)ZM->team.coder_role == role;
    const int dcod = (int)ZM->pipe.depth[n - 1];
    const int total_iters = (int)N + dmax;

    Enc4 C;
    C.low = 1; C.high = 0xffffffffu; C.out = A.out + B.out_off; C.out_pos = 0; C.out_cap = B.out_cap; C.status = ST_OK;
    Seg G;
    G.seg = 0; G.nseg = B.nseg; G.first = B.seg_first; G.seg_end = B.nseg ? A.seg_len[B.seg_first] : 0; G.seg_start_out = 0;
    if (has_coder) coder_flush_segments(C, G, A, lane, 0);       // empty leading segments

    // role 0 owns the shared-memory ICM / ISSE tables of the block: initialise them
    if (role == 0) {
      for (int l = 0; l < 32; ++l) {
        const int ci = ZM->team.comp_at[0][l];
        if (ci == 0xff) break;
        const u32 tw = wbase + ZM->fast.smem_off[ci] / 4u;
        if (ZM->comp[ci].type == ZT_ICM) for (int i = lane; i < 256; i += 32) sm32[tw + i] = A.tables->icm_init[i];
        else for (int i = lane; i < 512; i += 32) sm32[tw + i] = A.tables->isse_init[i];
      }
      __syncwarp();
    }


    TeamCtx T;
    T.A = &A; T.ZM = ZM; T.arena = arena; T.ringv = ringv; T.in = in; T.ctrl = ctrl; T.sm32c = sm32; T.sm32 = sm32;
    T.n = n; T.N = (int)N; T.total_iters = total_iters; T.dcod = dcod; T.lane = lane; T.role = role;
    T.pmask = pmask; T.cmask = cmask; T.wbase = wbase; T.has_coder = has_coder;
    if (role == 0) lane_role<0>(T, S, C, G);
    else if (role == 2) lane_role<2>(T, S, C, G);
    else mix_role(T, S, C, G);
    if (has_coder && lane == 0) { A.status[blk] = C.status; A.out_len[blk] = C.out_pos; }
    TEAM_BAR(team);
  }
}

// ---------------------------------------------------------------------------
extern "C" cudaError_t zpq_launch_enc4(const ZKernelArgs* a, int nblk, int sm_count, int teams, int pw_bytes, cudaStream_t st) {
  if (nblk <= 0) return cudaSuccess;
  const size_t smem = (size_t)S2_STATIC + (size_t)teams * pw_bytes + (size_t)teams * 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(zpq_enc_kernel4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZPQ_SMEM_MAX);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int ctas = (nblk + teams - 1) / teams;
  if (ctas > sm_count) ctas = sm_count;
  zpq_enc_kernel4<<<ctas, teams * ZPQ_ROLES * 32, smem, st>>>(*a, pw_bytes, teams);
  return cudaGetLastError();
}
