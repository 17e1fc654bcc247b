// zpq_kernels3.cu -- skewed-pipeline ENCODE kernel for sm_100a.
//
// Why a separate encode kernel.  Predictor::predict()/update() is one long
// dependency chain per coded bit (ICM -> 7 ISSE levels -> MIX -> MIX -> MIX2 ->
// SSE -> MIX2 -> coder at level 5).  When DEcoding that chain is real: bit t+1
// cannot start before bit t is known.  When ENcoding it is not: the coded bit is
// input data, and no component reads a later component's output or the coder's
// state (libzpaq.cpp:1954-2052: every update uses only y, the component's own
// state and its inputs p[j], j < i).  The COMP chain is therefore a feed-forward
// pipeline, and this kernel runs it as one:
//
//   * lane i (= component i) works `depth[i]` BYTES behind the head of the
//     stream (depth = longest input path, in components).  In byte-iteration K
//     lane i codes byte K - depth[i]; all lanes are at the same bit position, so
//     nibble and byte boundary code stays warp-uniform.
//   * a component's inputs p[j] for that byte were produced depth[i]-depth[j]
//     byte-iterations earlier; each lane keeps its own last 16 bytes x 8 bits of
//     outputs in a private delay ring and serves them by shuffle.  There is no
//     same-bit cross-lane dependency left: an ISSE chain of any length costs one
//     shuffle per bit.
//   * context hashes are the lane-parallel recipes of zpq_hcomp.cpp, evaluated
//     at mid-byte for the NEXT byte so that the ICM/ISSE hash rows of the next
//     nibble are already in flight (3 x 16 B per lane) while the current nibble
//     is coded: HBM latency is off the critical path.
//   * the arithmetic coder trails the last component and is executed uniformly.
//
// Results are bit-identical to the sequential order because every table is
// private to one lane (or, for MIX rows, touched only in the owner's bit order)
// and every value read across lanes is the exact p[j] of the same bit.
// Decoding keeps the sequential kernel (zpq_kernels2.cu).
#include <cuda_runtime.h>
#include <stdint.h>
#include "zpq_dev.h"
#include "zpq_kernels.h"
#include "zpq_kcommon.cuh"

struct Enc3 {
  u32 low, high;
  u8* out; u64 out_pos, out_cap;
  int status;
};

__device__ __forceinline__ void enc3_bit(Enc3& C, int lane, int y, u32 p16) {
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

#define RING_IDX(kb, bit) ((((kb) & 15) << 3) | (bit))

__global__ void __launch_bounds__(256, 1) zpq_enc_kernel3(const ZKernelArgs A, const int pw_bytes) {
  extern __shared__ __align__(16) unsigned char smem[];
  stage_tables(smem, A.tables);
  __syncthreads();
  const Tabs S = make_tabs(smem);
  u32* sm32 = reinterpret_cast<u32*>(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const u32 wbase = (u32)(S2_STATIC + warp * pw_bytes) / 4u;
  short ring[128];                       // this lane's outputs: [byte-iteration & 15][bit]

  for (;;) {
    int blk = 0;
    if (lane == 0) blk = (int)atomicAdd(A.counter, 1u);
    blk = __shfl_sync(FULL, blk, 0);
    if (blk >= A.nblk) break;

    const ZBlock B = A.blocks[blk];
    const ZModel* __restrict__ ZM = A.models + B.model;
    u8* const arena = A.arena + B.arena_off;
    const int n = (int)ZM->n;
    const int nsp = (int)ZM->pipe.nsp, dmax = (int)ZM->pipe.dmax;
    u32 sd[ZPQ_PIPE_STEPS];
#pragma unroll
    for (int s = 0; s < ZPQ_PIPE_STEPS; ++s) sd[s] = ZM->pipe.sd[s];
    // MIX / SSE register slots
    int* mix_tab[ZPQ_FAST_MIX]; const u32* sse_tab[ZPQ_FAST_SSE];
    {
      int nm = 0, ns = 0;
#pragma unroll
      for (int q = 0; q < ZPQ_FAST_MIX; ++q) mix_tab[q] = nullptr;
#pragma unroll
      for (int r = 0; r < ZPQ_FAST_SSE; ++r) sse_tab[r] = nullptr;
#pragma unroll
      for (int s = 0; s < ZPQ_PIPE_STEPS; ++s) {
        if (s < nsp) {
          const u32 kind = sd[s] & 7u; const int o = (int)((sd[s] >> 3) & 31u);
          if (kind == ZK_MIX) {
            int* t = reinterpret_cast<int*>(arena + (u64)ZM->comp[o].cm_off16 * 16);
            if (nm == 0) mix_tab[0] = t; else if (nm == 1) mix_tab[1] = t; else mix_tab[2] = t;
            ++nm;
          } else if (kind == ZK_SSE) {
            const u32* t = reinterpret_cast<const u32*>(arena + (u64)ZM->comp[o].cm_off16 * 16);
            if (ns == 0) sse_tab[0] = t; else sse_tab[1] = t;
            ++ns;
          }
        }
      }
    }

    // ---- lane <-> component ----
    u32 type = ZT_NONE, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, cm_mask = 0, ht_mask = 0, limit = 0;
    u32* gtab = nullptr; u8* ht = nullptr;
    int depth = 0; u32 delay = 0;
    if (lane < n) {
      const ZComp cd = ZM->comp[lane];
      type = cd.type; a1 = cd.a1; a2 = cd.a2; a3 = cd.a3; a4 = cd.a4; a5 = cd.a5;
      cm_mask = cd.cm_mask; ht_mask = cd.ht_mask; limit = cd.limit;
      gtab = reinterpret_cast<u32*>(arena + (u64)cd.cm_off16 * 16);
      ht = arena + (u64)cd.ht_off16 * 16;
      depth = ZM->pipe.depth[lane]; delay = ZM->pipe.delay[lane];
    }
    const u32 tabw = wbase + ((lane < n) ? ZM->fast.smem_off[lane] : 0u) / 4u;
    __syncwarp();
    for (int src = 0; src < n; ++src) {
      const u32 ty = __shfl_sync(FULL, type, src);
      const u32 tw = __shfl_sync(FULL, tabw, src);
      if (ty == ZT_ICM) for (int i = lane; i < 256; i += 32) sm32[tw + i] = A.tables->icm_init[i];
      else if (ty == ZT_ISSE) for (int i = lane; i < 512; i += 32) sm32[tw + i] = A.tables->isse_init[i];
    }
    __syncwarp();
    const bool is_hash = (type == ZT_ICM || type == ZT_ISSE);
    int p = (type == ZT_CONS) ? ((int)a1 - 128) * 4 : 0;
    for (int i = 0; i < 128; ++i) ring[i] = (short)p;

    // recipes (lane-private evaluation, one stage = one byte behind its seed lane)
    u32 rop[ZPQ_FAST_RCP];
#pragma unroll
    for (int t = 0; t < ZPQ_FAST_RCP; ++t) rop[t] = (lane < n) ? ZM->rcp.ops[t][lane] : 0u;
    u32 rstart = 0, rbase = 0, rlen = 0;
    if (lane < n) { rstart = ZM->rcp.start[lane]; rbase = ZM->rcp.base[lane]; rlen = ZM->rcp.len[lane]; }
    const u32 dist_byte = (lane < n) ? ZM->pipe.dist_byte[lane] : 0u;
    u32 dist_last = 0;

    // stream
    u64 N = 0;
    for (u32 sg = 0; sg < B.nseg; ++sg) N += A.seg_len[B.seg_first + sg];
    const u8* const in = A.in + B.in_off;

    // per-lane state
    u32 h = 0, h_next = 0, h_prev = 0;
    u64 hist = 0;
    uint4 row = make_uint4(0, 0, 0, 0); u32 row_off = 0;
    uint4 pf0 = row, pf1 = row, pf2 = row; u32 pf_h0 = 0; bool pf_valid = false;
    u32 bh = 0, v0 = 0, v1 = 0, cxt = 0; int xj = 0, xk = 0;
    u32 ma = 0, mb = 0, mpos = 0, mpred = 0, mc = 0;
    int mixw0 = 0, mixw1 = 0, mixw2 = 0, mixx0 = 0, mixx1 = 0, mixx2 = 0;
    u32 mixb0 = 0, mixb1 = 0, mixb2 = 0;

    // coder (uniform), trails the last component
    Enc3 C;
    C.low = 1; C.high = 0xffffffffu; C.out = A.out + B.out_off; C.out_pos = 0; C.out_cap = B.out_cap; C.status = ST_OK;
    const int dcod = __shfl_sync(FULL, depth, n - 1);
    u32 seg = 0; u64 seg_end = B.nseg ? A.seg_len[B.seg_first] : 0; u64 seg_start_out = 0;
    // segments that end before any byte (empty segments)
    while (seg < B.nseg && seg_end == 0) {
      enc3_bit(C, lane, 1, 0);
      if (lane == 0) A.seg_out[B.seg_first + seg] = C.out_pos - seg_start_out;
      seg_start_out = C.out_pos; ++seg;
      if (seg < B.nseg) seg_end += A.seg_len[B.seg_first + seg];
    }

    const long long total_iters = (long long)N + dmax;
    for (long long K = 0; K < total_iters; ++K) {
      const long long bpos = K - depth;                      // this lane's byte index
      const bool active = (lane < n) && bpos >= 0 && bpos < (long long)N;
      const u32 cur = active ? (u32)in[bpos] : 0u;
      const long long cpos = K - dcod;
      const bool cactive = cpos >= 0 && cpos < (long long)N;
      const u32 ccur = __shfl_sync(FULL, cur, n - 1);
      if (cactive) enc3_bit(C, lane, 0, 0);                  // per-byte EOS flag, p = 0 (libzpaq.cpp:2426)

#pragma unroll 1
      for (int bit = 0; bit < 8; ++bit) {
        // ---- bit context (closed form of the c8 / hmap4 bookkeeping, libzpaq.cpp:2054-2066) ----
        const u32 c8 = (0x100u | cur) >> (8 - bit);
        const u32 hmap4 = bit < 4 ? c8 : ((((0x10u | (cur >> 4)) << 4)) | ((0x10u | (cur & 15u)) >> (8 - bit)));
        const int y = (int)((cur >> (7 - bit)) & 1u);
        const int slot = (int)(hmap4 & 15u);

        if ((bit & 3) == 0) {
          // ================= nibble start =================
          if (bit == 4) {
            // contexts for the NEXT byte (this byte is complete input data)
            const u32 hn_old = h_next, hp_old = h_prev;
            const u32 seed = __shfl_sync(FULL, hn_old, (int)rbase);
            const u32 wprev = __shfl_up_sync(FULL, hp_old, 1);
            if (active) {
              const u32 vc = (u32)(-(bpos + 1));            // the VM's C register after this byte
              const u64 hs = (hist << 8) | (u64)cur;
              if (cur == dist_byte) dist_last = vc;
              u32 acc = rstart == ZB_ZERO ? 0u : (rstart == ZB_OWN ? hn_old : seed);
              bool skip = false;
#pragma unroll
              for (int t = 0; t < ZPQ_FAST_RCP; ++t) {
                if (t < (int)rlen && !skip) {
                  const u32 w = rop[t];
                  const u32 op = w & 15u, src = (w >> 4) & 15u, arg = (w >> 8) & 255u, off = w >> 16;
                  u32 x = 0;
                  if (src == ZS_BYTE) {
                    if (off < 8) x = (u32)(hs >> (off * 8)) & arg;
                    else x = ((long long)off <= bpos) ? ((u32)in[bpos - off] & arg) : 0u;
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
              h_prev = hn_old; h_next = acc; hist = hs;
            }
          }
          if (is_hash && active) {
            // Predictor::find (libzpaq.cpp:2072-2088) on rows that were requested one nibble ago
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
            // request the rows of the next nibble now; skip it if they share the 64-byte group of
            // the row being modified (then the next find reads them after the write-back)
            const u32 ncx = bit == 0 ? (h + 16u * (0x10u | (cur >> 4))) : (h_next + 16u);
            const u32 nh0 = (ncx * 16u) & ht_mask;
            pf_valid = ((nh0 ^ row_off) & ~63u) != 0u && (bit == 0 || bpos + 1 < (long long)N);
            if (pf_valid) {
              pf_h0 = nh0;
              pf0 = *reinterpret_cast<const uint4*>(ht + nh0);
              pf1 = *reinterpret_cast<const uint4*>(ht + (nh0 ^ 16u));
              pf2 = *reinterpret_cast<const uint4*>(ht + (nh0 ^ 32u));
            }
          }
        }

        // ---- delayed outputs of earlier stages (private rings, served by shuffle) ----
        const int r_isse = (int)ring[RING_IDX(K - 1, bit)];
        int rv[ZPQ_PIPE_STEPS];
#pragma unroll
        for (int s = 0; s < ZPQ_PIPE_STEPS; ++s)
          rv[s] = (s < nsp) ? (int)ring[RING_IDX(K - (long long)((delay >> (4 * s)) & 15u), bit)] : 0;
        const int pj_isse = __shfl_sync(FULL, r_isse, (int)a2 & 31);

        // ---- own component: predict ----
        if (active) {
          if (is_hash) {
            bh = row_get(row, slot);
            if (type == ZT_ISSE) {
              const uint2 w = *reinterpret_cast<const uint2*>(sm32 + tabw + bh * 2);
              v0 = w.x; v1 = w.y; xj = pj_isse;
              p = clamp2k(((int)v0 * pj_isse + (int)v1 * 64) >> 16);
            } else {
              v0 = sm32[tabw + bh];
              p = t_stretch(S, v0 >> 8);
            }
          }
          else if (type == ZT_CM) {
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
            v0 = reinterpret_cast<const u16*>(gtab)[cxt];
          }
          else if (type == ZT_MIX) cxt = ((h + (c8 & a5)) & cm_mask) * a3;
          else if (type == ZT_SSE) cxt = ((h + c8) * 32u) & cm_mask;
        }

        // ---- special steps (uniform): MIX / MIX2 / SSE / AVG, each on its owner's bit ----
#pragma unroll
        for (int s = 0; s < ZPQ_PIPE_STEPS; ++s) {
          if (s < nsp) {
            const u32 d = sd[s];
            const u32 kind = d & 7u;
            const int owner = (int)((d >> 3) & 31u), sj = (int)((d >> 8) & 31u), sk = (int)((d >> 13) & 63u);
            const bool oact = __shfl_sync(FULL, (int)active, owner) != 0;
            if (kind == ZK_MIX) {
              const u32 q = d >> 27;
              const u32 base = __shfl_sync(FULL, cxt, owner);
              const int x = __shfl_sync(FULL, rv[s], (sj + lane) & 31);
              int* tab = q == 0 ? mix_tab[0] : (q == 1 ? mix_tab[1] : mix_tab[2]);
              int w = 0;
              if (oact && lane < sk) w = tab[base + lane];
              const int sum = __reduce_add_sync(FULL, (w >> 8) * x);
              if (lane == owner && active) p = clamp2k(sum >> 8);
              if (q == 0) { mixw0 = w; mixx0 = x; mixb0 = base; }
              else if (q == 1) { mixw1 = w; mixx1 = x; mixb1 = base; }
              else { mixw2 = w; mixx2 = x; mixb2 = base; }
            }
            else if (kind == ZK_MIX2) {
              const int pj = __shfl_sync(FULL, rv[s], sj);
              const int pk = __shfl_sync(FULL, rv[s], sk & 31);
              if (lane == owner && active) {
                xj = pj; xk = pk;
                const int w = (int)v0;
                p = (w * pj + (65536 - w) * pk) >> 16;
              }
            }
            else if (kind == ZK_SSE) {
              const u32 r = d >> 27;
              const int pj = __shfl_sync(FULL, rv[s], sj);
              const u32 rb = __shfl_sync(FULL, cxt, owner);
              const u32* tab = r == 0 ? sse_tab[0] : sse_tab[1];
              u32 rowv = 0;
              if (oact) rowv = tab[rb + lane];
              int pq = min(max(pj + 992, 0), 1983);
              const int wt = pq & 63;
              pq >>= 6;
              const u32 q0 = __shfl_sync(FULL, rowv, pq), q1 = __shfl_sync(FULL, rowv, pq + 1);
              if (lane == owner && active) {
                p = t_stretch(S, (u32)(((int)(q0 >> 10) * (64 - wt) + (int)(q1 >> 10) * wt) >> 13));
                if (wt >> 5) { cxt += (u32)pq + 1u; v0 = q1; } else { cxt += (u32)pq; v0 = q0; }
              }
            }
            else {   // ZK_AVG
              const int pj = __shfl_sync(FULL, rv[s], sj);
              const int pk = __shfl_sync(FULL, rv[s], sk & 31);
              const int wt = (int)((d >> 19) & 255u);
              if (lane == owner && active) p = (pj * wt + pk * (256 - wt)) >> 8;
            }
          }
        }

        // ---- coder: the last component's bit ----
        {
          const int pfin = __shfl_sync(FULL, p, n - 1);
          if (cactive) enc3_bit(C, lane, (int)((ccur >> (7 - bit)) & 1u), (u32)t_squash(S, pfin) * 2u + 1u);
        }

        // ---- own component: update (libzpaq.cpp:1954-2052) ----
        if (active) {
          if (is_hash) {
            const u32 nx = ((u32)S.ns[bh] >> (8 * y)) & 255u;
            row_set(row, slot, nx);
            if (type == ZT_ISSE) {
              const int err = y * 32767 - t_squash(S, p);
              const int w0 = clamp512k((int)v0 + ((err * xj + (1 << 12)) >> 13));
              const int w1 = clamp512k((int)v1 + ((err + 16) >> 5));
              *reinterpret_cast<uint2*>(sm32 + tabw + bh * 2) = make_uint2((u32)w0, (u32)w1);
            } else {
              sm32[tabw + bh] = v0 + (u32)(((int)(y * 32767 - (int)(v0 >> 8))) >> 2);
            }
            if ((bit & 3) == 3) *reinterpret_cast<uint4*>(ht + row_off) = row;      // nibble done: row goes home
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
              u32* idx = gtab + (h & cm_mask);              // context of the byte just coded (gotcha 6)
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
          ring[RING_IDX(K, bit)] = (short)p;
        }
        // MIX rows: lane l trains weight l with the owner's error
#pragma unroll
        for (int s = 0; s < ZPQ_PIPE_STEPS; ++s) {
          if (s < nsp && (sd[s] & 7u) == ZK_MIX) {
            const u32 d = sd[s];
            const int owner = (int)((d >> 3) & 31u), sk = (int)((d >> 13) & 63u);
            const int rate = (int)((d >> 19) & 255u);
            const u32 q = d >> 27;
            const int e0 = active ? (((y * 32767 - t_squash(S, p)) * rate) >> 4) : 0;
            const int err = __shfl_sync(FULL, e0, owner);
            const bool oact = __shfl_sync(FULL, (int)active, owner) != 0;
            int* tab = q == 0 ? mix_tab[0] : (q == 1 ? mix_tab[1] : mix_tab[2]);
            const int w = q == 0 ? mixw0 : (q == 1 ? mixw1 : mixw2);
            const int x = q == 0 ? mixx0 : (q == 1 ? mixx1 : mixx2);
            const u32 base = q == 0 ? mixb0 : (q == 1 ? mixb1 : mixb2);
            if (oact && lane < sk) tab[base + lane] = clamp512k(w + ((err * x + (1 << 12)) >> 13));
          }
        }
      }  // bit

      // ---- byte boundary ----
      if (active) h = h_next;
      if (cactive) {
        // Compressor::endSegment: EOS after the last byte of a segment (libzpaq.cpp:2950-2952)
        while (seg < B.nseg && seg_end == (u64)(cpos + 1)) {
          enc3_bit(C, lane, 1, 0);
          if (lane == 0) A.seg_out[B.seg_first + seg] = C.out_pos - seg_start_out;
          seg_start_out = C.out_pos; ++seg;
          if (seg < B.nseg) seg_end += A.seg_len[B.seg_first + seg];
        }
      }
    }  // K
    if (lane == 0) { A.status[blk] = C.status; A.out_len[blk] = C.out_pos; }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
extern "C" cudaError_t zpq_launch_enc3(const ZKernelArgs* a, int nblk, int sm_count, int warps, int pw_bytes, cudaStream_t st) {
  if (nblk <= 0) return cudaSuccess;
  const size_t smem = (size_t)S2_STATIC + (size_t)warps * pw_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(zpq_enc_kernel3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZPQ_SMEM_MAX);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int ctas = (nblk + warps - 1) / warps;
  if (ctas > sm_count) ctas = sm_count;
  zpq_enc_kernel3<<<ctas, warps * 32, smem, st>>>(*a, pw_bytes);
  return cudaGetLastError();
}
