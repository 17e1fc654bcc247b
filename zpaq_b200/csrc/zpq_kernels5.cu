// zpq_kernels5.cu -- wavefront ENCODE for sm_100a: one thread per (block, component).
//
// Encoding is a feed-forward pipeline over components: while compressing, the coded bit is input data, so no component reads a
// later component's output or the coder's state.  Here the pipeline is unrolled in TIME over kernel
// launches instead of over lanes of one warp:
//
//   * the input of every block is cut into chunks of `chunk_bytes`; in wave w a component at pipeline
//     depth d codes chunk w-d, the coder (depth of the last component + 1) codes chunk w-dc;
//   * a thread owns ONE component of ONE block for one chunk: its tables are private, its state (context
//     hashes, MATCH pointers, coder range) is kept in a 64-byte record between waves;
//   * the lanes of a warp are the same component of `gsize` different blocks with the same model, so the
//     per-type code path is warp-uniform (no SIMT divergence); the loop is latency bound per warp, so
//     the group size trades lanes for warps: 8 blocks per warp puts ~3 000 component warps on the GPU
//     and lets every SM switch among 16 resident warps instead of 5;
//   * predictions travel through HBM streams (one 16-byte vector of eight int16 per component per
//     byte, laid out [group][component][chunk slot][byte][lane] so that a warp reads and writes 512
//     contiguous bytes); a stream written in wave w is read in waves w+1 .. w+15;
//   * MIX components with a full c8 mask are component threads too (eight independent weight rows per byte, see
//     the ZT_MIX branch); other mixers get a whole warp per block (lane = input, coalesced weight row, redux.sync);
//   * kernel boundaries are the only synchronisation.
//
// Context hashes come from the flattened recipes of zpq_hcomp.cpp (build_wave_plan): each thread
// carries private copies of the states its hash is seeded from.  Bit-exactness: every table and every
// state word is touched by exactly one thread in stream order; tests compare this tier with the oracle.
#include <cuda_runtime.h>
#include <stdint.h>
#include "zpq_dev.h"
#include "zpq_kernels.h"
#include "zpq_kcommon.cuh"

struct Enc5 {
  u32 low, high;
  u8* out; u64 out_pos, out_cap;
  int status;
};

__device__ __forceinline__ void enc5_bit(Enc5& C, int y, u32 p16) {
  const u32 mid = C.low + (u32)(((u64)(C.high - C.low) * p16) >> 16);
  if (y) C.high = mid; else C.low = mid + 1;
  while ((C.high ^ C.low) < 0x1000000u) {
    if (C.out_pos < C.out_cap) C.out[C.out_pos] = (u8)(C.high >> 24);
    else C.status = ST_FULL;
    ++C.out_pos;
    C.high = C.high << 8 | 255u;
    C.low = C.low << 8;
    C.low += (C.low == 0);
  }
}

#define CS_WORDS 16     // per (block, component) state record, in u32

// ROLE 0: component threads and MIX warps in one launch; 1: component threads only; 2: MIX warps only
// (the split launches have their own register budgets).  Returns the profiling slot of the warp
// (component type, 10 = coder, 11 + q = mixer q) or -1 for a warp without work in this wave.
template <int ROLE, bool MIXPF, int UNR>
__device__ __forceinline__ int wave_body(const ZWaveArgs& W, unsigned char* smem) {
  const Tabs S = make_tabs(smem);
  const ZKernelArgs& A = W.A;

  const int lane = threadIdx.x & 31;
  const int CB = W.chunk_bytes, NCS = W.ncs;
  const int nA = W.ngroups * NCS;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) + (ROLE == 2 ? nA : 0);   // global warp index

  if (ROLE != 2 && gw < nA) {
    // =======================================================================================
    // thread = (block, component) ; warp = one component of the 32 blocks of a group
    // =======================================================================================
    // role-major order (W.role_major): consecutive warps -- and so the warps of a CTA and of neighbouring CTAs -- run the
    // SAME component of different groups, i.e. the same code: the instruction caches of an SM then hold one or two roles
    // instead of all of them (ncu of the group-major order: 7 of 19.6 stall cycles per instruction were instruction fetches)
    const int g = W.role_major ? gw % W.ngroups : gw / NCS, c = W.role_major ? gw / W.ngroups : gw - g * NCS;
    const int cnt = (int)W.group_count[g];
    if (lane >= cnt) return -1;
    const int bi = (int)W.group_first[g] + lane;
    const ZBlock B = A.blocks[bi];
    const ZModel* __restrict__ ZM = A.models + B.model;
    const int n = (int)ZM->n;
    u8* const arena = A.arena + B.arena_off;
    int N = 0;
    for (u32 sg = 0; sg < B.nseg; ++sg) N += (int)A.seg_len[B.seg_first + sg];
    const u8* const in = A.in + B.in_off;
    u32* const cs = W.cstate + ((size_t)bi * NCS + c) * CS_WORDS;
    const int GS = W.gsize;
    uint4* const pb = W.pbuf + (size_t)g * NCS * 16 * CB * GS;        // this group's streams
#define PBIDX(comp, slot, i) ((((size_t)(comp) * 16 + (slot)) * CB + (i)) * GS + lane)

    if (c == NCS - 1) {
      // ---------------- coder stage ----------------
      const int ch = W.wave - (int)ZM->wave.dcoder;
      if (ch < 0) return -1;
      const int b0 = ch * CB;
      if (b0 >= N && !(ch == 0)) return -1;
      const int b1 = min(N, b0 + CB);
      Enc5 C;
      C.out = A.out + B.out_off; C.out_cap = B.out_cap;
      u32 seg; u64 seg_end, seg_start_out;
      if (ch == 0) {
        C.low = 1; C.high = 0xffffffffu; C.out_pos = 0; C.status = ST_OK;
        seg = 0; seg_end = B.nseg ? A.seg_len[B.seg_first] : 0; seg_start_out = 0;
      } else {
        C.low = cs[0]; C.high = cs[1]; C.out_pos = (u64)cs[2] | (u64)cs[3] << 32; C.status = (int)cs[9];
        seg = cs[4]; seg_end = (u64)cs[5] | (u64)cs[6] << 32; seg_start_out = (u64)cs[7] | (u64)cs[8] << 32;
      }
      const int slot = ch & 15;
      for (int pos = b0;; ++pos) {
        // Compressor::endSegment: EOS after the last byte of a segment (libzpaq.cpp:2950-2952)
        while (seg < B.nseg && seg_end == (u64)pos) {
          enc5_bit(C, 1, 0);
          A.seg_out[B.seg_first + seg] = C.out_pos - seg_start_out;
          seg_start_out = C.out_pos; ++seg;
          if (seg < B.nseg) seg_end += A.seg_len[B.seg_first + seg];
        }
        if (pos >= b1) break;
        const u32 cur = in[pos];
        const uint4 v = pb[PBIDX(n - 1, slot, pos - b0)];
        enc5_bit(C, 0, 0);                                    // per-byte EOS flag, p = 0 (libzpaq.cpp:2426)
#pragma unroll
        for (int bit = 0; bit < 8; ++bit)
          enc5_bit(C, (int)((cur >> (7 - bit)) & 1u), (u32)t_squash(S, vget(v, bit)) * 2u + 1u);
      }
      cs[0] = C.low; cs[1] = C.high; cs[2] = (u32)C.out_pos; cs[3] = (u32)(C.out_pos >> 32); cs[9] = (u32)C.status;
      cs[4] = seg; cs[5] = (u32)seg_end; cs[6] = (u32)(seg_end >> 32); cs[7] = (u32)seg_start_out; cs[8] = (u32)(seg_start_out >> 32);
      if (b1 >= N) { A.status[bi] = C.status; A.out_len[bi] = C.out_pos; }
      return 10;
    }
    if (c >= n) return -1;
    const ZComp cd = ZM->comp[c];
    const u32 type = cd.type;
    if (type == ZT_MIX && !W.mix_threads) return -1;                // mixers have their own warps below
    const int ch = W.wave - (int)ZM->pipe.depth[c];
    if (ch < 0) return -1;
    const int b0 = ch * CB;
    if (b0 >= N) return -1;
    const int b1 = min(N, b0 + CB);
    const int slot = ch & 15;

    const u32 a1 = cd.a1, a2 = cd.a2, a3 = cd.a3, a4 = cd.a4, a5 = cd.a5;
    const u32 cm_mask = cd.cm_mask, ht_mask = cd.ht_mask, limit = cd.limit;
    u32* const gtab = reinterpret_cast<u32*>(arena + (u64)cd.cm_off16 * 16);
    u8* const ht = arena + (u64)cd.ht_off16 * 16;
    const u32* const fops = ZM->wave.flat[c];
    const u32 flen = ZM->wave.flat_len[c];
    const u32 fold1 = ZM->wave.fold_n[c];
    const u32 dist_byte = ZM->pipe.dist_byte[c];
    const bool is_hash = (type == ZT_ICM || type == ZT_ISSE);

    u32 h = cs[0];
    u32 st[ZPQ_FLAT_SEGS] = {cs[1], cs[2], cs[3], cs[4]};
    u64 hist = (u64)cs[5] | (u64)cs[6] << 32;
    u32 dist_last = cs[7];
    u32 ma = cs[8], mb = cs[9], mpos = cs[10], mpred = cs[11], mc = 0;
    u32 h_next = h;
    if (type == ZT_MIX) {
      // ---------------- MIX as a component thread (W.mix_threads) ----------------
      // With a full c8 mask and >= 256 weight sets the eight bits of a byte use eight DIFFERENT weight rows, and a row
      // trained in this byte is next read in a later byte: the eight dot products of a byte are independent.  One thread
      // then codes a whole block's mixer: per byte it walks the m inputs twice (sums, then training), eight rows at a
      // time, instead of a warp spending ~150 instructions per bit on one block.  Same operations on the same weights
      // in an order that differs only between rows that do not interact -> same bytes.
      const int m = (int)a3, j0 = (int)a2, rate = (int)a4;
      int* const tab = reinterpret_cast<int*>(gtab);
      for (int pos = b0; pos < b1; ++pos) {
        const u32 cur = in[pos];
        const int i = pos - b0;
        u32 bb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) bb[k] = ((h + (((0x100u | cur) >> (8 - k)) & a5)) & cm_mask) * (u32)m;
        int sum[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) sum[k] = 0;
        for (int j = 0; j < m; ++j) {
          const uint4 v = pb[PBIDX(j0 + j, slot, i)];
#pragma unroll
          for (int k = 0; k < 8; ++k) sum[k] += (tab[bb[k] + (u32)j] >> 8) * vget(v, k);
        }
        int err[8];
        uint4 vout = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int pk = clamp2k(sum[k] >> 8);
          const int y = (int)((cur >> (7 - k)) & 1u);
          err[k] = ((y * 32767 - t_squash(S, pk)) * rate) >> 4;
          vset(vout, k, pk);
        }
        for (int j = 0; j < m; ++j) {
          const uint4 v = pb[PBIDX(j0 + j, slot, i)];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            int* const wp = tab + bb[k] + (u32)j;
            *wp = clamp512k(*wp + ((err[k] * vget(v, k) + (1 << 12)) >> 13));
          }
        }
        pb[PBIDX(c, slot, i)] = vout;
        const u64 hs = (hist << 8) | (u64)cur;
        h = wave_hash<UNR == 1>(fold1, fops, flen, st, cur, hs, pos, in, 0u);
        hist = hs;
        if (pos + 1 < b1) {                      // the next byte's eight rows are known now: pull them towards L2
          const u32 nxt = in[pos + 1];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int* const r = tab + ((h + (((0x100u | nxt) >> (8 - k)) & a5)) & cm_mask) * (u32)m;
            prefetch_l2(r); prefetch_l2(r + m - 1);
          }
        }
      }
      cs[0] = h; cs[1] = st[0]; cs[2] = st[1]; cs[3] = st[2]; cs[4] = st[3];
      cs[5] = (u32)hist; cs[6] = (u32)(hist >> 32);
      return 11;
    }
    if (is_hash && W.hash_loop) {
      // ---------------- ICM / ISSE: the loop of the longest chain, written out per nibble ----------------
      // Same operations as the generic loop below (which stays the path of the other component types and the A/B
      // reference for this one: ZPQ_WAVE_VARIANT bit 4 off), but with the two nibbles and their four bits unrolled:
      // the bit index is a constant (vector lane extraction, "last bit of the nibble" tests and the type switches fold
      // away) and the slot inside the hash row is a running value (1, 1y, 1yy, 1yyy).
      const bool isse = type == ZT_ISSE;
      uint4 row = make_uint4(0, 0, 0, 0); u32 row_off = 0;
      uint4 pf0 = row, pf1 = row, pf2 = row; u32 pf_h0 = 0; bool pf_valid = false;
      u32 cur_n = in[b0];
      for (int pos = b0; pos < b1; ++pos) {
        const u32 cur = cur_n;
        const int i = pos - b0;
        uint4 vin_j = make_uint4(0, 0, 0, 0);
        if (isse) vin_j = pb[PBIDX(a2, slot, i)];
        const bool more = pos + 1 < b1;
        if (more) {
          cur_n = in[pos + 1];
          if (isse) prefetch_l1(&pb[PBIDX(a2, slot, i + 1)]);
        }
        uint4 vout = make_uint4(0, 0, 0, 0);
#pragma unroll UNR
        for (int nb = 0; nb < 2; ++nb) {
          if (nb == 1) {
            const u64 hs = (hist << 8) | (u64)cur;
            if (cur == dist_byte) dist_last = (u32)(-(pos + 1));
            h_next = wave_hash<UNR == 1>(fold1, fops, flen, st, cur, hs, pos, in, dist_last);
            hist = hs;
          }
          // Predictor::find (libzpaq.cpp:2072-2088) on rows requested one nibble ago
          const u32 c8 = nb == 0 ? 1u : (0x10u | (cur >> 4));
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
          const u32 ncx = nb == 0 ? (h + 16u * (0x10u | (cur >> 4))) : (h_next + 16u);
          const u32 nh0 = (ncx * 16u) & ht_mask;
          pf_valid = ((nh0 ^ row_off) & ~63u) != 0u && (nb == 0 || more);
          if (pf_valid) {
            pf_h0 = nh0;
            pf0 = *reinterpret_cast<const uint4*>(ht + nh0);
            pf1 = *reinterpret_cast<const uint4*>(ht + (nh0 ^ 16u));
            pf2 = *reinterpret_cast<const uint4*>(ht + (nh0 ^ 32u));
          }
          // the four bits of the nibble; hsl = slot inside the row = hmap4 & 15
          const u32 nibv = nb == 0 ? (cur >> 4) : (cur & 15u);
          u32 hsl = 1u;
          u32 bh = row_get(row, 1);
          u32 t0, t1 = 0;                                   // table entry of the current bit (ISSE: weight pair)
          if (isse) { const uint2 w = *reinterpret_cast<const uint2*>(gtab + bh * 2); t0 = w.x; t1 = w.y; }
          else t0 = gtab[bh];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int bit = nb * 4 + b;
            const int y = (int)((nibv >> (3 - b)) & 1u);
            // entry of the NEXT bit of this nibble (its slot is input data), requested before this bit's arithmetic
            const u32 hsn = hsl * 2u + (u32)y;
            u32 nbh = 0, n0 = 0, n1 = 0;
            if (b < 3) {
              nbh = row_get(row, (int)hsn);
              if (isse) { const uint2 wn = *reinterpret_cast<const uint2*>(gtab + nbh * 2); n0 = wn.x; n1 = wn.y; }
              else n0 = gtab[nbh];
            }
            int pr;
            u32 u0, u1 = 0;
            if (isse) {
              const int xj = vget(vin_j, bit);
              pr = clamp2k(((int)t0 * xj + (int)t1 * 64) >> 16);
              const int err = y * 32767 - t_squash(S, pr);
              u0 = (u32)clamp512k((int)t0 + ((err * xj + (1 << 12)) >> 13));
              u1 = (u32)clamp512k((int)t1 + ((err + 16) >> 5));
              *reinterpret_cast<uint2*>(gtab + bh * 2) = make_uint2(u0, u1);
            } else {
              pr = t_stretch(S, t0 >> 8);
              u0 = t0 + (u32)(((int)(y * 32767 - (int)(t0 >> 8))) >> 2);
              gtab[bh] = u0;
            }
            row_set(row, (int)hsl, ((u32)S.ns[bh] >> (8 * y)) & 255u);
            vset(vout, bit, pr);
            if (b < 3) {
              // the next bit's entry is the one just trained when both slots hold the same bit-history state
              const bool same = nbh == bh;
              t0 = same ? u0 : n0; t1 = same ? u1 : n1;
              bh = nbh; hsl = hsn;
            }
          }
          *reinterpret_cast<uint4*>(ht + row_off) = row;
        }
        pb[PBIDX(c, slot, i)] = vout;
        h = h_next;
      }
      cs[0] = h; cs[1] = st[0]; cs[2] = st[1]; cs[3] = st[2]; cs[4] = st[3];
      cs[5] = (u32)hist; cs[6] = (u32)(hist >> 32); cs[7] = dist_last;
      cs[8] = ma; cs[9] = mb; cs[10] = mpos; cs[11] = mpred;
      return (int)type;
    }
    if (type == ZT_SSE && W.hash_loop) {
      // ---------------- SSE: the eight bits of a byte unrolled (same switch as the ICM / ISSE loop above) ----------------
      // Same operations as the generic loop below; with a constant bit index c8, the input extraction and the
      // "is there a next bit in this byte" tests are free, and the two interpolation buckets of the NEXT bit (known from
      // input data and from the input prediction that is already here) are requested before this bit's arithmetic.
      u32 cur_n = in[b0];
      for (int pos = b0; pos < b1; ++pos) {
        const u32 cur = cur_n;
        const int i = pos - b0;
        const uint4 vin_j = pb[PBIDX(a2, slot, i)];
        if (pos + 1 < b1) {
          cur_n = in[pos + 1];
          prefetch_l1(&pb[PBIDX(a2, slot, i + 1)]);
        }
        uint4 vout = make_uint4(0, 0, 0, 0);
        u32 q0 = 0, q1 = 0, key = 0;
        {
          const u32 rb = ((h + 1u) * 32u) & cm_mask;
          key = rb + (u32)(min(max(vget(vin_j, 0) + 992, 0), 1983) >> 6);
          q0 = gtab[key]; q1 = gtab[key + 1u];
        }
#pragma unroll (UNR == 2 ? 8 : 2)
        for (int bit = 0; bit < 8; ++bit) {
          const int y = (int)((cur >> (7 - bit)) & 1u);
          if ((bit & 3) == 0) {
            // rows of the bits after next: into L2 now (their addresses are input data)
#pragma unroll
            for (int k = 2; k <= 5; ++k)
              if (bit + k < 8) prefetch_l2(gtab + (((h + ((0x100u | cur) >> (8 - bit - k))) * 32u) & cm_mask));
            if (bit == 4) {
              const u64 hs = (hist << 8) | (u64)cur;
              if (cur == dist_byte) dist_last = (u32)(-(pos + 1));
              h_next = wave_hash<UNR == 1>(fold1, fops, flen, st, cur, hs, pos, in, dist_last);
              hist = hs;
            }
          }
          const int pin = min(max(vget(vin_j, bit) + 992, 0), 1983);
          const int wt = pin & 63;
          // buckets of the next bit of this byte
          u32 nkey = 0, n0 = 0, n1 = 0;
          if (bit < 7) {
            const u32 rbn = ((h + ((0x100u | cur) >> (7 - bit))) * 32u) & cm_mask;
            nkey = rbn + (u32)(min(max(vget(vin_j, bit + 1) + 992, 0), 1983) >> 6);
            n0 = gtab[nkey]; n1 = gtab[nkey + 1u];
          }
          const int pr = t_stretch(S, (u32)(((int)(q0 >> 10) * (64 - wt) + (int)(q1 >> 10) * wt) >> 13));
          const u32 cxt = key + (u32)(wt >> 5);                 // the nearer bucket is trained (libzpaq.cpp:1930-1939)
          const u32 v0 = (wt >> 5) ? q1 : q0;
          const u32 count = v0 & 0x3ffu;
          const int err = y * 32767 - (int)(v0 >> 17);
          const u32 prod = (u32)err * ((u32)S.dt[count] * 2u);
          const u32 up = v0 + (prod & 0xfffffc00u) + (count < limit ? 1u : 0u);
          gtab[cxt] = up;
          vset(vout, bit, pr);
          if (bit < 7) {
            // a prefetched bucket that was just trained takes the trained value
            if (nkey == cxt) n0 = up;
            if (nkey + 1u == cxt) n1 = up;
            key = nkey; q0 = n0; q1 = n1;
          }
        }
        pb[PBIDX(c, slot, i)] = vout;
        h = h_next;
      }
      cs[0] = h; cs[1] = st[0]; cs[2] = st[1]; cs[3] = st[2]; cs[4] = st[3];
      cs[5] = (u32)hist; cs[6] = (u32)(hist >> 32); cs[7] = dist_last;
      cs[8] = ma; cs[9] = mb; cs[10] = mpos; cs[11] = mpred;
      return (int)type;
    }
    int p = (type == ZT_CONS) ? ((int)a1 - 128) * 4 : 0;
    uint4 row = make_uint4(0, 0, 0, 0); u32 row_off = 0;
    uint4 pf0 = row, pf1 = row, pf2 = row; u32 pf_h0 = 0; bool pf_valid = false;
    u32 bh = 0, v0 = 0, v1 = 0, cxt = 0; int xj = 0, xk = 0;
    // next-bit prefetch: the table entry the NEXT bit will use is known from input data, so it is
    // requested while the current bit is computed; nx_* hold it, nx_key identifies it, and an entry that
    // equals the one trained in this bit is replaced by the trained value (nx_same)
    u32 nx_key = 0xffffffffu, nx_a = 0, nx_b = 0, up_a = 0, up_b = 0; bool nx_ok = false, nx_same = false;
    u32 nx_key1 = 0xffffffffu;   // SSE: index of the second bucket (first is nx_key)

    const u32 in_j = (type == ZT_AVG) ? a1 : a2, in_k = (type == ZT_AVG) ? a2 : a3;
    const bool has_j = (type == ZT_ISSE || type == ZT_SSE || type == ZT_MIX2 || type == ZT_AVG);
    const bool has_k = (type == ZT_MIX2 || type == ZT_AVG);
    // The input vectors of byte pos+1 are pulled into L1 while byte pos is coded and loaded at the top of their own byte.
    // (Holding them in registers instead costs eight registers across the whole bit loop; at the 128-register budget of
    // 4 CTAs per SM the compiler spilled them right behind their loads, i.e. waited for the stream read every byte.)
    u32 cur_n = in[b0];
    for (int pos = b0; pos < b1; ++pos) {
      const u32 cur = cur_n;
      const int i = pos - b0;
      uint4 vin_j = make_uint4(0, 0, 0, 0), vin_k = vin_j;
      if (has_j) vin_j = pb[PBIDX(in_j, slot, i)];
      if (has_k) vin_k = pb[PBIDX(in_k, slot, i)];
      if (pos + 1 < b1) {
        cur_n = in[pos + 1];
        if (has_j) prefetch_l1(&pb[PBIDX(in_j, slot, i + 1)]);
        if (has_k) prefetch_l1(&pb[PBIDX(in_k, slot, i + 1)]);
      }
      uint4 vout = make_uint4(0, 0, 0, 0);

      // c8 / hmap4 as Predictor::update maintains them (libzpaq.cpp:2054-2066), carried from bit to bit: the bit
      // index is a run-time value in this loop, so the closed forms would cost variable shifts and selects every bit
      u32 c8 = 1u, hmap4 = 1u, curs = cur << 24;
#pragma unroll 1
      for (int bit = 0; bit < 8; ++bit) {
        const int y = (int)(curs >> 31);
        const int hslot = (int)(hmap4 & 15u);
        // context of the NEXT bit (bit 7: first bit of the next byte, c8 = hmap4 = 1)
        const u32 c8n = bit < 7 ? c8 * 2u + (u32)y : 1u;
        const u32 hmap4n = bit == 7 ? 1u : (bit == 3 ? ((hmap4 & 0xfu) << 5 | (u32)y << 4 | 1u)
                                                     : ((hmap4 & 0x1f0u) | (((hmap4 & 0xfu) * 2u + (u32)y) & 0xfu)));

        if ((bit & 3) == 0) {
          // tables indexed per bit: pull the lines of the NEXT nibble into L2 now (addresses are input data)
          if (type == ZT_CM) {
            if (bit == 0) prefetch_l2(gtab + ((h ^ ((0x10u | (cur >> 4)) << 4)) & cm_mask));
          }
          else if (type == ZT_SSE) {
#pragma unroll
            for (int k = 2; k <= 5; ++k) {
              if (bit + k < 8) prefetch_l2(gtab + (((h + ((0x100u | cur) >> (8 - bit - k))) * 32u) & cm_mask));
            }
          }
          if (bit == 4) {
            const u64 hs = (hist << 8) | (u64)cur;
            if (cur == dist_byte) dist_last = (u32)(-(pos + 1));
            h_next = wave_hash<UNR == 1>(fold1, fops, flen, st, cur, hs, pos, in, dist_last);
            hist = hs;
          }
          if (is_hash) {
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
            pf_valid = ((nh0 ^ row_off) & ~63u) != 0u && (bit == 0 || pos + 1 < b1);
            if (pf_valid) {
              pf_h0 = nh0;
              pf0 = *reinterpret_cast<const uint4*>(ht + nh0);
              pf1 = *reinterpret_cast<const uint4*>(ht + (nh0 ^ 16u));
              pf2 = *reinterpret_cast<const uint4*>(ht + (nh0 ^ 32u));
            }
          }
        }

        // ---- predict (libzpaq.cpp:1854-1951) ----
        switch (type) {
          case ZT_ICM:
            bh = row_get(row, hslot);
            if (nx_ok && nx_key == bh) v0 = nx_same ? up_a : nx_a;
            else v0 = gtab[bh];
            nx_ok = false;
            if ((bit & 3) != 3) {                       // same row next bit: its slot is already final
              nx_key = row_get(row, (int)(hmap4n & 15u)); nx_a = gtab[nx_key]; nx_ok = true;
            }
            p = t_stretch(S, v0 >> 8);
            break;
          case ZT_ISSE: {
            bh = row_get(row, hslot);
            if (nx_ok && nx_key == bh) { v0 = nx_same ? up_a : nx_a; v1 = nx_same ? up_b : nx_b; }
            else { const uint2 w = *reinterpret_cast<const uint2*>(gtab + bh * 2); v0 = w.x; v1 = w.y; }
            nx_ok = false;
            if ((bit & 3) != 3) {
              nx_key = row_get(row, (int)(hmap4n & 15u));
              const uint2 wn = *reinterpret_cast<const uint2*>(gtab + nx_key * 2);
              nx_a = wn.x; nx_b = wn.y; nx_ok = true;
            }
            xj = vget(vin_j, bit);
            p = clamp2k(((int)v0 * xj + (int)v1 * 64) >> 16);
            break;
          }
          case ZT_CM:
            cxt = (h ^ hmap4) & cm_mask;
            if (nx_ok && nx_key == cxt) v0 = nx_same ? up_a : nx_a;
            else v0 = gtab[cxt];
            nx_ok = false;
            if (bit < 7 || pos + 1 < b1) {
              nx_key = ((bit < 7 ? h : h_next) ^ hmap4n) & cm_mask; nx_a = gtab[nx_key]; nx_ok = true;
            }
            p = t_stretch(S, v0 >> 17);
            break;
          case ZT_MATCH:
            if (ma == 0) p = 0;
            else {
              mc = (mpred >> (7 - bit)) & 1u;
              p = t_stretch(S, (u32)((int)S.dt2k[ma] * (1 - 2 * (int)mc)) & 32767u);
            }
            break;
          case ZT_MIX2: {
            cxt = (h + (c8 & a5)) & cm_mask;
            const int w = (int)reinterpret_cast<const u16*>(gtab)[cxt];
            v0 = (u32)w; xj = vget(vin_j, bit); xk = vget(vin_k, bit);
            p = (w * xj + (65536 - w) * xk) >> 16;
            break;
          }
          case ZT_AVG:
            p = (vget(vin_j, bit) * (int)a3 + vget(vin_k, bit) * (256 - (int)a3)) >> 8;
            break;
          case ZT_SSE: {
            const u32 rb = ((h + c8) * 32u) & cm_mask;
            int pq = min(max(vget(vin_j, bit) + 992, 0), 1983);
            const int wt = pq & 63;
            pq >>= 6;
            u32 q0, q1;
            if (nx_ok && nx_key == rb + (u32)pq) { q0 = nx_a; q1 = nx_b; }
            else { q0 = gtab[rb + (u32)pq]; q1 = gtab[rb + (u32)pq + 1u]; }
            nx_ok = false;
            if (bit < 7) {                               // next bit of the same byte: input prediction already here
              const u32 rbn = ((h + c8n) * 32u) & cm_mask;
              const int pqn = min(max(vget(vin_j, bit + 1) + 992, 0), 1983) >> 6;
              nx_key = rbn + (u32)pqn; nx_key1 = nx_key + 1u;
              nx_a = gtab[nx_key]; nx_b = gtab[nx_key1]; nx_ok = true;
            }
            p = t_stretch(S, (u32)(((int)(q0 >> 10) * (64 - wt) + (int)(q1 >> 10) * wt) >> 13));
            if (wt >> 5) { cxt = rb + (u32)pq + 1u; v0 = q1; } else { cxt = rb + (u32)pq; v0 = q0; }
            break;
          }
          default: break;   // CONS keeps its constant
        }

        // ---- update (libzpaq.cpp:1954-2052) ----
        switch (type) {
          case ZT_ICM: case ZT_ISSE: {
            const u32 nx = ((u32)S.ns[bh] >> (8 * y)) & 255u;
            row_set(row, hslot, nx);
            if (type == ZT_ISSE) {
              const int err = y * 32767 - t_squash(S, p);
              const int w0 = clamp512k((int)v0 + ((err * xj + (1 << 12)) >> 13));
              const int w1 = clamp512k((int)v1 + ((err + 16) >> 5));
              *reinterpret_cast<uint2*>(gtab + bh * 2) = make_uint2((u32)w0, (u32)w1);
              up_a = (u32)w0; up_b = (u32)w1;
            } else {
              up_a = v0 + (u32)(((int)(y * 32767 - (int)(v0 >> 8))) >> 2);
              gtab[bh] = up_a;
            }
            nx_same = nx_ok && nx_key == bh;
            if ((bit & 3) == 3) *reinterpret_cast<uint4*>(ht + row_off) = row;
            break;
          }
          case ZT_CM: case ZT_SSE: {
            const u32 count = v0 & 0x3ffu;
            const int err = y * 32767 - (int)(v0 >> 17);
            const u32 prod = (u32)err * ((u32)S.dt[count] * 2u);
            up_a = v0 + (prod & 0xfffffc00u) + (count < limit ? 1u : 0u);
            gtab[cxt] = up_a;
            if (type == ZT_CM) nx_same = nx_ok && nx_key == cxt;
            else if (nx_ok) {                            // SSE: patch whichever prefetched bucket was just trained
              if (nx_key == cxt) nx_a = up_a;
              if (nx_key1 == cxt) nx_b = up_a;
            }
            break;
          }
          case ZT_MIX2: {
            const int err = ((y * 32767 - t_squash(S, p)) * (int)a4) >> 5;
            int w = (int)v0;
            w += (err * (xj - xk) + (1 << 12)) >> 13;
            w = min(max(w, 0), 65535);
            reinterpret_cast<u16*>(gtab)[cxt] = (u16)w;
            break;
          }
          case ZT_MATCH:
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
            break;
          default: break;
        }
        vset(vout, bit, p);
        c8 = c8n; hmap4 = hmap4n; curs <<= 1;
      }
      pb[PBIDX(c, slot, i)] = vout;
      h = h_next;
    }
    cs[0] = h; cs[1] = st[0]; cs[2] = st[1]; cs[3] = st[2]; cs[4] = st[3];
    cs[5] = (u32)hist; cs[6] = (u32)(hist >> 32); cs[7] = dist_last;
    cs[8] = ma; cs[9] = mb; cs[10] = mpos; cs[11] = mpred;
    return (int)type;
  }

  // =========================================================================================
  // MIX warps: warp = (block, mixer); lane l = input l
  // =========================================================================================
  if (ROLE != 1 && !W.mix_threads) {
    const int idx = gw - nA;
    const int bi = idx / ZPQ_FAST_MIX, q = idx - bi * ZPQ_FAST_MIX;
    if (bi >= A.nblk) return -1;
    const ZBlock B = A.blocks[bi];
    const ZModel* __restrict__ ZM = A.models + B.model;
    if (q >= (int)ZM->team.nmix) return -1;
    const int cq = (int)ZM->team.mix_comp[q];
    const ZComp cd = ZM->comp[cq];
    const int ch = W.wave - (int)ZM->pipe.depth[cq];
    if (ch < 0) return -1;
    int N = 0;
    for (u32 sg = 0; sg < B.nseg; ++sg) N += (int)A.seg_len[B.seg_first + sg];
    const int b0 = ch * CB;
    if (b0 >= N) return -1;
    const int b1 = min(N, b0 + CB);
    const int slot = ch & 15;
    const u8* const in = A.in + B.in_off;
    u8* const arena = A.arena + B.arena_off;
    int* const tab = reinterpret_cast<int*>(arena + (u64)cd.cm_off16 * 16);
    const int g = (int)W.blk_group[bi], bl = (int)W.blk_lane[bi];
    const int GS = W.gsize;
    uint4* const pb = W.pbuf + (size_t)g * NCS * 16 * CB * GS;
#define PBIDX2(comp, slot, i) ((((size_t)(comp) * 16 + (slot)) * CB + (i)) * GS + bl)
    u32* const cs = W.cstate + ((size_t)bi * NCS + cq) * CS_WORDS;
    const int m = (int)cd.a3, j0 = (int)cd.a2, rate = (int)cd.a4;
    const u32 a5 = cd.a5, cmm = cd.cm_mask;
    const u32* const fops = ZM->wave.flat[cq];
    const u32 flen = ZM->wave.flat_len[cq];
    const u32 fold1 = ZM->wave.fold_n[cq];

    u32 h = cs[0];
    u32 st[ZPQ_FLAT_SEGS] = {cs[1], cs[2], cs[3], cs[4]};
    u64 hist = (u64)cs[5] | (u64)cs[6] << 32;
    u32 h_next = h;
    int w_ld = 0, w_upd = 0; u32 wbase = 0; bool wvalid = false, wsame = false;

    u32 cur_n = in[b0];
    uint4 vin_n = make_uint4(0, 0, 0, 0);
    if (lane < m) vin_n = pb[PBIDX2(j0 + lane, slot, 0)];
    if (W.mix_batch && (a5 & 255u) == 255u && cmm >= 255u) {
      // Byte-batched mixing.  Under a full c8 mask (and >= 256 weight sets) the eight bits of a byte use eight DIFFERENT
      // weight rows and a row trained in this byte is next read in a later byte, so the eight dot products of a byte do
      // not depend on each other: load the eight rows, reduce the eight sums back to back, let lane k do bit k's scalar
      // work (clamp, squash, error) for all eight bits at once, broadcast the eight errors, train the eight rows.
      // ~35 warp instructions per coded bit instead of ~90; same operations on every weight, in an order that differs
      // only between rows that do not interact.
      const int kk = lane & 7;
      for (int pos = b0; pos < b1; ++pos) {
        const u32 cur = cur_n;
        const int i = pos - b0;
        const uint4 vin = vin_n;
        const bool more = pos + 1 < b1;
        if (more) {
          cur_n = in[pos + 1];
          if (lane < m) vin_n = pb[PBIDX2(j0 + lane, slot, i + 1)];
        }
        u32 bb[8]; int wr[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          bb[k] = ((h + (((0x100u | cur) >> (8 - k)) & a5)) & cmm) * (u32)m;
          wr[k] = lane < m ? tab[bb[k] + lane] : 0;
        }
        // the next byte's hash needs only input data: evaluate it now and pull the next byte's rows towards L2
        const u64 hs = (hist << 8) | (u64)cur;
        h_next = wave_hash<UNR == 1>(fold1, fops, flen, st, cur, hs, pos, in, 0u);
        hist = hs;
        if (more && lane < 16) {
          const int* const r = tab + ((h_next + (((0x100u | cur_n) >> (8 - kk)) & a5)) & cmm) * (u32)m;
          prefetch_l2(lane < 8 ? r : r + m - 1);
        }
        int sum[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) sum[k] = __reduce_add_sync(FULL, lane < m ? (wr[k] >> 8) * vget(vin, k) : 0);
        int s = sum[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) if (kk == k) s = sum[k];
        const int p = clamp2k(s >> 8);
        const int y = (int)((cur >> (7 - kk)) & 1u);
        const int e = ((y * 32767 - t_squash(S, p)) * rate) >> 4;
        if (lane < 8) reinterpret_cast<short*>(&pb[PBIDX2(cq, slot, i)])[kk] = (short)p;     // int16 k of the byte's vector
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int ek = __shfl_sync(FULL, e, k);
          if (lane < m) tab[bb[k] + lane] = clamp512k(wr[k] + ((ek * vget(vin, k) + (1 << 12)) >> 13));
        }
        h = h_next;
      }
    } else
    if (MIXPF && (a5 & 255u) == 255u && cmm >= 255u) {
      // The eight weight rows of a byte are distinct (c8 takes eight different values under a full mask)
      // and their indexes are input data: the rows of byte pos+1 are requested during bit 4 of byte pos,
      // as soon as its context hash is known, so a warp has eight row loads in flight instead of one.
      // A row of the next byte that is still to be trained in bits 4..7 of this byte ("late") is loaded
      // at the start of its own byte instead, after the stores (same thread, program order).
      int wr[8], wn[8];
      u32 latem = 0xffu;
#pragma unroll
      for (int k = 0; k < 8; ++k) { wr[k] = 0; wn[k] = 0; }
      for (int pos = b0; pos < b1; ++pos) {
        const u32 cur = cur_n;
        const int i = pos - b0;
        const uint4 vin = vin_n;
        const bool more = pos + 1 < b1;
        if (more) {
          cur_n = in[pos + 1];
          if (lane < m) vin_n = pb[PBIDX2(j0 + lane, slot, i + 1)];
        }
        u32 bb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          bb[k] = ((h + (((0x100u | cur) >> (8 - k)) & a5)) & cmm) * (u32)m;
          if ((latem >> k) & 1u) wr[k] = lane < m ? tab[bb[k] + lane] : 0;
          else wr[k] = wn[k];
        }
        latem = 0xffu;
        uint4 vout = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int bit = 0; bit < 8; ++bit) {
          if (bit == 4) {
            const u64 hs = (hist << 8) | (u64)cur;
            h_next = wave_hash<UNR == 1>(fold1, fops, flen, st, cur, hs, pos, in, 0u);
            hist = hs;
            if (more) {
              latem = 0u;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const u32 nb = ((h_next + (((0x100u | cur_n) >> (8 - k)) & a5)) & cmm) * (u32)m;
                if (nb == bb[4] || nb == bb[5] || nb == bb[6] || nb == bb[7]) latem |= 1u << k;
                else wn[k] = lane < m ? tab[nb + lane] : 0;
              }
            }
          }
          const int y = (int)((cur >> (7 - bit)) & 1u);
          const int wc = wr[bit];
          const int x = vget(vin, bit);
          const int sum = __reduce_add_sync(FULL, lane < m ? (wc >> 8) * x : 0);
          const int p = clamp2k(sum >> 8);
          const int err = ((y * 32767 - t_squash(S, p)) * rate) >> 4;
          if (lane < m) tab[bb[bit] + lane] = clamp512k(wc + ((err * x + (1 << 12)) >> 13));
          vset(vout, bit, p);
        }
        if (lane == 0) pb[PBIDX2(cq, slot, i)] = vout;
        h = h_next;
      }
    } else
    for (int pos = b0; pos < b1; ++pos) {
      const u32 cur = cur_n;
      const int i = pos - b0;
      const uint4 vin = vin_n;
      if (pos + 1 < b1) {
        cur_n = in[pos + 1];
        if (lane < m) vin_n = pb[PBIDX2(j0 + lane, slot, i + 1)];
      }
      uint4 vout = make_uint4(0, 0, 0, 0);
#pragma unroll 1
      for (int bit = 0; bit < 8; ++bit) {
        if (bit == 4) {
          const u64 hs = (hist << 8) | (u64)cur;
          h_next = wave_hash<UNR == 1>(fold1, fops, flen, st, cur, hs, pos, in, 0u);
          hist = hs;
        }
        const u32 c8 = (0x100u | cur) >> (8 - bit);
        const int y = (int)((cur >> (7 - bit)) & 1u);
        const u32 base = ((h + (c8 & a5)) & cmm) * (u32)m;
        if (bit < 5 && lane < 2) {                      // row of bit+3 (same byte) into L2: covers HBM + TLB latency
          const u32 c83 = (0x100u | cur) >> (5 - bit);
          const u32 b3 = ((h + (c83 & a5)) & cmm) * (u32)m;
          prefetch_l2(tab + b3 + (lane ? (u32)m - 1u : 0u));
        }
        int wc;
        if (wvalid && wbase == base) wc = wsame ? w_upd : w_ld;   // fetched (or trained) during the previous bit
        else wc = lane < m ? tab[base + lane] : 0;
        // request the next bit's weight row now: its index depends only on input data
        u32 nbase; bool nvalid;
        if (bit < 7) { const u32 nc8 = (0x100u | cur) >> (7 - bit); nbase = ((h + (nc8 & a5)) & cmm) * (u32)m; nvalid = true; }
        else { nbase = ((h_next + (1u & a5)) & cmm) * (u32)m; nvalid = pos + 1 < b1; }
        int wn = 0;
        if (nvalid && nbase != base && lane < m) wn = tab[nbase + lane];
        const int x = vget(vin, bit);
        const int sum = __reduce_add_sync(FULL, lane < m ? (wc >> 8) * x : 0);
        const int p = clamp2k(sum >> 8);
        const int err = ((y * 32767 - t_squash(S, p)) * rate) >> 4;
        const int wu = clamp512k(wc + ((err * x + (1 << 12)) >> 13));
        if (lane < m) tab[base + lane] = wu;
        w_ld = wn; w_upd = wu; wsame = (nbase == base); wbase = nbase; wvalid = nvalid;   // wn is first read next bit
        vset(vout, bit, p);
      }
      if (lane == 0) pb[PBIDX2(cq, slot, i)] = vout;
      h = h_next;
    }
    if (lane == 0) {
      cs[0] = h; cs[1] = st[0]; cs[2] = st[1]; cs[3] = st[2]; cs[4] = st[3];
      cs[5] = (u32)hist; cs[6] = (u32)(hist >> 32);
    }
    return 11 + q;
  }
  return -1;
}

template <int ROLE, bool MIXPF, int UNR>
__global__ void __launch_bounds__(128) zpq_wave_kernel(const ZWaveArgs W) {
  extern __shared__ __align__(16) unsigned char smem[];
  stage_tables(smem, W.A.tables);
  __syncthreads();
  const long long t0 = W.prof ? clock64() : 0;
  const int slot = wave_body<ROLE, MIXPF, UNR>(W, smem);
  if (W.prof && slot >= 0) {
    if ((threadIdx.x & 31) == 0) { atomicAdd(W.prof + 2 * slot, (unsigned long long)(clock64() - t0)); atomicAdd(W.prof + 2 * slot + 1, 1ull); }
  }
}

// ---------------------------------------------------------------------------
extern "C" cudaError_t zpq_launch_wave(const ZWaveArgs* w, cudaStream_t st, int variant) {
  // variant bit 0: byte-ahead weight rows in the MIX warps; bit 1: component threads and MIX warps as two launches
  // (w->mix_threads: the mixers are component threads, there are no MIX warps to launch); bit 5: compact code (the
  // nibble loop of the ICM / ISSE threads rolled, the SSE byte loop unrolled by two instead of eight).  The fully unrolled
  // kernel is 190 KB of SASS and every role is resident on every SM: ncu showed 7 of 19.6 stall cycles per instruction
  // waiting for instruction fetch (profiles/r02a_wave_*).  The smallest hot code -- one generic per-bit loop for all
  // component threads (no bit 4) plus the byte-batched MIX warps -- measured fastest (profiles/r02e_session.log);
  // a single non-inlined copy of the recipe interpreter removed the fetch stalls altogether (0.4 cycles) but cost more
  // in spills around the call and extra instructions than it saved (profiles/r02f_*), so the interpreter stays inlined
  const int wa = w->ngroups * w->ncs, wm = w->mix_threads ? 0 : w->A.nblk * ZPQ_FAST_MIX;
  if (wa + wm <= 0) return cudaSuccess;
  const bool pf = (variant & 1) != 0, compact = (variant & 32) != 0;
  if (variant & 2) {
    if (wa > 0) {
      if (compact) zpq_wave_kernel<1, false, 1><<<(wa + 3) / 4, 128, S2_STATIC, st>>>(*w);
      else zpq_wave_kernel<1, false, 2><<<(wa + 3) / 4, 128, S2_STATIC, st>>>(*w);
    }
    if (wm > 0) {
      if (pf) zpq_wave_kernel<2, true, 2><<<(wm + 3) / 4, 128, S2_STATIC, st>>>(*w);
      else zpq_wave_kernel<2, false, 2><<<(wm + 3) / 4, 128, S2_STATIC, st>>>(*w);
    }
  } else {
    const int ctas = (wa + wm + 3) / 4;
    if (compact) {
      if (pf) zpq_wave_kernel<0, true, 1><<<ctas, 128, S2_STATIC, st>>>(*w);
      else zpq_wave_kernel<0, false, 1><<<ctas, 128, S2_STATIC, st>>>(*w);
    } else {
      if (pf) zpq_wave_kernel<0, true, 2><<<ctas, 128, S2_STATIC, st>>>(*w);
      else zpq_wave_kernel<0, false, 2><<<ctas, 128, S2_STATIC, st>>>(*w);
    }
  }
  return cudaGetLastError();
}
