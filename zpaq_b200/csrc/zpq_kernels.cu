// zpq_kernels.cu -- sm_100a kernels of the context-mixing core.
//
// Mapping:
//   * one WARP owns one archive block: its model arena, its HCOMP machine, its
//     arithmetic coder.  Blocks are independent, so a persistent grid of
//     4-warp CTAs pulls block indices from an atomic counter.
//   * inside the warp, LANE i owns COMP component i (n <= 32): its tables, its
//     context hash h[i], its prediction p[i] and its training state live in
//     lane i's registers.  The data-independent part of predict() runs on all
//     lanes at once; the dependent part (ISSE chains, AVG, MIX2, SSE, MIX) is a
//     short host-built step list, p[] values move by warp shuffle, and every
//     MIX is a lane-parallel dot product reduced with redux.sync.
//   * ICM/ISSE hash rows (16 B) are fetched once per nibble into registers
//     (Predictor::find, libzpaq.cpp:2072-2088), worked on for 4 bits and
//     written back once.
//   * the model-independent tables (stretch, squash, dt, dt2k, bit-history
//     next-state) are staged in shared memory once per CTA.
//   * the HCOMP program is interpreted uniformly by all lanes (same registers,
//     same addresses), so H[i] is directly readable by lane i afterwards.
//
// Reference semantics restated here (bit-exact, integer only):
//   Predictor::predict0 libzpaq.cpp:1854-1951, update0 1954-2066, find 2072-2088,
//   train libzpaq.h:1151-1157, Encoder::encode 2402-2416, Decoder::decode
//   2104-2124, Decoder::decompress 2127-2155, ZPAQL::execute 1041-1262.
#include <cuda_runtime.h>
#include <stdint.h>
#include "zpq_dev.h"
#include "zpq_kernels.h"

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

#define FULL 0xffffffffu

// status codes mirrored from include/zpq.h (kept numeric here to stay header-light)
#define ST_OK 0
#define ST_VM (-3)
#define ST_FULL (-4)
#define ST_CORRUPT (-5)
#define ST_EOF (-6)
#define ST_EOS (-11)

// ---------------------------------------------------------------------------
// Arena initialisation: Predictor::init's table fills (libzpaq.cpp:1768-1850)
// as one coalesced 16-byte-per-thread store stream over all blocks.
__global__ void __launch_bounds__(256) zpq_init_kernel(const ZKernelArgs A) {
  const ZBlock B = A.blocks[blockIdx.y];
  const ZModel* __restrict__ M = A.models + B.model;
  u8* arena = A.arena + B.arena_off;
  const ZTables* __restrict__ T = A.tables;
  const u32 nf = M->nfills;
  const u32 stride = gridDim.x * blockDim.x;
  const ZFill* __restrict__ fills = M->wide ? reinterpret_cast<const ZFill*>(A.progs + M->prog_off + M->wide_fill_off) : M->fills;
  for (u32 f = 0; f < nf; ++f) {
    const ZFill F = fills[f];
    uint4* dst = reinterpret_cast<uint4*>(arena + (u64)F.off16 * 16);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < F.n16; i += stride) {
      uint4 v = make_uint4(0, 0, 0, 0);
      switch (F.kind) {
        case ZF_CONST32: v = make_uint4(F.value, F.value, F.value, F.value); break;
        case ZF_SSE: {
          // cm[j] = squash((j&31)*64-992)<<17 | start   (libzpaq.cpp:1844)
          u32 j = (i * 4) & 31;
          v.x = ((u32)T->squash[(j + 0) * 64 - 992 + 2048] << 17) | F.value;
          v.y = ((u32)T->squash[(j + 1) * 64 - 992 + 2048] << 17) | F.value;
          v.z = ((u32)T->squash[(j + 2) * 64 - 992 + 2048] << 17) | F.value;
          v.w = ((u32)T->squash[(j + 3) * 64 - 992 + 2048] << 17) | F.value;
          break;
        }
        case ZF_ICM: {
          u32 j = (i * 4) & 255;
          v = make_uint4(T->icm_init[j], T->icm_init[j + 1], T->icm_init[j + 2], T->icm_init[j + 3]);
          break;
        }
        case ZF_ISSE: {
          u32 j = (i * 4) & 511;
          v = make_uint4(T->isse_init[j], T->isse_init[j + 1], T->isse_init[j + 2], T->isse_init[j + 3]);
          break;
        }
        case ZF_BYTE1: if (i == 0) v.x = 1; break;   // MATCH buf: ht(0)=1 (libzpaq.cpp:1796)
        default: break;
      }
      dst[i] = v;
    }
  }
}

// ---------------------------------------------------------------------------
struct SharedTables {
  const short* stretch_hi;
  const u16* squash;
  const int* dt;
  const u32* ns;
  const u16* dt2k;
};

__device__ __forceinline__ int d_stretch(const SharedTables& S, u32 x) {
  return x >= 16384u ? (int)S.stretch_hi[x - 16384u] : -(int)S.stretch_hi[16383u - x];
}
__device__ __forceinline__ int d_squash(const SharedTables& S, int x) { return S.squash[x + 2048]; }
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

// Predictor::find (libzpaq.cpp:2072-2088).  The three candidate rows h0,
// h0^16, h0^32 share one 64-byte line; the chosen row is returned in
// registers and only written back at the end of the nibble.
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

// ---------------------------------------------------------------------------
// HCOMP interpreter (ZPAQL::run0/execute, libzpaq.cpp:1027-1262).  Run by lane 0
// of the block's warp (the call sites follow it with __syncwarp, after which
// lane i reads H[i]); the machine registers a,b,c,d,f live in that lane.
// `code` points at the first program byte inside an image with zero guards on
// both sides, so runaway jumps fetch opcode 0 = error exactly as the
// reference's guard gaps do (libzpaq.cpp:913).
struct VmState {
  u32 a, b, c, d, f;
  u8* m; u32* h; u32* r;
  u32 mmask, hmask;
  const u8* code; int len;
};

#define VM_MAX_STEPS (1 << 24)   // watchdog: the reference has none; a GPU must not hang

__device__ __noinline__ int vm_run(VmState& S, u32 input) {
  u32 a = input, b = S.b, c = S.c, d = S.d, f = S.f;
  u8* const m = S.m; u32* const h = S.h; u32* const r = S.r;
  const u32 mmask = S.mmask, hmask = S.hmask;
  const u8* const code = S.code;
  int pc = 0, rc = ST_OK;
  for (int guard = 0;; ++guard) {
    if (guard >= VM_MAX_STEPS) { rc = ST_VM; break; }
    const u32 op = code[pc++];
    if (op == 56) break;                                   // HALT
    if (op < 56) {
      const u32 tgt = op >> 3, act = op & 7;
      if (act == 7) {
        const u32 n = code[pc];
        if (tgt == 0) { a = r[n]; ++pc; }
        else if (tgt == 1) { b = r[n]; ++pc; }
        else if (tgt == 2) { c = r[n]; ++pc; }
        else if (tgt == 3) { d = r[n]; ++pc; }
        else if (tgt == 4) { if (f) pc += (int)((n + 128) & 255) - 127; else ++pc; }    // JT
        else if (tgt == 5) { if (!f) pc += (int)((n + 128) & 255) - 127; else ++pc; }   // JF
        else { r[n] = a; ++pc; }                                                          // R=A
        continue;
      }
      if (act > 4 || op == 0) { rc = ST_VM; break; }
      u32 v;
      if (tgt == 0) v = a; else if (tgt == 1) v = b; else if (tgt == 2) v = c; else if (tgt == 3) v = d;
      else if (tgt == 4) v = m[b & mmask]; else if (tgt == 5) v = m[c & mmask]; else v = h[d & hmask];
      if (act == 0) {                      // X<>A; *B and *C swap only the low byte of A
        const u32 olda = a;
        if (tgt == 4 || tgt == 5) { a = (olda & 0xffffff00u) | (v & 255u); v = olda & 255u; }
        else { a = v; v = olda; }
      }
      else if (act == 1) ++v;
      else if (act == 2) --v;
      else if (act == 3) v = ~v;
      else v = 0;
      if (tgt == 0) a = v; else if (tgt == 1) b = v; else if (tgt == 2) c = v; else if (tgt == 3) d = v;
      else if (tgt == 4) m[b & mmask] = (u8)v; else if (tgt == 5) m[c & mmask] = (u8)v; else h[d & hmask] = v;
      continue;
    }
    if (op < 64) {
      if (op == 57) continue;                                        // OUT: no output in HCOMP
      if (op == 59) { a = (a + m[b & mmask] + 512u) * 773u; continue; }          // HASH
      if (op == 60) { h[d & hmask] = (h[d & hmask] + a + 512u) * 773u; continue; } // HASHD
      if (op == 63) { const u32 n = code[pc]; pc += (int)((n + 128) & 255) - 127; continue; }  // JMP
      rc = ST_VM; break;
    }
    if (op == 255) {                                                 // LJ
      const int t = (int)code[pc] + 256 * (int)code[pc + 1];
      if (t >= S.len) { rc = ST_VM; break; }
      pc = t; continue;
    }
    const u32 grp = op >> 3, src = op & 7;
    if (grp == 15 || grp >= 30) { rc = ST_VM; break; }
    u32 v;
    if (src == 0) v = a; else if (src == 1) v = b; else if (src == 2) v = c; else if (src == 3) v = d;
    else if (src == 4) v = m[b & mmask]; else if (src == 5) v = m[c & mmask]; else if (src == 6) v = h[d & hmask];
    else v = code[pc++];
    switch (grp) {
      case 8: a = v; break;
      case 9: b = v; break;
      case 10: c = v; break;
      case 11: d = v; break;
      case 12: m[b & mmask] = (u8)v; break;
      case 13: m[c & mmask] = (u8)v; break;
      case 14: h[d & hmask] = v; break;
      case 16: a += v; break;
      case 17: a -= v; break;
      case 18: a *= v; break;
      case 19: a = v ? a / v : 0u; break;
      case 20: a = v ? a % v : 0u; break;
      case 21: a &= v; break;
      case 22: a &= ~v; break;
      case 23: a |= v; break;
      case 24: a ^= v; break;
      case 25: a <<= (v & 31u); break;
      case 26: a >>= (v & 31u); break;
      case 27: f = (a == v); break;
      case 28: f = (a < v); break;
      default: f = (a > v); break;   // 29
    }
  }
  S.a = a; S.b = b; S.c = c; S.d = d; S.f = f;
  return rc;
}

// ---------------------------------------------------------------------------
// Per-lane model state (lane i <-> component i).
struct Lane {
  // descriptor
  u32 type, a1, a2, a3, a4, a5, step;
  u32* cm; u8* ht;
  u32 cm_mask, ht_mask, limit;
  // dynamic state
  u32 h;          // context hash h[i]
  int p;          // stretch-domain prediction p[i]
  int xj, xk;     // inputs p[j], p[k] latched in the step that computed p
  u32 cxt;        // CM: table index; MIX2: index; MIX: row start; SSE: trained bucket
  u32 v0, v1;     // CM/ICM: pn; ISSE: wt[0], wt[1]; SSE: the two interpolated buckets; MIX2: w
  u32 bh;         // ICM/ISSE: bit history under the current slot
  uint4 row; u32 row_off;   // ICM/ISSE: hash row of the current nibble
  // MATCH
  u32 ma, mb, mpos, mbit, mpred, mc;
};

struct Coder {
  u32 low, high, curr;
  const u8* in; u64 in_pos, in_end;
  u8* out; u64 out_pos, out_cap;
  int status;
};

// Encoder::encode (libzpaq.cpp:2402-2416)
__device__ __forceinline__ void enc_bit(Coder& C, int lane, int y, u32 p16) {
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

// Decoder::decode (libzpaq.cpp:2104-2124)
__device__ __forceinline__ int dec_bit(Coder& C, u32 p16) {
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

struct Model {
  int n, nsteps;
  const ZStep* steps;     // shared memory copy
  u8* arena;
  int c8, hmap4;
};

// Predictor::predict0 (libzpaq.cpp:1854-1951).  Returns squash(p[n-1]) on all lanes.
__device__ __forceinline__ int predict(Lane& L, const Model& M, const SharedTables& S, int lane) {
  const int c8 = M.c8, hmap4 = M.hmap4;
  const bool nib = (c8 == 1) || ((c8 & 0xf0) == 16);
  // ---- phase A: everything that does not depend on another component -----
  switch (L.type) {
    case ZT_CM:
      L.cxt = (L.h ^ (u32)hmap4) & L.cm_mask;
      L.v0 = L.cm[L.cxt];
      L.p = d_stretch(S, L.v0 >> 17);
      break;
    case ZT_ICM:
      if (nib) d_find(L.ht, L.ht_mask, (int)L.a1 + 2, L.h + 16u * (u32)c8, L.row, L.row_off);
      L.bh = row_get(L.row, hmap4 & 15);
      L.v0 = L.cm[L.bh];
      L.p = d_stretch(S, L.v0 >> 8);
      break;
    case ZT_ISSE: {
      if (nib) d_find(L.ht, L.ht_mask, (int)L.a1 + 2, L.h + 16u * (u32)c8, L.row, L.row_off);
      L.bh = row_get(L.row, hmap4 & 15);
      const uint2 w = *reinterpret_cast<const uint2*>(L.cm + L.bh * 2);
      L.v0 = w.x; L.v1 = w.y;
      break;
    }
    case ZT_MATCH:
      if (L.ma == 0) L.p = 0;
      else {
        L.mc = (L.mpred >> (7 - L.mbit)) & 1u;
        L.p = d_stretch(S, (u32)((int)S.dt2k[L.ma] * (1 - 2 * (int)L.mc)) & 32767u);
      }
      break;
    case ZT_MIX2:
      L.cxt = (L.h + (u32)(c8 & (int)L.a5)) & L.cm_mask;
      L.v0 = reinterpret_cast<const u16*>(L.cm)[L.cxt];
      break;
    case ZT_MIX:
      L.cxt = ((L.h + (u32)(c8 & (int)L.a5)) & L.cm_mask) * L.a3;
      break;
    case ZT_SSE:
      L.cxt = (L.h + (u32)c8) * 32u;
      break;
    default: break;   // CONS keeps its constant p; AVG has no state
  }
  // ---- phase B: dependent components, in host-built steps -----------------
  for (int s = 0; s < M.nsteps; ++s) {
    const ZStep st = M.steps[s];
    if (st.kind == ZPQ_STEP_PAR) {
      const int pj = __shfl_sync(FULL, L.p, (int)L.a2 & 31);
      const int pk = __shfl_sync(FULL, L.p, (int)L.a3 & 31);
      const int pj1 = __shfl_sync(FULL, L.p, (int)L.a1 & 31);   // AVG: j is the first argument
      if ((int)L.step == s + 1) {
        switch (L.type) {
          case ZT_ISSE:
            L.xj = pj;
            L.p = clamp2k(((int)L.v0 * pj + (int)L.v1 * 64) >> 16);
            break;
          case ZT_AVG:   // j k wt = a1 a2 a3
            L.p = (pj1 * (int)L.a3 + pj * (256 - (int)L.a3)) >> 8;
            break;
          case ZT_MIX2: {  // sizebits j k rate mask
            L.xj = pj; L.xk = pk;
            const int w = (int)L.v0;
            L.p = (w * pj + (65536 - w) * pk) >> 16;
            break;
          }
          case ZT_SSE: {   // sizebits j start limit
            int pq = pj + 992;
            pq = min(max(pq, 0), 1983);
            const int wt = pq & 63;
            pq >>= 6;
            const u32 c0 = (L.cxt + (u32)pq) & L.cm_mask, c1 = (L.cxt + (u32)pq + 1u) & L.cm_mask;
            const u32 q0 = L.cm[c0], q1 = L.cm[c1];
            L.p = d_stretch(S, (u32)(((int)(q0 >> 10) * (64 - wt) + (int)(q1 >> 10) * wt) >> 13));
            if (wt >> 5) { L.cxt = c1; L.v0 = q1; } else { L.cxt = c0; L.v0 = q0; }
            break;
          }
          default: break;
        }
      }
    } else {
      // MIX component st.kind: lane l multiplies weight l by input p[j0+l]
      const int im = st.kind;
      const u32 base = __shfl_sync(FULL, L.cxt, im);
      const int x = __shfl_sync(FULL, L.p, ((int)st.j0 + lane) & 31);
      int term = 0;
      if (lane < (int)st.m) {
        const int* wt = reinterpret_cast<const int*>(M.arena + (u64)st.cm_off16 * 16) + base;
        term = (wt[lane] >> 8) * x;
      }
      const int sum = __reduce_add_sync(FULL, term);
      if (lane == im) L.p = clamp2k(sum >> 8);
    }
  }
  const int pfinal = __shfl_sync(FULL, L.p, M.n - 1);
  return d_squash(S, pfinal);
}

// Predictor::update0 (libzpaq.cpp:1954-2052), component part only.
__device__ __forceinline__ void update_components(Lane& L, const Model& M, const SharedTables& S, int lane, int y,
                                                  u32 byte_done) {
  const int hmap4 = M.hmap4;
  switch (L.type) {
    case ZT_CM: case ZT_SSE: {   // train(), libzpaq.h:1151-1157
      const u32 pn = L.v0;
      const u32 count = pn & 0x3ffu;
      const int err = y * 32767 - (int)(pn >> 17);
      const u32 prod = (u32)err * (u32)S.dt[count];          // 32-bit wrap, as x86 imul
      L.cm[L.cxt] = pn + (prod & 0xfffffc00u) + (count < L.limit ? 1u : 0u);
      break;
    }
    case ZT_ICM: {
      row_set(L.row, hmap4 & 15, (S.ns[L.bh] >> (8 * y)) & 255u);
      const u32 pn = L.v0;
      L.cm[L.bh] = pn + (u32)(((int)(y * 32767 - (int)(pn >> 8))) >> 2);
      break;
    }
    case ZT_ISSE: {
      const int err = y * 32767 - d_squash(S, L.p);
      const int w0 = clamp512k((int)L.v0 + ((err * L.xj + (1 << 12)) >> 13));
      const int w1 = clamp512k((int)L.v1 + ((err + 16) >> 5));
      *reinterpret_cast<uint2*>(L.cm + L.bh * 2) = make_uint2((u32)w0, (u32)w1);
      row_set(L.row, hmap4 & 15, (S.ns[L.bh] >> (8 * y)) & 255u);
      break;
    }
    case ZT_MIX2: {
      const int err = ((y * 32767 - d_squash(S, L.p)) * (int)L.a4) >> 5;
      int w = (int)L.v0;
      w += (err * (L.xj - L.xk) + (1 << 12)) >> 13;
      w = min(max(w, 0), 65535);
      reinterpret_cast<u16*>(L.cm)[L.cxt] = (u16)w;
      break;
    }
    case ZT_MATCH: {
      if ((int)L.mc != y) L.ma = 0;
      if (++L.mbit == 8) {
        L.mbit = 0;
        L.ht[L.mpos] = (u8)byte_done;                // buf(limit) has shifted in all 8 bits
        L.mpos = (L.mpos + 1) & L.ht_mask;
        u32* idx = L.cm + (L.h & L.cm_mask);         // old h[i]: the VM has not run yet (gotcha 6)
        if (L.ma == 0) {
          L.mb = L.mpos - *idx;
          if (L.mb & L.ht_mask)
            while (L.ma < 255 && L.ht[(L.mpos - L.ma - 1) & L.ht_mask] == L.ht[(L.mpos - L.ma - L.mb - 1) & L.ht_mask])
              ++L.ma;
        }
        else L.ma += (L.ma < 255);
        *idx = L.mpos;
        L.mpred = L.ht[(L.mpos - L.mb) & L.ht_mask];
      }
      break;
    }
    default: break;
  }
  // MIX weight rows: lane l trains weight l
  for (int s = 0; s < M.nsteps; ++s) {
    const ZStep st = M.steps[s];
    if (st.kind == ZPQ_STEP_PAR) continue;
    const int im = st.kind;
    const int e0 = ((y * 32767 - d_squash(S, L.p)) * (int)st.rate) >> 4;
    const int err = __shfl_sync(FULL, e0, im);
    const u32 base = __shfl_sync(FULL, L.cxt, im);
    const int x = __shfl_sync(FULL, L.p, ((int)st.j0 + lane) & 31);
    if (lane < (int)st.m) {
      int* wt = reinterpret_cast<int*>(M.arena + (u64)st.cm_off16 * 16) + base;
      wt[lane] = clamp512k(wt[lane] + ((err * x + (1 << 12)) >> 13));
    }
  }
}

// ---------------------------------------------------------------------------
#define ZPQ_PROG_SMEM 1024
#define ZPQ_STATIC_SMEM (32768 + 8192 + 4096 + 1024 + 512)
#define ZPQ_WARP_SMEM (ZPQ_PROG_SMEM + ZPQ_MAX_STEPS * 8)

template <bool DEC>
__global__ void __launch_bounds__(ZPQ_WARPS * 32, 2) zpq_code_kernel(const ZKernelArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  short* s_stretch = reinterpret_cast<short*>(smem);
  u16* s_squash = reinterpret_cast<u16*>(smem + 32768);
  int* s_dt = reinterpret_cast<int*>(smem + 32768 + 8192);
  u32* s_ns = reinterpret_cast<u32*>(smem + 32768 + 8192 + 4096);
  u16* s_dt2k = reinterpret_cast<u16*>(smem + 32768 + 8192 + 4096 + 1024);
  {
    const ZTables* T = A.tables;
    const uint4* src = reinterpret_cast<const uint4*>(T->stretch_hi);
    uint4* dst = reinterpret_cast<uint4*>(s_stretch);
    for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) dst[i] = src[i];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_squash[i] = T->squash[i];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_dt[i] = T->dt[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_ns[i] = T->ns[i]; s_dt2k[i] = T->dt2k[i]; }
  }
  __syncthreads();
  SharedTables S;
  S.stretch_hi = s_stretch; S.squash = s_squash; S.dt = s_dt; S.ns = s_ns; S.dt2k = s_dt2k;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  u8* s_prog = smem + ZPQ_STATIC_SMEM + warp * ZPQ_WARP_SMEM;
  ZStep* s_steps = reinterpret_cast<ZStep*>(s_prog + ZPQ_PROG_SMEM);

  for (;;) {
    int blk = 0;
    if (lane == 0) blk = (int)atomicAdd(A.counter, 1u);
    blk = __shfl_sync(FULL, blk, 0);
    if (blk >= A.nblk) break;

    const ZBlock B = A.blocks[blk];
    const ZModel* __restrict__ ZM = A.models + B.model;
    Model M;
    M.n = (int)ZM->n; M.nsteps = (int)ZM->nsteps; M.steps = s_steps;
    M.arena = A.arena + B.arena_off;
    M.c8 = 1; M.hmap4 = 1;
    __syncwarp();
    if (lane < M.nsteps) s_steps[lane] = ZM->steps[lane];

    // HCOMP machine
    VmState V;
    V.a = V.b = V.c = V.d = V.f = 0;
    V.m = M.arena + (u64)ZM->m_off16 * 16;
    V.h = reinterpret_cast<u32*>(M.arena + (u64)ZM->h_off16 * 16);
    V.r = reinterpret_cast<u32*>(M.arena + (u64)ZM->r_off16 * 16);
    V.mmask = ZM->m_mask; V.hmask = ZM->h_mask;
    V.len = (int)ZM->prog_len;
    {
      const u8* img = A.progs + ZM->prog_off;
      const int total = ZPQ_GUARD_LO + V.len + ZPQ_GUARD_HI;
      if (total <= ZPQ_PROG_SMEM) {
        for (int i = lane; i < total; i += 32) s_prog[i] = img[i];
        V.code = s_prog + ZPQ_GUARD_LO;
      } else V.code = img + ZPQ_GUARD_LO;
    }
    __syncwarp();

    // lane <-> component
    Lane L;
    {
      ZComp cd;
      if (lane < M.n) cd = ZM->comp[lane];
      else { cd.type = ZT_NONE; cd.a1 = cd.a2 = cd.a3 = cd.a4 = cd.a5 = 0; cd.step = 0; cd.cm_off16 = cd.ht_off16 = 0;
             cd.cm_mask = cd.ht_mask = cd.limit = 0; }
      L.type = cd.type; L.a1 = cd.a1; L.a2 = cd.a2; L.a3 = cd.a3; L.a4 = cd.a4; L.a5 = cd.a5; L.step = cd.step;
      L.cm = reinterpret_cast<u32*>(M.arena + (u64)cd.cm_off16 * 16);
      L.ht = M.arena + (u64)cd.ht_off16 * 16;
      L.cm_mask = cd.cm_mask; L.ht_mask = cd.ht_mask; L.limit = cd.limit;
      L.h = 0;
      L.p = (cd.type == ZT_CONS) ? ((int)cd.a1 - 128) * 4 : 0;    // libzpaq.cpp:1775
      L.xj = L.xk = 0; L.cxt = 0; L.v0 = L.v1 = 0; L.bh = 0;
      L.row = make_uint4(0, 0, 0, 0); L.row_off = 0;
      L.ma = L.mb = L.mpos = L.mbit = L.mc = 0; L.mpred = 0;
    }

    Coder C;
    C.low = 1; C.high = 0xffffffffu; C.curr = 0; C.status = ST_OK;
    C.in = A.in + B.in_off; C.in_pos = 0; C.in_end = 0;
    C.out = A.out + B.out_off; C.out_pos = 0; C.out_cap = B.out_cap;

    int status = ST_OK;
    for (u32 sg = 0; sg < B.nseg && status == ST_OK; ++sg) {
      const u64 seg_n = A.seg_len[B.seg_first + sg];
      const u64 out_start = C.out_pos;
      C.in_end = C.in_pos + seg_n;
      if constexpr (DEC) {
        // Decoder::decompress: curr==0 at segment start -> load 4 bytes (libzpaq.cpp:2129-2132)
        C.curr = 0;
        for (int i = 0; i < 4; ++i) {
          if (C.in_pos >= C.in_end) { status = ST_EOF; break; }
          C.curr = C.curr << 8 | (u32)C.in[C.in_pos++];
        }
      }
      for (;;) {
        if (status != ST_OK) break;
        int cbyte = 0;
        if constexpr (!DEC) {
          if (C.in_pos >= C.in_end) break;
          cbyte = C.in[C.in_pos++];
          enc_bit(C, lane, 0, 0);
        } else {
          const int eos = dec_bit(C, 0);
          if (C.status != ST_OK) { status = C.status; break; }
          if (eos) { if (C.curr != 0) status = ST_EOS; break; }
        }
        for (int bit = 7; bit >= 0; --bit) {
          const u32 p16 = (u32)predict(L, M, S, lane) * 2u + 1u;
          int y;
          if constexpr (!DEC) { y = (cbyte >> bit) & 1; enc_bit(C, lane, y, p16); }
          else {
            y = dec_bit(C, p16);
            if (C.status != ST_OK) { status = C.status; break; }
          }
          const int c8n = M.c8 * 2 + y;
          update_components(L, M, S, lane, y, (u32)c8n & 255u);
          // tail of Predictor::update (libzpaq.cpp:2054-2066)
          if (c8n >= 256 || (c8n >= 16 && c8n < 32)) {
            if (L.type == ZT_ICM || L.type == ZT_ISSE) *reinterpret_cast<uint4*>(L.ht + L.row_off) = L.row;
          }
          if (c8n >= 256) {
            // lane 0 interprets the program (every lane would issue the same loads and stores; running it once does not
            // lean on lock-step execution), __syncwarp orders its H[] stores before the reads below
            int rc = ST_OK;
            if (lane == 0) rc = vm_run(V, (u32)(c8n - 256));
            __syncwarp();
            rc = __shfl_sync(FULL, rc, 0);
            if (rc != ST_OK) { status = rc; break; }
            M.hmap4 = 1; M.c8 = 1;
            L.h = V.h[(u32)lane & V.hmask];
            if (DEC) cbyte = c8n - 256;
          }
          else {
            M.c8 = c8n;
            if (c8n >= 16 && c8n < 32) M.hmap4 = (M.hmap4 & 0xf) << 5 | y << 4 | 1;
            else M.hmap4 = (M.hmap4 & 0x1f0) | (((M.hmap4 & 0xf) * 2 + y) & 0xf);
          }
        }
        if (DEC && status == ST_OK) {
          if (C.out_pos < C.out_cap) { if (lane == 0) C.out[C.out_pos] = (u8)cbyte; }
          else status = ST_FULL;
          ++C.out_pos;
        }
      }
      if (!DEC && status == ST_OK) {
        enc_bit(C, lane, 1, 0);     // EOS: flushes 4 bytes (libzpaq.cpp:2422)
        if (C.status != ST_OK) status = C.status;
      }
      if (DEC && status == ST_OK && C.in_pos != C.in_end) status = ST_CORRUPT;  // segment must end with its 4 zeros
      if (lane == 0) A.seg_out[B.seg_first + sg] = C.out_pos - out_start;
    }
    if (!DEC && status == ST_OK && C.status != ST_OK) status = C.status;
    if (lane == 0) { A.status[blk] = status; A.out_len[blk] = C.out_pos; }
  }
}

// ---------------------------------------------------------------------------
// WIDE models: 33 .. 255 components (Predictor holds comp[256], libzpaq.h:1131-1137).  The lane = component mapping of
// the kernels above stops at 32, so these blocks are coded by ONE THREAD each, walking the components in order exactly as
// Predictor::predict0 / update0 do (libzpaq.cpp:1854-2066); every table, the hash rows and the per-component registers
// (Component::cxt, a, b, c, limit) live in the block's arena, the static tables are read through the read-only path.  It is
// a completeness tier (makeConfig stops at ~30 components; hand-written configs can go further), not a fast one: lanes
// of a warp are different blocks, possibly of different models.
struct WideState { u32 cxt, a, b, c, limit, row, pad0, pad1; };    // 8 words per component in the arena (ZModel.wst_off16)

__device__ __forceinline__ int w_stretch(const ZTables* T, u32 x) {
  return x >= 16384u ? (int)__ldg(&T->stretch_hi[x - 16384u]) : -(int)__ldg(&T->stretch_hi[16383u - x]);
}
__device__ __forceinline__ int w_squash(const ZTables* T, int x) { return (int)__ldg(&T->squash[x + 2048]); }

// Predictor::find on the table itself (libzpaq.cpp:2072-2088): returns the row's byte offset
__device__ __forceinline__ u32 w_find(u8* ht, u32 ht_mask, int shift, u32 cxt) {
  const u32 chk = (cxt >> shift) & 255u;
  const u32 h0 = (cxt * 16u) & ht_mask, h1 = h0 ^ 16u, h2 = h0 ^ 32u;
  if (ht[h0] == chk) return h0;
  if (ht[h1] == chk) return h1;
  if (ht[h2] == chk) return h2;
  u32 r;
  if (ht[h0 + 1] <= ht[h1 + 1] && ht[h0 + 1] <= ht[h2 + 1]) r = h0;
  else if (ht[h1 + 1] < ht[h2 + 1]) r = h1;
  else r = h2;
  *reinterpret_cast<uint4*>(ht + r) = make_uint4(chk, 0, 0, 0);
  return r;
}

template <bool DEC>
__global__ void __launch_bounds__(32) zpq_wide_kernel(const ZKernelArgs A) {
  const int blk = blockIdx.x * blockDim.x + threadIdx.x;
  if (blk >= A.nblk) return;
  const ZBlock B = A.blocks[blk];
  const ZModel* __restrict__ ZM = A.models + B.model;
  const ZTables* __restrict__ T = A.tables;
  u8* const arena = A.arena + B.arena_off;
  const int n = (int)ZM->n;
  const ZComp* __restrict__ comp = reinterpret_cast<const ZComp*>(A.progs + ZM->prog_off + ZM->wide_comp_off);
  WideState* const ws = reinterpret_cast<WideState*>(arena + (u64)ZM->wst_off16 * 16);
  int p[256];
  for (int i = 0; i < n; ++i) p[i] = comp[i].type == ZT_CONS ? ((int)comp[i].a1 - 128) * 4 : 0;      // libzpaq.cpp:1775

  VmState V;
  V.a = V.b = V.c = V.d = V.f = 0;
  V.m = arena + (u64)ZM->m_off16 * 16;
  V.h = reinterpret_cast<u32*>(arena + (u64)ZM->h_off16 * 16);
  V.r = reinterpret_cast<u32*>(arena + (u64)ZM->r_off16 * 16);
  V.mmask = ZM->m_mask; V.hmask = ZM->h_mask;
  V.len = (int)ZM->prog_len;
  V.code = A.progs + ZM->prog_off + ZPQ_GUARD_LO;

  Coder C;
  C.low = 1; C.high = 0xffffffffu; C.curr = 0; C.status = ST_OK;
  C.in = A.in + B.in_off; C.in_pos = 0; C.in_end = 0;
  C.out = A.out + B.out_off; C.out_pos = 0; C.out_cap = B.out_cap;
  int c8 = 1, hmap4 = 1;
  int status = ST_OK;

  for (u32 sg = 0; sg < B.nseg && status == ST_OK; ++sg) {
    const u64 seg_n = A.seg_len[B.seg_first + sg];
    const u64 out_start = C.out_pos;
    C.in_end = C.in_pos + seg_n;
    if constexpr (DEC) {
      C.curr = 0;
      for (int i = 0; i < 4; ++i) {
        if (C.in_pos >= C.in_end) { status = ST_EOF; break; }
        C.curr = C.curr << 8 | (u32)C.in[C.in_pos++];
      }
    }
    for (;;) {
      if (status != ST_OK) break;
      int cbyte = 0;
      if constexpr (!DEC) {
        if (C.in_pos >= C.in_end) break;
        cbyte = C.in[C.in_pos++];
        enc_bit(C, 0, 0, 0);
      } else {
        const int eos = dec_bit(C, 0);
        if (C.status != ST_OK) { status = C.status; break; }
        if (eos) { if (C.curr != 0) status = ST_EOS; break; }
      }
      for (int bit = 7; bit >= 0; --bit) {
        // ---------------- predict0 ----------------
        const bool nib = (c8 == 1) || ((c8 & 0xf0) == 16);
        for (int i = 0; i < n; ++i) {
          const ZComp cd = comp[i];
          WideState& W = ws[i];
          u32* cm = reinterpret_cast<u32*>(arena + (u64)cd.cm_off16 * 16);
          u8* ht = arena + (u64)cd.ht_off16 * 16;
          const u32 h = V.h[(u32)i & V.hmask];               // H(i) wraps modulo 2^hh (libzpaq.h:1040)
          switch (cd.type) {
            case ZT_CM:
              W.cxt = (h ^ (u32)hmap4) & cd.cm_mask;
              p[i] = w_stretch(T, cm[W.cxt] >> 17);
              break;
            case ZT_ICM:
              if (nib) W.row = w_find(ht, cd.ht_mask, (int)cd.a1 + 2, h + 16u * (u32)c8);
              W.cxt = ht[W.row + (u32)(hmap4 & 15)];
              p[i] = w_stretch(T, cm[W.cxt] >> 8);
              break;
            case ZT_MATCH:
              if (W.a == 0) p[i] = 0;
              else {
                W.c = ((u32)ht[(W.limit - W.b) & cd.ht_mask] >> (7 - W.cxt)) & 1u;
                p[i] = w_stretch(T, (u32)((int)__ldg(&T->dt2k[W.a]) * (1 - 2 * (int)W.c)) & 32767u);
              }
              break;
            case ZT_AVG:
              p[i] = (p[cd.a1] * (int)cd.a3 + p[cd.a2] * (256 - (int)cd.a3)) >> 8;
              break;
            case ZT_MIX2: {
              W.cxt = (h + (u32)(c8 & (int)cd.a5)) & cd.cm_mask;
              const int w = (int)reinterpret_cast<const u16*>(cm)[W.cxt];
              p[i] = (w * p[cd.a2] + (65536 - w) * p[cd.a3]) >> 16;
              break;
            }
            case ZT_MIX: {
              const int m = (int)cd.a3;
              W.cxt = ((h + (u32)(c8 & (int)cd.a5)) & cd.cm_mask) * (u32)m;
              const int* wt = reinterpret_cast<const int*>(cm) + W.cxt;
              int sum = 0;
              for (int j = 0; j < m; ++j) sum += (wt[j] >> 8) * p[(int)cd.a2 + j];
              p[i] = clamp2k(sum >> 8);
              break;
            }
            case ZT_ISSE: {
              if (nib) W.row = w_find(ht, cd.ht_mask, (int)cd.a1 + 2, h + 16u * (u32)c8);
              W.cxt = ht[W.row + (u32)(hmap4 & 15)];
              const int* wt = reinterpret_cast<const int*>(cm) + W.cxt * 2;
              p[i] = clamp2k((wt[0] * p[cd.a2] + wt[1] * 64) >> 16);
              break;
            }
            case ZT_SSE: {
              W.cxt = (h + (u32)c8) * 32u;
              int pq = min(max(p[cd.a2] + 992, 0), 1983);
              const int wt = pq & 63;
              pq >>= 6;
              W.cxt += (u32)pq;
              p[i] = w_stretch(T, (u32)(((int)(cm[W.cxt & cd.cm_mask] >> 10) * (64 - wt) + (int)(cm[(W.cxt + 1u) & cd.cm_mask] >> 10) * wt) >> 13));
              W.cxt += (u32)(wt >> 5);
              break;
            }
            default: break;   // CONS
          }
        }
        const u32 p16 = (u32)w_squash(T, p[n - 1]) * 2u + 1u;
        int y;
        if constexpr (!DEC) { y = (cbyte >> bit) & 1; enc_bit(C, 0, y, p16); }
        else {
          y = dec_bit(C, p16);
          if (C.status != ST_OK) { status = C.status; break; }
        }
        // ---------------- update0 ----------------
        for (int i = 0; i < n; ++i) {
          const ZComp cd = comp[i];
          WideState& W = ws[i];
          u32* cm = reinterpret_cast<u32*>(arena + (u64)cd.cm_off16 * 16);
          u8* ht = arena + (u64)cd.ht_off16 * 16;
          switch (cd.type) {
            case ZT_CM: case ZT_SSE: {     // train(), libzpaq.h:1151-1157
              u32& pn = cm[W.cxt & cd.cm_mask];
              const u32 count = pn & 0x3ffu;
              const int err = y * 32767 - (int)(pn >> 17);
              pn += (((u32)err * (u32)__ldg(&T->dt[count])) & 0xfffffc00u) + (count < cd.limit ? 1u : 0u);
              break;
            }
            case ZT_ICM: {
              ht[W.row + (u32)(hmap4 & 15)] = (u8)((__ldg(&T->ns[W.cxt]) >> (8 * y)) & 255u);
              u32& pn = cm[W.cxt];
              pn += (u32)(((int)(y * 32767 - (int)(pn >> 8))) >> 2);
              break;
            }
            case ZT_MATCH: {
              // libzpaq.cpp:1990-2008; h[i] is still the previous byte's hash here (gotcha 6)
              if ((int)W.c != y) W.a = 0;
              u8& cur = ht[W.limit & cd.ht_mask];
              cur = (u8)(cur * 2 + y);
              if (++W.cxt == 8) {
                W.cxt = 0;
                W.limit = (W.limit + 1) & cd.ht_mask;
                u32& idx = cm[V.h[(u32)i & V.hmask] & cd.cm_mask];
                if (W.a == 0) {
                  W.b = W.limit - idx;
                  if (W.b & cd.ht_mask)
                    while (W.a < 255 && ht[(W.limit - W.a - 1) & cd.ht_mask] == ht[(W.limit - W.a - W.b - 1) & cd.ht_mask]) ++W.a;
                }
                else W.a += (W.a < 255);
                idx = W.limit;
              }
              break;
            }
            case ZT_MIX2: {
              const int err = ((y * 32767 - w_squash(T, p[i])) * (int)cd.a4) >> 5;
              int w = (int)reinterpret_cast<u16*>(cm)[W.cxt];
              w += (err * (p[cd.a2] - p[cd.a3]) + (1 << 12)) >> 13;
              w = min(max(w, 0), 65535);
              reinterpret_cast<u16*>(cm)[W.cxt] = (u16)w;
              break;
            }
            case ZT_MIX: {
              const int m = (int)cd.a3;
              const int err = ((y * 32767 - w_squash(T, p[i])) * (int)cd.a4) >> 4;
              int* wt = reinterpret_cast<int*>(cm) + W.cxt;
              for (int j = 0; j < m; ++j) wt[j] = clamp512k(wt[j] + ((err * p[(int)cd.a2 + j] + (1 << 12)) >> 13));
              break;
            }
            case ZT_ISSE: {
              const int err = y * 32767 - w_squash(T, p[i]);
              int* wt = reinterpret_cast<int*>(cm) + W.cxt * 2;
              wt[0] = clamp512k(wt[0] + ((err * p[cd.a2] + (1 << 12)) >> 13));
              wt[1] = clamp512k(wt[1] + ((err + 16) >> 5));
              ht[W.row + (u32)(hmap4 & 15)] = (u8)((__ldg(&T->ns[W.cxt]) >> (8 * y)) & 255u);
              break;
            }
            default: break;
          }
        }
        const int c8n = c8 * 2 + y;
        if (c8n >= 256) {
          const int rc = vm_run(V, (u32)(c8n - 256));
          if (rc != ST_OK) { status = rc; break; }
          hmap4 = 1; c8 = 1;
          if (DEC) cbyte = c8n - 256;
        } else {
          c8 = c8n;
          if (c8n >= 16 && c8n < 32) hmap4 = (hmap4 & 0xf) << 5 | y << 4 | 1;
          else hmap4 = (hmap4 & 0x1f0) | (((hmap4 & 0xf) * 2 + y) & 0xf);
        }
      }
      if (DEC && status == ST_OK) {
        if (C.out_pos < C.out_cap) C.out[C.out_pos] = (u8)cbyte;
        else status = ST_FULL;
        ++C.out_pos;
      }
    }
    if (!DEC && status == ST_OK) {
      enc_bit(C, 0, 1, 0);
      if (C.status != ST_OK) status = C.status;
    }
    if (DEC && status == ST_OK && C.in_pos != C.in_end) status = ST_CORRUPT;
    A.seg_out[B.seg_first + sg] = C.out_pos - out_start;
  }
  if (!DEC && status == ST_OK && C.status != ST_OK) status = C.status;
  A.status[blk] = status; A.out_len[blk] = C.out_pos;
}

extern "C" cudaError_t zpq_launch_wide(const ZKernelArgs* a, int nblk, int decode, cudaStream_t st) {
  if (nblk <= 0) return cudaSuccess;
  const int ctas = (nblk + 31) / 32;
  if (decode) zpq_wide_kernel<true><<<ctas, 32, 0, st>>>(*a);
  else zpq_wide_kernel<false><<<ctas, 32, 0, st>>>(*a);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
extern "C" size_t zpq_code_kernel_smem(void) { return ZPQ_STATIC_SMEM + ZPQ_WARPS * ZPQ_WARP_SMEM; }

extern "C" cudaError_t zpq_launch_init(const ZKernelArgs* a, int nblk, cudaStream_t st) {
  if (nblk <= 0) return cudaSuccess;
  dim3 grid(64, (unsigned)nblk);
  zpq_init_kernel<<<grid, 256, 0, st>>>(*a);
  return cudaGetLastError();
}

extern "C" cudaError_t zpq_launch_code(const ZKernelArgs* a, int nblk, int decode, int sm_count, cudaStream_t st) {
  if (nblk <= 0) return cudaSuccess;
  const size_t smem = zpq_code_kernel_smem();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(zpq_code_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(zpq_code_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int ctas = (nblk + ZPQ_WARPS - 1) / ZPQ_WARPS;
  const int max_ctas = sm_count * 2;            // persistent: 2 CTAs (8 warps = 8 blocks) per SM
  if (ctas > max_ctas) ctas = max_ctas;
  if (decode) zpq_code_kernel<true><<<ctas, ZPQ_WARPS * 32, smem, st>>>(*a);
  else zpq_code_kernel<false><<<ctas, ZPQ_WARPS * 32, smem, st>>>(*a);
  return cudaGetLastError();
}
