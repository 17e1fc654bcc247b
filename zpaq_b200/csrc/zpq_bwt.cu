// zpq_bwt.cu -- the forward transforms in front of the model, on the device, sm_100a.
//
//   zpq_e8e9_forward   libzpaq's e8e9() (libzpaq.cpp:6450-6459): x86 CALL / JMP operands made absolute.
//   zpq_bwt_forward    LZBuffer level 3 (libzpaq.cpp:6630-6658): suffix array of the block (divsufsort there), then
//                      last byte | BWT with the end-of-string slot coded as 255 | index of that slot, 4 bytes LSB first.
//
// The suffix array of a text is unique, so any correct suffix sorter gives the reference's bytes.  This one is prefix
// doubling written for a batch: all blocks of a call are sorted together in ONE index space (a block is a range of
// slots), every round is a handful of streaming kernels over the still-unsorted suffixes only:
//
//   round 0  key = (slot of the block's first suffix) << 32 | first three symbols (9 bits each: byte + 1, 0 = past the
//            end, so a suffix that is a prefix of another one sorts first, as in divsufsort)
//   round k  key = (head slot of the suffix's group) << 32 | rank of the suffix h positions further on (0 = past the end),
//            h = 3, 6, 12, ...
//   sort     LSD radix sort of (key, suffix) pairs, 8 bits per pass, only over the bits that can differ; a pass is
//            histogram per 2 048-key tile -> exclusive scan -> stable scatter (ranks inside a warp from match.any, inside
//            the tile from per-warp digit counters in shared memory)
//   regroup  a suffix keeps the position range of its group (sorting only permutes inside groups), so its slot is known
//            without a search; equal keys form the new groups (head slot by a max-scan); groups of one are final and
//            leave the active list (sum-scan + compaction).  The round costs O(active suffixes).
//
// All of it is HBM-bound integer work: the roofline of a pass is (8 + 4 bytes) x 3 per active suffix.  Bit-exactness is
// checked against blocks written by the unmodified reference (tests/test_gpu_transforms.py, emulator pre-flight on CPU).
#include <cuda_runtime.h>
#include <stdint.h>
#include "zpq_kernels.h"

namespace {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
#define FULLM 0xffffffffu
#define TILE 2048u                     // keys per CTA in the sort and scan kernels (256 threads x 8)

struct BwtBlocks {                     // per call, in device memory
  const u8* const* in;                 // payload of block b
  u8* const* out;                      // len + 5 bytes
  const u32* off;                      // nb + 1 slot offsets
  int nb;
};

__device__ __forceinline__ int block_of(const u32* __restrict__ off, int nb, u32 i) {
  int lo = 0, hi = nb - 1;             // largest b with off[b] <= i and off[b + 1] > i (empty blocks are skipped by the >)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(off + mid) <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ------------------------------------------------------------------ E8E9
// Position i (descending) is rewritten when in[i] is E8 / E9 and the byte four further on is 00 / FF; the test of the
// opcode byte always sees original data (only bytes ABOVE a candidate are written), the other reads may see operands
// rewritten by candidates up to three positions above.  Three consecutive non-opcode positions therefore cut every
// dependency chain: a thread owns the positions between the first cut at or above its slice's two ends.
__device__ __forceinline__ bool e8_op(u8 c) { return (c & 254) == 0xe8; }

__device__ u64 e8_cut(const u8* __restrict__ in, u64 x, u64 top) {
  for (u64 q = x; q < top; ++q) {
    const bool o0 = e8_op(in[q]), o1 = q + 1 < top && e8_op(in[q + 1]), o2 = q + 2 < top && e8_op(in[q + 2]);
    if (!o0 && !o1 && !o2) return q;
  }
  return top;
}

__global__ void __launch_bounds__(128) zpq_e8e9_kernel(const u8* const* in_p, u8* const* out_p, const u64* len, int slice) {
  const int b = blockIdx.y;
  const u64 n = len[b];
  if (n < 5) return;
  const u8* __restrict__ in = in_p[b];
  u8* out = out_p[b];
  const u64 top = n - 4;                              // first position that can never be rewritten
  const u64 nsl = (top + slice - 1) / slice;
  for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < nsl; t += (u64)gridDim.x * blockDim.x) {
    const u64 a = t * slice, e = a + slice < top ? a + slice : top;
    const u64 hi = e8_cut(in, e, top), lo = t == 0 ? 0 : e8_cut(in, a, top);
    for (u64 p = hi; p-- > lo;) {
      if (e8_op(in[p]) && (((u32)out[p + 4] + 1u) & 254u) == 0u) {
        const u32 v = ((u32)out[p + 1] | (u32)out[p + 2] << 8 | (u32)out[p + 3] << 16) + (u32)p;
        out[p + 1] = (u8)v; out[p + 2] = (u8)(v >> 8); out[p + 3] = (u8)(v >> 16);
      }
    }
  }
}

// ------------------------------------------------------------------ scans over u32 (sum or max), 2 048 elements per CTA
template <bool MAXOP> __device__ __forceinline__ u32 comb(u32 a, u32 b) { return MAXOP ? (a > b ? a : b) : a + b; }

template <bool MAXOP>
__global__ void __launch_bounds__(256) zpq_scan_reduce(const u32* __restrict__ in, u32 n, u32* __restrict__ partial) {
  __shared__ u32 ws[8];
  const u32 base = blockIdx.x * TILE;
  u32 acc = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { const u32 idx = base + k * 256u + threadIdx.x; if (idx < n) acc = comb<MAXOP>(acc, in[idx]); }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc = comb<MAXOP>(acc, __shfl_xor_sync(FULLM, acc, o));
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { u32 t = 0; for (int w = 0; w < 8; ++w) t = comb<MAXOP>(t, ws[w]); partial[blockIdx.x] = t; }
}

// exclusive scan of the per-tile totals, one CTA, in place
template <bool MAXOP>
__global__ void __launch_bounds__(1024) zpq_scan_partials(u32* partial, u32 m) {
  __shared__ u32 ws[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 carry = 0;
  for (u32 base = 0; base < m; base += 1024u) {
    const u32 idx = base + threadIdx.x;
    const u32 v = idx < m ? partial[idx] : 0u;
    u32 x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULLM, x, o); if (lane >= o) x = comb<MAXOP>(x, y); }
    u32 xe = __shfl_up_sync(FULLM, x, 1); if (lane == 0) xe = 0;
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) {
      u32 t = ws[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULLM, t, o); if (lane >= o) t = comb<MAXOP>(t, y); }
      ws[lane] = t;
    }
    __syncthreads();
    const u32 pre = warp ? ws[warp - 1] : 0u;
    if (idx < m) partial[idx] = comb<MAXOP>(comb<MAXOP>(carry, pre), xe);
    const u32 total = ws[31];
    __syncthreads();
    carry = comb<MAXOP>(carry, total);
  }
}

// per tile: local scan + the tile's carry; a thread owns eight consecutive elements; in == out is allowed
template <bool MAXOP, bool INCL>
__global__ void __launch_bounds__(256) zpq_scan_apply(const u32* in, u32 n, const u32* __restrict__ partial, u32* out) {
  __shared__ u32 ws[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 base = blockIdx.x * TILE + threadIdx.x * 8u;
  u32 v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = base + k < n ? in[base + k] : 0u;
  u32 tot = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot = comb<MAXOP>(tot, v[k]);
  u32 x = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(FULLM, x, o); if (lane >= o) x = comb<MAXOP>(x, y); }
  u32 xe = __shfl_up_sync(FULLM, x, 1); if (lane == 0) xe = 0;
  if (lane == 31) ws[warp] = x;
  __syncthreads();
  u32 pre = partial[blockIdx.x];
  for (int w = 0; w < warp; ++w) pre = comb<MAXOP>(pre, ws[w]);
  u32 run = comb<MAXOP>(pre, xe);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const u32 incl = comb<MAXOP>(run, v[k]);
    if (base + k < n) out[base + k] = INCL ? incl : run;
    run = incl;
  }
}

// ------------------------------------------------------------------ LSD radix sort pass (8 bits) over (u64 key, u32 value)
__global__ void __launch_bounds__(256) zpq_radix_hist(const u64* __restrict__ key, u32 n, int shift, u32* __restrict__ counts, u32 ntiles) {
  __shared__ u32 cnt[256];
  cnt[threadIdx.x] = 0;
  __syncthreads();
  const u32 base = blockIdx.x * TILE;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const u32 idx = base + k * 256u + threadIdx.x;
    if (idx < n) atomicAdd(&cnt[(u32)(key[idx] >> shift) & 255u], 1u);
  }
  __syncthreads();
  counts[threadIdx.x * ntiles + blockIdx.x] = cnt[threadIdx.x];           // digit-major: the scan yields global offsets
}

__global__ void __launch_bounds__(256) zpq_radix_scatter(const u64* __restrict__ kin, const u32* __restrict__ vin, u64* __restrict__ kout,
                                                         u32* __restrict__ vout, u32 n, int shift, const u32* __restrict__ offs, u32 ntiles) {
  __shared__ u32 wc[8][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 2048; i += 256) (&wc[0][0])[i] = 0;
  __syncthreads();
  u64 key[8]; u32 val[8], lr[8];
  const u32 base = blockIdx.x * TILE + (u32)w * 256u;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const u32 idx = base + r * 32u + lane;
    const bool valid = idx < n;
    key[r] = 0; val[r] = 0;
    if (valid) { key[r] = kin[idx]; val[r] = vin[idx]; }
    const u32 d = valid ? ((u32)(key[r] >> shift) & 255u) : 256u;
    const u32 peers = __match_any_sync(FULLM, d);
    const int leader = __ffs((int)peers) - 1;
    u32 b = 0;
    if (valid && lane == leader) { b = wc[w][d]; wc[w][d] = b + (u32)__popc(peers); }
    b = __shfl_sync(FULLM, b, leader);
    lr[r] = b + (u32)__popc(peers & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  {
    const int d = threadIdx.x;
    u32 run = offs[(u32)d * ntiles + blockIdx.x];
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) { const u32 t = wc[w2][d]; wc[w2][d] = run; run += t; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const u32 idx = base + r * 32u + lane;
    if (idx < n) {
      const u32 pos = wc[w][(u32)(key[r] >> shift) & 255u] + lr[r];
      kout[pos] = key[r]; vout[pos] = val[r];
    }
  }
}

// ------------------------------------------------------------------ prefix doubling
// round 0: every suffix is active, its group is its block
__global__ void __launch_bounds__(256) zpq_bwt_setup(BwtBlocks B, u32 n, u8* __restrict__ txt, u16* __restrict__ bid,
                                                     u64* __restrict__ key, u32* __restrict__ val, u32* __restrict__ slot) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int b = block_of(B.off, B.nb, i);
    const u32 o = __ldg(B.off + b), e = __ldg(B.off + b + 1);
    const u8* __restrict__ p = B.in[b];
    const u32 j = i - o, len = e - o;
    const u32 c0 = (u32)p[j] + 1u, c1 = j + 1 < len ? (u32)p[j + 1] + 1u : 0u, c2 = j + 2 < len ? (u32)p[j + 2] + 1u : 0u;
    txt[i] = p[j]; bid[i] = (u16)b;
    key[i] = (u64)o << 32 | (u64)(c0 << 18 | c1 << 9 | c2);
    val[i] = i; slot[i] = i;
  }
}

// low half of the keys of round k >= 1: the rank of the suffix h positions further on
__global__ void __launch_bounds__(256) zpq_bwt_keylow(u64* __restrict__ key, const u32* __restrict__ val, u32 m, u32 h, const u32* __restrict__ rank_of,
                                                      const u16* __restrict__ bid, const u32* __restrict__ off) {
  for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
    const u32 i = val[k];
    const u32 end = __ldg(off + (u32)bid[i] + 1u);
    const u32 low = (u64)i + h < (u64)end ? rank_of[i + h] : 0u;
    key[k] = (key[k] & 0xffffffff00000000ull) | (u64)low;
  }
}

// after the sort: a position whose key differs from its left neighbour opens a new group at its own slot
__global__ void __launch_bounds__(256) zpq_bwt_heads(const u64* __restrict__ key, const u32* __restrict__ slot, u32 m, u32* __restrict__ head) {
  for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x)
    head[k] = (k == 0 || key[k] != key[k - 1]) ? slot[k] : 0u;
}

// suffix into its slot, rank = head slot + 1; still active = its group has a second member
__global__ void __launch_bounds__(256) zpq_bwt_update(const u64* __restrict__ key, const u32* __restrict__ val, const u32* __restrict__ slot,
                                                      const u32* __restrict__ headscan, u32 m, u32* __restrict__ sa, u32* __restrict__ rank_of,
                                                      u32* __restrict__ active) {
  for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
    const u32 i = val[k];
    sa[slot[k]] = i;
    rank_of[i] = headscan[k] + 1u;
    const bool first = k == 0 || key[k] != key[k - 1];
    const bool next_first = k + 1 == m || key[k + 1] != key[k];
    active[k] = (first && next_first) ? 0u : 1u;
  }
}

__global__ void __launch_bounds__(256) zpq_bwt_compact(const u64* __restrict__ key, const u32* __restrict__ val, const u32* __restrict__ slot,
                                                       const u32* __restrict__ headscan, const u32* __restrict__ pos, u32 m,
                                                       u64* __restrict__ key_next, u32* __restrict__ val_next, u32* __restrict__ slot_next, u32* __restrict__ m_next) {
  for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
    const bool first = k == 0 || key[k] != key[k - 1];
    const bool next_first = k + 1 == m || key[k + 1] != key[k];
    const bool act = !(first && next_first);
    const u32 p = pos[k];
    if (act) { key_next[p] = (u64)headscan[k] << 32; val_next[p] = val[k]; slot_next[p] = slot[k]; }
    if (k + 1 == m) *m_next = p + (act ? 1u : 0u);
  }
}

// LZBuffer::fill level 3 (libzpaq.cpp:6649-6658)
__global__ void __launch_bounds__(256) zpq_bwt_emit(BwtBlocks B, const u8* __restrict__ txt, const u32* __restrict__ sa) {
  const int b = blockIdx.y;
  const u32 o = B.off[b], len = B.off[b + 1] - o;
  u8* out = B.out[b];
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j <= len; j += gridDim.x * blockDim.x) {
    if (j == 0) {
      out[0] = len ? txt[o + len - 1] : (u8)255;
      if (len == 0) { out[1] = 0; out[2] = 0; out[3] = 0; out[4] = 0; }
      continue;
    }
    const u32 s = sa[o + j - 1] - o;
    if (s == 0) {
      out[j] = 255;
      out[len + 1] = (u8)j; out[len + 2] = (u8)(j >> 8); out[len + 3] = (u8)(j >> 16); out[len + 4] = (u8)(j >> 24);
    }
    else out[j] = txt[o + s - 1];
  }
}

// ------------------------------------------------------------------ inverse BWT (what the block's PCOMP program computes)
// Rows = the sorted suffixes of T$ (row 0 = "$"); L[row] = the byte in front of that suffix, the row `idx` holds the
// end-of-string mark.  A stable sort of the rows by L (the mark first) is the LF mapping: the row at sorted position p is
// the one whose suffix, extended by one byte to the left, is row p -- read the other way round, psi(p) = the row that follows p
// in text order.  The text is the list idx -> psi(idx) -> ... -> 0; list ranking by pointer jumping gives every row its
// distance to row 0 and so its place in the output.
struct IbwtBlocks {
  const u8* const* in;                 // BWT bytes of block b: rows 0 .. n, then the index (4 bytes)
  u8* const* out;                      // n bytes
  const u32* off;                      // nb + 1 row offsets (block b has n_b + 1 rows)
  const u32* idx;                      // row of the end-of-string mark, per block
  int nb;
};

__global__ void __launch_bounds__(256) zpq_ibwt_setup(IbwtBlocks B, u32 rows, u64* __restrict__ key, u32* __restrict__ val) {
  for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    const int b = block_of(B.off, B.nb, r);
    const u32 o = __ldg(B.off + b), j = r - o;
    const u32 sym = j == __ldg(B.idx + b) ? 0u : (u32)B.in[b][j] + 1u;
    key[r] = (u64)o << 32 | (u64)sym;
    val[r] = r;
  }
}

// psi from the sorted rows; row 0 of every block ends its list
__global__ void __launch_bounds__(256) zpq_ibwt_links(IbwtBlocks B, u32 rows, const u64* __restrict__ key, const u32* __restrict__ val,
                                                      u32* __restrict__ nxt, u32* __restrict__ dst) {
  for (u32 p = blockIdx.x * blockDim.x + threadIdx.x; p < rows; p += gridDim.x * blockDim.x) {
    const u32 o = (u32)(key[p] >> 32);
    const bool last = p == o;
    nxt[p] = last ? p : val[p];
    dst[p] = last ? 0u : 1u;
  }
}

__global__ void __launch_bounds__(256) zpq_ibwt_jump(u32 rows, const u32* __restrict__ nxt, const u32* __restrict__ dst, u32* __restrict__ nxt2, u32* __restrict__ dst2) {
  for (u32 p = blockIdx.x * blockDim.x + threadIdx.x; p < rows; p += gridDim.x * blockDim.x) {
    const u32 q = nxt[p];
    dst2[p] = dst[p] + dst[q];
    nxt2[p] = nxt[q];
  }
}

__global__ void __launch_bounds__(256) zpq_ibwt_emit(IbwtBlocks B, u32 rows, const u32* __restrict__ dst) {
  for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    const int b = block_of(B.off, B.nb, r);
    const u32 o = __ldg(B.off + b), n = __ldg(B.off + b + 1) - o - 1u, j = r - o;
    if (j == __ldg(B.idx + b)) continue;
    const u32 d = dst[r];
    if (d < n) B.out[b][n - 1u - d] = B.in[b][j];
  }
}

inline u64 al256(u64 x) { return (x + 255) & ~(u64)255; }
inline int bitlen(u64 x) { int r = 0; for (; x; x >>= 1) ++r; return r; }

template <bool MAXOP, bool INCL>
cudaError_t scan_u32(const u32* in, u32 n, u32* out, u32* partial, cudaStream_t st, u64* launches) {
  if (n == 0) return cudaSuccess;
  const u32 tiles = (n + TILE - 1) / TILE;
  zpq_scan_reduce<MAXOP><<<tiles, 256, 0, st>>>(in, n, partial);
  zpq_scan_partials<MAXOP><<<1, 1024, 0, st>>>(partial, tiles);
  zpq_scan_apply<MAXOP, INCL><<<tiles, 256, 0, st>>>(in, n, partial, out);
  if (launches) *launches += 3;
  return cudaGetLastError();
}

}  // namespace

// ------------------------------------------------------------------ host drivers
// device workspace for a batch of `total` payload bytes in nb blocks
extern "C" uint64_t zpq_bwt_workspace(uint64_t total, int nb) {
  const u64 n = total ? total : 1;
  const u64 tiles = (n + TILE - 1) / TILE;
  u64 w = 0;
  w += al256(n);                 // txt
  w += al256(n * 2);             // bid
  w += al256(n * 4) * 2;         // sa, rank_of
  w += al256(n * 8) * 2;         // key A / B
  w += al256(n * 4) * 2;         // val A / B
  w += al256(n * 4) * 2;         // slot A / B
  w += al256(n * 4) * 2;         // head / scan, active / pos
  w += al256(tiles * 256 * 4);   // radix counts
  w += al256((tiles + 2) * 4 + 4096);               // scan partials
  w += al256((u64)(nb + 1) * 4) + al256((u64)nb * 8) * 2 + 256;     // offsets, pointers, counter
  return w;
}

namespace {
struct SortView { const u8* txt; const u32* sa; const u32* rank_of; const u32* d_off; BwtBlocks B; u32 n; int maxg; };

// suffix arrays of all blocks in one index space (see the header): `out` may be null (no BWT emitted by the caller).
// alg_bytes (optional): algorithmic HBM bytes of the kernels launched, for the roofline: per sort pass 8 (histogram read) +
// 12 + 12 (scatter read + write) per active suffix, per round 100 more (keys, heads, two scans, update, compaction).
cudaError_t suffix_sort_core(const uint8_t* const* in, const uint64_t* len, uint8_t* const* out, int nb,
                             void* ws, uint64_t ws_bytes, int sm_count, cudaStream_t st, uint64_t* launches, uint64_t* alg_bytes, SortView* V) {
  if (nb > 65535) return cudaErrorInvalidValue;
  u64 total = 0;
  for (int b = 0; b < nb; ++b) total += len[b];
  if (total >= ((u64)1 << 31) || zpq_bwt_workspace(total, nb) > ws_bytes) return cudaErrorInvalidValue;
  const u32 n = (u32)total;
  const u64 n1 = n ? n : 1, tiles1 = (n1 + TILE - 1) / TILE;
  u8* p = (u8*)ws;
  auto take = [&](u64 bytes) { u8* r = p; p += al256(bytes); return r; };
  u8* txt = take(n1);
  u16* bid = (u16*)take(n1 * 2);
  u32* sa = (u32*)take(n1 * 4);
  u32* rank_of = (u32*)take(n1 * 4);
  u64* keyb[2] = {(u64*)take(n1 * 8), (u64*)take(n1 * 8)};
  u32* valb[2] = {(u32*)take(n1 * 4), (u32*)take(n1 * 4)};
  u32* slotb[2] = {(u32*)take(n1 * 4), (u32*)take(n1 * 4)};
  u32* t1 = (u32*)take(n1 * 4);
  u32* t2 = (u32*)take(n1 * 4);
  u32* counts = (u32*)take(tiles1 * 256 * 4);
  u32* partial = (u32*)take((tiles1 + 2) * 4 + 4096);
  u32* d_off = (u32*)take((u64)(nb + 1) * 4);
  const u8** d_in = (const u8**)take((u64)nb * 8);
  u8** d_out = (u8**)take((u64)nb * 8);
  u32* d_m = (u32*)take(256);

  // offsets, pointers (host staging lives until the sync below)
  u32* h_off = new u32[nb + 1];
  h_off[0] = 0;
  for (int b = 0; b < nb; ++b) h_off[b + 1] = h_off[b] + (u32)len[b];
  cudaError_t e;
  if ((e = cudaMemcpyAsync(d_off, h_off, (size_t)(nb + 1) * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) { delete[] h_off; return e; }
  if ((e = cudaMemcpyAsync(d_in, in, (size_t)nb * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) { delete[] h_off; return e; }
  if (out && (e = cudaMemcpyAsync(d_out, out, (size_t)nb * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) { delete[] h_off; return e; }
  e = cudaStreamSynchronize(st);
  delete[] h_off;
  if (e != cudaSuccess) return e;

  BwtBlocks B; B.in = d_in; B.out = d_out; B.off = d_off; B.nb = nb;
  const int maxg = sm_count * 8;
  auto grid = [&](u32 m) { const u32 g = (m + 255) / 256; return (int)(g < (u32)maxg ? (g ? g : 1) : (u32)maxg); };
  u64 nl = 0, ab = (u64)n * 24;                  // set-up: payload read, text / block id / key / value / slot written
  int kcur = 0, scur = 0;                        // key / value buffer and slot buffer that hold the active list
  u32 m = n;
  if (n) { zpq_bwt_setup<<<grid(n), 256, 0, st>>>(B, n, txt, bid, keyb[0], valb[0], slotb[0]); ++nl; }
  const int bits_hi = bitlen(n ? n - 1 : 0), bits_rank = bitlen(n);
  u32 h = 3;
  for (int round = 0; m > 0; ++round) {
    // ---- sort the active list by (group head, low key); the slots stay where they are (sorting permutes inside groups) ----
    const int bits_lo = round == 0 ? 27 : bits_rank;
    const u32 tiles = (m + TILE - 1) / TILE;
    for (int part = 0; part < 2; ++part) {
      const int nbits = part == 0 ? bits_lo : bits_hi, base = part == 0 ? 0 : 32;
      for (int sh = 0; sh < nbits; sh += 8) {
        zpq_radix_hist<<<tiles, 256, 0, st>>>(keyb[kcur], m, base + sh, counts, tiles);
        ++nl;
        if ((e = scan_u32<false, false>(counts, tiles * 256u, counts, partial, st, &nl)) != cudaSuccess) return e;
        zpq_radix_scatter<<<tiles, 256, 0, st>>>(keyb[kcur], valb[kcur], keyb[kcur ^ 1], valb[kcur ^ 1], m, base + sh, counts, tiles);
        ++nl;
        ab += (u64)m * 32;
        kcur ^= 1;
      }
    }
    ab += (u64)m * 100;
    // ---- regroup ----
    zpq_bwt_heads<<<grid(m), 256, 0, st>>>(keyb[kcur], slotb[scur], m, t1); ++nl;
    if ((e = scan_u32<true, true>(t1, m, t1, partial, st, &nl)) != cudaSuccess) return e;
    zpq_bwt_update<<<grid(m), 256, 0, st>>>(keyb[kcur], valb[kcur], slotb[scur], t1, m, sa, rank_of, t2); ++nl;
    if ((e = scan_u32<false, false>(t2, m, t2, partial, st, &nl)) != cudaSuccess) return e;
    zpq_bwt_compact<<<grid(m), 256, 0, st>>>(keyb[kcur], valb[kcur], slotb[scur], t1, t2, m, keyb[kcur ^ 1], valb[kcur ^ 1], slotb[scur ^ 1], d_m); ++nl;
    u32 m_next = 0;
    if ((e = cudaMemcpyAsync(&m_next, d_m, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    kcur ^= 1; scur ^= 1;
    if (m_next > 0) {
      zpq_bwt_keylow<<<grid(m_next), 256, 0, st>>>(keyb[kcur], valb[kcur], m_next, h, rank_of, bid, d_off); ++nl;
      h = h < 0x40000000u ? h * 2 : h;
    }
    m = m_next;
  }
  V->txt = txt; V->sa = sa; V->rank_of = rank_of; V->d_off = d_off; V->B = B; V->n = n; V->maxg = maxg;
  if (launches) *launches += nl;
  if (alg_bytes) *alg_bytes += ab;
  return cudaGetLastError();
}
}  // namespace

// in[b] / out[b]: DEVICE pointers (host arrays of them); len[b] host.  out[b] receives len[b] + 5 bytes.
extern "C" cudaError_t zpq_bwt_forward(const uint8_t* const* in, const uint64_t* len, uint8_t* const* out, int nb,
                                       void* ws, uint64_t ws_bytes, int sm_count, cudaStream_t st, uint64_t* launches, uint64_t* alg_bytes) {
  if (nb <= 0) return cudaSuccess;
  SortView V;
  cudaError_t e = suffix_sort_core(in, len, out, nb, ws, ws_bytes, sm_count, st, launches, alg_bytes, &V);
  if (e != cudaSuccess) return e;
  u32 maxlen = 0;
  for (int b = 0; b < nb; ++b) maxlen = len[b] > maxlen ? (u32)len[b] : maxlen;
  u32 gx = (maxlen + 1 + 255) / 256;
  if (gx > (u32)V.maxg) gx = (u32)V.maxg;
  if (gx * (u32)nb > (u32)V.maxg * 4u) gx = ((u32)V.maxg * 4u + nb - 1) / nb;
  if (gx == 0) gx = 1;
  zpq_bwt_emit<<<dim3(gx, (unsigned)nb), 256, 0, st>>>(V.B, V.txt, V.sa);
  if (launches) *launches += 1;
  if (alg_bytes) *alg_bytes += (u64)V.n * 6;
  return cudaGetLastError();
}

// The same sort without the BWT: device pointers to the suffix array (slot -> text index), the 1-based ranks (text index ->
// slot + 1) and the nb + 1 slot offsets, all in the batch's one index space, valid until `ws` is written again (zpq_lz77.cu).
extern "C" cudaError_t zpq_suffix_arrays(const uint8_t* const* in, const uint64_t* len, int nb, void* ws, uint64_t ws_bytes, int sm_count,
                                         cudaStream_t st, uint64_t* launches, uint64_t* alg_bytes,
                                         const uint32_t** sa, const uint32_t** rank, const uint32_t** off) {
  if (nb <= 0) return cudaSuccess;
  SortView V;
  cudaError_t e = suffix_sort_core(in, len, nullptr, nb, ws, ws_bytes, sm_count, st, launches, alg_bytes, &V);
  if (e != cudaSuccess) return e;
  *sa = V.sa; *rank = V.rank_of; *off = V.d_off;
  return cudaSuccess;
}

// in[b] -> out[b] (DEVICE pointers, host arrays of them), out[b] already holds a copy of in[b]; ws: 3 * nb * 8 bytes of device memory
extern "C" cudaError_t zpq_e8e9_forward(const uint8_t* const* in, uint8_t* const* out, const uint64_t* len, int nb, void* ws,
                                        int sm_count, cudaStream_t st, uint64_t* launches) {
  if (nb <= 0) return cudaSuccess;
  if (nb > 65535) return cudaErrorInvalidValue;
  u8* p = (u8*)ws;
  const u8** d_in = (const u8**)p; u8** d_out = (u8**)(p + al256((u64)nb * 8)); u64* d_len = (u64*)(p + 2 * al256((u64)nb * 8));
  cudaError_t e;
  if ((e = cudaMemcpyAsync(d_in, in, (size_t)nb * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(d_out, out, (size_t)nb * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(d_len, len, (size_t)nb * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;          // the host arrays may go out of scope
  u64 maxlen = 0;
  for (int b = 0; b < nb; ++b) maxlen = len[b] > maxlen ? len[b] : maxlen;
  const int slice = 512;
  u64 gx = (maxlen / slice + 127) / 128;
  if (gx < 1) gx = 1;
  if (gx > (u64)sm_count * 16) gx = (u64)sm_count * 16;
  zpq_e8e9_kernel<<<dim3((unsigned)gx, (unsigned)nb), 128, 0, st>>>(d_in, d_out, d_len, slice);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// ------------------------------------------------------------------ inverse BWT driver
extern "C" uint64_t zpq_ibwt_workspace(uint64_t total_rows, int nb) {
  const u64 n = total_rows ? total_rows : 1;
  const u64 tiles = (n + TILE - 1) / TILE;
  return al256(n * 8) * 2 + al256(n * 4) * 6 + al256(tiles * 256 * 4) + al256((tiles + 2) * 4 + 4096) +
         al256((u64)(nb + 1) * 4) + al256((u64)nb * 4) + al256((u64)nb * 8) * 2 + 256;
}

// in[b]: DEVICE pointer to the n_b + 5 bytes a BWT block's model emits after the program (rows, then the index);
// n[b]: payload length; idx[b]: the index (host, already range-checked: 1 .. n_b); out[b]: DEVICE pointer, n_b bytes.
extern "C" cudaError_t zpq_bwt_inverse(const uint8_t* const* in, const uint64_t* n, const uint32_t* idx, uint8_t* const* out, int nb,
                                       void* ws, uint64_t ws_bytes, int sm_count, cudaStream_t st, uint64_t* launches) {
  if (nb <= 0) return cudaSuccess;
  if (nb > 65535) return cudaErrorInvalidValue;
  u64 total = 0, maxn = 0;
  for (int b = 0; b < nb; ++b) { total += n[b] + 1; maxn = n[b] > maxn ? n[b] : maxn; }
  if (total >= ((u64)1 << 31) || zpq_ibwt_workspace(total, nb) > ws_bytes) return cudaErrorInvalidValue;
  const u32 rows = (u32)total;
  const u64 tiles1 = ((u64)rows + TILE - 1) / TILE;
  u8* p = (u8*)ws;
  auto take = [&](u64 bytes) { u8* r = p; p += al256(bytes); return r; };
  u64* keyb[2] = {(u64*)take((u64)rows * 8), (u64*)take((u64)rows * 8)};
  u32* valb[2] = {(u32*)take((u64)rows * 4), (u32*)take((u64)rows * 4)};
  u32* nxtb[2] = {(u32*)take((u64)rows * 4), (u32*)take((u64)rows * 4)};
  u32* dstb[2] = {(u32*)take((u64)rows * 4), (u32*)take((u64)rows * 4)};
  u32* counts = (u32*)take(tiles1 * 256 * 4);
  u32* partial = (u32*)take((tiles1 + 2) * 4 + 4096);
  u32* d_off = (u32*)take((u64)(nb + 1) * 4);
  u32* d_idx = (u32*)take((u64)nb * 4);
  const u8** d_in = (const u8**)take((u64)nb * 8);
  u8** d_out = (u8**)take((u64)nb * 8);
  u32* h_off = new u32[nb + 1];
  h_off[0] = 0;
  for (int b = 0; b < nb; ++b) h_off[b + 1] = h_off[b] + (u32)n[b] + 1u;
  cudaError_t e;
  if ((e = cudaMemcpyAsync(d_off, h_off, (size_t)(nb + 1) * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) { delete[] h_off; return e; }
  if ((e = cudaMemcpyAsync(d_idx, idx, (size_t)nb * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) { delete[] h_off; return e; }
  if ((e = cudaMemcpyAsync(d_in, in, (size_t)nb * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) { delete[] h_off; return e; }
  if ((e = cudaMemcpyAsync(d_out, out, (size_t)nb * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) { delete[] h_off; return e; }
  e = cudaStreamSynchronize(st);
  delete[] h_off;
  if (e != cudaSuccess) return e;
  IbwtBlocks B; B.in = d_in; B.out = d_out; B.off = d_off; B.idx = d_idx; B.nb = nb;
  const int maxg = sm_count * 8;
  auto grid = [&](u32 m) { const u32 g = (m + 255) / 256; return (int)(g < (u32)maxg ? (g ? g : 1) : (u32)maxg); };
  u64 nl = 0;
  zpq_ibwt_setup<<<grid(rows), 256, 0, st>>>(B, rows, keyb[0], valb[0]); ++nl;
  // stable sort of the rows by (block, symbol): 9 symbol bits, then the bits of the block's first row
  int kcur = 0;
  const u32 tiles = (u32)tiles1;
  const int bits_hi = bitlen(rows - 1);
  for (int part = 0; part < 2; ++part) {
    const int nbits = part == 0 ? 9 : bits_hi, base = part == 0 ? 0 : 32;
    for (int sh = 0; sh < nbits; sh += 8) {
      zpq_radix_hist<<<tiles, 256, 0, st>>>(keyb[kcur], rows, base + sh, counts, tiles); ++nl;
      if ((e = scan_u32<false, false>(counts, tiles * 256u, counts, partial, st, &nl)) != cudaSuccess) return e;
      zpq_radix_scatter<<<tiles, 256, 0, st>>>(keyb[kcur], valb[kcur], keyb[kcur ^ 1], valb[kcur ^ 1], rows, base + sh, counts, tiles); ++nl;
      kcur ^= 1;
    }
  }
  zpq_ibwt_links<<<grid(rows), 256, 0, st>>>(B, rows, keyb[kcur], valb[kcur], nxtb[0], dstb[0]); ++nl;
  int lc = 0;
  for (u64 span = 1; span < maxn + 1; span <<= 1) {
    zpq_ibwt_jump<<<grid(rows), 256, 0, st>>>(rows, nxtb[lc], dstb[lc], nxtb[lc ^ 1], dstb[lc ^ 1]); ++nl;
    lc ^= 1;
  }
  zpq_ibwt_emit<<<grid(rows), 256, 0, st>>>(B, rows, dstb[lc]); ++nl;
  if (launches) *launches += nl;
  return cudaGetLastError();
}
