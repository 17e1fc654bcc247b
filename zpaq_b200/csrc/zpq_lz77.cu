// zpq_lz77.cu -- LZBuffer levels 1 and 2 (libzpaq.cpp:6473-6883): the LZ77 parsers in front of the model, on the device.
//
// The parse is a sequential greedy search: every position probes a hash table (or the neighbourhood of its suffix in the
// suffix array) that earlier positions filled, scores the candidates, emits a match or extends the literal run, and
// indexes the bytes it steps over.  Each decision depends on the previous one and an archive is only byte-identical when
// every decision is, so there is no parallelism inside a block to use: the kernel runs ONE THREAD PER BLOCK and a batch is
// as fast as it is wide.  (That is the honest shape of this stage; the heavy pre-processing stage, suffix sorting, is the
// batched sorter of zpq_bwt.cu, which this file reuses for the suffix-array variant.)
//
//   level 1  variable-length codes, LSB first:  00,n,L[n] = n literals ;  mm,mmm,n,ll,r,q = match of 4n+ll at offset
//            (q << rb) + r - 1 ; lengths interleaved Elias-gamma (libzpaq.cpp:6546-6556)
//   level 2  byte codes: 00xxxxxx = x + 1 literals ; yyxxxxxx = match of x + minMatch with y + 1 offset bytes (6558-6561)
//   search   hash table of 2^bits entries, each (position << checkbits | check byte), buckets of 2^b neighbours, rolling
//            hash of the next minMatch bytes -- or, when the method asks for it, the 2^b neighbours of the position's suffix
//            in the suffix array on both sides, through a windowed inverse suffix array rebuilt every 2^(17 + size class)
//            positions, with an optional look-ahead of one byte
#include <cuda_runtime.h>
#include <stdint.h>
#include "zpq_kernels.h"

namespace {

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;

__device__ __forceinline__ int lz_lg(u32 x) { return 32 - __clz((int)x); }          // bits without leading zeros, lg(0) = 0

struct LzOut {
  u8* out; u64 pos, cap;
  u32 bits, nbits;
  bool full;
  __device__ __forceinline__ void put(u32 c) { if (pos < cap) out[pos] = (u8)c; else full = true; ++pos; }
  __device__ __forceinline__ void putb(u32 x, int k) {
    x &= (1u << k) - 1u;
    bits |= x << nbits;
    nbits += (u32)k;
    while (nbits > 7) { put(bits & 255u); bits >>= 8; nbits -= 8; }
  }
  __device__ __forceinline__ void flush() { if (nbits > 0) put(bits & 255u); bits = 0; nbits = 0; }
};

struct LzCfg {
  int level;                 // 1 or 2
  u32 min_match, bucket, shift1, min_both, rb, lookahead, checkbits, htsize, max_match, max_literal;
  int use_sa;
};

// literals in[i - lit .. i - 1]
__device__ void lz_literals(const LzCfg& C, LzOut& O, const u8* __restrict__ in, u32 i, u32& lit) {
  if (C.level == 1) {
    if (lit < 1) return;
    int ll = lz_lg(lit);
    O.putb(0, 2);
    --ll;
    while (--ll >= 0) { O.putb(1, 1); O.putb((lit >> ll) & 1u, 1); }
    O.putb(0, 1);
    while (lit) { O.putb(in[i - lit], 8); --lit; }
  } else {
    while (lit > 0) {
      const u32 run = lit > 64 ? 64u : lit;
      O.put(run - 1);
      for (u32 j = i - lit; j < i - lit + run; ++j) O.put(in[j]);
      lit -= run;
    }
  }
}

__device__ void lz_match(const LzCfg& C, LzOut& O, u32 len, u32 off) {
  if (C.level == 1) {
    int ll = lz_lg(len) - 1;
    off += (1u << C.rb) - 1u;
    const int lo = lz_lg(off) - 1 - (int)C.rb;
    O.putb((u32)(lo + 8) >> 3, 2);
    O.putb((u32)lo & 7u, 3);
    while (--ll >= 2) { O.putb(1, 1); O.putb((len >> ll) & 1u, 1); }
    O.putb(0, 1);
    O.putb(len & 3u, 2);
    O.putb(off, (int)C.rb);
    O.putb(off >> C.rb, lo);
  } else {
    --off;
    const u32 m = C.min_match;
    while (len > 0) {                      // long matches in pieces of m .. m + 63
      const u32 piece = len > m * 2 + 63 ? m + 63 : (len > m + 63 ? len - m : len);
      if (off < (1u << 16)) { O.put(64 + piece - m); O.put(off >> 8); O.put(off); }
      else if (off < (1u << 24)) { O.put(128 + piece - m); O.put(off >> 16); O.put(off >> 8); O.put(off); }
      else { O.put(192 + piece - m); O.put(off >> 24); O.put(off >> 16); O.put(off >> 8); O.put(off); }
      len -= piece;
    }
  }
}

// per block: in / n / out / cap ; table = hash table (htsize words, zeroed) or the ISA window (2^checkbits words, zeroed);
// sa / rank = the block's suffix array slice and the 1-based global ranks (suffix-array variant), base = its first slot
struct LzBlock { const u8* in; u8* out; u32* table; const u32* sa; const u32* rank; u64 cap; u32 n, base; };

__global__ void __launch_bounds__(32) zpq_lz77_kernel(const LzBlock* blocks, int nb, LzCfg C, u64* out_len, int* status) {
  const int b = blockIdx.x;
  if (b >= nb || threadIdx.x != 0) return;
  const LzBlock B = blocks[b];
  const u8* __restrict__ in = B.in;
  const u32 n = B.n;
  u32* const tab = B.table;
  LzOut O; O.out = B.out; O.pos = 0; O.cap = B.cap; O.bits = 0; O.nbits = 0; O.full = false;
  const u32 mask = (1u << C.checkbits) - 1u;
  u32 i = 0, lit = 0, h1 = 0;
  while (i < n) {
    u32 best_len = C.min_match - 1, best_pos = 0, best_back = 0;
    int best_score = 0;
    if (C.use_sa) {
      // the ISA window of position i: entry x holds the rank of position (i & ~mask) + x once the window was rebuilt for it
      {
        const u32 e = tab[i & mask];
        if (B.sa[B.base + e] - B.base != i) {
          const u32 w0 = i & ~mask;
          for (u32 x = 0; x <= mask && w0 + x < n; ++x) tab[x] = B.rank[B.base + w0 + x] - 1u - B.base;
        }
      }
      for (u32 h = 0; h <= C.lookahead; ++h) {
        const u32 q = tab[(h + i) & mask];
        if (B.sa[B.base + q] - B.base != h + i) continue;
        for (int dir = -1; dir <= 1; dir += 2) {
          for (u32 k = 1; k <= C.bucket; ++k) {
            const u32 nbr = q + (u32)(dir * (int)k);
            u32 p;
            if (nbr < n && (p = B.sa[B.base + nbr] - B.base - h) < i) {
              u32 run, back;
              for (run = h; i + run < n && run < C.max_match && in[p + run] == in[i + run]; ++run) { }
              for (back = h; back > 0 && in[p + back - 1] == in[i + back - 1]; --back) { }
              int score = (int)(run - back) * 8 - lz_lg(i - p) - 4 * (lit == 0 && back > 0 ? 1 : 0) - 11;
              for (u32 a = 0; a < h; ++a) score = score * 5 / 8;
              if (score > best_score) { best_len = run; best_pos = p; best_back = back; best_score = score; }
              if (run < best_len || run < C.min_match || run > 255) break;
            }
          }
        }
        if (best_score <= 0 || best_len < C.min_match) break;
      }
    }
    else if (C.level == 1 || C.min_match <= 64) {
      for (u32 k = 0; k <= C.bucket; ++k) {
        u32 p = tab[h1 ^ k];
        if (p && i + 3 < n && (p & mask) == ((u32)in[i + 3] & mask)) {
          p >>= C.checkbits;
          if (p < i && i + best_len <= n && in[p + best_len - 1] == in[i + best_len - 1]) {
            u32 run;
            for (run = 0; i + run < n && run < C.max_match && in[p + run] == in[i + run]; ++run) { }
            const int score = (int)run * 8 - lz_lg(i - p) - 2 * (lit > 0 ? 1 : 0) - 11;
            if (score > best_score) { best_len = run; best_pos = p; best_back = 0; best_score = score; }
          }
        }
        if (best_len >= 128) break;
      }
    }
    // a match that pays for its code, else one more literal
    const u32 off = i - best_pos;
    if (off > 0 && best_score > 0 && best_len - best_back >= C.min_match + (C.level == 2 ? ((off >= (1u << 16)) + (off >= (1u << 24))) : 0u)) {
      lit += best_back;
      lz_literals(C, O, in, i + best_back, lit);
      lz_match(C, O, best_len - best_back, off);
    } else {
      best_len = 1;
      ++lit;
    }
    // index what is stepped over
    if (C.use_sa) i += best_len;
    else {
      while (best_len--) {
        if (i + C.min_both < n) {
          const u32 ih = ((i * 1234547u) >> 19) & C.bucket;
          tab[h1 ^ ih] = (i << C.checkbits) | ((u32)in[i + 3] & mask);
          h1 = (((h1 * 5u) << C.shift1) + ((u32)in[i + C.min_match] + 1u) * 123456791u) & (C.htsize - 1u);
        }
        ++i;
      }
    }
    if (lit >= C.max_literal) lz_literals(C, O, in, i, lit);
  }
  lz_literals(C, O, in, n, lit);
  O.flush();
  out_len[b] = O.pos;
  status[b] = O.full ? 1 : 0;
}

// ------------------------------------------------------------------ the other direction: what the blocks' PCOMP programs compute
// One thread per block again (copies depend on what earlier codes wrote).  Anything a well-formed stream cannot contain --
// an offset before the start of the block, more output than the caller's size hint -- ends the block with status 1 and the
// host ZPAQL machine takes it, so a damaged archive still behaves as libzpaq's post-processor would.
struct UnBlock { const u8* in; u8* out; u64 in_len, out_cap; };

// E8E9 inverse: ascending positions, each test on the bytes as they are at that moment (libzpaq.cpp:7296-7320 and the
// "e8e9 transform to out" passes of the other programs)
__device__ void un_e8e9(u8* buf, u64 n) {
  if (n < 5) return;
  for (u64 i = 0; i + 4 < n; ++i) {
    if ((buf[i] & 254) == 0xe8 && (((u32)buf[i + 4] + 1u) & 254u) == 0u) {
      const u32 v = ((u32)buf[i + 1] | (u32)buf[i + 2] << 8 | (u32)buf[i + 3] << 16) - (u32)i;
      buf[i + 1] = (u8)v; buf[i + 2] = (u8)(v >> 8); buf[i + 3] = (u8)(v >> 16);
    }
  }
}

// kind: 1 = variable-length codes, 2 = byte codes, 4 = E8E9 only (in = the rewritten bytes); e8: undo E8E9 at the end
__global__ void __launch_bounds__(32) zpq_unlz_kernel(const UnBlock* blocks, int nb, int kind, int e8, u32 rb, u32 min_match, u64* out_len, int* status) {
  const int b = blockIdx.x;
  if (b >= nb || threadIdx.x != 0) return;
  const UnBlock B = blocks[b];
  const u8* __restrict__ in = B.in;
  u8* out = B.out;
  u64 ptr = 0;
  bool bad = false;
  if (kind == 4) {
    if (B.in_len > B.out_cap) bad = true;
    else { if (B.in != B.out) for (u64 k = 0; k < B.in_len; ++k) out[k] = B.in[k]; ptr = B.in_len; }      // in == out: E8E9 undone in place
  }
  else if (kind == 1) {
    u64 bits = 0; u32 n = 0, state = 0, len = 0, m = 0, r5 = 0;
    for (u64 k = 0; k < B.in_len && !bad; ++k) {
      bits += (u64)in[k] << n; n += 8;
      if (n > 40) { bad = true; break; }
      if (state == 0) {
        len = 1;
        if (bits & 3) { m = (((u32)bits & 3u) - 1u) * 8u; bits >>= 2; m += (u32)bits & 7u; bits >>= 3; n -= 5; state = 1; }
        else { bits >>= 2; n -= 2; state = 3; }
      }
      while (state == 1 && n >= 3) {
        if (bits & 1) { bits >>= 1; len += len + ((u32)bits & 1u); bits >>= 1; n -= 2; }
        else { bits >>= 1; len = (len << 2) + ((u32)bits & 3u); bits >>= 2; n -= 3; state = rb ? 5u : 2u; }
      }
      if (rb && state == 5 && n >= rb) { r5 = (u32)bits & ((1u << rb) - 1u); bits >>= rb; n -= rb; state = 2; }
      if (state == 2 && m <= n) {
        u64 off = (bits & (((u64)1 << m) - 1)) + ((u64)1 << m);
        if (rb) off = (off << rb) + r5 - ((1u << rb) - 1u);
        if (off > ptr || off == 0 || ptr + len > B.out_cap) { bad = true; break; }
        for (u32 j = 0; j < len; ++j) { out[ptr] = out[ptr - off]; ++ptr; }
        bits >>= m; n -= m; state = 0;
      }
      while (state == 3 && n >= 2) {
        if (bits & 1) { bits >>= 1; len += len + ((u32)bits & 1u); bits >>= 1; n -= 2; }
        else { bits >>= 1; n -= 1; state = 4; }
      }
      if (state == 4 && n >= 8) {
        if (ptr >= B.out_cap) { bad = true; break; }
        out[ptr++] = (u8)bits; bits >>= 8; n -= 8;
        if (--len == 0) state = 0;
      }
    }
  }
  else {
    u32 d = 0, len = 0; u64 off = 0;
    for (u64 k = 0; k < B.in_len && !bad; ++k) {
      const u32 c = in[k];
      if (d == 0) {
        d = (c >> 6) + 1;
        if (d == 1) len = c + 1;
        else { ++d; len = (c & 63u) + min_match; }
        off = 0;
      }
      else if (d == 1) {
        if (ptr >= B.out_cap) { bad = true; break; }
        out[ptr++] = (u8)c;
        if (--len == 0) d = 0;
      }
      else if (d > 2) { off = (off << 8) | c; --d; }
      else {
        off = ((off << 8) | c) + 1;
        if (off > ptr || ptr + len > B.out_cap || len == 0) { bad = true; break; }
        for (u32 j = 0; j < len; ++j) { out[ptr] = out[ptr - off]; ++ptr; }
        d = 0;
      }
    }
  }
  if (!bad && e8) un_e8e9(out, ptr);
  out_len[b] = ptr;
  status[b] = bad ? 1 : 0;
}

inline u64 al256(u64 x) { return (x + 255) & ~(u64)255; }

}  // namespace

// numbers = makeConfig's $1..$7 for this block family (size class, kind, minimum match, second order, log2 bucket, table bits,
// look-ahead); every block of a call has the same numbers.  The second context order (numbers[3]) must be 0.
extern "C" int zpq_lz77_uses_sa(const int numbers[7]) { return numbers[5] - numbers[0] >= 21; }

// device bytes of per-block tables for a block of n bytes
extern "C" uint64_t zpq_lz77_table_bytes(const int numbers[7]) {
  const bool sa = numbers[5] - numbers[0] >= 21;
  return sa ? ((u64)4 << (17 + numbers[0])) : ((u64)4 << numbers[5]);
}

// in[b] / out[b]: DEVICE pointers (host arrays); tables: nb * zpq_lz77_table_bytes of device memory; sa / rank: the batch's suffix
// array and 1-based ranks in one index space with base[b] = first slot of block b (suffix-array variant; else null);
// meta: device scratch of nb * 64 + 64 bytes.  out_len / status: host arrays, valid after the stream is synchronised HERE.
extern "C" cudaError_t zpq_lz77_forward(const uint8_t* const* in, const uint64_t* len, uint8_t* const* out, const uint64_t* out_cap, int nb,
                                        const int numbers[7], void* tables, const uint32_t* sa, const uint32_t* rank, const uint32_t* base,
                                        void* meta, uint64_t* out_len, int* status, cudaStream_t st, uint64_t* launches) {
  if (nb <= 0) return cudaSuccess;
  LzCfg C;
  C.level = numbers[1] & 3;
  C.min_match = (u32)numbers[2];
  C.bucket = (1u << numbers[4]) - 1u;
  C.use_sa = numbers[5] - numbers[0] >= 21 ? 1 : 0;
  C.checkbits = (u32)(C.use_sa ? 17 + numbers[0] : 12 - numbers[0]);
  C.htsize = C.use_sa ? 0u : (1u << numbers[5]);
  C.shift1 = C.min_match > 0 ? (u32)(numbers[5] - 1) / C.min_match + 1u : 1u;
  C.lookahead = (u32)numbers[6];
  C.min_both = (C.min_match > C.lookahead ? C.min_match : C.lookahead) + 4u;
  C.rb = numbers[0] > 4 ? (u32)(numbers[0] - 4) : 0u;
  C.max_match = (1u << 14) * 3u;
  C.max_literal = (1u << 14) / 4u;
  const u64 tb = zpq_lz77_table_bytes(numbers);
  cudaError_t e;
  if ((e = cudaMemsetAsync(tables, 0, (size_t)(tb * (u64)nb), st)) != cudaSuccess) return e;
  LzBlock* hb = new LzBlock[nb];
  for (int b = 0; b < nb; ++b) {
    hb[b].in = in[b]; hb[b].out = out[b]; hb[b].cap = out_cap[b]; hb[b].n = (u32)len[b];
    hb[b].table = (u32*)((u8*)tables + tb * (u64)b);
    hb[b].sa = sa; hb[b].rank = rank; hb[b].base = base ? base[b] : 0u;
  }
  u8* dm = (u8*)meta;
  LzBlock* d_blocks = (LzBlock*)dm;
  u64* d_len = (u64*)(dm + al256((u64)nb * sizeof(LzBlock)));
  int* d_st = (int*)(dm + al256((u64)nb * sizeof(LzBlock)) + al256((u64)nb * 8));
  e = cudaMemcpyAsync(d_blocks, hb, (size_t)nb * sizeof(LzBlock), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  delete[] hb;
  if (e != cudaSuccess) return e;
  zpq_lz77_kernel<<<nb, 32, 0, st>>>(d_blocks, nb, C, d_len, d_st);
  if (launches) *launches += 1;
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(out_len, d_len, (size_t)nb * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(status, d_st, (size_t)nb * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
  return cudaStreamSynchronize(st);
}

extern "C" uint64_t zpq_lz77_meta_bytes(int nb) { return al256((u64)nb * sizeof(LzBlock)) + al256((u64)nb * 8) + al256((u64)nb * 4) + 256; }

// Post-processing of blocks whose program is one this library writes: kind 1 / 2 = LZ77 decoders, 4 = E8E9 only, e8 = undo E8E9
// after the decode; in[b]: the model bytes after the program (DEVICE pointers, host arrays), out[b]: out_cap[b] bytes.
// meta: zpq_unlz_meta_bytes(nb) of device memory.  out_len / status (host) are valid on return; status 1 = not a stream this
// kernel takes (the caller falls back to the host ZPAQL machine).
extern "C" uint64_t zpq_unlz_meta_bytes(int nb) { return al256((u64)nb * sizeof(UnBlock)) + al256((u64)nb * 8) + al256((u64)nb * 4) + 256; }

extern "C" cudaError_t zpq_unlz_forward(const uint8_t* const* in, const uint64_t* in_len, uint8_t* const* out, const uint64_t* out_cap, int nb,
                                        int kind, int e8, int rb, int min_match, void* meta, uint64_t* out_len, int* status,
                                        cudaStream_t st, uint64_t* launches) {
  if (nb <= 0) return cudaSuccess;
  UnBlock* hb = new UnBlock[nb];
  for (int b = 0; b < nb; ++b) { hb[b].in = in[b]; hb[b].out = out[b]; hb[b].in_len = in_len[b]; hb[b].out_cap = out_cap[b]; }
  u8* dm = (u8*)meta;
  UnBlock* d_blocks = (UnBlock*)dm;
  u64* d_len = (u64*)(dm + al256((u64)nb * sizeof(UnBlock)));
  int* d_st = (int*)(dm + al256((u64)nb * sizeof(UnBlock)) + al256((u64)nb * 8));
  cudaError_t e = cudaMemcpyAsync(d_blocks, hb, (size_t)nb * sizeof(UnBlock), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  delete[] hb;
  if (e != cudaSuccess) return e;
  zpq_unlz_kernel<<<nb, 32, 0, st>>>(d_blocks, nb, kind, e8, (u32)rb, (u32)min_match, d_len, d_st);
  if (launches) *launches += 1;
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(out_len, d_len, (size_t)nb * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
  if ((e = cudaMemcpyAsync(status, d_st, (size_t)nb * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
  return cudaStreamSynchronize(st);
}
