// zpq_prep.cu -- per-block analysis that precedes the coding path, on the device, sm_100a.
//
//   zpq_sha1_kernel   SHA-1 of every block payload (libzpaq::SHA1, libzpaq.cpp:101-177, called from compressBlock at
//                     7568-7577 and by the decompresser's verify): one THREAD per message.  SHA-1 is a serial chain of
//                     64-byte compressions, so the parallelism is across blocks -- which is what a batch has (1 029 for the
//                     1 GiB benchmark); a lane streams its own block with aligned 32-bit loads and a byte permute.
//   zpq_gaps_kernel   the level-5 model selection statistics (libzpaq.cpp:7655-7688): histogram of the distance from every
//                     byte to the previous byte of the same value (1 .. 4095), then up to two dominant gaps.  One WARP per
//                     block: lane l scans the l-th 1/32 of the block with private "last seen" tables in shared memory and
//                     a shared histogram (shared-memory atomics); the first occurrences inside each slice are stitched to
//                     the slices before them afterwards; lane 0 then repeats the reference's two scans over the 4 096 bins
//                     in IEEE double arithmetic (same expression, same comparisons -> same choice).
// Both results are bit-exact by construction and are checked against the host implementations (zpq_frame.cpp,
// zpq_method.cpp) in tests/test_gpu_prep.py.
#include <cuda_runtime.h>
#include <stdint.h>
#include "zpq_kernels.h"

namespace {

__device__ __forceinline__ uint32_t rol32(uint32_t x, int n) { return __funnelshift_l(x, x, n); }

__device__ __forceinline__ void sha1_compress(uint32_t h[5], uint32_t w[16]) {
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
#pragma unroll
  for (int i = 0; i < 80; ++i) {
    uint32_t wi;
    if (i < 16) wi = w[i];
    else {
      wi = rol32(w[(i + 13) & 15] ^ w[(i + 8) & 15] ^ w[(i + 2) & 15] ^ w[i & 15], 1);
      w[i & 15] = wi;
    }
    uint32_t f, k;
    if (i < 20) { f = (b & c) | (~b & d); k = 0x5A827999u; }
    else if (i < 40) { f = b ^ c ^ d; k = 0x6ED9EBA1u; }
    else if (i < 60) { f = (b & c) | (b & d) | (c & d); k = 0x8F1BBCDCu; }
    else { f = b ^ c ^ d; k = 0xCA62C1D6u; }
    const uint32_t t = rol32(a, 5) + f + e + k + wi;
    e = d; d = c; c = rol32(b, 30); b = a; a = t;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e;
}

}  // namespace

// ptr[i] / len[i]: message i; out: 20 bytes per message
__global__ void __launch_bounds__(32) zpq_sha1_kernel(const uint8_t* const* ptr, const uint64_t* len, uint8_t* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* p = ptr[i];
  const uint64_t L = len[i];
  uint32_t h[5] = {0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u};
  uint32_t w[16];
  // whole 64-byte chunks: 17 aligned words, re-assembled big-endian at the message's byte offset
  const uint32_t sh = (uint32_t)((uintptr_t)p & 3u);
  const uint32_t sel = (sh + 3u) | (sh + 2u) << 4 | (sh + 1u) << 8 | sh << 12;
  const uint32_t* q = reinterpret_cast<const uint32_t*>(p - sh);
  uint64_t done = 0;
  // the aligned window reads up to 3 bytes past the chunk: keep it to chunks that are followed by at least 4 more bytes
  for (; done + 68 <= L; done += 64) {
    uint32_t lo = q[0];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint32_t hi = q[k + 1];
      w[k] = __byte_perm(lo, hi, sel);
      lo = hi;
    }
    q += 16;
    sha1_compress(h, w);
  }
  // tail (at most 67 + padding bytes): byte loads into a separately indexed buffer (keeps w[] in registers above)
  const uint8_t* t = p + done;
  uint64_t rest = L - done;
  bool pad_done = false, len_done = false;
  while (!len_done) {
    uint32_t wt[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) wt[k] = 0;
    int fill = 0;
    for (; fill < 64 && rest; ++fill, --rest) wt[fill >> 2] |= (uint32_t)*t++ << (24 - 8 * (fill & 3));
    if (fill < 64 && !pad_done) { wt[fill >> 2] |= 0x80u << (24 - 8 * (fill & 3)); ++fill; pad_done = true; }
    if (pad_done && fill <= 56) { const uint64_t bits = L * 8; wt[14] = (uint32_t)(bits >> 32); wt[15] = (uint32_t)bits; len_done = true; }
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = wt[k];
    sha1_compress(h, w);
  }
  uint8_t* o = out + (size_t)i * 20;
#pragma unroll
  for (int k = 0; k < 5; ++k) { o[k * 4] = (uint8_t)(h[k] >> 24); o[k * 4 + 1] = (uint8_t)(h[k] >> 16); o[k * 4 + 2] = (uint8_t)(h[k] >> 8); o[k * 4 + 3] = (uint8_t)h[k]; }
}

#define GAP_BINS 4096
// shared memory of one warp: hist[4096] | last[32][256] | first[32][256]      (all int32; -1 = not seen in the slice)
#define GAPS_SMEM ((GAP_BINS + 2 * 32 * 256) * 4)

__global__ void __launch_bounds__(32) zpq_gaps_kernel(const uint8_t* const* ptr, const uint64_t* len, int32_t* gaps, int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* hist = reinterpret_cast<int*>(smem_raw);
  int* last = hist + GAP_BINS;
  int* first = last + 32 * 256;
  const int lane = threadIdx.x;
  for (int blk = blockIdx.x; blk < n; blk += gridDim.x) {
    const uint8_t* p = ptr[blk];
    const int L = (int)len[blk];
    for (int i = lane; i < GAP_BINS; i += 32) hist[i] = 0;
    for (int i = lane; i < 32 * 256; i += 32) { last[i] = -1; first[i] = -1; }
    __syncwarp();
    // slices of whole 16-byte lines so that the lanes' loads do not straddle
    const int slice = ((L + 31) / 32 + 15) & ~15;
    const int b0 = min(L, lane * slice), b1 = min(L, b0 + slice);
    int* mylast = last + lane * 256;
    int* myfirst = first + lane * 256;
    for (int i = b0; i < b1; ++i) {
      const int c = p[i];
      const int prev = mylast[c];
      if (prev >= 0) { const int g = i - prev; if (g < GAP_BINS) atomicAdd(&hist[g], 1); }
      else myfirst[c] = i;
      mylast[c] = i;
    }
    __syncwarp();
    // stitch: the first occurrence of a byte value in a slice measures its gap to the last occurrence in the slices before
    // it; before any occurrence the reference's table holds position 0 (libzpaq.cpp:7657: pt starts as zeros), so the very
    // first occurrence at position i counts as a gap of i
    for (int c = lane; c < 256; c += 32) {
      int prev = 0;
      for (int s = 0; s < 32; ++s) {
        const int f = first[s * 256 + c];
        if (f >= 0) {
          const int g = f - prev;
          if (g > 0 && g < GAP_BINS) atomicAdd(&hist[g], 1);
          prev = last[s * 256 + c];
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      int pool = L - hist[1] - hist[2] - hist[3];
      int found = 0;
      int32_t* o = gaps + 2 * (size_t)blk;
      o[0] = 0; o[1] = 0;
      while (found < 2) {
        int best = 0, seen = 0;
        double best_share = 0;
        for (int g = 5; g < GAP_BINS && seen < pool; ++g) {
          const double share = (double)hist[g] / (256.0 + (double)pool - (double)seen);
          if (share > best_share) { best_share = share; best = g; }
          seen += hist[g];
        }
        if (best <= 4 || !(best_share > 0.1)) break;
        o[found++] = best;
        pool -= hist[best];
        hist[best] = 0;
      }
    }
    __syncwarp();
  }
}

// one thread per record: copy `n` bytes from src to dst (block prologues and trailers written around the coded data)
__global__ void zpq_scatter_kernel(const uint8_t* pool, const uint64_t* src_off, uint8_t* const* dst, const uint32_t* nbytes, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* s = pool + src_off[i];
  uint8_t* d = dst[i];
  for (uint32_t k = 0; k < nbytes[i]; ++k) d[k] = s[k];
}

extern "C" cudaError_t zpq_launch_sha1(const uint8_t* const* ptr, const uint64_t* len, uint8_t* out, int n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  zpq_sha1_kernel<<<(n + 31) / 32, 32, 0, st>>>(ptr, len, out, n);
  return cudaGetLastError();
}

extern "C" cudaError_t zpq_launch_gaps(const uint8_t* const* ptr, const uint64_t* len, int32_t* gaps, int n, int sm_count, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(zpq_gaps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GAPS_SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int ctas = n < sm_count * 2 ? n : sm_count * 2;
  zpq_gaps_kernel<<<ctas, 32, GAPS_SMEM, st>>>(ptr, len, gaps, n);
  return cudaGetLastError();
}

extern "C" cudaError_t zpq_launch_scatter(const uint8_t* pool, const uint64_t* src_off, uint8_t* const* dst, const uint32_t* nbytes, int n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  zpq_scatter_kernel<<<(n + 127) / 128, 128, 0, st>>>(pool, src_off, dst, nbytes, n);
  return cudaGetLastError();
}
