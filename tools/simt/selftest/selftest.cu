// tools/simt/selftest/selftest.cu -- semantics check of the SIMT emulator itself (tools/simt), independent of the product
// kernels: every primitive the product uses, with results that can be stated in closed form.  Built by selftest.py with
// the same translator as the product sources and run on the host; prints "simt selftest ok" or the first mismatch.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define FULL 0xffffffffu
#define PAIR_BAR(id) asm volatile("bar.sync %0, %1;" ::"r"((id) + 1), "r"(64) : "memory")

// out[0..] per thread: a few words each
__global__ void __launch_bounds__(128) k_collectives(unsigned* out, unsigned* counter) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned* sh = reinterpret_cast<unsigned*>(smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
  unsigned* o = out + ((size_t)blockIdx.x * blockDim.x + tid) * 8;
  // shuffles
  o[0] = __shfl_sync(FULL, (unsigned)(lane * 3 + warp), (lane + 5) & 31);             // value of lane+5
  o[1] = __shfl_up_sync(FULL, (unsigned)lane, 2);                                       // lane-2 (own for lanes 0,1)
  o[2] = __ballot_sync(FULL, (lane % 3) == 0);
  o[3] = (unsigned)__reduce_add_sync(FULL, lane - 10);                                  // sum(0..31) - 320 = 176
  o[4] = (unsigned)__all_sync(FULL, lane < 32) | (unsigned)__any_sync(FULL, lane == 40) << 1;
  // block barrier + shared memory: every thread writes, then reads its neighbour
  sh[tid] = tid * 7u;
  __syncthreads();
  o[5] = sh[(tid + 33) % blockDim.x];
  __syncthreads();
  // named barrier between warp pairs + spin on a volatile flag: warp 2k produces, warp 2k+1 consumes
  volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(sh + 256 + (warp >> 1) * 4);
  volatile unsigned* box = flag + 1;
  if (lane == 0 && (warp & 1) == 0) { *flag = 0; *box = 0; }
  PAIR_BAR(warp >> 1);
  unsigned got = 0;
  for (unsigned round = 1; round <= 5; ++round) {
    if ((warp & 1) == 0) {
      if (lane == 0) { *box = round * 100u + (unsigned)(warp >> 1); __threadfence_block(); *flag = round; }
      while (*flag != round + 1000u) { }          // wait for the consumer's acknowledgement
      __syncwarp();                               // every lane has seen it before lane 0 overwrites the word
    } else {
      while (*flag != round) { }
      __threadfence_block();
      got += *box;
      __syncwarp();
      if (lane == 0) *flag = round + 1000u;
    }
  }
  o[6] = got;                                      // consumers: sum(round*100 + pair) = 1500 + 5*pair
  // early exit of some lanes, then a collective among the rest (exited lanes count as arrived)
  if (lane >= 24) { o[7] = 0xdead; atomicAdd(counter, 1u); return; }
  o[7] = __ballot_sync(FULL, 1);                   // the 24 remaining lanes
  atomicAdd(counter, 100u);
}

extern "C" int simt_selftest(void) {
  const int ctas = 3, threads = 128;
  unsigned *out = nullptr, *counter = nullptr;
  cudaMalloc((void**)&out, (size_t)ctas * threads * 8 * 4);
  cudaMalloc((void**)&counter, 4);
  cudaMemset(counter, 0, 4);
  k_collectives<<<ctas, threads, 4096, 0>>>(out, counter);
  cudaDeviceSynchronize();
  unsigned h[ctas * threads * 8], hc = 0;
  cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
  cudaMemcpy(&hc, counter, 4, cudaMemcpyDeviceToHost);
  int bad = 0;
#define EXPECT(cond, what) do { if (!(cond)) { if (!bad) printf("simt selftest MISMATCH: %s (cta %d thread %d)\n", what, c, t); ++bad; } } while (0)
  for (int c = 0; c < ctas; ++c)
    for (int t = 0; t < threads; ++t) {
      const unsigned* o = h + ((size_t)c * threads + t) * 8;
      const int lane = t & 31, warp = t >> 5;
      EXPECT(o[0] == (unsigned)(((lane + 5) & 31) * 3 + warp), "__shfl_sync");
      EXPECT(o[1] == (unsigned)(lane >= 2 ? lane - 2 : lane), "__shfl_up_sync");
      EXPECT(o[2] == 0x49249249u, "__ballot_sync");
      EXPECT(o[3] == 176u, "__reduce_add_sync");
      EXPECT(o[4] == 1u, "__all_sync / __any_sync");
      EXPECT(o[5] == (unsigned)((t + 33) % threads) * 7u, "__syncthreads + shared memory");
      EXPECT(o[6] == ((warp & 1) ? 1500u + 5u * (unsigned)(warp >> 1) : 0u), "named barrier + spin hand-shake");
      EXPECT(o[7] == (lane >= 24 ? 0xdeadu : 0x00ffffffu), "collective after lanes exited");
    }
  if (hc != (unsigned)(ctas * 4 * (8 * 1 + 24 * 100))) { if (!bad) printf("simt selftest MISMATCH: atomicAdd total %u\n", hc); ++bad; }
  cudaFree(out); cudaFree(counter);
  if (!bad) printf("simt selftest ok\n");
  return bad;
}
