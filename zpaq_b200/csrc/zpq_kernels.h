// zpq_kernels.h -- launch entry points of zpq_kernels.cu (internal).
#ifndef ZPQ_KERNELS_H
#define ZPQ_KERNELS_H
#include <cuda_runtime.h>
#include "zpq_dev.h"

#define ZPQ_WARPS 4   // v1 kernel: warps (= archive blocks in flight) per CTA
#define ZPQ_SMEM_MAX 232448   // 227 KB: per-CTA dynamic shared memory limit on sm_100
#define ZPQ_MIXREG_HOST 3     // must equal ZPQ_MIXREG in zpq_kernels2.cu

extern "C" {
size_t zpq_code_kernel_smem(void);
cudaError_t zpq_launch_init(const ZKernelArgs* a, int nblk, cudaStream_t st);
cudaError_t zpq_launch_code(const ZKernelArgs* a, int nblk, int decode, int sm_count, cudaStream_t st);
cudaError_t zpq_launch_wide(const ZKernelArgs* a, int nblk, int decode, cudaStream_t st);
// v2 (production) kernel: one CTA per SM, `warps` blocks in flight per CTA,
// `pw_bytes` of shared memory per warp for the ICM/ISSE tables.
int zpq_code2_static_smem(void);
int zpq_dec6_static_smem(void);
int zpq_dec6_ctrl_bytes(void);
cudaError_t zpq_launch_dec6(const ZKernelArgs* a, int nblk, int sm_count, int teams, int pw_bytes, unsigned long long sig, int variant, cudaStream_t st);
int zpq_dec7_static_smem(void);
int zpq_dec7_ctrl_bytes(void);
int zpq_dec7_takes(unsigned long long sig);
cudaError_t zpq_launch_dec7(const ZKernelArgs* a, int nblk, int sm_count, int teams, int pw_bytes, unsigned long long sig, cudaStream_t st);
int zpq_dec8_static_smem(void);
int zpq_dec8_xchg_bytes(void);
cudaError_t zpq_launch_dec8(const ZKernelArgs* a, const uint32_t* d_group_first, const uint32_t* d_group_count, int ngroups,
                            int nbk, int pw_bytes, int nwarps, int sm_count, cudaStream_t st);
cudaError_t zpq_launch_sha1(const uint8_t* const* ptr, const uint64_t* len, uint8_t* out, int n, cudaStream_t st);
cudaError_t zpq_launch_gaps(const uint8_t* const* ptr, const uint64_t* len, int32_t* gaps, int n, int sm_count, cudaStream_t st);
cudaError_t zpq_launch_scatter(const uint8_t* pool, const uint64_t* src_off, uint8_t* const* dst, const uint32_t* nbytes, int n, cudaStream_t st);
cudaError_t zpq_launch_wave(const ZWaveArgs* w, cudaStream_t st, int variant);
// zpq_bwt.cu: forward transforms in front of the model; in / out are HOST arrays of DEVICE pointers, len is host memory
uint64_t zpq_bwt_workspace(uint64_t total, int nb);
cudaError_t zpq_bwt_forward(const uint8_t* const* in, const uint64_t* len, uint8_t* const* out, int nb,
                            void* ws, uint64_t ws_bytes, int sm_count, cudaStream_t st, uint64_t* launches, uint64_t* alg_bytes);
// inverse BWT of blocks whose PCOMP program is the BWT post-processor: in[b] = the n[b] + 5 model bytes after the program
uint64_t zpq_ibwt_workspace(uint64_t total_rows, int nb);
cudaError_t zpq_bwt_inverse(const uint8_t* const* in, const uint64_t* n, const uint32_t* idx, uint8_t* const* out, int nb,
                            void* ws, uint64_t ws_bytes, int sm_count, cudaStream_t st, uint64_t* launches);
cudaError_t zpq_suffix_arrays(const uint8_t* const* in, const uint64_t* len, int nb, void* ws, uint64_t ws_bytes, int sm_count,
                              cudaStream_t st, uint64_t* launches, uint64_t* alg_bytes,
                              const uint32_t** sa, const uint32_t** rank, const uint32_t** off);
// zpq_lz77.cu: LZBuffer levels 1 / 2, one thread per block; numbers = makeConfig's $1..$7
int zpq_lz77_uses_sa(const int numbers[7]);
uint64_t zpq_lz77_table_bytes(const int numbers[7]);
uint64_t zpq_lz77_meta_bytes(int nb);
cudaError_t zpq_lz77_forward(const uint8_t* const* in, const uint64_t* len, uint8_t* const* out, const uint64_t* out_cap, int nb,
                             const int numbers[7], void* tables, const uint32_t* sa, const uint32_t* rank, const uint32_t* base,
                             void* meta, uint64_t* out_len, int* status, cudaStream_t st, uint64_t* launches);
uint64_t zpq_unlz_meta_bytes(int nb);
cudaError_t zpq_unlz_forward(const uint8_t* const* in, const uint64_t* in_len, uint8_t* const* out, const uint64_t* out_cap, int nb,
                             int kind, int e8, int rb, int min_match, void* meta, uint64_t* out_len, int* status,
                             cudaStream_t st, uint64_t* launches);
cudaError_t zpq_e8e9_forward(const uint8_t* const* in, uint8_t* const* out, const uint64_t* len, int nb, void* ws,
                             int sm_count, cudaStream_t st, uint64_t* launches);
unsigned long long zpq_code2_signature(const ZFast* F);
cudaError_t zpq_launch_code2(const ZKernelArgs* a, int nblk, int decode, int sm_count, int warps, int pw_bytes,
                             unsigned long long sig, cudaStream_t st);
}
#endif
