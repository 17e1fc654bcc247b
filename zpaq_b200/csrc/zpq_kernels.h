// zpq_kernels.h -- launch entry points of zpq_kernels.cu (internal).
#ifndef ZPQ_KERNELS_H
#define ZPQ_KERNELS_H
#include <cuda_runtime.h>
#include "zpq_dev.h"

#define ZPQ_WARPS 4   // warps (= archive blocks in flight) per CTA

extern "C" {
size_t zpq_code_kernel_smem(void);
cudaError_t zpq_launch_init(const ZKernelArgs* a, int nblk, cudaStream_t st);
cudaError_t zpq_launch_code(const ZKernelArgs* a, int nblk, int decode, int sm_count, cudaStream_t st);
}
#endif
