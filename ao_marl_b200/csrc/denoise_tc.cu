// Translation unit of the tcgen05 denoiser (denoise_tc.cuh).
#define DT_DEFINE_KERNELS
#include "denoise_tc.cuh"
#include "denoise_tc_host.h"

cudaError_t denoise_tc_launch(const DnTcParams& P, int num_sms, cudaStream_t st) {
  if (P.n_spots <= 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(denoise_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  const long long passes = (P.n_spots + DT_G - 1) / DT_G;
  const int grid = (int)(passes < num_sms ? passes : num_sms);
  denoise_tc_kernel<<<grid, DT_THREADS, DT_SMEM_BYTES, st>>>(P);
  return cudaGetLastError();
}
