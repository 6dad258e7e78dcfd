// Host entry of the tcgen05 denoiser (denoise_tc.cu).
#pragma once
#include "denoise_tc.cuh"
cudaError_t denoise_tc_launch(const DnTcParams& P, int num_sms, cudaStream_t st);
