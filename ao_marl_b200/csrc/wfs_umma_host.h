// Host interface of the tcgen05 Shack-Hartmann frame kernel (wfs_umma.cuh), compiled in its own translation unit
// (wfs_umma.cu) and linked into libaomarl.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "wfs_umma.cuh"

struct WfsUmmaHost {
  WfsUmmaTables f;
  CUtensorMap maps[WU_MAX_LAYERS];
};

// Launches wfs_frame_umma_kernel for the frame described by p.  full != 0: fp32-grade products in both stages.
// Returns cudaSuccess or the error of the attribute / launch call.
// ws != 0: the warp-specialised kernel (wfs_umma_ws.cuh) on the same tables.
cudaError_t wfs_umma_launch(const WfsParams& p, const WfsUmmaHost& h, int num_sms, int full, int ws, cudaStream_t st);
