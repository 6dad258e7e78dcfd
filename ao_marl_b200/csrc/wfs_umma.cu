// Translation unit of the tcgen05 Shack-Hartmann frame kernel: instantiations per layer count + launch geometry.
#include <string.h>
#include "wfs_umma_host.h"
#include "wfs_umma_ws.cuh"

template <int NL, int FULL, int WS>
static cudaError_t launch_t(const WfsParams& p, const WfsUmmaHost& h, int num_sms, cudaStream_t st) {
  const long long total = (long long)p.E * p.nvalid;
  // persistent CTAs over contiguous ranges of work items: two CTAs per SM, about 4 waves, at least 8 iterations each
  long long grid = (long long)num_sms * 2 * 4;
  long long ipc = (total + grid - 1) / grid;
  if (ipc < 8 * WU_WARPS) ipc = 8 * WU_WARPS;
  ipc = (ipc + WU_WARPS - 1) / WU_WARPS * WU_WARPS;
  grid = (total + ipc - 1) / ipc;
  WfsUmmaParams P;
  memset(&P, 0, sizeof(P));
  P.p = p;
  P.f = h.f;
  P.f.items_per_cta = ipc;
  for (int l = 0; l < NL; ++l) P.maps[l] = h.maps[l];
  const size_t smem = wu_smem_bytes<NL>(P.f.GW);
  if (WS) {
    cudaError_t e = cudaFuncSetAttribute(wfs_frame_ws_kernel<NL, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    wfs_frame_ws_kernel<NL, FULL><<<(unsigned)grid, WS_THREADS, smem, st>>>(P);
  } else {
    cudaError_t e = cudaFuncSetAttribute(wfs_frame_umma_kernel<NL, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    wfs_frame_umma_kernel<NL, FULL><<<(unsigned)grid, WU_WARPS * 32, smem, st>>>(P);
  }
  return cudaGetLastError();
}

cudaError_t wfs_umma_launch(const WfsParams& p, const WfsUmmaHost& h, int num_sms, int full, int ws, cudaStream_t st) {
#define WU_GO(NL) return ws ? launch_t<NL, 1, 1>(p, h, num_sms, st) : full ? launch_t<NL, 1, 0>(p, h, num_sms, st) : launch_t<NL, 0, 0>(p, h, num_sms, st)
  switch (p.n_layers) {
    case 0: WU_GO(0);
    case 1: WU_GO(1);
    case 2: WU_GO(2);
    case 3: WU_GO(3);
    case 4: WU_GO(4);
  }
#undef WU_GO
  return cudaErrorInvalidValue;
}
