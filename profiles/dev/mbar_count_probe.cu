// Development probe: mbarrier.arrive returns the state before the arrival; mbarrier.pending_count of that state tells
// the last arriver (pending == 1) without a separate atomic counter.  8 warps arrive on a count-8 barrier, 4 rounds.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint32_t* out) {
  __shared__ uint64_t bar;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(b) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int round = 0; round < 4; ++round) {
    // stagger the warps differently every round
    const int delay = ((warp * 5 + round * 3) & 7) * 2000;
    const long long t0 = clock64();
    while (clock64() - t0 < delay) {}
    uint32_t pend = 0;
    if (lane == 0) {
      uint64_t st;
      asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(b) : "memory");
      asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pend) : "l"(st));
    }
    pend = __shfl_sync(0xffffffffu, pend, 0);
    uint32_t ok = 2;
    if (pend == 1 && lane == 0) {   // last arriver: the phase it completed must test as complete at once
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(b), "r"((uint32_t)(round & 1)) : "memory");
    }
    if (lane == 0) out[(round * 8 + warp) * 2] = pend, out[(round * 8 + warp) * 2 + 1] = ok;
    __syncthreads();
  }
}
int main() {
  uint32_t* d; cudaMalloc(&d, 64 * 4); cudaMemset(d, 0xff, 64 * 4);
  k<<<1, 256>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  uint32_t h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("kernel: %s\n", cudaGetErrorString(e));
  for (int r = 0; r < 4; ++r) { printf("round %d pending:", r); for (int w = 0; w < 8; ++w) printf(" %u%s", h[(r * 8 + w) * 2], h[(r * 8 + w) * 2 + 1] == 1 ? "*ok" : h[(r * 8 + w) * 2 + 1] == 0 ? "*FAIL" : ""); printf("\n"); }
  return 0;
}
