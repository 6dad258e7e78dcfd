// (second half: effect of a start address that is not 128-byte aligned)
// Issue-to-completion time of back-to-back tcgen05.mma kind::f16 (M = 128, K = 16) for several N: how small an MMA may be
// before the tensor pipe stops scaling.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate_probe umma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ bool elect() {
  uint32_t p; asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(p)); return p != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, {%6, %6, %6, %6}, p;\n\t}"
      :: "r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(0x4008u), "r"(0u) : "memory");
}
template <int N, int REP>
__global__ void k(long long* out, uint32_t a_shift, uint32_t n_acc) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0;
  const uint32_t barp = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barp)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t sbq = __shfl_sync(0xffffffffu, sb >> 4, 0);
  if (warp == 0 && elect()) {
    const long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REP; ++r) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        mma(tmem + (j % n_acc) * 64, sbq + 8 + a_shift + j * 16 + (144u << 16), sbq + 2048 + j * 8 + ((uint32_t)(N * 16 >> 4) << 16), idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barp) : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(barp), "r"(0u) : "memory");
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
template <int N> void run(long long* d, unsigned a_shift = 0, unsigned n_acc = 2) {
  const int REP = 64;
  cudaFuncSetAttribute(k<N, REP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<N, REP><<<1, 128, 65536>>>(d, a_shift, n_acc); cudaDeviceSynchronize();
  k<N, REP><<<1, 128, 65536>>>(d, a_shift, n_acc);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("A start + %u x 16 B, %u accumulator(s) in turn  N=%3d: issue %.1f cycles/MMA, complete %.1f cycles/MMA (%d MMAs, M=128 K=16 f16) -> %.0f MAC/clk   %s\n", a_shift, n_acc, N, h[0] / (16.0 * REP), h[1] / (16.0 * REP),
         16 * REP, 128.0 * N * 16 / (h[1] / (16.0 * REP)), cudaGetErrorString(cudaGetLastError()));
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  run<16>(d); run<32>(d); run<64>(d); run<128>(d); run<256>(d);
  // the same with the A tile starting 1, 3, 4 positions (16 B each) off a 128-byte boundary: shifted implicit-GEMM taps
  // dependent chains: every MMA accumulates into the same TMEM columns (1), or 2 / 4 accumulators in turn
  run<16>(d, 0, 1); run<32>(d, 0, 1); run<64>(d, 0, 1); run<32>(d, 0, 4); run<64>(d, 0, 4);
  run<16>(d, 1); run<32>(d, 1); run<32>(d, 3); run<32>(d, 4); run<64>(d, 1); run<128>(d, 1);
  return 0;
}
