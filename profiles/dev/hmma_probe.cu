// Development probe: issue cost of legacy mma.sync (HMMA.16816.F32) on sm_100a and how it shares a scheduler
// with FFMA work.  Per SM: W_MMA warps run N independent-chain HMMAs, W_FMA warps run FFMA chains.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int ilp>
__global__ void k(int w_mma, int iters, int distinct, float* out, long long* cyc) {
  const int warp = threadIdx.x >> 5;
  float acc[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float f[16];
  for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 0.001f + i;
  uint32_t a = 0x3c003c00u + threadIdx.x, b = 0x38003800u;
  uint32_t av[16], bv[8];
  for (int i = 0; i < 16; ++i) av[i] = a + 7 * i + (distinct ? i * threadIdx.x : 0);
  for (int i = 0; i < 8; ++i) bv[i] = b + 3 * i + (distinct ? i * threadIdx.x : 0);
  __syncthreads();
  long long t0 = clock64();
  if (warp < w_mma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (distinct) mma(acc[i], av[(4 * i) & 15], av[(4 * i + 1) & 15], av[(4 * i + 2) & 15], av[(4 * i + 3) & 15], bv[(2 * i) & 7], bv[(2 * i + 1) & 7]);
        else mma(acc[i], a, a, a, a, b, b);
      }
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i % ilp] = fmaf(f[i % ilp], 1.0001f, 0.5f);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  for (int i = 0; i < 16; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + warp] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 32 * 8);
  long long h[32];
  const int iters = 2000;
  // configurations: (warps per CTA, of which MMA warps); warp w runs on scheduler w % 4
  int cfg[][2] = {{4, 4}, {8, 8}, {16, 16}, {4, 0}, {8, 0}, {16, 0}, {8, 4}, {16, 4}, {16, 8}, {12, 4}, {16, 12}};
  for (int mode = 0; mode < 4; ++mode)
  for (auto& c : cfg) {
    const int W = c[0], wm = c[1];
    const int distinct = mode & 1, lowilp = mode >> 1;
    if (c == cfg[0]) printf("---- distinct operand registers: %d, FFMA chains per warp: %d\n", distinct, lowilp ? 2 : 16);
    if (lowilp) k<2><<<148, W * 32>>>(wm, iters, distinct, out, cyc);
    else k<16><<<148, W * 32>>>(wm, iters, distinct, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, W * 8, cudaMemcpyDeviceToHost);
    double cm = 0, cf = 0;
    for (int w = 0; w < W; ++w) (w < wm ? cm : cf) += (double)h[w];
    if (wm) cm /= wm;
    if (W - wm) cf /= (W - wm);
    // per scheduler: wm/4 MMA warps, (W-wm)/4 FMA warps
    printf("warps %2d (mma %2d): %s  mma warp: %.1f cyc/HMMA  (%.2f cyc/HMMA per scheduler)   fma warp: %.2f cyc/FFMA (%.2f per scheduler)\n",
           W, wm, cudaGetErrorString(e), wm ? cm / (iters * 8.0) : 0.0, wm ? cm / (iters * 8.0) / (wm / 4.0) : 0.0,
           (W - wm) ? cf / (iters * 64.0) : 0.0, (W - wm) ? cf / (iters * 64.0) / ((W - wm) / 4.0) : 0.0);
  }
  return 0;
}
