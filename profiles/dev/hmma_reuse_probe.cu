// Development probe: HMMA.16816.F32 rate on sm_100a as a function of operand-register reuse between consecutive
// MMAs (8 independent accumulators; all 16 warps of every SM issue MMAs only).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void mma(float (&c)[4], const uint4& a, const uint2& b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y));
}
// MODE 0: same A, same B | 1: same A, distinct B | 2: distinct A, same B | 3: distinct A, distinct B
// MODE 4: A shared by consecutive pairs, B distinct | 5: A shared by groups of 4, B distinct
template <int MODE>
__global__ void k(int iters, const uint4* src, float* out, long long* cyc) {
  float acc[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  uint4 A[8]; uint2 B[8];
  for (int i = 0; i < 8; ++i) { A[i] = src[threadIdx.x * 8 + i]; uint4 t = src[4096 + threadIdx.x * 8 + i]; B[i] = make_uint2(t.x, t.y); }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ia = (MODE == 0 || MODE == 1) ? 0 : (MODE == 4 ? i / 2 : (MODE == 5 ? i / 4 : i));
      const int ib = (MODE == 0 || MODE == 2) ? 0 : i;
      mma(acc[i], A[ia], B[ib]);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps, const uint4* src, float* out, long long* cyc) {
  const int iters = 2000;
  k<MODE><<<148, warps * 32>>>(iters, src, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s warps/SM %2d: %s  %.2f cycles per HMMA per scheduler\n", name, warps, cudaGetErrorString(e),
         (double)h / (iters * 8.0) / (warps / 4.0));
}
int main() {
  uint4* src; float* out; long long* cyc;
  cudaMalloc(&src, 8192 * 16 * 2); cudaMemset(src, 0x3c, 8192 * 16 * 2);
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int warps : {4, 16}) {
    run<0>("same A, same B", warps, src, out, cyc);
    run<1>("same A, distinct B", warps, src, out, cyc);
    run<2>("distinct A, same B", warps, src, out, cyc);
    run<3>("distinct A, distinct B", warps, src, out, cyc);
    run<4>("A shared by pairs, distinct B", warps, src, out, cyc);
    run<5>("A shared by groups of 4, distinct B", warps, src, out, cyc);
  }
  return 0;
}
