// Development probe: 3-D TMA box loads of small tiles (what wfs_tma.cuh relies on).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct Params { int c0, c1, c2, bw, bh; float* out; CUtensorMap map; };
__global__ void k(const __grid_constant__ Params P) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = (uint64_t*)(sm + 8192);
  uint32_t bar32 = (uint32_t)__cvta_generic_to_shared(bar), dst = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar32) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar32), "r"(P.bw * P.bh * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(&P.map), "r"(bar32), "r"(P.c0), "r"(P.c1), "r"(P.c2) : "memory");
  }
  uint32_t ok = 0;
  for (int it = 0; it < (1 << 22) && !ok; ++it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar32), "r"(0) : "memory");
  const float* t = (const float*)sm;
  for (int i = threadIdx.x; i < P.bw * P.bh; i += blockDim.x) P.out[i] = ok ? t[i] : -1.f;
}
int main(int argc, char** argv) {
  int N = 168, E = 4, bw = argc > 1 ? atoi(argv[1]) : 24, bh = argc > 2 ? atoi(argv[2]) : 17;
  int c0 = argc > 3 ? atoi(argv[3]) : 5, c1 = argc > 4 ? atoi(argv[4]) : 7, c2 = argc > 5 ? atoi(argv[5]) : 2;
  float* h = (float*)malloc((size_t)E * N * N * 4);
  for (int i = 0; i < E * N * N; ++i) h[i] = (float)i;
  float *d, *out;
  cudaMalloc(&d, (size_t)E * N * N * 4); cudaMemcpy(d, h, (size_t)E * N * N * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 4096 * 4);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  Params P; P.c0 = c0; P.c1 = c1; P.c2 = c2; P.bw = bw; P.bh = bh; P.out = out;
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)E}, strides[2] = {(cuuint64_t)N * 4, (cuuint64_t)N * N * 4};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = ((enc_fn)fn)(&P.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d box %dx%d at (%d,%d,%d)\n", (int)r, bw, bh, c0, c1, c2);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  k<<<1, 32, 16384>>>(P);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    float* ho = (float*)malloc(4096 * 4); cudaMemcpy(ho, out, 4096 * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < bh; ++rr) for (int cc = 0; cc < bw; ++cc) {
      float want = (c0 + cc < N && c1 + rr < N) ? (float)((size_t)c2 * N * N + (size_t)(c1 + rr) * N + c0 + cc) : 0.f;
      if (ho[rr * bw + cc] != want) { if (bad < 5) printf("mismatch (%d,%d): %f vs %f\n", rr, cc, ho[rr * bw + cc], want); ++bad; }
    }
    printf("mismatches: %d\n", bad);
  }
  return 0;
}
