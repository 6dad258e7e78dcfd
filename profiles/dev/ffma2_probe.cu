// Development probe: which packed float32 (f32x2) instruction forms and which shared-memory atomic forms run on B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
struct P { float w; const float2* in; float2* out; };
__global__ void k_pair(const __grid_constant__ P p) { float2 v = p.in[threadIdx.x]; f2 r = fma2(pk(v.x, v.y), pk(v.y, v.x), pk(v.x, v.x)); float a, b; upk(r, a, b); p.out[threadIdx.x] = make_float2(a, b); }
__global__ void k_ubc(const __grid_constant__ P p) { float2 v = p.in[threadIdx.x]; f2 r = fma2(pk(v.x, v.y), pk(p.w, p.w), pk(v.y, v.x)); float a, b; upk(r, a, b); p.out[threadIdx.x] = make_float2(a, b); }
__global__ void k_rbc(const __grid_constant__ P p) { float2 v = p.in[threadIdx.x]; float s = v.x * v.y; f2 r = fma2(pk(v.x, v.y), pk(s, s), pk(v.y, v.x)); float a, b; upk(r, a, b); p.out[threadIdx.x] = make_float2(a, b); }
__global__ void k_imm(const __grid_constant__ P p) { float2 v = p.in[threadIdx.x]; f2 r = fma2(pk(v.x, v.y), pk(-1.f, -1.f), pk(v.y, v.x)); float a, b; upk(r, a, b); p.out[threadIdx.x] = make_float2(a, b); }
__global__ void k_add(const __grid_constant__ P p) { float2 v = p.in[threadIdx.x]; f2 r = add2(pk(v.x, v.y), pk(v.y, v.x)); r = sub2(r, pk(v.x, 1.f)); r = mul2(r, pk(v.x, v.y)); r = mul2(r, pk(p.w, p.w)); float a, b; upk(r, a, b); p.out[threadIdx.x] = make_float2(a, b); }
__global__ void k_atom(const __grid_constant__ P p) {
  __shared__ uint32_t cnt[2];
  if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
  __syncthreads();
  uint32_t old = 0, a = (uint32_t)__cvta_generic_to_shared(cnt);
  if ((threadIdx.x & 31) == 0) asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(a + 4u) : "memory");
  __syncthreads();
  p.out[threadIdx.x] = make_float2((float)old, (float)cnt[1]);
}
template <typename F> static void run(const char* name, F f, const P& p) {
  f<<<1, 64>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  float2 h[2]; cudaMemcpy(h, p.out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-8s %s  out[1] = %g %g\n", name, cudaGetErrorString(e), h[1].x, h[1].y);
}
int main() {
  P p; p.w = 3.f; float2 h[64]; for (int i = 0; i < 64; ++i) h[i] = make_float2(1.f + i, 2.f + i);
  float2 *in, *out; cudaMalloc(&in, sizeof(h)); cudaMalloc(&out, sizeof(h)); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  p.in = in; p.out = out;
  run("pair", k_pair, p); run("ubc", k_ubc, p); run("rbc", k_rbc, p); run("imm", k_imm, p); run("addsub", k_add, p); run("atom", k_atom, p);
  return 0;
}
