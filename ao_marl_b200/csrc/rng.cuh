// Counter-based RNG and deterministic float32 transforms shared by every kernel.
//
// Same arithmetic, operation for operation, as oracle/rng.py: Philox4x32-10 plus log / sincos / exp
// written with ONE IEEE rounding per step (no FMA contraction: every product/sum goes through
// aom_mul / aom_add, which are __fmul_rn / __fadd_rn on the device), so the GPU and the numpy
// oracle agree bit for bit and Poisson photon counts agree as integers.
// Replaces the cuRAND streams hidden behind the reference's Atmos.set_seed / Sensors.set_noise
// (shesha/supervisor/components/atmosCompass.py:137-145, wfsCompass.py:297-310, 345-350).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define AOM_HD __host__ __device__ __forceinline__
#else
#define AOM_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define aom_mul(a, b) __fmul_rn((a), (b))
#define aom_add(a, b) __fadd_rn((a), (b))
#define aom_sub(a, b) __fsub_rn((a), (b))
#define aom_div(a, b) __fdiv_rn((a), (b))
#define aom_sqrt(a) __fsqrt_rn((a))
#else
// host build (tests/cpu_kernels harness): compile with -ffp-contract=off
static inline float aom_mul(float a, float b) { volatile float r = a * b; return r; }
static inline float aom_add(float a, float b) { volatile float r = a + b; return r; }
static inline float aom_sub(float a, float b) { volatile float r = a - b; return r; }
static inline float aom_div(float a, float b) { volatile float r = a / b; return r; }
static inline float aom_sqrt(float a) { volatile float r = sqrtf(a); return r; }
#endif

#define AOM_TAG_ATMOS 1u
#define AOM_TAG_WFS 2u
#define AOM_TAG_ACTOR 4u
#define AOM_POISSON_SWITCH 30.0f
#define AOM_POISSON_MAXK 200

struct aom_u4 { uint32_t x, y, z, w; };

AOM_HD void aom_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  hi = __umulhi(a, b);
  lo = a * b;
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
#endif
}

AOM_HD aom_u4 aom_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    aom_mulhilo(0xD2511F53u, c0, hi0, lo0);
    aom_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    if (r < 9) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  }
  aom_u4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

AOM_HD float aom_u01(uint32_t x) {
  return aom_mul(aom_add((float)(x >> 8), 0.5f), 5.9604644775390625e-08f);  // 2^-24
}

AOM_HD float aom_as_float(int32_t i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  union { int32_t i; float f; } u; u.i = i; return u.f;
#endif
}
AOM_HD int32_t aom_as_int(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  union { int32_t i; float f; } u; u.f = f; return u.i;
#endif
}

AOM_HD float aom_det_log(float x) {
  int32_t bits = aom_as_int(x);
  int32_t e = ((bits >> 23) & 0xFF) - 127;
  float m = aom_as_float((bits & 0x007FFFFF) | 0x3F800000);
  if (m > 1.41421354f) { m = aom_mul(m, 0.5f); e += 1; }
  float f = aom_sub(m, 1.0f);
  float s = aom_div(f, aom_add(2.0f, f));
  float z = aom_mul(s, s);
  float p = (float)(2.0 / 9.0);
  p = aom_add(aom_mul(p, z), (float)(2.0 / 7.0));
  p = aom_add(aom_mul(p, z), (float)(2.0 / 5.0));
  p = aom_add(aom_mul(p, z), (float)(2.0 / 3.0));
  float logm = aom_add(aom_mul(2.0f, s), aom_mul(aom_mul(s, z), p));
  return aom_add(aom_mul((float)e, 0.6931471805599453f), logm);
}

AOM_HD void aom_det_sincos2pi(float u, float& c_out, float& s_out) {
  float t = aom_mul(u, 4.0f);
  float q = floorf(aom_add(t, 0.5f));
  float r = aom_sub(t, q);
  float a = aom_mul(r, 1.5707963267948966f);
  float a2 = aom_mul(a, a);
  float ps = (float)(1.0 / 362880.0);
  ps = aom_add(aom_mul(ps, a2), (float)(-1.0 / 5040.0));
  ps = aom_add(aom_mul(ps, a2), (float)(1.0 / 120.0));
  ps = aom_add(aom_mul(ps, a2), (float)(-1.0 / 6.0));
  float s = aom_add(a, aom_mul(aom_mul(a, a2), ps));
  float pc = (float)(-1.0 / 3628800.0);
  pc = aom_add(aom_mul(pc, a2), (float)(1.0 / 40320.0));
  pc = aom_add(aom_mul(pc, a2), (float)(-1.0 / 720.0));
  pc = aom_add(aom_mul(pc, a2), (float)(1.0 / 24.0));
  pc = aom_add(aom_mul(pc, a2), -0.5f);
  float c = aom_add(1.0f, aom_mul(a2, pc));
  int qi = ((int)q) & 3;
  c_out = (qi == 0) ? c : (qi == 1) ? -s : (qi == 2) ? -c : s;
  s_out = (qi == 0) ? s : (qi == 1) ? c : (qi == 2) ? -s : -c;
}

AOM_HD float aom_det_exp(float x) {
  float n = rintf(aom_mul(x, 1.4426950408889634f));
  float r = aom_sub(x, aom_mul(n, 0.693145751953125f));
  r = aom_sub(r, aom_mul(n, 1.42860682030941723212e-6f));
  float p = (float)(1.0 / 720.0);
  p = aom_add(aom_mul(p, r), (float)(1.0 / 120.0));
  p = aom_add(aom_mul(p, r), (float)(1.0 / 24.0));
  p = aom_add(aom_mul(p, r), (float)(1.0 / 6.0));
  p = aom_add(aom_mul(p, r), 0.5f);
  p = aom_add(aom_mul(p, r), 1.0f);
  p = aom_add(aom_mul(p, r), 1.0f);
  float scale = aom_as_float((((int32_t)n) + 127) << 23);
  return aom_mul(p, scale);
}

// Box-Muller on two words -> two standard normals
AOM_HD void aom_normal_pair(uint32_t x0, uint32_t x1, float& z0, float& z1) {
  float u1 = aom_u01(x0), u2 = aom_u01(x1);
  float r = aom_sqrt(aom_mul(-2.0f, aom_det_log(u1)));
  float c, s;
  aom_det_sincos2pi(u2, c, s);
  z0 = aom_mul(r, c);
  z1 = aom_mul(r, s);
}

// element `j` (0..3) of the 4 normals produced by one Philox block
AOM_HD float aom_normal_of_block(const aom_u4& w, int j) {
  float z0, z1;
  if (j < 2) aom_normal_pair(w.x, w.y, z0, z1); else aom_normal_pair(w.z, w.w, z0, z1);
  return (j & 1) ? z1 : z0;
}

// uniform for the CDF inversion: ((x >> 9) + 0.5) * 2^-23, exact in float32 and strictly inside (0, 1)
// (aom_u01's (x >> 8) + 0.5 rounds to 2^24 for the largest words, i.e. u == 1.0, which no float32 CDF reaches)
AOM_HD float aom_u01_open(uint32_t x) {
  return aom_mul(aom_add((float)(x >> 9), 0.5f), 1.1920928955078125e-07f);  // 2^-23
}

AOM_HD int32_t aom_poisson(float lam, uint32_t x0, uint32_t x1) {
  if (!(lam > 0.0f)) return 0;
  if (lam < AOM_POISSON_SWITCH) {
    float u = aom_u01_open(x0);
    float p = aom_det_exp(-lam);
    float F = p;
    int k = 0;
    while (u > F && k < AOM_POISSON_MAXK) {
      k += 1;
      p = aom_div(aom_mul(p, lam), (float)k);
      const float Fn = aom_add(F, p);
      // the float32 CDF has stopped growing beyond the mode: this k is the tail sample (never the loop cap)
      if (Fn == F && (float)k > lam) break;
      F = Fn;
    }
    return k;
  }
  float z0, z1;
  aom_normal_pair(x0, x1, z0, z1);
  float v = aom_add(lam, aom_mul(aom_sqrt(lam), z0));
  v = floorf(aom_add(v, 0.5f));
  return v > 0.0f ? (int32_t)v : 0;
}

// photon + read noise on one detector pixel (noise < 0: unchanged)
AOM_HD float aom_pixel_noise(float lam, float noise, uint32_t pixel_index, uint32_t frame, uint32_t wfs,
                             uint32_t k0, uint32_t k1) {
  if (noise < 0.0f) return lam;
  aom_u4 w = aom_philox(pixel_index, frame, AOM_TAG_WFS, wfs, k0, k1);
  float cnt = (float)aom_poisson(lam, w.x, w.y);
  if (noise > 0.0f) {
    float z0, z1;
    aom_normal_pair(w.z, w.w, z0, z1);
    cnt = aom_add(cnt, aom_mul(noise, z0));
  }
  return cnt;
}
