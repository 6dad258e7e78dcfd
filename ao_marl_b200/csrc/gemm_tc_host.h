// Host side of the pre-tiled B operand of gemm_tc_kernel (gemm_tc.cuh): plain C++, no CUDA, so that the CPU test harness
// (tests/cpu_kernels/harness.cpp) checks the same code the library runs at aom_set_table.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifndef GTC_BN
#define GTC_BN 128
#define GTC_BK 16
#endif

// Host: TF32 hi / lo planes of a K-major operator [batch][rows][ldb] (K valid columns) in the tile order gemm_tc_kernel
// reads with one bulk copy per stage: [batch][n tile of 128][k block of 16][hi, lo][128 x 16 floats, 8-row x 16-byte core
// matrices].  Rows beyond `rows` and columns beyond K are zero.  Returns the number of floats per batch element.
static inline size_t gtc_pretile_floats(int n_cols_ld, int K) {
  const size_t NT = (size_t)(n_cols_ld + GTC_BN - 1) / GTC_BN, nkb = (size_t)(K + GTC_BK - 1) / GTC_BK;
  return NT * nkb * 2 * GTC_BN * GTC_BK;
}
static inline void gtc_pretile_host(const float* W, int batch, long long sB, int ldb, int rows, int n_cols_ld, int K, float* out) {
  const int NT = (n_cols_ld + GTC_BN - 1) / GTC_BN, nkb = (K + GTC_BK - 1) / GTC_BK;
  const size_t per = gtc_pretile_floats(n_cols_ld, K);
  for (int b = 0; b < batch; ++b)
    for (int nt = 0; nt < NT; ++nt)
      for (int kb = 0; kb < nkb; ++kb) {
        float* hi = out + (size_t)b * per + ((size_t)(nt * nkb + kb) * 2) * (GTC_BN * GTC_BK);
        float* lo = hi + GTC_BN * GTC_BK;
        for (int r = 0; r < GTC_BN; ++r)
          for (int kk = 0; kk < GTC_BK; ++kk) {
            const int row = nt * GTC_BN + r, k = kb * GTC_BK + kk;
            const float x = (row < rows && k < K) ? W[(size_t)b * sB + (size_t)row * ldb + k] : 0.f;
            uint32_t u;
            memcpy(&u, &x, 4);
            u &= 0xFFFFE000u;
            float h;
            memcpy(&h, &u, 4);
            const size_t off = (size_t)((r >> 3) * 512 + (kk >> 2) * 128 + (r & 7) * 16 + (kk & 3) * 4) / 4;
            hi[off] = h;
            lo[off] = x - h;
          }
      }
}

