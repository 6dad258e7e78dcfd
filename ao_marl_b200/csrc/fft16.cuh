// Pruned small DFT used by the Shack-Hartmann kernel.
//
// A subaperture holds pd = 16 phase points per axis, zero-padded to Nfft = 16*R (R = 4 -> 64,
// R = 8 -> 128); the detector only reads the central Nfft/2 frequencies (k in [0, Nfft/4) and
// [3Nfft/4, Nfft)), because `binmap` (reference geom_init.py:731-758) covers npix*nrebin = Nfft/2
// high-resolution pixels per axis.  With k = R*a + b:
//       X[R a + b] = sum_n (x[n] W_N^{b n}) W_16^{a n}  =  FFT16(x . W_N^{b .})[a]
// and the wanted k are exactly a in {0,1,2,3,12,13,14,15} for every b.  One lane computes one
// (vector, b): 15 twiddle products, a radix-4 first stage (16 outputs) and the 8 wanted outputs of
// the second radix-4 stage.  Forward sign exp(-2 pi i k n / N), as numpy.fft / cuFFT forward.
#pragma once
#include "rng.cuh"
#include "twiddles.cuh"

#define AOM_C_PI 3.14159265358979323846

// W_16^k = exp(-2 pi i k / 16), k = 1, 2, 3, 6, 9 (k = 4 is -i and handled by swaps)
#define AOM_W16_C1 0.92387953251128674f
#define AOM_W16_S1 0.38268343236508977f
#define AOM_W16_C2 0.70710678118654752f

// (ar + i ai) * (c - i s)
AOM_HD void aom_cmul_conj(float ar, float ai, float c, float s, float& orr, float& oi) {
  orr = ar * c + ai * s;
  oi = ai * c - ar * s;
}

// Radix-4 x radix-4 16-point DFT of u, pruned to the outputs a in {0,1,2,3,12,13,14,15}:
// out[o], o = q      -> a = q       (non-negative frequencies)
//         o = 4 + q  -> a = 12 + q  (negative frequencies)
AOM_HD void aom_fft16_core(const float* ur, const float* ui, float* outr, float* outi) {
  float yr[4][4], yi[4][4];  // [q][m]
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    float s02r = ur[m] + ur[m + 8], s02i = ui[m] + ui[m + 8];
    float d02r = ur[m] - ur[m + 8], d02i = ui[m] - ui[m + 8];
    float s13r = ur[m + 4] + ur[m + 12], s13i = ui[m + 4] + ui[m + 12];
    float d13r = ur[m + 4] - ur[m + 12], d13i = ui[m + 4] - ui[m + 12];
    yr[0][m] = s02r + s13r; yi[0][m] = s02i + s13i;
    yr[2][m] = s02r - s13r; yi[2][m] = s02i - s13i;
    yr[1][m] = d02r + d13i; yi[1][m] = d02i - d13r;   // d02 - i d13
    yr[3][m] = d02r - d13i; yi[3][m] = d02i + d13r;   // d02 + i d13
  }
  // twiddles W_16^{m q}
  float tr, ti;
  // m = 1: q = 1,2,3 -> W^1, W^2, W^3
  aom_cmul_conj(yr[1][1], yi[1][1], AOM_W16_C1, AOM_W16_S1, tr, ti); yr[1][1] = tr; yi[1][1] = ti;
  aom_cmul_conj(yr[2][1], yi[2][1], AOM_W16_C2, AOM_W16_C2, tr, ti); yr[2][1] = tr; yi[2][1] = ti;
  aom_cmul_conj(yr[3][1], yi[3][1], AOM_W16_S1, AOM_W16_C1, tr, ti); yr[3][1] = tr; yi[3][1] = ti;
  // m = 2: q = 1,2,3 -> W^2, W^4 = -i, W^6 = (-c2, -c2) i.e. cos = -c2, sin = +c2
  aom_cmul_conj(yr[1][2], yi[1][2], AOM_W16_C2, AOM_W16_C2, tr, ti); yr[1][2] = tr; yi[1][2] = ti;
  tr = yi[2][2]; ti = -yr[2][2]; yr[2][2] = tr; yi[2][2] = ti;          // * (-i)
  aom_cmul_conj(yr[3][2], yi[3][2], -AOM_W16_C2, AOM_W16_C2, tr, ti); yr[3][2] = tr; yi[3][2] = ti;
  // m = 3: q = 1,2,3 -> W^3, W^6, W^9 = (cos = -c1, sin = -s1)
  aom_cmul_conj(yr[1][3], yi[1][3], AOM_W16_S1, AOM_W16_C1, tr, ti); yr[1][3] = tr; yi[1][3] = ti;
  aom_cmul_conj(yr[2][3], yi[2][3], -AOM_W16_C2, AOM_W16_C2, tr, ti); yr[2][3] = tr; yi[2][3] = ti;
  aom_cmul_conj(yr[3][3], yi[3][3], -AOM_W16_C1, -AOM_W16_S1, tr, ti); yr[3][3] = tr; yi[3][3] = ti;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    outr[q] = (yr[q][0] + yr[q][2]) + (yr[q][1] + yr[q][3]);
    outi[q] = (yi[q][0] + yi[q][2]) + (yi[q][1] + yi[q][3]);
    // a = 12 + q: sum_m W_4^{3 m} y[m] = (y0 - y2) + i (y1 - y3)
    outr[4 + q] = (yr[q][0] - yr[q][2]) - (yi[q][1] - yi[q][3]);
    outi[4 + q] = (yi[q][0] - yi[q][2]) + (yr[q][1] - yr[q][3]);
  }
}

// x . W_N^{b n} with the twiddles read from memory (row pass: b differs between lanes), then the core.
// twr/twi[16] hold the complex value W_N^{b n} = (cos, -sin).
AOM_HD void aom_fft16_pruned(const float* xr, const float* xi, const float* twr, const float* twi,
                             float* outr, float* outi) {
  float ur[16], ui[16];
  ur[0] = xr[0];
  ui[0] = xi[0];
#pragma unroll
  for (int n = 1; n < 16; ++n) {
    ur[n] = xr[n] * twr[n] - xi[n] * twi[n];
    ui[n] = xr[n] * twi[n] + xi[n] * twr[n];
  }
  aom_fft16_core(ur, ui, outr, outi);
}

// Same with a compile-time phase class B of an N = 16 R point transform: the twiddles fold into immediates
// (column pass: every lane of the warp works on the same b).
template <int R, int B>
AOM_HD void aom_fft16_pruned_const(const float* xr, const float* xi, float* outr, float* outi) {
  if (B == 0) {
    aom_fft16_core(xr, xi, outr, outi);
    return;
  }
  float ur[16], ui[16];
  ur[0] = xr[0];
  ui[0] = xi[0];
#pragma unroll
  for (int n = 1; n < 16; ++n) {
    const float c = aom_tw_c((128 / (16 * R)) * B * n), s = aom_tw_s((128 / (16 * R)) * B * n);
    ur[n] = xr[n] * c + xi[n] * s;       // (xr + i xi)(c - i s)
    ui[n] = xi[n] * c - xr[n] * s;
  }
  aom_fft16_core(ur, ui, outr, outi);
}
