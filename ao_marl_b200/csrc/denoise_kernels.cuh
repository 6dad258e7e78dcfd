// denoise_kernels.cuh -- per-subaperture convolutional denoiser of the *_d0_noise configurations.
//
// Reference: DenoisingAutoencoderCNN2DSingleSubapeture (src/autoencoder/autoencoder_models.py:130-197, MSE variant) applied
// to every 16x16 spot of the frame by RlSupervisor.autoencoder_denoising (shesha/supervisor/rlSupervisor.py:876-891):
//   conv3x3(1->16) relu pool2 | conv3x3(16->32) relu pool2 | conv3x3(32->64) relu |
//   convT4x4s2(64->32) relu | convT4x4s2(32->16) relu | convT3x3(16->1)
// 1.71 M multiply-adds per spot, 64 449 parameters shared by all spots.
//
// One fused kernel: a CTA of 384 threads carries DN_G = 6 spots through all six layers with the activations in shared
// memory (two ping-pong buffers per spot, zero borders instead of bounds checks); nothing but the 1 KB spot and the
// 1 KB result touches HBM.  (Measured on B200: 4 spots / 256 threads / 233 registers 34.3 TFLOP/s, 6 / 384 / 167
// 37.8 TFLOP/s, 7 / 448 / 128 with spills 34.2 TFLOP/s.)  Arithmetic is plain float32 FFMA -- the result has to match the reference's float32 module
// (tests/golden/ref_autoencoder.npz), and the spots are raw photo-electron counts, so no reduced-precision operand
// format is safe without a per-spot scale.  Every thread owns a register tile (2 rows x 4 columns x CB channels); the
// threads of a layer are indexed (tile position, spot, channel group) with the channel group slowest, so the
// lanes of a warp read the same weights: weight loads are warp-uniform broadcasts from L1/L2 in the prepacked layout
// [input channel][tap][output channel] (aom_table AOM_T_DENOISER, packed by ao_marl_b200/denoiser.py::pack_weights).
// A stride-2 transposed 4x4 convolution is four independent 2x2 convolutions, one per output parity class:
//   out[2i+py][2j+px] = sum_{a,b in {0,1}} W[1-py+2a][1-px+2b] . in[i+py-a][j+px-b].
#pragma once
#include <cuda_runtime.h>

#define DN_G 6
#define DN_THREADS 384
// Activation layouts [channel][row][pitch] and the distance between the spots of a batch, chosen so that the 32 lanes
// of a warp (tile positions x spots) hit 32 different banks when they read their input windows (model of the address
// patterns: profiles/dev/denoise_banks.py; the first version with pitch = side + 2 and strides that were multiples of
// 32 replayed 77 % of its shared-memory wavefronts):
//   buffer A: x [18][18] stride 325 -> e2p [32][6][6] stride 1153 -> d1 [32][10][12] stride 3841
//   buffer B: e1p [16][10][12] stride 1921 -> e3 [64][6][6] stride 2312 -> d2 [16][16][16] stride 4100
#define DN_X_S 325
#define DN_E1_P 12
#define DN_E1_S 1921
#define DN_E2_S 1153
#define DN_E3_S 2312
#define DN_D1_P 12
#define DN_D1_S 3841
#define DN_D2_S 4100
#define DN_BUF_A (DN_G * DN_D1_S)
#define DN_BUF_B (DN_G * DN_D2_S)
#define DN_SMEM_BYTES ((DN_BUF_A + DN_BUF_B) * 4)

// offsets (floats) inside the packed parameter block
#define DN_E1W 0
#define DN_E1B (DN_E1W + 1 * 9 * 16)
#define DN_E2W (DN_E1B + 16)
#define DN_E2B (DN_E2W + 16 * 9 * 32)
#define DN_E3W (DN_E2B + 32)
#define DN_E3B (DN_E3W + 32 * 9 * 64)
#define DN_D1W (DN_E3B + 64)
#define DN_D1B (DN_D1W + 64 * 16 * 32)
#define DN_D2W (DN_D1B + 32)
#define DN_D2B (DN_D2W + 32 * 16 * 16)
#define DN_D3W (DN_D2B + 16)
#define DN_D3B (DN_D3W + 16 * 9)
#define DN_PARAM_FLOATS (((DN_D3B + 1) + 3) & ~3)

template <int CB>
__device__ __forceinline__ void dn_load_w(const float* __restrict__ w, float (&v)[CB]) {
  if (CB == 8) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w)), b = __ldg(reinterpret_cast<const float4*>(w) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else if (CB == 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else {
    const float2 a = __ldg(reinterpret_cast<const float2*>(w));
    v[0] = a.x; v[1] = a.y;
  }
}

// 3x3 same-padding convolution + relu (+ 2x2 max-pool).  src [G][CIN][H+2][SP] (zero border), dst padded the same way:
// [COUT][HO+2][DP], HO = H/2 when pooling, H otherwise.  The weights of the next input channel are fetched while the
// current one is multiplied (two register sets, the loop is unrolled by two so that their index is static).
template <int CIN, int COUT, int H, int CB, bool POOL, int SP, int DP>
__device__ __forceinline__ void dn_conv3(const float* __restrict__ src, int src_stride, float* __restrict__ dst,
                                         int dst_stride, const float* __restrict__ w, const float* __restrict__ bias) {
  constexpr int TX = H / 4, PT = (H / 2) * TX, NCG = COUT / CB, SC = (H + 2) * SP, HO = POOL ? H / 2 : H, DC = (HO + 2) * DP;
  static_assert(DN_G * PT * NCG == DN_THREADS, "layer does not fill the CTA");
  const int q = threadIdx.x;
  const int pt = q % PT, s = (q / PT) % DN_G, cg = q / (PT * DN_G);
  const int ty = (pt / TX) * 2, tx = (pt % TX) * 4;
  float acc[CB][2][4];
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    const float b = __ldg(bias + cg * CB + c);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int x = 0; x < 4; ++x) acc[c][r][x] = b;
  }
  const float* in0 = src + s * src_stride + ty * SP + tx;
  const float* wp = w + cg * CB;
  float wv[2][9][CB];
#pragma unroll
  for (int t = 0; t < 9; ++t) dn_load_w<CB>(wp + t * COUT, wv[0][t]);
#pragma unroll 2
  for (int ci = 0; ci < CIN; ++ci) {
    if (ci + 1 < CIN) {
#pragma unroll
      for (int t = 0; t < 9; ++t) dn_load_w<CB>(wp + ((ci + 1) * 9 + t) * COUT, wv[(ci + 1) & 1][t]);
    }
    float in[4][6];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int x = 0; x < 6; ++x) in[r][x] = in0[ci * SC + r * SP + x];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int c = 0; c < CB; ++c)
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int x = 0; x < 4; ++x)
              acc[c][r][x] = fmaf(wv[ci & 1][ky * 3 + kx][c], in[r + ky][x + kx], acc[c][r][x]);
  }
  float* out = dst + s * dst_stride + (cg * CB) * DC;
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    if (POOL) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float m = fmaxf(fmaxf(acc[c][0][2 * j], acc[c][0][2 * j + 1]), fmaxf(acc[c][1][2 * j], acc[c][1][2 * j + 1]));
        out[c * DC + (ty / 2 + 1) * DP + tx / 2 + j + 1] = fmaxf(m, 0.f);
      }
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int x = 0; x < 4; ++x) out[c * DC + (ty + r + 1) * DP + tx + x + 1] = fmaxf(acc[c][r][x], 0.f);
    }
  }
}

// 4x4 stride-2 padding-1 transposed convolution + relu, one output parity class per thread.
// src [G][CIN][HIN+2][SP] (zero border); dst [G][COUT][2 HIN + 2 PAD][DP], PAD = 1 keeps a zero border.
template <int CIN, int COUT, int HIN, int CB, int PAD, int SP, int DP>
__device__ __forceinline__ void dn_convt4(const float* __restrict__ src, int src_stride, float* __restrict__ dst,
                                          int dst_stride, const float* __restrict__ w, const float* __restrict__ bias) {
  constexpr int TX = HIN / 4, PT = (HIN / 2) * TX, NCG = COUT / CB, SC = (HIN + 2) * SP, DC = (2 * HIN + 2 * PAD) * DP;
  static_assert(DN_G * PT * NCG * 4 == DN_THREADS, "layer does not fill the CTA");
  const int q = threadIdx.x;
  const int pt = q % PT, s = (q / PT) % DN_G, combo = q / (PT * DN_G);
  const int cls = combo & 3, cg = combo >> 2;
  const int py = cls >> 1, px = cls & 1;
  const int ty = (pt / TX) * 2, tx = (pt % TX) * 4;
  float acc[CB][2][4];
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    const float b = __ldg(bias + cg * CB + c);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int x = 0; x < 4; ++x) acc[c][r][x] = b;
  }
  // window rows (ty + py - 1 .. + 2), columns (tx + px - 1 .. + 4) of the input; +1 for the border
  const float* in0 = src + s * src_stride + (ty + py) * SP + tx + px;
  const float* wp = w + cg * CB + ((1 - py) * 4 + (1 - px)) * COUT;
  float wv[2][4][CB];                                                  // tap (1 - py + 2a, 1 - px + 2b) at index 2a + b
#pragma unroll
  for (int t = 0; t < 4; ++t) dn_load_w<CB>(wp + (8 * (t >> 1) + 2 * (t & 1)) * COUT, wv[0][t]);
#pragma unroll 2
  for (int ci = 0; ci < CIN; ++ci) {
    if (ci + 1 < CIN) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        dn_load_w<CB>(wp + ((ci + 1) * 16 + 8 * (t >> 1) + 2 * (t & 1)) * COUT, wv[(ci + 1) & 1][t]);
    }
    float in[3][5];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int x = 0; x < 5; ++x) in[r][x] = in0[ci * SC + r * SP + x];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < CB; ++c)
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int x = 0; x < 4; ++x)
              acc[c][r][x] = fmaf(wv[ci & 1][2 * a + b][c], in[r + 1 - a][x + 1 - b], acc[c][r][x]);
  }
  float* out = dst + s * dst_stride + (cg * CB) * DC;
#pragma unroll
  for (int c = 0; c < CB; ++c)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int x = 0; x < 4; ++x)
        out[c * DC + (2 * (ty + r) + py + PAD) * DP + 2 * (tx + x) + px + PAD] = fmaxf(acc[c][r][x], 0.f);
}

__device__ __forceinline__ void dn_zero(float* buf, int stride, int count) {
  for (int i = threadIdx.x; i < DN_G * count; i += DN_THREADS) buf[(i / count) * stride + (i % count)] = 0.f;
}

// in / out: [n_spots][256] (row-major 16 x 16); par: packed parameters (DN_PARAM_FLOATS)
__global__ void __launch_bounds__(DN_THREADS, 1) denoise_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                long long n_spots, const float* __restrict__ par) {
  extern __shared__ __align__(16) float dn_smem[];
  float* A = dn_smem;                         // DN_BUF_A floats
  float* B = dn_smem + DN_BUF_A;              // DN_BUF_B floats
  const long long n_batches = (n_spots + DN_G - 1) / DN_G;
  for (long long batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const long long spot0 = batch * DN_G;
    dn_zero(A, DN_X_S, 18 * 18);
    dn_zero(B, DN_E1_S, 16 * 10 * DN_E1_P);
    __syncthreads();
    for (int i = threadIdx.x; i < DN_G * 256; i += DN_THREADS) {
      const int s = i >> 8, pix = i & 255;
      if (spot0 + s < n_spots) A[s * DN_X_S + ((pix >> 4) + 1) * 18 + (pix & 15) + 1] = in[(spot0 + s) * 256 + pix];
    }
    __syncthreads();
    dn_conv3<1, 16, 16, 8, true, 18, DN_E1_P>(A, DN_X_S, B, DN_E1_S, par + DN_E1W, par + DN_E1B);
    __syncthreads();
    dn_zero(A, DN_E2_S, 32 * 6 * 6);
    __syncthreads();
    dn_conv3<16, 32, 8, 4, true, DN_E1_P, 6>(B, DN_E1_S, A, DN_E2_S, par + DN_E2W, par + DN_E2B);
    __syncthreads();
    dn_zero(B, DN_E3_S, 64 * 6 * 6);
    __syncthreads();
    dn_conv3<32, 64, 4, 2, false, 6, 6>(A, DN_E2_S, B, DN_E3_S, par + DN_E3W, par + DN_E3B);
    __syncthreads();
    dn_zero(A, DN_D1_S, 32 * 10 * DN_D1_P);
    __syncthreads();
    dn_convt4<64, 32, 4, 4, 1, 6, DN_D1_P>(B, DN_E3_S, A, DN_D1_S, par + DN_D1W, par + DN_D1B);
    __syncthreads();
    dn_convt4<32, 16, 8, 8, 0, DN_D1_P, 16>(A, DN_D1_S, B, DN_D2_S, par + DN_D2W, par + DN_D2B);
    __syncthreads();
    {
      // transposed 3x3 stride-1 convolution (16 -> 1): out[oy][ox] = b + sum W[ci][ky][kx] d2[ci][oy+1-ky][ox+1-kx]
      const int q = threadIdx.x, s = q >> 6, t = q & 63, oy = t >> 2, ox0 = (t & 3) * 4;
      const float* d2 = B + s * DN_D2_S;
      const float* w3 = par + DN_D3W;
      float acc[4];
      const float b3 = __ldg(par + DN_D3B);
#pragma unroll
      for (int x = 0; x < 4; ++x) acc[x] = b3;
      for (int ci = 0; ci < 16; ++ci) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int iy = oy + 1 - ky;
          if (iy < 0 || iy > 15) continue;
          float row[6];
#pragma unroll
          for (int x = 0; x < 6; ++x) {
            const int ix = ox0 - 1 + x;
            row[x] = (ix >= 0 && ix <= 15) ? d2[ci * 256 + iy * 16 + ix] : 0.f;
          }
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float wv = __ldg(w3 + ci * 9 + ky * 3 + kx);
#pragma unroll
            for (int x = 0; x < 4; ++x) acc[x] = fmaf(wv, row[x + 2 - kx], acc[x]);      // ix = ox + 1 - kx
          }
        }
      }
      if (spot0 + s < n_spots)
        *reinterpret_cast<float4*>(out + (spot0 + s) * 256 + oy * 16 + ox0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
    __syncthreads();
  }
}
