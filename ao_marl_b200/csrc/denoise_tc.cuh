// denoise_tc.cuh -- the per-subaperture CNN denoiser with its four wide layers on tcgen05 (row a-6, config 4).
//
// Reference: DenoisingAutoencoderCNN2DSingleSubapeture (src/autoencoder/autoencoder_models.py:130-197), applied to every
// 16 x 16 spot of the detector cube between the sensor frame and the centroider (rlSupervisor.py:876-891, 968-979):
//   e1 conv3x3 1->16 + relu + pool | e2 conv3x3 16->32 + relu + pool | e3 conv3x3 32->64 + relu
//   d1 convT4x4/2 64->32 + relu    | d2 convT4x4/2 32->16 + relu     | d3 convT3x3 16->1
// 96 % of the 3.42 MFLOP per spot are in e2, e3, d1, d2.  They run here as implicit GEMMs on `tcgen05.mma kind::f16`
// (fp16 hi + lo operands, three products each -- two instructions, see the issue groups -- float32 accumulators in TMEM); e1 and d3 (K = 9 / N = 1) stay on the
// CUDA cores, fused into the prologue and the last epilogue.  denoise_kernels.cuh keeps the float32 FFMA form
// (cross-check, AOM_DENOISE_SIMT).
//
// One geometry for every layer.  A pass takes G = 5 spots.  Every feature map is stored *space-to-depth* on the 4 x 4
// coarse grid of a spot: an 8 x 8 map as 4 phase planes, a 16 x 16 map as 16.  Coarse positions of the five spots are
// laid out linearly with a shared zero pad column and pad row (pitch 5: q = (1 + 5 s + Y) 5 + 1 + X < 128), so
//   * the M dimension of every GEMM is the same 128 positions = one UMMA tile = TMEM lane = one epilogue thread,
//   * a convolution tap, a pooling phase or a transposed-convolution parity is a (phase plane, linear shift) pair, i.e.
//     only a different *start address* of the A descriptor: activations live in shared memory as planes
//     [channel group of 8][position][8 channels] (16 bytes per position), which is the no-swizzle K-major core-matrix
//     order (8 consecutive positions x 16 bytes = one core matrix, SBO = 128 B, LBO = plane stride) for any shift,
//   * max-pooling = the maximum over four accumulators (the four output phases of e2) in the epilogue thread; the four
//     parities of a transposed convolution are four accumulators (d1) / sixteen (d2, two levels) -- no data movement.
// Zero padding = the pad positions of the planes, zeroed once and never written.  Activations carry a factor 2^-4 (exact)
// so that fp16 cannot overflow on bright spots; biases are pre-scaled, the last layer undoes it.
// d3 (transposed 3 x 3, 16 -> 1) is a scatter from the d2 epilogue: the thread that holds a 4 x 4 x 16 block of d2's
// output adds its 6 x 6 window of partial sums into the output image in shared memory.
// Weights (256 KB as fp16 hi / lo tiles, pre-ordered by ao_marl_b200/denoiser.py::pack_weights_tc) stream through two
// 36 KB buffers by cp.async.bulk, eight chunks per pass, prefetched one chunk ahead.
// Warps 0-15: prologue (e1), epilogues (four warps per TMEM lane quarter split the columns); warp 16: one elected lane
// issues the bulk copies and the 510 MMAs of a pass.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <type_traits>

#define DT_G 5
#define DT_PITCH 5
#define DT_LEAD 8
#define DT_NPOS 144
#define DT_PLANE (DT_NPOS * 16)            // 2304 B: one channel group of 8 (fp16) over all positions
#define DT_WBUF 36864
#define DT_EPI_WARPS 16
#define DT_THREADS (32 * DT_EPI_WARPS + 32)
#define DT_TMEM_COLS 512
#define DT_NPARAM 452
#define DT_WBLOB_BYTES 256000

// float parameters (kernel argument = constant bank): offsets
#define DT_P_W1 0      // [16][9]
#define DT_P_B1 144
#define DT_P_B2 160    // scaled
#define DT_P_B3 192
#define DT_P_B4 256
#define DT_P_B5 288
#define DT_P_W6 304    // [16][9]
#define DT_P_B6 448
#define DT_P_S 449
#define DT_P_INVS 450

// shared memory map
#define DT_OFF_A14 0                                  // 32 planes: A1s [phase][part][2 groups] / A4s [phase][part][4 groups]
#define DT_OFF_A2 (32 * DT_PLANE)                     // 8 planes  [part][4 groups]
#define DT_OFF_A3 (40 * DT_PLANE)                     // 16 planes [part][8 groups]
#define DT_OFF_W (56 * DT_PLANE)                      // two weight buffers
#define DT_OFF_IN (DT_OFF_W + 2 * DT_WBUF)            // [G][18][18] float, zero border
#define DT_OFF_OUT (DT_OFF_IN + DT_G * 324 * 4)       // d3 windows [G][16 cells][4 quarters][4][4] float
#define DT_OFF_BAR (DT_OFF_OUT + DT_G * 1024 * 4)     // mbarriers: w0, w1, mma, cA, cB ; tmem slot
#define DT_OFF_SPRM (DT_OFF_BAR + 64)                 // CUDA-core weights as vector-register operands: w1 [16][12], w6 [9][16]
#define DT_SPRM_FLOATS (16 * 12 + 9 * 16)
#define DT_SMEM_BYTES (DT_OFF_SPRM + DT_SPRM_FLOATS * 4)

struct DnTcParams {
  const float* in;          // [n_spots][256]
  float* out;               // [n_spots][256] (may alias in)
  long long n_spots;
  const uint8_t* wblob;     // DT_WBLOB_BYTES, 16-byte aligned
  int* err;
  long long* dbg;           // optional [32]: per-phase clock totals of CTA 0: [0..8] an epilogue thread, [16..30] the issuing lane
  float prm[DT_NPARAM];
};

#ifdef DT_DEFINE_KERNELS
#define DT_WAIT_SPINS (1u << 22)

__device__ __forceinline__ uint32_t dt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool dt_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait (a nanosleep back-off cost up to 0.5 us of wake-up latency per phase here: nine warps, nothing to yield to);
// on expiry the error word is raised and the kernel traps
__device__ __forceinline__ void dt_mbar_wait(uint32_t bar, uint32_t parity, int* err) {
#pragma unroll 1
  for (uint32_t it = 0; it < DT_WAIT_SPINS; ++it)
    if (dt_mbar_try(bar, parity)) return;        // try_wait itself suspends the thread for a bounded time
  atomicExch(err, 5);
  __threadfence_system();
  __trap();
}
__device__ __forceinline__ bool dt_elect() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}
// one weight chunk as 4 KB bulk copies on the same mbarrier: a single 36 KB copy streamed at ~6 B per cycle
__device__ __forceinline__ void dt_bulk(uint32_t dst, const uint8_t* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#pragma unroll 1
  for (uint32_t o = 0; o < bytes; o += 4096u) {
    const uint32_t n = bytes - o < 4096u ? bytes - o : 4096u;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst + o), "l"(src + o), "r"(n), "r"(bar) : "memory");
  }
}
__device__ __forceinline__ uint64_t dt_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ void dt_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
// The same with the descriptors given as their low words (start address >> 4 | LBO >> 4 << 16); the high word is the
// constant (SBO = 128 B) >> 4 | version 1 << 14.  A tap / phase / K step is then one add on the low word.
__device__ __forceinline__ void dt_mma2(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, {%6, %6, %6, %6}, p;\n\t}"
      :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(0x4008u), "r"(0u) : "memory");
}
// Fully immediate form: the descriptor low words are assembler constants (shared-window offset of the dynamic array
// >> 4, plus template offsets), built next to the instruction.  With run-time (even uniform) low words the compiler CSE'd
// and hoisted hundreds of descriptor values out of the unrolled issue loops and then spilled uniform registers
// (R2UR / MOV.SPILL chains: 63-98 cycles per MMA instead of the 40 the tensor pipe needs, profiles/dev/umma_rate_probe.cu).
template <uint32_t AQ, uint32_t BQ, uint32_t ACC>
__device__ __forceinline__ void dt_mma_c(uint32_t d_tmem, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 sa, ta, tb;\n\t.reg .b64 da, db;\n\t"
      "mov.u32 sa, dt_sm;\n\t"
      "shr.u32 sa, sa, 4;\n\t"
      "and.b32 sa, sa, 0x3FFF;\n\t"
      "add.u32 ta, sa, %2;\n\t"
      "add.u32 tb, sa, %3;\n\t"
      "mov.b64 da, {ta, %4};\n\t"
      "mov.b64 db, {tb, %4};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %1, {%6, %6, %6, %6}, p;\n\t}"
      :: "r"(d_tmem), "r"(idesc), "n"(AQ), "n"(BQ), "n"(0x4008), "n"(ACC), "r"(0u) : "memory");
}
template <int I, int N, class F>
__device__ __forceinline__ void dt_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    dt_for<I + 1, N>(f);
  }
}
__device__ __forceinline__ void dt_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void dt_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void dt_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void dt_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void dt_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// eight consecutive channels of one position -> the hi and lo planes of their channel group
__device__ __forceinline__ void dt_store8(uint32_t hi_addr, uint32_t lo_addr, const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) dt_split2(v[2 * j], v[2 * j + 1], h[j], l[j]);
  dt_sts128(hi_addr, h[0], h[1], h[2], h[3]);
  dt_sts128(lo_addr, l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ constexpr int dt_floor2(int a) { return (a + 4) / 2 - 2; }    // floor(a / 2), a >= -4
// stride-2 4 x 4 transposed convolution, output parity p, tap t: kernel index and input offset
__device__ __forceinline__ constexpr int dt_ct_k(int p, int t) { return p == 0 ? (t == 0 ? 1 : 3) : (t == 0 ? 0 : 2); }
__device__ __forceinline__ constexpr int dt_ct_d(int p, int t) { return p == 0 ? (t == 0 ? 0 : -1) : (t == 0 ? 1 : 0); }

// Sixteen epilogue warps: TMEM lane quarter = warp & 3 (the position), column quarter Q = warp >> 2 (which quarter of the
// channels / classes of that position).  With eight warps (two column halves) the CUDA-core phases ran at 0.25 instructions
// per cycle and scheduler: two warps per scheduler cannot hide their own LDS / LDTM / constant-load latencies.

// e1 on the CUDA cores: conv3x3 1 -> 16 + relu + pool for this thread's coarse position and four channels, written as
// the four phase planes of the 8 x 8 map (8-byte halves of the 16-byte channel group)
template <int Q>
__device__ __forceinline__ void dt_layer1(const DnTcParams& P, const float* sprm, const float* img, int Y, int X, uint32_t a1_pos) {
  float win[6][6];
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = 0; b < 6; ++b) win[a][b] = img[(4 * Y + a) * 18 + 4 * X + b];
  float res[4][4];
  const float S = P.prm[DT_P_S];
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int c = 4 * Q + cc;
    float cv[4][4];
#pragma unroll
    for (int fy = 0; fy < 4; ++fy)
#pragma unroll
      for (int fx = 0; fx < 4; ++fx) cv[fy][fx] = 0.f;
    float wk[12];
#pragma unroll
    for (int q4 = 0; q4 < 3; ++q4) {
      const float4 t = *reinterpret_cast<const float4*>(sprm + c * 12 + 4 * q4);
      wk[4 * q4] = t.x; wk[4 * q4 + 1] = t.y; wk[4 * q4 + 2] = t.z; wk[4 * q4 + 3] = t.w;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float w = wk[ky * 3 + kx];                             // one fetch, sixteen independent FMAs
#pragma unroll
        for (int fy = 0; fy < 4; ++fy)
#pragma unroll
          for (int fx = 0; fx < 4; ++fx) cv[fy][fx] = fmaf(w, win[fy + ky][fx + kx], cv[fy][fx]);
      }
    const float b = P.prm[DT_P_B1 + c];
#pragma unroll
    for (int py = 0; py < 2; ++py)
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        const float m = fmaxf(fmaxf(cv[2 * py][2 * px], cv[2 * py][2 * px + 1]), fmaxf(cv[2 * py + 1][2 * px], cv[2 * py + 1][2 * px + 1]));
        res[py * 2 + px][cc] = fmaxf(m + b, 0.f) * S;
      }
  }
#pragma unroll
  for (int ph = 0; ph < 4; ++ph) {
    uint32_t h0, l0, h1, l1;
    dt_split2(res[ph][0], res[ph][1], h0, l0);
    dt_split2(res[ph][2], res[ph][3], h1, l1);
    const uint32_t hi = a1_pos + ((ph * 2 + 0) * 2 + (Q >> 1)) * DT_PLANE + (Q & 1) * 8;
    const uint32_t lo = a1_pos + ((ph * 2 + 1) * 2 + (Q >> 1)) * DT_PLANE + (Q & 1) * 8;
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(hi), "r"(h0), "r"(h1) : "memory");
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(lo), "r"(l0), "r"(l1) : "memory");
  }
}

// epilogue of e2: max over the four phase accumulators (= 2 x 2 pooling), bias, relu -> A2; this quarter's 8 channels
template <int Q>
__device__ __forceinline__ void dt_epi2(const DnTcParams& P, uint32_t tlane, bool real, uint32_t a2_pos) {
  float m[8], v[8], u[8];
  dt_ld8(tlane + 0 * 64 + 8 * Q, m);
  dt_ld8(tlane + 0 * 64 + 32 + 8 * Q, u);                     // the A_hi W_lo half of the accumulator
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] += u[j];
#pragma unroll
  for (int ph = 1; ph < 4; ++ph) {
    dt_ld8(tlane + ph * 64 + 8 * Q, v);
    dt_ld8(tlane + ph * 64 + 32 + 8 * Q, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j] + u[j]);
  }
  if (real) {
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j] + P.prm[DT_P_B2 + 8 * Q + j], 0.f);
    dt_store8(a2_pos + (0 * 4 + Q) * DT_PLANE, a2_pos + (1 * 4 + Q) * DT_PLANE, m);
  }
}

// epilogue of e3: bias, relu -> A3; this quarter's 16 channels
template <int Q>
__device__ __forceinline__ void dt_epi3(const DnTcParams& P, uint32_t tlane, bool real, uint32_t a3_pos) {
  float v[16];
  dt_ld16(tlane + 256 + 16 * Q, v);
  if (real) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + P.prm[DT_P_B3 + 16 * Q + j], 0.f);
#pragma unroll
    for (int g = 0; g < 2; ++g)
      dt_store8(a3_pos + (0 * 8 + 2 * Q + g) * DT_PLANE, a3_pos + (1 * 8 + 2 * Q + g) * DT_PLANE, v + 8 * g);
  }
}

// epilogue of d1: the four parity classes are the four phase planes of the 8 x 8 map; this quarter's class
template <int Q>
__device__ __forceinline__ void dt_epi4(const DnTcParams& P, uint32_t tlane, bool real, uint32_t a4_pos) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float v[16], u[16];
    dt_ld16(tlane + Q * 64 + 16 * q, v);
    dt_ld16(tlane + Q * 64 + 32 + 16 * q, u);
    if (real) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j] + u[j] + P.prm[DT_P_B4 + 16 * q + j], 0.f);
#pragma unroll
      for (int g = 0; g < 2; ++g)
        dt_store8(a4_pos + ((Q * 2 + 0) * 4 + 2 * q + g) * DT_PLANE, a4_pos + ((Q * 2 + 1) * 4 + 2 * q + g) * DT_PLANE, v + 8 * g);
    }
  }
}

// epilogue of d2 + d3: this quarter Q = (py, px) holds the 2 x 2 fine pixels fy = 2 py + qy, fx = 2 px + qx of its 4 x 4
// block (x 16 channels); every fine pixel adds its nine taps of the transposed 3 x 3 convolution into a 4 x 4 window of
// partial sums, which is then added to the output image of the spot
template <int Q>
__device__ __forceinline__ void dt_epi5(const DnTcParams& P, const float* sprm, uint32_t tlane, bool real, float* out_img, int Y, int X) {
  // all four fine pixels first, then every d3 weight is fetched once and feeds four independent FMAs
  float v[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i) {                                               // i = (qy, qx)
    float u[16];
    dt_ld16(tlane + (Q * 4 + i) * 32, v[i]);
    dt_ld16(tlane + (Q * 4 + i) * 32 + 16, u);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[i][j] += u[j];
  }
  float wnd[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) wnd[a][b] = 0.f;
  if (real) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) v[i][j] = fmaxf(v[i][j] + P.prm[DT_P_B5 + j], 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 t = *reinterpret_cast<const float4*>(sprm + 192 + (ky * 3 + kx) * 16 + 4 * j4);
          const float w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[i][4 * j4 + jj], w[jj], acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) wnd[(i >> 1) + ky][(i & 1) + kx] += acc[i];
      }
  }
  // the windows of neighbouring threads overlap: each thread parks its own in shared memory, and every output pixel then
  // sums its (at most four, one per quarter) contributions in a fixed order -- reproducible bit for bit, one barrier
  // (a first form added the windows in 16 colour phases separated by barriers: ~5 k cycles per pass)
  if (real) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
      *reinterpret_cast<float4*>(out_img + ((Y * 4 + X) * 4 + Q) * 16 + 4 * a) = make_float4(wnd[a][0], wnd[a][1], wnd[a][2], wnd[a][3]);
  }
}

// ---- MMA issue groups.  Each group is its own (non-inlined) function: with the 738 MMAs of a pass in one body ptxas hoisted
// the descriptor constants of the whole pass and spilled uniform registers between every pair of UTCHMMA. ----
constexpr uint32_t DT_PQ = DT_PLANE >> 4;
constexpr uint32_t DT_QA14 = (DT_OFF_A14 >> 4) + DT_LEAD + (DT_PQ << 16);
constexpr uint32_t DT_QA2 = (DT_OFF_A2 >> 4) + DT_LEAD + (DT_PQ << 16);
constexpr uint32_t DT_QA3 = (DT_OFF_A3 >> 4) + DT_LEAD + (DT_PQ << 16);
constexpr uint32_t DT_QW0 = (DT_OFF_W >> 4), DT_QW1 = DT_QW0 + (DT_WBUF >> 4);
constexpr uint32_t DT_ID16 = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t DT_ID32 = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t DT_ID64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// Weight tiles of e2, d1, d2 hold the hi rows and the lo rows of W side by side (2N rows): A_hi x tile gives A_hi W_hi and
// A_hi W_lo in ONE instruction of width 2N (columns [0, N) and [N, 2N) of the accumulator), A_lo x its first N rows adds
// A_lo W_hi to [0, N) -- two MMAs per product instead of three at the same ~40-cycle floor; the epilogue adds the halves.

// e2, output phase PH: nine taps, K = 16, weights in buffer 0; accumulator columns PH * 64 .. + 63
template <int PH>
__device__ __noinline__ void dt_issue_e2(uint32_t tmem_u) {
  constexpr int py = PH >> 1, px = PH & 1;
  const uint32_t d = tmem_u + PH * 64;
  dt_for<0, 9>([&](auto tc) {
    constexpr int tap = decltype(tc)::value, dy = tap / 3 - 1, dx = tap % 3 - 1;
    constexpr int iph = ((py + dy) & 1) * 2 + ((px + dx) & 1);
    constexpr int sh = dt_floor2(py + dy) * DT_PITCH + dt_floor2(px + dx);
    constexpr uint32_t a_hi = DT_QA14 + ((iph * 2 + 0) * 2) * DT_PQ + sh, a_lo = DT_QA14 + ((iph * 2 + 1) * 2) * DT_PQ + sh;
    constexpr uint32_t b = DT_QW0 + tap * 128 + (64u << 16);
    dt_mma_c<a_hi, b, (tap ? 1u : 0u)>(d, DT_ID64);
    dt_mma_c<a_lo, b, 1u>(d, DT_ID32);
  });
}
// e3, taps T0 .. T0 + 2: high weight parts from buffer 1 (LO = 0: two products) or low parts from buffer 0 (LO = 1)
template <int T0, int LO>
__device__ __noinline__ void dt_issue_e3(uint32_t tmem_u) {
  const uint32_t d = tmem_u + 256;
  dt_for<0, 6>([&](auto ic) {
    constexpr int tap = T0 + (decltype(ic)::value >> 1), ks = decltype(ic)::value & 1;
    constexpr int sh = (tap / 3 - 1) * DT_PITCH + tap % 3 - 1;
    constexpr uint32_t a_hi = DT_QA2 + (0 * 4 + 2 * ks) * DT_PQ + sh, a_lo = DT_QA2 + (1 * 4 + 2 * ks) * DT_PQ + sh;
    constexpr uint32_t b = (LO ? DT_QW0 : DT_QW1) + tap * 256 + ks * 128 + (64u << 16);
    if constexpr (LO == 0) {
      dt_mma_c<a_hi, b, ((tap | ks) ? 1u : 0u)>(d, DT_ID64);
      dt_mma_c<a_lo, b, 1u>(d, DT_ID64);
    } else {
      dt_mma_c<a_hi, b, 1u>(d, DT_ID64);
    }
  });
}
// d1, parity class CL, tap T: K = 64; accumulator columns CL * 64 .. + 63
template <int CL, int T>
__device__ __noinline__ void dt_issue_d1(uint32_t tmem_u) {
  constexpr int py = CL >> 1, px = CL & 1;
  constexpr uint32_t wq = (CL & 1) ? DT_QW0 : DT_QW1;
  const uint32_t d = tmem_u + CL * 64;
  constexpr int sh = dt_ct_d(py, T >> 1) * DT_PITCH + dt_ct_d(px, T & 1);
  dt_for<0, 4>([&](auto kc) {
    constexpr int ks = decltype(kc)::value;
    constexpr uint32_t a_hi = DT_QA3 + (0 * 8 + 2 * ks) * DT_PQ + sh, a_lo = DT_QA3 + (1 * 8 + 2 * ks) * DT_PQ + sh;
    constexpr uint32_t b = wq + T * 512 + ks * 128 + (64u << 16);
    dt_mma_c<a_hi, b, ((T | ks) ? 1u : 0u)>(d, DT_ID64);
    dt_mma_c<a_lo, b, 1u>(d, DT_ID32);
  });
}
// d2, (input phase, output parity) class SC: four taps, K = 32, weights in buffer 1; accumulator columns SC * 32 .. + 31
template <int SC>
__device__ __noinline__ void dt_issue_d2(uint32_t tmem_u) {
  constexpr int py = SC >> 3, px = (SC >> 2) & 1, qy = (SC >> 1) & 1, qx = SC & 1;
  const uint32_t d = tmem_u + SC * 32;
  dt_for<0, 8>([&](auto ic) {
    constexpr int t = decltype(ic)::value >> 1, ks = decltype(ic)::value & 1;
    constexpr int ky = dt_ct_k(qy, t >> 1), dy = dt_ct_d(qy, t >> 1), kx = dt_ct_k(qx, t & 1), dx = dt_ct_d(qx, t & 1);
    constexpr int iph = ((py + dy) & 1) * 2 + ((px + dx) & 1);
    constexpr int sh = dt_floor2(py + dy) * DT_PITCH + dt_floor2(px + dx);
    constexpr uint32_t b = DT_QW1 + (ky * 4 + kx) * 128 + ks * 64 + (32u << 16);
    constexpr uint32_t a_hi = DT_QA14 + ((iph * 2 + 0) * 4 + 2 * ks) * DT_PQ + sh, a_lo = DT_QA14 + ((iph * 2 + 1) * 4 + 2 * ks) * DT_PQ + sh;
    dt_mma_c<a_hi, b, ((t | ks) ? 1u : 0u)>(d, DT_ID32);
    dt_mma_c<a_lo, b, 1u>(d, DT_ID16);
  });
}

// the spots of pass `pass` into the padded input images, asynchronously (cp.async): HBM latency hides behind a whole pass
__device__ __forceinline__ void dt_fetch_input(const DnTcParams& P, long long pass, int tid, uint32_t img_u32) {
  const long long spot0 = pass * DT_G;
  const int ns = (int)((P.n_spots - spot0) < DT_G ? (P.n_spots - spot0) : DT_G);
  for (int i = tid; i < ns * 256; i += 32 * DT_EPI_WARPS) {
    const int sp = i >> 8, px = i & 255;
    const uint32_t dst = img_u32 + 4u * (uint32_t)(sp * 324 + (1 + (px >> 4)) * 18 + 1 + (px & 15));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(P.in + (spot0 + sp) * 256 + px) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(DT_THREADS, 1) denoise_tc_kernel(const __grid_constant__ DnTcParams P) {
  extern __shared__ __align__(128) uint8_t dt_sm[];
  const uint32_t sb = dt_smem_u32(dt_sm);
  // warp index through a shuffle and the issuing lane through elect.sync: the compiler then knows the control path is
  // warp-uniform and keeps the UMMA descriptors in uniform registers (UTCHMMA back to back); with `lane == 0` every MMA
  // sat in a per-thread waterfall loop (ELECT / R2UR / BRA.U.ANY, ~13 instructions and ~160 cycles per MMA)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = tid & 31;
  const uint32_t bar_w0 = sb + DT_OFF_BAR, bar_w1 = bar_w0 + 8, bar_mma = bar_w0 + 16, bar_ca = bar_w0 + 24, bar_cb = bar_w0 + 32;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(dt_sm + DT_OFF_BAR + 40);

  for (int i = tid; i < DT_OFF_W / 16; i += DT_THREADS) reinterpret_cast<uint4*>(dt_sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < (DT_OFF_BAR - DT_OFF_IN) / 4; i += DT_THREADS) reinterpret_cast<float*>(dt_sm + DT_OFF_IN)[i] = 0.f;
  // e1 / d3 weights into shared memory: as kernel-argument constants they became uniform-register operands, and FFMA with
  // a uniform operand issued at a quarter of the rate (both phases took 9 k cycles for 2.3 k FFMA per scheduler)
  {
    float* sp = reinterpret_cast<float*>(dt_sm + DT_OFF_SPRM);
    for (int i = tid; i < 16 * 12; i += DT_THREADS) sp[i] = (i % 12) < 9 ? P.prm[DT_P_W1 + (i / 12) * 9 + i % 12] : 0.f;
    for (int i = tid; i < 9 * 16; i += DT_THREADS) sp[192 + i] = P.prm[DT_P_W6 + (i % 16) * 9 + i / 16];
  }
  if (tid == 0) {
    for (int b = 0; b < 5; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_w0 + 8 * b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == DT_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb + DT_OFF_BAR + 40), "r"((uint32_t)DT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const long long n_pass = (P.n_spots + DT_G - 1) / DT_G;
  const bool w8 = warp == DT_EPI_WARPS;      // the issuing lane is re-elected at every control block (see above)
  const uint32_t wb0 = sb + DT_OFF_W, wb1 = wb0 + DT_WBUF;
  // mbarrier phases: each weight buffer completes four chunks per pass and bar_ca two, so their parities are fixed per
  // use; bar_cb completes once per pass, bar_mma four times (tracked by every thread)
  uint32_t pm = 0, pcb = 0;

  // this thread's coarse position (epilogue warps): TMEM lane m = UMMA row = linear position q
  const int m = 32 * (warp & 3) + lane, quarter = (warp >> 2) & 3;
  const int r = m / DT_PITCH, c = m - r * DT_PITCH;
  const int s = r >= 1 ? (r - 1) / 5 : 0, Y = r >= 1 ? (r - 1) % 5 : 4, X = c - 1;
  const bool is_pos = warp < DT_EPI_WARPS && r >= 1 && c >= 1 && Y < 4 && s < DT_G;
  const uint32_t tlane = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
  const uint32_t pos_off = (uint32_t)(DT_LEAD + m) * 16;
  float* img_in = reinterpret_cast<float*>(dt_sm + DT_OFF_IN);
  float* img_out = reinterpret_cast<float*>(dt_sm + DT_OFF_OUT);
  const float* sprm = reinterpret_cast<const float*>(dt_sm + DT_OFF_SPRM);

  if (w8 && (long long)blockIdx.x < n_pass && dt_elect()) {
    dt_bulk(wb0, P.wblob + 0, 18432, bar_w0);
    dt_bulk(wb1, P.wblob + 18432, 36864, bar_w1);
  }

  // descriptor low words in 16-byte units, warp-uniform: start address >> 4 | LBO >> 4 << 16.  Position shift s = +s,
  // plane p = +p * PQ, weight offset o bytes = +o / 16.
  // (the masked shared-window offset is a compile-time constant for the compiler -- no shuffle here, it would hide that --
  // so every descriptor low word folds to an immediate)
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);

  long long tacc[12], tprev = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) tacc[i] = 0;
#define DT_TICK(i) if (P.dbg) { const long long tn = clock64(); tacc[i] += tn - tprev; tprev = tn; }
  long long cacc[15], cprev = 0;
#pragma unroll
  for (int i = 0; i < 15; ++i) cacc[i] = 0;
#define DT_CT0 if (P.dbg) cprev = clock64();
#define DT_CT(i) if (P.dbg) { const long long tn = clock64(); cacc[i] += tn - cprev; cprev = tn; }
  if (warp < DT_EPI_WARPS && (long long)blockIdx.x < n_pass) dt_fetch_input(P, blockIdx.x, tid, sb + DT_OFF_IN);
#pragma unroll 1
  for (long long pass = blockIdx.x; pass < n_pass; pass += gridDim.x) {
    if (P.dbg) tprev = clock64();
    const long long spot0 = pass * DT_G;
    const int ns = (int)((P.n_spots - spot0) < DT_G ? (P.n_spots - spot0) : DT_G);
    const bool has_next = pass + gridDim.x < n_pass;
    const bool real = is_pos && s < ns;

    // ---- input spots -> padded images; e1 -> A1s
    if (warp < DT_EPI_WARPS) {
      asm volatile("cp.async.wait_all;" ::: "memory");        // this pass's spots, requested a pass ago (dt_fetch_input)
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (real) {
        switch (quarter) {
          case 0: dt_layer1<0>(P, sprm, img_in + s * 324, Y, X, sb + DT_OFF_A14 + pos_off); break;
          case 1: dt_layer1<1>(P, sprm, img_in + s * 324, Y, X, sb + DT_OFF_A14 + pos_off); break;
          case 2: dt_layer1<2>(P, sprm, img_in + s * 324, Y, X, sb + DT_OFF_A14 + pos_off); break;
          default: dt_layer1<3>(P, sprm, img_in + s * 324, Y, X, sb + DT_OFF_A14 + pos_off); break;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DT_TICK(0)
    if (warp < DT_EPI_WARPS && has_next) dt_fetch_input(P, pass + gridDim.x, tid, sb + DT_OFF_IN);   // e1 has read this pass's images

    // ---- e2: four output phases x nine taps, K = 16
    if (w8 && dt_elect()) {
      DT_CT0
      dt_mbar_wait(bar_w0, 0u, P.err);
      DT_CT(0)
      dt_issue_e2<0>(tmem_u); dt_issue_e2<1>(tmem_u); dt_issue_e2<2>(tmem_u); dt_issue_e2<3>(tmem_u);
      dt_commit(bar_mma);
      DT_CT(1)
      dt_mbar_wait(bar_mma, pm, P.err);
      DT_CT(2)
    }
    dt_mbar_wait(bar_mma, pm, P.err);
    pm ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DT_TICK(1)
    if (w8 && dt_elect()) dt_bulk(wb0, P.wblob + 55296, 36864, bar_w0);                  // e3 low parts
    if (warp < DT_EPI_WARPS) {
      switch (quarter) {
        case 0: dt_epi2<0>(P, tlane, real, sb + DT_OFF_A2 + pos_off); break;
        case 1: dt_epi2<1>(P, tlane, real, sb + DT_OFF_A2 + pos_off); break;
        case 2: dt_epi2<2>(P, tlane, real, sb + DT_OFF_A2 + pos_off); break;
        default: dt_epi2<3>(P, tlane, real, sb + DT_OFF_A2 + pos_off); break;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DT_TICK(2)

    // ---- e3: nine taps, K = 32; high weight parts from buffer 1, low parts from buffer 0
    if (w8 && dt_elect()) {
      DT_CT0
      dt_mbar_wait(bar_w1, 0u, P.err);
      DT_CT(3)
      dt_issue_e3<0, 0>(tmem_u); dt_issue_e3<3, 0>(tmem_u); dt_issue_e3<6, 0>(tmem_u);
      DT_CT(4)
      dt_mbar_wait(bar_w0, 1u, P.err);
      DT_CT(5)
      dt_issue_e3<0, 1>(tmem_u); dt_issue_e3<3, 1>(tmem_u); dt_issue_e3<6, 1>(tmem_u);
      dt_commit(bar_mma);
      DT_CT(6)
      dt_mbar_wait(bar_mma, pm, P.err);
      DT_CT(7)
    }
    dt_mbar_wait(bar_mma, pm, P.err);
    pm ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DT_TICK(3)
    if (w8 && dt_elect()) {
      dt_bulk(wb1, P.wblob + 92160, 32768, bar_w1);                           // d1 class 0
      dt_bulk(wb0, P.wblob + 124928, 32768, bar_w0);                          // d1 class 1
    }
    if (warp < DT_EPI_WARPS) {
      switch (quarter) {
        case 0: dt_epi3<0>(P, tlane, real, sb + DT_OFF_A3 + pos_off); break;
        case 1: dt_epi3<1>(P, tlane, real, sb + DT_OFF_A3 + pos_off); break;
        case 2: dt_epi3<2>(P, tlane, real, sb + DT_OFF_A3 + pos_off); break;
        default: dt_epi3<3>(P, tlane, real, sb + DT_OFF_A3 + pos_off); break;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DT_TICK(4)

    // ---- d1: four parity classes x four taps, K = 64; one weight chunk per class
    if (w8 && dt_elect()) {
      dt_for<0, 4>([&](auto clc) {
        constexpr int cl = decltype(clc)::value;
        DT_CT0
        dt_mbar_wait((cl & 1) ? bar_w0 : bar_w1, (cl & 1) ? (uint32_t)(cl >> 1) : (uint32_t)(1 - (cl >> 1)), P.err);
        DT_CT(8)
        dt_issue_d1<cl, 0>(tmem_u); dt_issue_d1<cl, 1>(tmem_u); dt_issue_d1<cl, 2>(tmem_u); dt_issue_d1<cl, 3>(tmem_u);
        DT_CT(9)
        if (cl == 0) dt_commit(bar_ca);
        if (cl == 1) {
          dt_commit(bar_cb);
          dt_mbar_wait(bar_ca, 0u, P.err);
          dt_bulk(wb1, P.wblob + 157696, 32768, bar_w1);                      // class 2
          dt_mbar_wait(bar_cb, pcb, P.err);
          dt_bulk(wb0, P.wblob + 190464, 32768, bar_w0);                      // class 3
          DT_CT(10)
        }
        if (cl == 2) dt_commit(bar_ca);
        if (cl == 3) {
          dt_commit(bar_mma);
          dt_mbar_wait(bar_ca, 1u, P.err);
          dt_bulk(wb1, P.wblob + 223232, 32768, bar_w1);                      // d2
          DT_CT(10)
          dt_mbar_wait(bar_mma, pm, P.err);
          DT_CT(11)
        }
      });
    }
    dt_mbar_wait(bar_mma, pm, P.err);
    pm ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    pcb ^= 1;
    DT_TICK(5)
    if (w8 && has_next && dt_elect()) dt_bulk(wb0, P.wblob + 0, 18432, bar_w0);           // next pass: e2
    if (warp < DT_EPI_WARPS) {
      switch (quarter) {
        case 0: dt_epi4<0>(P, tlane, real, sb + DT_OFF_A14 + pos_off); break;
        case 1: dt_epi4<1>(P, tlane, real, sb + DT_OFF_A14 + pos_off); break;
        case 2: dt_epi4<2>(P, tlane, real, sb + DT_OFF_A14 + pos_off); break;
        default: dt_epi4<3>(P, tlane, real, sb + DT_OFF_A14 + pos_off); break;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DT_TICK(6)

    // ---- d2: sixteen (input phase, output parity) classes x four taps, K = 32
    if (w8 && dt_elect()) {
      DT_CT0
      dt_mbar_wait(bar_w1, 1u, P.err);
      DT_CT(12)
      dt_for<0, 16>([&](auto scc) { dt_issue_d2<decltype(scc)::value>(tmem_u); });
      dt_commit(bar_mma);
      DT_CT(13)
      dt_mbar_wait(bar_mma, pm, P.err);
      DT_CT(14)
    }
    dt_mbar_wait(bar_mma, pm, P.err);
    pm ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DT_TICK(7)
    if (w8 && has_next && dt_elect()) dt_bulk(wb1, P.wblob + 18432, 36864, bar_w1);       // next pass: e3 high parts
    if (warp < DT_EPI_WARPS) {
      switch (quarter) {
        case 0: dt_epi5<0>(P, sprm, tlane, real, img_out + s * 1024, Y, X); break;
        case 1: dt_epi5<1>(P, sprm, tlane, real, img_out + s * 1024, Y, X); break;
        case 2: dt_epi5<2>(P, sprm, tlane, real, img_out + s * 1024, Y, X); break;
        default: dt_epi5<3>(P, sprm, tlane, real, img_out + s * 1024, Y, X); break;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      const float invs = P.prm[DT_P_INVS], b6 = P.prm[DT_P_B6];
      for (int i = tid; i < ns * 256; i += 32 * DT_EPI_WARPS) {
        const int sp = i >> 8, row = (i >> 4) & 15, col = i & 15;
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {                          // window row a of cell Yc, quarter (py, px): row = 4 Yc + 2 py + a - 1
          const int ty = row + 1 - 2 * (q >> 1), tx = col + 1 - 2 * (q & 1);
          if (ty >= 0 && tx >= 0 && ty < 16 && tx < 16)
            sum += img_out[sp * 1024 + (((ty >> 2) * 4 + (tx >> 2)) * 4 + q) * 16 + (ty & 3) * 4 + (tx & 3)];
        }
        P.out[spot0 * 256 + i] = fmaf(sum, invs, b6);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    DT_TICK(8)
  }
  if (P.dbg && blockIdx.x == 0 && tid == 0)
    for (int i = 0; i < 12; ++i) P.dbg[i] = tacc[i];
  if (P.dbg && blockIdx.x == 0 && w8 && dt_elect())
    for (int i = 0; i < 15; ++i) P.dbg[16 + i] = cacc[i];

  if (warp == DT_EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)DT_TMEM_COLS) : "memory");
}
#endif  // DT_DEFINE_KERNELS
