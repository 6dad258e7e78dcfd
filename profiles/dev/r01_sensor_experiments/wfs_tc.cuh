// Shack-Hartmann frame, fourth generation: stage 2 of the DFT on the tcgen05 tensor cores.
//
// Everything up to and including stage 1 is wfs_frame_tma_kernel (TMA-staged tiles, lane (g, q) owns pixels rows
// {2g, 2g+1} x columns {4q..4q+3}, stage 1 = 48 warp-level HMMA in registers).  Stage 2 -- two thirds of the MMA
// work -- leaves the legacy mma.sync path, whose 96 HMMA per subaperture cost 8 tensor-pipe cycles each on top of
// the FP32 instruction stream (the two add up):
//
//   * four consecutive warps form a group; each warp splits its 32 x 16 complex T into fp16 hi / lo and stores it
//     as 32 rows of a 128 x 32 K-major A tile in shared memory (row = 32 warp + kept-frequency index of fx,
//     K = {Tr(y), Ti(y)}; UMMA canonical no-swizzle layout, 8-row x 16-byte core matrices);
//   * the last warp to arrive issues six tcgen05.mma (kind::f16, M = 128, N = 64, K = 16):
//         D[128][64] = Th . Wh + Tl . Wh + Th . Wl,     W = [[Wr, Wi], [-Wi, Wr]]  (64 x 32, constant, in smem)
//     into one of two TMEM accumulators of the group, and commits to an mbarrier: 6 x 32 = 192 tensor cycles for
//     four subapertures instead of 4 x 768;
//   * the accumulator of the PREVIOUS item is read back one iteration later (tcgen05.ld, lane = fx row, 64
//     columns = {Yr, Yi} of the 32 kept fy), so the MMA and its commit latency hide behind the field arithmetic
//     of the next subaperture; |.|^2, the fy half of the binning in-thread, the fx half with one shuffle.
//
// One CTA of 16 warps per SM (four groups, 4 x 128 TMEM columns, single-stage tiles).  Every wait is bounded.
#pragma once
#include "gemm_tc.cuh"
#include "wfs_tma.cuh"

#define WTC_WARPS 16
#define WTC_A_BYTES (128 * 32 * 2)            // one hi or lo A tile: 128 rows x 32 fp16
#define WTC_B_BYTES (64 * 32 * 2)             // one hi or lo B tile: 64 rows x 32 fp16

struct WfsTcParams {
  WfsParams p;
  WfsFast f;
  const uint4* b2;            // [2][WTC_B_BYTES / 16] stage-2 B tiles (hi, lo) in the UMMA canonical layout
  CUtensorMap maps[WFT_MAX_LAYERS];
};

template <int NL>
constexpr size_t wtc_smem_bytes(int gw, int nvalid_smem) {
  return (size_t)WTC_WARPS * (NL > 0 ? NL : 1) * WFT_TILE_STRIDE        // tiles, one stage per warp
         + 4 * 2 * WTC_A_BYTES + 2 * WTC_B_BYTES                         // A tiles per group (hi, lo), B tiles
         + 8 * 32 * 16 + 256 * 4 + 2 * WFT_NG * 16 * 4 + WTC_WARPS * 2 * 32 * 4
         + (WTC_WARPS + 8) * 8 + 64                                      // mbarriers, counters, TMEM slot
         + (((size_t)gw * gw * 2 + 15) & ~(size_t)15) + (size_t)nvalid_smem * 8;
}

__device__ __forceinline__ void wtc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

template <int NL>
__global__ void __launch_bounds__(WTC_WARPS * 32, 1) wfs_frame_tc_kernel(const __grid_constant__ WfsTcParams P) {
  const WfsParams& p = P.p;
  const WfsFast& f = P.f;
  extern __shared__ __align__(1024) unsigned char wtc_smem_raw[];
  unsigned char* sm = wtc_smem_raw;
  constexpr int NLS = NL > 0 ? NL : 1;
  unsigned char* s_tiles = sm;                                                   // [warp][layer][WFT_TILE_STRIDE]
  unsigned char* s_a = s_tiles + (size_t)WTC_WARPS * NLS * WFT_TILE_STRIDE;      // [group][hi, lo][WTC_A_BYTES]
  unsigned char* s_b = s_a + 4 * 2 * WTC_A_BYTES;                                // [hi, lo][WTC_B_BYTES]
  uint4* s_c1 = (uint4*)(s_b + 2 * WTC_B_BYTES);                                 // [8][32] stage-1 A fragments
  float* s_half = (float*)(s_c1 + 8 * 32);
  float* s_fx = s_half + 256;
  float* s_fy = s_fx + WFT_NG * 16;
  float* s_vall = s_fy + WFT_NG * 16;                                            // [warp][2][32]
  uint64_t* s_bar = (uint64_t*)(s_vall + WTC_WARPS * 2 * 32);                    // [warp] tile barrier, then [group][2] D_full
  uint32_t* s_cnt = (uint32_t*)(s_bar + WTC_WARPS + 8);                          // [group] arrival counters, [8] = TMEM slot
  short* s_amap = (short*)(s_cnt + 16);
  uint2* s_sub = (uint2*)((unsigned char*)s_amap + ((f.GW * f.GW * 2 + 15) & ~15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int grp = warp >> 2, wi = warp & 3;

  for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) s_c1[i] = f.c1[i];
  for (int i = threadIdx.x; i < 2 * WTC_B_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4*>(s_b)[i] = P.b2[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_half[i] = p.halfxy[i];
  for (int i = threadIdx.x; i < 2 * WFT_NG * 16; i += blockDim.x) s_fx[i] = f.fxy[i];
  for (int i = threadIdx.x; i < f.GW * f.GW; i += blockDim.x) s_amap[i] = f.amap[i];
  if (f.sub_in_smem)
    for (int i = threadIdx.x; i < p.nvalid; i += blockDim.x) s_sub[i] = f.sub[i];
  if (threadIdx.x < WTC_WARPS + 8)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(wft_smem_u32(s_bar + threadIdx.x)) : "memory");
  if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0u;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the B tiles are read by the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wft_smem_u32(s_cnt + 8)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_cnt[8];
  const uint32_t tmem_grp = tmem_base + (uint32_t)(grp * 128);          // two 64-column accumulators per group

  const uint2* sub = f.sub_in_smem ? s_sub : f.sub;
  unsigned char* my_tiles = s_tiles + (size_t)warp * NLS * WFT_TILE_STRIDE;
  const uint32_t my_tiles_u32 = wft_smem_u32(my_tiles);
  const uint32_t my_bar_u32 = wft_smem_u32(s_bar + warp);
  const uint32_t dfull_u32 = wft_smem_u32(s_bar + WTC_WARPS + grp * 2);
  unsigned char* my_a = s_a + (size_t)grp * 2 * WTC_A_BYTES;
  float* s_v = s_vall + warp * 2 * 32;
  const int lane_off = (2 * g) * WFT_TILE_W + 4 * q;
  const int amap_lane = (lane >> 2) * f.GW + (lane & 3);
  const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32

  const long long total = (long long)p.E * p.nvalid;
  const long long base = (long long)blockIdx.x * f.items_per_cta;
  long long end = base + f.items_per_cta;
  if (end > total) end = total;
  const int n_cta = (int)(end - base);                         // > 0 by construction of the grid
  const int n_iter = (n_cta + WTC_WARPS - 1) / WTC_WARPS;      // every warp runs the same number of iterations
  long long w = base + warp;

  int e = (int)(w / p.nvalid), k = (int)(w % p.nvalid);
  int ring_e = -1;
  int rx[NLS], ry[NLS];
  uint32_t tile_phase = 0;

  uint32_t n_xy = 0, n_pm = 0, n_d = 0;
  float n_v = 0.f;
  bool n_seam = false;
  auto prefetch = [&](int pe, int pk, bool tiles, bool aux) {
    const uint2 sb = sub[pk];
    n_xy = sb.x;
    if (NL > 0 && tiles) {
      if (pe != ring_e) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
          const int N = p.layer[l].N;
          int a = p.layer[l].ix + p.layer[l].ox[pe];  a -= (a >= N) ? N : 0;
          int b = p.layer[l].iy + p.layer[l].oy[pe];  b -= (b >= N) ? N : 0;
          rx[l] = a; ry[l] = b;
        }
        ring_e = pe;
      }
      const int x0 = (int)(sb.x & 0xffffu), y0 = (int)(sb.x >> 16);
      int tc[NLS], tr[NLS];
      bool seam = false;
      uint32_t dbits = 0;
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const int N = p.layer[l].N;
        int c = x0 + rx[l];  c -= (c >= N) ? N : 0;
        int r = y0 + ry[l];  r -= (r >= N) ? N : 0;
        tc[l] = c & ~3; tr[l] = r;
        dbits |= (uint32_t)(c & 3) << (2 * l);
        seam |= (c + WFT_TILE_H > N) | (r + WFT_TILE_H > N);
      }
      n_seam = seam;
      n_d = dbits;
      if (!seam && lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(my_bar_u32), "r"(NL * WFT_TILE_BYTES) : "memory");
#pragma unroll
        for (int l = 0; l < NL; ++l)
          wft_tma_load_3d(my_tiles_u32 + l * WFT_TILE_STRIDE, &P.maps[l], tc[l], tr[l], pe, my_bar_u32);
      }
    }
    if (aux) {
      n_v = 0.f;
      if (p.use_dm && lane < 18) {
        int idx = (lane < 16) ? (int)s_amap[(int)sb.y + amap_lane] : p.pzt_nact + lane - 16;
        if (idx >= 0) n_v = __ldg(p.volts + (size_t)pe * p.ldv + idx);
      }
      n_pm = f.pmask[(size_t)pk * 32 + lane];
    }
  };

  // ---- epilogue of one item from a TMEM accumulator: lane = fx row (kept index), columns = {Yr, Yi} x 32 fy ----
  auto epilogue = [&](int buf, int ie, int ik) {
    float pixv[16];        // this fx row: 16 fy bins, first the 8 of kept fy 0..15 (py 8..15), then py 0..7
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      const uint32_t taddr = tmem_grp + ((uint32_t)(wi * 32) << 16) + (uint32_t)(buf * 64 + half * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int m = 0; m < 8; ++m) {                 // columns 4m .. 4m+3 = {Yr, Yi}(fy 2m'), {Yr, Yi}(fy 2m'+1)
        const float a0 = __uint_as_float(v[4 * m]), a1 = __uint_as_float(v[4 * m + 1]);
        const float a2 = __uint_as_float(v[4 * m + 2]), a3 = __uint_as_float(v[4 * m + 3]);
        pixv[half * 8 + m] = fmaf(a3, a3, fmaf(a2, a2, fmaf(a1, a1, a0 * a0)));
      }
    }
    // fx half of the binning: rows 2m, 2m+1 are lanes 2m, 2m+1; the even lane keeps slots 0..7, the odd one 8..15
    float mine[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float give = (lane & 1) ? pixv[j] : pixv[8 + j];
      const float keep = (lane & 1) ? pixv[8 + j] : pixv[j];
      mine[j] = keep + __shfl_xor_sync(0xffffffffu, give, 1);
    }
    const int pr = lane >> 1;                                  // fx pair: kept indices 2 pr, 2 pr + 1
    const int px = (pr < 8) ? 8 + pr : pr - 8;
    const int py0 = (lane & 1) ? 0 : 8;                        // slots 0..7 -> py 8..15, slots 8..15 -> py 0..7
    const bool plain = (p.noise < 0.f) && (p.bincube == nullptr);
    float s0 = 0.f, sx = 0.f, sy = 0.f;
    if (plain) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s0 += mine[j];
        sy = fmaf(mine[j], (float)(py0 + j), sy);
      }
      sx = s0 * (float)px;
    } else {
      float tot = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) tot += mine[j];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, sft);
      const float scale = p.nphotons * p.flux[ik] / tot;
      const uint32_t k0 = p.k0[ie], k1 = p.k1[ie];
      float* cube = p.bincube ? p.bincube + ((size_t)ie * p.nvalid + ik) * 256 : nullptr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int py = py0 + j;
        const int pidx = py * 16 + px;
        float v = mine[j] * scale;
        v = aom_pixel_noise(v, p.noise, (uint32_t)(ik * 256 + pidx), p.frame, p.wfs_index, k0, k1);
        if (cube) cube[pidx] = v;
        s0 += v;
        sx = fmaf(v, (float)px, sx);
        sy = fmaf(v, (float)py, sy);
      }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, sft);
      sx += __shfl_xor_sync(0xffffffffu, sx, sft);
      sy += __shfl_xor_sync(0xffffffffu, sy, sft);
    }
    if (lane == 0) {
      const float inv = __frcp_rn(s0);
      const float gx = (s0 > 0.f) ? sx * inv : p.cog_offset;
      const float gy = (s0 > 0.f) ? sy * inv : p.cog_offset;
      float* sl = p.slopes + (size_t)ie * p.lds;
      sl[ik] = (gx - p.cog_offset) * p.pixsize;
      sl[p.nvalid + ik] = (gy - p.cog_offset) * p.pixsize;
    }
  };

  bool valid = w < end;
  if (valid) prefetch(e, k, true, true);
  s_v[lane] = n_v;
  __syncwarp();

  int pe = 0, pk = 0;          // item whose accumulator is pending in TMEM
  bool pvalid = false;

  for (int it = 0; it < n_iter; ++it) {
    const int s = it & 1;
    const uint32_t c_xy = n_xy, c_pm = n_pm, c_d = n_d;
    const bool c_seam = n_seam, c_valid = valid;
    const int ce = e, ck = k;
    const int x0 = (int)(c_xy & 0xffffu), y0 = (int)(c_xy >> 16);

    // ---- next work item of this warp: aux loads now, tiles as soon as the current ones are sampled ----
    w += WTC_WARPS;
    valid = w < end;
    if (valid) {
      k += WTC_WARPS;
      if (k >= p.nvalid) { k -= p.nvalid; e += 1; }
      prefetch(e, k, false, true);
    }

    uint32_t th[2][2][2], tl[2][2][2], uh[2][2][2], ul[2][2][2];   // [u][row g / g+8][k-pair]: Tr hi, lo, Ti hi, lo
    if (c_valid) {
      float4 tta[2], ttb[2];
      if (p.use_dm) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const size_t to = (size_t)(y0 + 2 * g + r + p.tt_off) * p.tt_dim + (x0 + 4 * q + p.tt_off);
          tta[r] = __ldg(reinterpret_cast<const float4*>(p.tt_planes + to));
          ttb[r] = __ldg(reinterpret_cast<const float4*>(p.tt_planes + (size_t)p.tt_dim * p.tt_dim + to));
        }
      }
      float ph[2][4];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) ph[r][c] = 0.f;

      // ---- atmosphere ----
      if (NL > 0) {
        if (!c_seam) {
          wft_mbar_wait(my_bar_u32, tile_phase, f.err);
          tile_phase ^= 1u;
        } else {
#pragma unroll 1
          for (int l = 0; l < NL; ++l) {
            const int N = p.layer[l].N;
            const float* scr = p.layer[l].screen + (size_t)ce * N * N;
            int c0 = x0 + p.layer[l].ix + p.layer[l].ox[ce];  c0 -= (c0 >= N) ? N : 0;  c0 -= (c0 >= N) ? N : 0;
            int r0 = y0 + p.layer[l].iy + p.layer[l].oy[ce];  r0 -= (r0 >= N) ? N : 0;  r0 -= (r0 >= N) ? N : 0;
            float* tile = reinterpret_cast<float*>(my_tiles + l * WFT_TILE_STRIDE);
            for (int i = lane; i < WFT_TILE_H * WFT_TILE_H; i += 32) {
              const int r = i / WFT_TILE_H, c = i - r * WFT_TILE_H;
              int rr = r0 + r;  rr -= (rr >= N) ? N : 0;
              int cc = c0 + c;  cc -= (cc >= N) ? N : 0;
              tile[r * WFT_TILE_W + (c0 & 3) + c] = __ldg(scr + (size_t)rr * N + cc);
            }
          }
          __syncwarp();
        }
#pragma unroll
        for (int l = 0; l < NL; ++l) {
          const float* t = reinterpret_cast<const float*>(my_tiles + l * WFT_TILE_STRIDE) + lane_off;
          switch ((c_d >> (2 * l)) & 3u) {
            case 0: wft_layer<0>(t, p.layer[l], ph); break;
            case 1: wft_layer<1>(t, p.layer[l], ph); break;
            case 2: wft_layer<2>(t, p.layer[l], ph); break;
            default: wft_layer<3>(t, p.layer[l], ph); break;
          }
        }
        __syncwarp();                                   // the stage is drained: re-arm it with the next item
        if (valid) prefetch(e, k, true, false);
      }

      // ---- mirrors ----
      if (p.use_dm) {
        const float* V = s_v + s * 32;
        float u[2][WFT_NG];
#pragma unroll
        for (int jx = 0; jx < WFT_NG; ++jx) u[0][jx] = u[1][jx] = 0.f;
#pragma unroll
        for (int jy = 0; jy < WFT_NG; ++jy) {
          const float4 vr = *reinterpret_cast<const float4*>(V + jy * 4);
          const float2 fyv = *reinterpret_cast<const float2*>(s_fy + jy * 16 + 2 * g);
          u[0][0] = fmaf(fyv.x, vr.x, u[0][0]); u[0][1] = fmaf(fyv.x, vr.y, u[0][1]);
          u[0][2] = fmaf(fyv.x, vr.z, u[0][2]); u[0][3] = fmaf(fyv.x, vr.w, u[0][3]);
          u[1][0] = fmaf(fyv.y, vr.x, u[1][0]); u[1][1] = fmaf(fyv.y, vr.y, u[1][1]);
          u[1][2] = fmaf(fyv.y, vr.z, u[1][2]); u[1][3] = fmaf(fyv.y, vr.w, u[1][3]);
        }
        float dm[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) dm[r][c] = 0.f;
#pragma unroll
        for (int jx = 0; jx < WFT_NG; ++jx) {
          const float4 fxv = *reinterpret_cast<const float4*>(s_fx + jx * 16 + 4 * q);
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            dm[r][0] = fmaf(u[r][jx], fxv.x, dm[r][0]);
            dm[r][1] = fmaf(u[r][jx], fxv.y, dm[r][1]);
            dm[r][2] = fmaf(u[r][jx], fxv.z, dm[r][2]);
            dm[r][3] = fmaf(u[r][jx], fxv.w, dm[r][3]);
          }
        }
        const float tt0 = V[16], tt1 = V[17];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          ph[r][0] += fmaf(tt1, ttb[r].x, fmaf(tt0, tta[r].x, dm[r][0]));
          ph[r][1] += fmaf(tt1, ttb[r].y, fmaf(tt0, tta[r].y, dm[r][1]));
          ph[r][2] += fmaf(tt1, ttb[r].z, fmaf(tt0, tta[r].z, dm[r][2]));
          ph[r][3] += fmaf(tt1, ttb[r].w, fmaf(tt0, tta[r].w, dm[r][3]));
        }
      }

      // ---- complex field -> stage-1 B fragments ----
      uint32_t xr_h[2][2], xr_l[2][2], xi_h[2][2], xi_l[2][2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float4 hf = *reinterpret_cast<const float4*>(s_half + (2 * g + r) * 16 + 4 * q);
        const float hv[4] = {hf.x, hf.y, hf.z, hf.w};
        float re[4], im[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float sn, cs;
          wfm_sincos(p.k2 * ph[r][c] - hv[c], sn, cs);
          const bool on = (c_pm >> (r * 4 + c)) & 1u;
          re[c] = on ? cs : 0.f;
          im[c] = on ? sn : 0.f;
        }
        wfm_split(re[0], re[1], xr_h[r][0], xr_l[r][0]);
        wfm_split(re[2], re[3], xr_h[r][1], xr_l[r][1]);
        wfm_split(im[0], im[1], xi_h[r][0], xi_l[r][0]);
        wfm_split(im[2], im[3], xi_h[r][1], xi_l[r][1]);
      }

      // ---- stage 1 on the warp-level tensor path (unchanged) ----
      float T[2][2][2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) T[a][mt][j][c] = 0.f;
      {
        uint32_t nxi_h[2][2], nxi_l[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int c = 0; c < 2; ++c) { nxi_h[j][c] = xi_h[j][c] ^ 0x80008000u; nxi_l[j][c] = xi_l[j][c] ^ 0x80008000u; }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint4 wr_h = s_c1[(mt * 4 + 0) * 32 + lane], wr_l = s_c1[(mt * 4 + 1) * 32 + lane];
          const uint4 wi_h = s_c1[(mt * 4 + 2) * 32 + lane], wi_l = s_c1[(mt * 4 + 3) * 32 + lane];
#define WTC_S1(WR, WI, XR, XI, NXI)                                    \
  _Pragma("unroll") for (int j = 0; j < 2; ++j) {                      \
    wfm_mma(T[0][mt][j], WR, XR[j][0], XR[j][1]);                      \
    wfm_mma(T[1][mt][j], WI, XR[j][0], XR[j][1]);                      \
  }                                                                    \
  _Pragma("unroll") for (int j = 0; j < 2; ++j) {                      \
    wfm_mma(T[0][mt][j], WI, NXI[j][0], NXI[j][1]);                    \
    wfm_mma(T[1][mt][j], WR, XI[j][0], XI[j][1]);                      \
  }
          WTC_S1(wr_h, wi_h, xr_h, xi_h, nxi_h)
          WTC_S1(wr_h, wi_h, xr_l, xi_l, nxi_l)
          WTC_S1(wr_l, wi_l, xr_h, xi_h, nxi_h)
#undef WTC_S1
        }
      }
      // ---- T of this lane's four fx rows as fp16 hi / lo, y = 4q .. 4q+3 contiguous: (j0,c0) (j1,c0) (j0,c1) (j1,c1) ----
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
          wfm_split(T[0][u][0][2 * rs], T[0][u][1][2 * rs], th[u][rs][0], tl[u][rs][0]);
          wfm_split(T[0][u][0][2 * rs + 1], T[0][u][1][2 * rs + 1], th[u][rs][1], tl[u][rs][1]);
          wfm_split(T[1][u][0][2 * rs], T[1][u][1][2 * rs], uh[u][rs][0], ul[u][rs][0]);
          wfm_split(T[1][u][0][2 * rs + 1], T[1][u][1][2 * rs + 1], uh[u][rs][1], ul[u][rs][1]);
        }
    }

    // ---- the group's previous MMAs are done: A tile reusable, accumulator (it-1) & 1 complete ----
    if (it > 0) {
      wft_mbar_wait(dfull_u32 + ((it - 1) & 1) * 8, (uint32_t)(((it - 1) >> 1) & 1), f.err);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (c_valid) {
      // row R = 32 wi + 16 u + 2 g + rs ; K chunk = 2 part + (q >> 1) ; 8 bytes at (q & 1) * 8
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
          const int R = 32 * wi + 16 * u + 2 * g + rs;
          const uint32_t off = (uint32_t)((R >> 3) * 512 + (q >> 1) * 128 + (R & 7) * 16 + (q & 1) * 8);
          *reinterpret_cast<uint2*>(my_a + off) = make_uint2(th[u][rs][0], th[u][rs][1]);
          *reinterpret_cast<uint2*>(my_a + off + 256) = make_uint2(uh[u][rs][0], uh[u][rs][1]);
          *reinterpret_cast<uint2*>(my_a + WTC_A_BYTES + off) = make_uint2(tl[u][rs][0], tl[u][rs][1]);
          *reinterpret_cast<uint2*>(my_a + WTC_A_BYTES + off + 256) = make_uint2(ul[u][rs][0], ul[u][rs][1]);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      const uint32_t old = atomicAdd(&s_cnt[grp], 1u);
      if ((old & 3u) == 3u) {
        // last of the four warps: all 128 rows are in place -> six MMAs into accumulator it & 1
        __threadfence_block();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = wft_smem_u32(my_a), a_lo = a_hi + WTC_A_BYTES;
        const uint32_t b_hi = wft_smem_u32(s_b), b_lo = b_hi + WTC_B_BYTES;
        const uint32_t d = tmem_grp + (uint32_t)((it & 1) * 64);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) wtc_mma_f16(d, gtc_desc(a_hi + ks * 256), gtc_desc(b_hi + ks * 256), idesc, ks ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) wtc_mma_f16(d, gtc_desc(a_lo + ks * 256), gtc_desc(b_hi + ks * 256), idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) wtc_mma_f16(d, gtc_desc(a_hi + ks * 256), gtc_desc(b_lo + ks * 256), idesc, 1u);
        gtc_commit(dfull_u32 + (it & 1) * 8);
      }
    }
    __syncwarp();

    // ---- epilogue of the previous item (its MMAs completed while this item's field was computed) ----
    if (it > 0 && pvalid) epilogue((it - 1) & 1, pe, pk);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    pe = ce; pk = ck; pvalid = c_valid;

    // ---- hand the prefetched neighbourhood volts to the next iteration ----
    if (valid) s_v[(s ^ 1) * 32 + lane] = n_v;
    __syncwarp();
  }
  // ---- drain: accumulator of the last iteration ----
  wft_mbar_wait(dfull_u32 + ((n_iter - 1) & 1) * 8, (uint32_t)(((n_iter - 1) >> 1) & 1), f.err);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (pvalid) epilogue((n_iter - 1) & 1, pe, pk);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}
