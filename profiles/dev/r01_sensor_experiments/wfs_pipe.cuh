// Shack-Hartmann frame, software-pipelined across subapertures inside every warp.
//
// Same tables, tiles, fragment mapping and arithmetic as wfs_frame_tma_kernel (wfs_tma.cuh); what changes is the
// order of the work inside a warp.  Measurements on B200: with 4 warps per scheduler each
// warp runs latency bound, and its field arithmetic (~6300 cycles per subaperture) and its two MMA stages
// (~3200 cycles, 144 HMMA at >= 8 pipe cycles each) simply add up, although the issue port (1322 slots) and the
// tensor pipe (1152 cycles) would both fit in a fraction of that.  Here iteration `it` of a warp
//
//     A  waits for the staged tiles of item it+1 and samples the three layers            (branchy: seam, offset)
//     B  issues the stage-1 MMAs of item it   || mirror surface + tip-tilt of item it+1  (one basic block)
//     C  splits T of item it into fp16 hi / lo fragments
//     D  issues the stage-2 MMAs of item it   || exp(i phi) + fp16 split of item it+1    (one basic block)
//     E  bins, normalises, centre of gravity of item it
//
// so that the long-latency HMMA chains of one subaperture are filled with the independent FP32 work of the next
// one from the same instruction stream.  The TMA boxes of item it+3 are issued as soon as A has drained the
// stage of item it+1 (two iterations of flight), the neighbourhood volts / pupil byte of item it+2 are loaded in
// iteration `it`.  Work that belongs to a non-existent next item (last iteration) is executed on the clamped
// last item and discarded, which keeps B and D free of branches.
#pragma once
#include "wfs_tma.cuh"

struct WftItem {
  int e, k;
  uint32_t xy;        // y0 << 16 | x0
  uint32_t ds;        // bit 0: tile set straddles the torus seam; bits 1..: 2 bits per layer = tile origin column & 3
};

template <int FULL, int NL, int DM>
__global__ void __launch_bounds__(WFT_WARPS * 32, 2) wfs_frame_pipe_kernel(const __grid_constant__ WfsTmaParams P) {
  const WfsParams& p = P.p;
  const WfsFast& f = P.f;
  extern __shared__ __align__(128) unsigned char wft_smem_raw[];
  unsigned char* sm = wft_smem_raw;
  constexpr int NLS = NL > 0 ? NL : 1;
  unsigned char* s_tiles = sm;
  uint64_t* s_bar = (uint64_t*)(s_tiles + (size_t)WFT_WARPS * 2 * NLS * WFT_TILE_STRIDE);
  uint4* s_c1 = (uint4*)(s_bar + WFT_WARPS * 2);
  uint4* s_c2 = s_c1 + 8 * 32;
  float* s_half = (float*)(s_c2 + 12 * 32);
  float* s_fx = s_half + 256;
  float* s_fy = s_fx + WFT_NG * 16;
  float* s_vall = s_fy + WFT_NG * 16;
  short* s_amap = (short*)(s_vall + WFT_WARPS * 2 * 32);
  uint2* s_sub = (uint2*)((unsigned char*)s_amap + ((f.GW * f.GW * 2 + 15) & ~15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;

  for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) s_c1[i] = f.c1[i];
  for (int i = threadIdx.x; i < 12 * 32; i += blockDim.x) s_c2[i] = f.c2[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_half[i] = p.halfxy[i];
  for (int i = threadIdx.x; i < 2 * WFT_NG * 16; i += blockDim.x) s_fx[i] = f.fxy[i];
  for (int i = threadIdx.x; i < f.GW * f.GW; i += blockDim.x) s_amap[i] = f.amap[i];
  if (f.sub_in_smem)
    for (int i = threadIdx.x; i < p.nvalid; i += blockDim.x) s_sub[i] = f.sub[i];
  if (threadIdx.x < WFT_WARPS * 2)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(wft_smem_u32(s_bar + threadIdx.x)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const uint2* sub = f.sub_in_smem ? s_sub : f.sub;
  unsigned char* my_tiles = s_tiles + (size_t)warp * 2 * NLS * WFT_TILE_STRIDE;
  const uint32_t my_tiles_u32 = wft_smem_u32(my_tiles);
  const uint32_t my_bar_u32 = wft_smem_u32(s_bar + warp * 2);
  float* s_v = s_vall + warp * 2 * 32;
  const int lane_off = (2 * g) * WFT_TILE_W + 4 * q;
  const int amap_lane = (lane >> 2) * f.GW + (lane & 3);

  const long long total = (long long)p.E * p.nvalid;
  const long long base = (long long)blockIdx.x * f.items_per_cta;
  long long end = base + f.items_per_cta;
  if (end > total) end = total;
  const long long w0 = base + warp;
  if (w0 >= end) return;
  const int n_items = (int)((end - w0 + WFT_WARPS - 1) / WFT_WARPS);

  // ---- item bookkeeping: items j = 0 .. n_items-1 of this warp are work items w0 + 8 j ----
  int ring_e = -1;
  int rx[NLS], ry[NLS];                       // (ix + ox[e]) mod N, (iy + oy[e]) mod N of the cached environment
  uint32_t phase_bits = 0;

  auto advance = [&](WftItem& it) {           // next item of this warp (clamped at the last one by the callers)
    it.k += WFT_WARPS;
    if (it.k >= p.nvalid) { it.k -= p.nvalid; it.e += 1; }
  };
  // tile origins of an item; returns seam flag | offset bits, fills tc / tr
  auto tile_coords = [&](const WftItem& it, int (&tc)[NLS], int (&tr)[NLS]) -> uint32_t {
    uint32_t ds = 0;
    if (NL > 0) {
      if (it.e != ring_e) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
          const int N = p.layer[l].N;
          int a = p.layer[l].ix + p.layer[l].ox[it.e];  a -= (a >= N) ? N : 0;
          int b = p.layer[l].iy + p.layer[l].oy[it.e];  b -= (b >= N) ? N : 0;
          rx[l] = a; ry[l] = b;
        }
        ring_e = it.e;
      }
      const int x0 = (int)(it.xy & 0xffffu), y0 = (int)(it.xy >> 16);
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const int N = p.layer[l].N;
        int c = x0 + rx[l];  c -= (c >= N) ? N : 0;
        int r = y0 + ry[l];  r -= (r >= N) ? N : 0;
        tc[l] = c; tr[l] = r;
        ds |= (uint32_t)(c & 3) << (1 + 2 * l);
        ds |= ((c + WFT_TILE_H > N) | (r + WFT_TILE_H > N)) ? 1u : 0u;
      }
    }
    return ds;
  };
  // TMA boxes of an item into its stage (item index parity)
  auto issue_tiles = [&](WftItem& it, int stage) {
    it.xy = sub[it.k].x;
    if (NL > 0) {
      int tc[NLS], tr[NLS];
      it.ds = tile_coords(it, tc, tr);
      if (!(it.ds & 1u) && lane == 0) {
        const uint32_t bar = my_bar_u32 + stage * 8;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(NL * WFT_TILE_BYTES) : "memory");
#pragma unroll
        for (int l = 0; l < NL; ++l)
          wft_tma_load_3d(my_tiles_u32 + (stage * NL + l) * WFT_TILE_STRIDE, &P.maps[l], tc[l] & ~3, tr[l], it.e, bar);
      }
    } else {
      it.ds = 0;
    }
  };
  // neighbourhood volts (lanes 0..15), tip-tilt volts (lanes 16, 17) and the pupil byte of an item
  auto load_aux = [&](const WftItem& it, float& v, uint32_t& pm) {
    v = 0.f;
    if (DM && lane < 18) {
      const int idx = (lane < 16) ? (int)s_amap[(int)sub[it.k].y + amap_lane] : p.pzt_nact + lane - 16;
      if (idx >= 0) v = __ldg(p.volts + (size_t)it.e * p.ldv + idx);
    }
    pm = f.pmask[(size_t)it.k * 32 + lane];
  };
  // region A: atmosphere of an item from its staged tiles
  auto sample_layers = [&](const WftItem& it, int stage, float (&ph)[2][4]) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) ph[r][c] = 0.f;
    if (NL > 0) {
      if (!(it.ds & 1u)) {
        wft_mbar_wait(my_bar_u32 + stage * 8, (phase_bits >> stage) & 1u, f.err);
        phase_bits ^= (1u << stage);
      } else {
        const int x0 = (int)(it.xy & 0xffffu), y0 = (int)(it.xy >> 16);
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
          const int N = p.layer[l].N;
          const float* scr = p.layer[l].screen + (size_t)it.e * N * N;
          int c0 = x0 + p.layer[l].ix + p.layer[l].ox[it.e];  c0 -= (c0 >= N) ? N : 0;  c0 -= (c0 >= N) ? N : 0;
          int r0 = y0 + p.layer[l].iy + p.layer[l].oy[it.e];  r0 -= (r0 >= N) ? N : 0;  r0 -= (r0 >= N) ? N : 0;
          float* tile = reinterpret_cast<float*>(my_tiles + (stage * NL + l) * WFT_TILE_STRIDE);
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
        const float* t = reinterpret_cast<const float*>(my_tiles + (stage * NL + l) * WFT_TILE_STRIDE) + lane_off;
        switch ((it.ds >> (1 + 2 * l)) & 3u) {
          case 0: wft_layer<0>(t, p.layer[l], ph); break;
          case 1: wft_layer<1>(t, p.layer[l], ph); break;
          case 2: wft_layer<2>(t, p.layer[l], ph); break;
          default: wft_layer<3>(t, p.layer[l], ph); break;
        }
      }
      __syncwarp();      // every lane has drained the stage before lane 0 re-arms it
    }
  };
  // tip-tilt planes of an item (L2-resident tables)
  auto load_tt = [&](const WftItem& it, float4 (&tta)[2], float4 (&ttb)[2]) {
    if (DM) {
      const int x0 = (int)(it.xy & 0xffffu), y0 = (int)(it.xy >> 16);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const size_t to = (size_t)(y0 + 2 * g + r + p.tt_off) * p.tt_dim + (x0 + 4 * q + p.tt_off);
        tta[r] = __ldg(reinterpret_cast<const float4*>(p.tt_planes + to));
        ttb[r] = __ldg(reinterpret_cast<const float4*>(p.tt_planes + (size_t)p.tt_dim * p.tt_dim + to));
      }
    }
  };
  // mirror surface (separable stamps of the 4 x 4 neighbourhood + tip-tilt) added to ph; branch free
  auto add_mirrors = [&](const float* V, const float4 (&tta)[2], const float4 (&ttb)[2], float (&ph)[2][4]) {
    if (DM) {
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
  };
  // complex field of one pixel row -> fp16 hi / lo B fragments of stage 1 (row 2g + r feeds n-tile r); branch free
  auto field_row = [&](int r, const float (&ph)[2][4], uint32_t pm, uint32_t (&xrh)[2][2], uint32_t (&xrl)[2][2],
                       uint32_t (&xih)[2][2], uint32_t (&xil)[2][2]) {
    const float4 hf = *reinterpret_cast<const float4*>(s_half + (2 * g + r) * 16 + 4 * q);
    const float hv[4] = {hf.x, hf.y, hf.z, hf.w};
    float re[4], im[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float sn, cs;
      wfm_sincos(p.k2 * ph[r][c] - hv[c], sn, cs);
      const bool on = (pm >> (r * 4 + c)) & 1u;
      re[c] = on ? cs : 0.f;
      im[c] = on ? sn : 0.f;
    }
    wfm_split(re[0], re[1], xrh[r][0], xrl[r][0]);
    wfm_split(re[2], re[3], xrh[r][1], xrl[r][1]);
    wfm_split(im[0], im[1], xih[r][0], xil[r][0]);
    wfm_split(im[2], im[3], xih[r][1], xil[r][1]);
  };

  // ---- prologue: tiles of items 0, 1, 2 in flight; field of item 0 ----
  WftItem cur, nx1, nx2;                       // items it, it+1, it+2 (clamped at the last item)
  cur.e = (int)(w0 / p.nvalid); cur.k = (int)(w0 % p.nvalid); cur.xy = 0; cur.ds = 0;
  issue_tiles(cur, 0);
  nx1 = cur;
  if (n_items > 1) { advance(nx1); issue_tiles(nx1, 1); }
  nx2 = nx1;
  if (n_items > 2) advance(nx2);               // its tiles go out once stage 0 has been drained
  uint32_t xr_h[2][2], xr_l[2][2], xi_h[2][2], xi_l[2][2];
  float v_n1, v_n2 = 0.f;                      // aux of items it+1, it+2
  uint32_t pm_n1, pm_n2 = 0;
  {
    float v0;
    uint32_t pm0;
    load_aux(cur, v0, pm0);
    load_aux(nx1, v_n1, pm_n1);
    s_v[lane] = v0;
    s_v[32 + lane] = v_n1;
    __syncwarp();
    float ph[2][4];
    float4 tta[2], ttb[2];
    load_tt(cur, tta, ttb);
    sample_layers(cur, 0, ph);
    if (n_items > 2) issue_tiles(nx2, 0); else nx2.xy = sub[nx2.k].x;
    add_mirrors(s_v, tta, ttb, ph);
    field_row(0, ph, pm0, xr_h, xr_l, xi_h, xi_l);
    field_row(1, ph, pm0, xr_h, xr_l, xi_h, xi_l);
  }

  for (int it = 0; it < n_items; ++it) {
    const int s1 = (it + 1) & 1;               // stage of item it+1 (and of item it+3)
    const bool has1 = it + 1 < n_items;

    // ---- A: atmosphere of item it+1; then re-arm its stage with item it+3; aux loads of item it+2 ----
    float ph[2][4];
    float4 tta[2], ttb[2];
    load_tt(nx1, tta, ttb);
    if (has1) {
      sample_layers(nx1, s1, ph);
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) ph[r][c] = 0.f;
    }
    WftItem nx3 = nx2;
    if (it + 3 < n_items) { advance(nx3); issue_tiles(nx3, s1); }
    if (it + 2 < n_items) load_aux(nx2, v_n2, pm_n2);

    // ---- B: stage-1 MMAs of item `it` || mirrors of item it+1 ----
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
        // Tr = Wr.Xr - Wi.Xi     Ti = Wi.Xr + Wr.Xi     (hi.hi, hi.lo, lo.hi); four tiles advance in lock step
#define WFP_S1(WR, WI, XR, XI, NXI)                                    \
  _Pragma("unroll") for (int j = 0; j < 2; ++j) {                      \
    wfm_mma(T[0][mt][j], WR, XR[j][0], XR[j][1]);                      \
    wfm_mma(T[1][mt][j], WI, XR[j][0], XR[j][1]);                      \
  }                                                                    \
  _Pragma("unroll") for (int j = 0; j < 2; ++j) {                      \
    wfm_mma(T[0][mt][j], WI, NXI[j][0], NXI[j][1]);                    \
    wfm_mma(T[1][mt][j], WR, XI[j][0], XI[j][1]);                      \
  }
        WFP_S1(wr_h, wi_h, xr_h, xi_h, nxi_h)
        WFP_S1(wr_h, wi_h, xr_l, xi_l, nxi_l)
        WFP_S1(wr_l, wi_l, xr_h, xi_h, nxi_h)
#undef WFP_S1
      }
    }
    add_mirrors(s_v + s1 * 32, tta, ttb, ph);

    // ---- C: stage-2 A fragments of item `it` ----
    uint4 tr_h[2], tr_l[2], ti_h[2], ti_l[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      wfm_split(T[0][u][0][0], T[0][u][0][1], tr_h[u].x, tr_l[u].x);
      wfm_split(T[0][u][0][2], T[0][u][0][3], tr_h[u].y, tr_l[u].y);
      wfm_split(T[0][u][1][0], T[0][u][1][1], tr_h[u].z, tr_l[u].z);
      wfm_split(T[0][u][1][2], T[0][u][1][3], tr_h[u].w, tr_l[u].w);
      wfm_split(T[1][u][0][0], T[1][u][0][1], ti_h[u].x, ti_l[u].x);
      wfm_split(T[1][u][0][2], T[1][u][0][3], ti_h[u].y, ti_l[u].y);
      wfm_split(T[1][u][1][0], T[1][u][1][1], ti_h[u].z, ti_l[u].z);
      wfm_split(T[1][u][1][2], T[1][u][1][3], ti_h[u].w, ti_l[u].w);
    }

    // ---- D: stage-2 MMAs + |.|^2 + binning of item `it` || field of item it+1 (one pixel row per two fy tiles) ----
    uint32_t yr_h[2][2], yr_l[2][2], yi_h[2][2], yi_l[2][2];      // field fragments of item it+1
    float pix[2][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const uint4 ch = s_c2[(b * 3 + 0) * 32 + lane];
      const uint4 cn = s_c2[(b * 3 + 2) * 32 + lane];
      uint4 cl = make_uint4(0u, 0u, 0u, 0u);
      if (FULL) cl = s_c2[(b * 3 + 1) * 32 + lane];
      float yr[2][4], yi[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int c = 0; c < 4; ++c) { yr[u][c] = 0.f; yi[u][c] = 0.f; }
      // Yr = Tr.Wr - Ti.Wi     Yi = Tr.Wi + Ti.Wr
#pragma unroll
      for (int u = 0; u < 2; ++u) { wfm_mma(yr[u], tr_h[u], ch.x, ch.y); wfm_mma(yi[u], tr_h[u], ch.z, ch.w); }
#pragma unroll
      for (int u = 0; u < 2; ++u) { wfm_mma(yr[u], ti_h[u], cn.x, cn.y); wfm_mma(yi[u], ti_h[u], ch.x, ch.y); }
#pragma unroll
      for (int u = 0; u < 2; ++u) { wfm_mma(yr[u], tr_l[u], ch.x, ch.y); wfm_mma(yi[u], tr_l[u], ch.z, ch.w); }
#pragma unroll
      for (int u = 0; u < 2; ++u) { wfm_mma(yr[u], ti_l[u], cn.x, cn.y); wfm_mma(yi[u], ti_l[u], ch.x, ch.y); }
      if (FULL) {
#pragma unroll
        for (int u = 0; u < 2; ++u) { wfm_mma(yr[u], tr_h[u], cl.x, cl.y); wfm_mma(yi[u], tr_h[u], cl.z, cl.w); }
#pragma unroll
        for (int u = 0; u < 2; ++u) { wfm_mma(yr[u], ti_h[u], cn.z, cn.w); wfm_mma(yi[u], ti_h[u], cl.x, cl.y); }
      }
      if (b == 0) field_row(0, ph, pm_n1, yr_h, yr_l, yi_h, yi_l);
      if (b == 2) field_row(1, ph, pm_n1, yr_h, yr_l, yi_h, yi_l);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float a = yr[u][0] * yr[u][0];
        a = fmaf(yi[u][0], yi[u][0], a);
        a = fmaf(yr[u][1], yr[u][1], a);
        a = fmaf(yi[u][1], yi[u][1], a);
        float c = yr[u][2] * yr[u][2];
        c = fmaf(yi[u][2], yi[u][2], c);
        c = fmaf(yr[u][3], yr[u][3], c);
        c = fmaf(yi[u][3], yi[u][3], c);
        pix[u][b] = a + c;
      }
    }

    // ---- E: flux normalisation, noise, centre of gravity of item `it` ----
    const bool plain = (p.noise < 0.f) && (p.bincube == nullptr);     // centre of gravity is scale invariant
    float s0 = 0.f, sx = 0.f, sy = 0.f;
    if (plain) {
      const float r0 = (pix[0][0] + pix[0][1]) + (pix[0][2] + pix[0][3]);   // px = 8 + g
      const float r1 = (pix[1][0] + pix[1][1]) + (pix[1][2] + pix[1][3]);   // px = g
      s0 = r0 + r1;
      sx = fmaf(r0, (float)(8 + g), r1 * (float)g);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int py = ((b < 2) ? 8 + 4 * b : 4 * (b - 2)) + q;
        sy = fmaf(pix[0][b] + pix[1][b], (float)py, sy);
      }
    } else {
      float tot = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int b = 0; b < 4; ++b) tot += pix[u][b];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, sft);
      const float scale = p.nphotons * p.flux[cur.k] / tot;
      const uint32_t k0 = p.k0[cur.e], k1 = p.k1[cur.e];
      float* cube = p.bincube ? p.bincube + ((size_t)cur.e * p.nvalid + cur.k) * 256 : nullptr;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int px = 8 * (1 - u) + g;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int py = ((b < 2) ? 8 + 4 * b : 4 * (b - 2)) + q;
          const int pidx = py * 16 + px;
          float v = pix[u][b] * scale;
          v = aom_pixel_noise(v, p.noise, (uint32_t)(cur.k * 256 + pidx), p.frame, p.wfs_index, k0, k1);
          if (cube) cube[pidx] = v;
          s0 += v;
          sx = fmaf(v, (float)px, sx);
          sy = fmaf(v, (float)py, sy);
        }
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
      float* sl = p.slopes + (size_t)cur.e * p.lds;
      sl[cur.k] = (gx - p.cog_offset) * p.pixsize;
      sl[p.nvalid + cur.k] = (gy - p.cog_offset) * p.pixsize;
    }

    // ---- rotate: item it+1 becomes current ----
    s_v[(it & 1) * 32 + lane] = v_n2;          // volts of item it+2 into the buffer item `it` used
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        xr_h[j][c] = yr_h[j][c]; xr_l[j][c] = yr_l[j][c]; xi_h[j][c] = yi_h[j][c]; xi_l[j][c] = yi_l[j][c];
      }
    pm_n1 = pm_n2;
    cur = nx1; nx1 = nx2; nx2 = nx3;
  }
}
