// Parameter blocks shared by every Shack-Hartmann frame kernel (wfs_kernels.cuh, wfs_mma.cuh, wfs_tma.cuh,
// wfs_umma.cuh): one turbulence layer as the raytrace sees it, and the per-frame sensor / mirror description.
#pragma once
#include <stdint.h>
#include "../../include/aomarl.h"

struct WfsLayer {
  const float* screen;   // [E][N][N]
  const int* ox; const int* oy;
  int N;
  int ix, iy;            // integer part of (offset + wind accumulator)
  float fx, fy;          // fractional part
  float w00, w01, w10, w11;   // bilinear weights (1-fx)(1-fy), fx(1-fy), (1-fx)fy, fx fy of the four taps
};

struct WfsParams {
  WfsLayer layer[AOM_MAX_LAYERS];
  int n_layers;          // 0 when the atmosphere is not traced
  int use_dm;
  int E, n, nvalid;
  const float* mpupil;   // [n][n]
  const float* halfxy;   // [16][16]
  const int* sub_x0; const int* sub_y0; const float* flux;
  // mirrors
  const float* volts; int ldv;         // [E][ldv]: pzt volts then 2 tip-tilt volts
  const float* stamp1d; int ss;
  const int* act_map; int grid_n, pitch, i1_0, j1_0, pzt_off, pzt_nact;
  const float* tt_planes; int tt_dim, tt_off;
  // sensor
  float k2;              // 2 pi / lambda
  float nphotons, noise;
  float cog_offset, pixsize;
  uint32_t frame, wfs_index;
  const uint32_t* k0; const uint32_t* k1;
  // outputs
  float* slopes; int lds;              // [E][lds]: x slopes then y slopes
  float* bincube;                      // [E][nvalid][256] or null
};

