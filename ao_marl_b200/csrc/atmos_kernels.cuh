// Turbulence screens: ring-buffered storage + batched extrusion.
//
// Reference maths: shesha/util/iterkolmo.py:255-288 (extrude), scheduling in sutra's Atmos.move_atmos
// behind shesha/supervisor/components/atmosCompass.py:158-161.  The reference shifts the whole
// N x N screen by one pixel per extrusion (read + write of 1.7 MB per layer); here every screen is a
// torus addressed through per-environment offsets (ox, oy):
//        logical (y, x)  ->  physical ((y + oy) mod N, (x + ox) mod N)
// so an extrusion only writes the N new pixels.  The four directions reuse the single +x stencil:
//        +x: (r, c)          -x: (N-1-r, N-1-c)        (iterkolmo.py:246-249 mirror)
//        +y: (c, r)          -y: (N-1-c, N-1-r)        (iterkolmo.py:241-244 transpose)
// One extrusion of all E environments = gather Z [E][S+N] -> GEMM with [A|B] -> scatter N pixels.
#pragma once
#include <cuda_runtime.h>
#include "rng.cuh"

struct ExtrudeParams {
  float* screen;          // [E][N][N]
  int* ox; int* oy;       // [E]
  uint32_t* count;        // [E] extrusions done since reset (RNG counter)
  const uint32_t* k0; const uint32_t* k1;   // [E] Philox key = env seed
  const int* stencil;     // [S]
  float* Z; int ldz;      // [E][ldz]
  float* zref;            // [E]
  float* newcol; int ldn; // [E][ldn]
  int N, S, E, layer, axis, sign;
  float amp;
};

__device__ __forceinline__ void extr_logical(int r, int c, int N, int axis, int sign, int& lr, int& lc) {
  if (axis == 0) { lr = r; lc = c; } else { lr = c; lc = r; }
  if (sign < 0) {
    if (axis == 0) { lr = N - 1 - r; lc = N - 1 - c; } else { lr = N - 1 - c; lc = N - 1 - r; }
  }
}

__device__ __forceinline__ int wrapN(int v, int N) { return v >= N ? v - N : v; }

// grid: (ceil(ldz/256), E)
__global__ void extrude_gather_kernel(ExtrudeParams p) {
  const int e = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.ldz) return;
  const int N = p.N;
  const float* scr = p.screen + (size_t)e * N * N;
  const int ox = p.ox[e], oy = p.oy[e];
  // reference pixel: p'[0, N-1] of the rotated/transposed screen
  int rr, rc;
  extr_logical(0, N - 1, N, p.axis, p.sign, rr, rc);
  const float zr = scr[(size_t)wrapN(rr + oy, N) * N + wrapN(rc + ox, N)];
  if (k == 0) p.zref[e] = zr;
  float v = 0.f;
  if (k < p.S) {
    int idx = p.stencil[k];
    int lr, lc;
    extr_logical(idx / N, idx % N, N, p.axis, p.sign, lr, lc);
    v = scr[(size_t)wrapN(lr + oy, N) * N + wrapN(lc + ox, N)] - zr;
  } else if (k < p.S + N) {
    int j = k - p.S;
    aom_u4 w = aom_philox((uint32_t)(j >> 2), p.count[e], AOM_TAG_ATMOS, (uint32_t)p.layer, p.k0[e], p.k1[e]);
    v = aom_mul(aom_normal_of_block(w, j & 3), p.amp);
  }
  p.Z[(size_t)e * p.ldz + k] = v;
}

// grid: E blocks of 256 threads
__global__ void extrude_scatter_kernel(ExtrudeParams p) {
  const int e = blockIdx.x;
  const int N = p.N;
  float* scr = p.screen + (size_t)e * N * N;
  const int ox = p.ox[e], oy = p.oy[e];
  const float zr = p.zref[e];
  __syncthreads();
  int nox = ox, noy = oy;
  if (p.axis == 0) nox = (p.sign > 0) ? wrapN(ox + 1, N) : (ox == 0 ? N - 1 : ox - 1);
  else             noy = (p.sign > 0) ? wrapN(oy + 1, N) : (oy == 0 ? N - 1 : oy - 1);
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    float v = p.newcol[(size_t)e * p.ldn + j] + zr;
    size_t addr;
    if (p.axis == 0) {
      int pc = (p.sign > 0) ? ox : nox;                       // new logical column N-1 (or 0)
      int lr = (p.sign > 0) ? j : N - 1 - j;
      addr = (size_t)wrapN(lr + oy, N) * N + pc;
    } else {
      int pr = (p.sign > 0) ? oy : noy;                       // new logical row N-1 (or 0)
      int lc = (p.sign > 0) ? j : N - 1 - j;
      addr = (size_t)pr * N + wrapN(lc + ox, N);
    }
    scr[addr] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    p.ox[e] = nox;
    p.oy[e] = noy;
    p.count[e] += 1;
  }
}

// grid-stride fill helpers
__global__ void fill_i32_kernel(int* p, int v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
