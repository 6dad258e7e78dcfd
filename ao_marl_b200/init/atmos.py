"""Turbulence-layer operands (host side, numpy, once per configuration).

The autoregressive "infinite phase screen" of Assemat/Fried&Clark: a new column x of an
N x N von-Karman screen is drawn from the S stencil values z already on the screen,

        x = A (z - z_ref) + B eps + z_ref ,      eps ~ N(0, I)

with A = C_xz C_zz^+ and B B^T = C_xx - A C_zx.  Restates the operand construction of the
reference's shesha/util/iterkolmo.py (create_stencil 41-73, Czz/Cxz/Cxx 115-187, AB 190-252,
phase_struct/rodconan/asymp_macdo/macdo_x56 291-415) and the per-layer set-up of
shesha/init/atmos_init.py:76-132.  Matrices are returned row-major [N, S] / [N, N] (the CUDA GEMM
reads them K-major), stencils as flat indices x + N*y for the +x direction; the kernels derive the
-x / +-y variants by index arithmetic (mirror = N*N-1-i, transpose = swap x,y) instead of storing
four copies.
"""
import numpy as np

from .geom import ARCSEC2RAD, DEG2RAD


def stencil_mask(n):
    """0/1 mask [y, x] of the screen pixels feeding the extrusion (last column full, then columns
    at power-of-two distances sampled with power-of-two row strides, plus a few far points)."""
    ns = int(np.log2(n + 1) + 1)
    m = np.zeros((n, n))
    flat = m.reshape(-1)
    m[:, 0] = 1
    for i in range(2, ns):
        m[::2 ** (i - 1), 2 ** (i - 2)] = 1
        flat[2 ** (i - 1) - 1] = 1
    flat[2 ** (ns - 1) - 1] = 1
    for i in range(0, n, 2 ** (ns - 1)):
        flat[2 ** (ns - 2) + i * n] = 1
    flat[2 ** ns - 1] = 1
    for i in range(0, n, 2 ** ns):
        flat[2 ** (ns - 1) + i * n] = 1
    return m


def stencil_size(n):
    return int(stencil_mask(n).sum())


def stencil_indices(n):
    m = np.fliplr(np.roll(stencil_mask(n), n // 2, axis=0))
    return np.nonzero(m.reshape(-1))[0]


def _macdonald_series(x, k=10):
    """x^(5/6) K_{5/6}(x) power series (first term dropped: it cancels in the structure function)."""
    a = 5.0 / 6.0
    fact = 1.0
    x2a = x ** (2.0 * a)
    x22 = x * x / 4.0
    x2n = 0.5
    Ga = 2.01126983599717856777
    Gma = -3.74878707653729348337
    s = np.zeros(x.shape)
    for n in range(k + 1):
        term = Gma * x2a
        if n:
            term = term + Ga
        term = term * x2n / fact
        s = s - term if n % 2 else s + term
        if n < k:
            fact *= n + 1
            Gma /= -a - n - 1
            Ga /= a - n - 1
            x2n = x2n * x22
    return s


def _macdonald_asymptotic(x):
    k2 = 1.00563491799858928388289314170833
    k3 = 1.25331413731550012081
    a1, a2, a3 = 0.22222222222222222222, -0.08641975308641974829, 0.08001828989483310284
    xi = 1.0 / x
    return k2 - k3 * np.exp(-x) * x ** (1 / 3.0) * (1.0 + xi * (a1 + xi * (a2 + xi * a3)))


def structure_function(r2, L0):
    """von-Karman phase structure function D_phi(sqrt(r2)) for r0 = 1 pixel (rodconan)."""
    r = np.sqrt(r2)
    k1 = 0.1716613621245709486
    x = (2 * np.pi / L0) * r
    lim = 0.75 * 2 * np.pi
    res = np.zeros_like(r)
    big = x > lim
    if big.any():
        res[big] = _macdonald_asymptotic(x[big])
    if (~big).any():
        res[~big] = -_macdonald_series(x[~big])
    return k1 * L0 ** (5.0 / 3.0) * res


def extrusion_operands(n, L0):
    """A [n, S], B [n, n] (float64) and the +x stencil for an n x n screen, outer scale L0 pixels."""
    ist = stencil_indices(n)
    zx = (ist % n + 1).astype(np.float64)      # 1-based pixel coordinates, as the covariances use
    zy = (ist // n + 1).astype(np.float64)
    xx = np.full(n, n + 1.0)
    xy = np.arange(n) + 1.0
    refx, refy = float(n), 1.0                  # reference pixel: row 0, last column

    def D(dx2dy2):
        return structure_function(dx2dy2, L0)

    dz = D((refx - zx) ** 2 + (refy - zy) ** 2)
    dxr = D((refx - xx) ** 2 + (refy - xy) ** 2)
    # the association of the three terms follows iterkolmo.py:128-131, 154-159, 180-185 so that the
    # (badly conditioned) pseudo-inverse sees the same rounding as the reference
    zz = ((-D((zx[:, None] - zx[None, :]) ** 2 + (zy[:, None] - zy[None, :]) ** 2)
           + dz[None, :]) + dz[:, None]) * 0.5
    xz = (-D((xx[:, None] - zx[None, :]) ** 2 + (xy[:, None] - zy[None, :]) ** 2)
          + (dxr[:, None] + dz[None, :])) * 0.5
    cxx = (-D((xx[:, None] - xx[None, :]) ** 2 + (xy[:, None] - xy[None, :]) ** 2)
           + (dxr[None, :] + dxr[:, None])) * 0.5
    U, s, Vt = np.linalg.svd(zz)
    sinv = np.zeros_like(s)
    sinv[:-1] = 1.0 / s[:-1]                    # the last singular value (piston-like) is dropped
    zz_pinv = (U * sinv) @ Vt
    A = xz @ zz_pinv
    bbt = cxx - A @ xz.T
    U1, l1, _ = np.linalg.svd(bbt)
    B = U1 * np.sqrt(l1)
    return A, B, ist


def transposed_stencil(ist, n):
    """The +y stencil in the reference's ordering (iterkolmo.py:241-244)."""
    return (ist % n) * n + ist // n


def atmos_init(p_atmos, p_tel, p_geom, ittime, p_wfss=None, p_targets=None):
    """Screen sizes, wind steps in pixels per frame, per-layer r0 in pixels (atmos_init.py:76-116)."""
    p_atmos.alt = p_atmos.alt / np.cos(p_geom.zenithangle * DEG2RAD)
    p_atmos.pupixsize = p_tel.diam / p_geom.pupdiam
    norms = [0.0]
    if p_wfss is not None:
        norms += [(w.xpos ** 2 + w.ypos ** 2) ** 0.5 for w in p_wfss]
    if p_targets is not None:
        norms += [(t.xpos ** 2 + t.ypos ** 2) ** 0.5 for t in p_targets]
    max_size = max(norms)
    patch = (p_geom._n + 2 * (max_size * ARCSEC2RAD * p_atmos.alt) / p_atmos.pupixsize
             + 4).astype(np.int64)
    p_atmos.dim_screens = patch + patch % 2
    lin = p_geom.pupdiam / p_tel.diam * p_atmos.windspeed * np.cos(DEG2RAD * p_geom.zenithangle) * ittime
    # the reference's setters hold the wind steps as float32 (PATMOS.py set_deltax/set_deltay)
    p_atmos._deltax = (lin * np.sin(DEG2RAD * p_atmos.winddir + np.pi)).astype(np.float32)
    p_atmos._deltay = (lin * np.cos(DEG2RAD * p_atmos.winddir + np.pi)).astype(np.float32)
    p_atmos.frac = p_atmos.frac / np.sum(p_atmos.frac)
    if p_atmos.L0 is None:
        p_atmos.L0 = np.ones(p_atmos.nscreens, dtype=np.float32) * 1e5
    p_atmos._L0_pix = p_atmos.L0 * p_geom.pupdiam / p_tel.diam
    if p_atmos.seeds is None:
        p_atmos.seeds = np.arange(p_atmos.nscreens, dtype=np.int64) + 1234
    p_atmos._r0_layers = p_atmos.r0 / (p_atmos.frac ** (3.0 / 5.0) * p_atmos.pupixsize)
    p_atmos._stencil_size = np.array([stencil_size(int(n)) for n in p_atmos.dim_screens])


def layer_operands(p_atmos, cache=None):
    """Per-layer (A, B, istx) as float32 / int32; identical (n, L0) layers share one computation."""
    out = []
    memo = {} if cache is None else cache
    for l in range(p_atmos.nscreens):
        key = (int(p_atmos.dim_screens[l]), float(p_atmos._L0_pix[l]))
        if key not in memo:
            A, B, ist = extrusion_operands(*key)
            memo[key] = (np.ascontiguousarray(A, dtype=np.float32),
                         np.ascontiguousarray(B, dtype=np.float32), ist.astype(np.int32))
        out.append(memo[key])
    return out
