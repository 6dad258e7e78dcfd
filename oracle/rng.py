"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Counter-based RNG shared, bit for bit, by the numpy oracle and the CUDA kernels
(ao_marl_b200/csrc/rng.cuh restates exactly the same arithmetic).

The reference draws its noise inside the un-vendored COMPASS simulator
(cuRAND behind ``Atmos.set_seed`` / ``Sensors.set_noise``; call sites
shesha/supervisor/components/atmosCompass.py:137-145 and wfsCompass.py:297-310,
345-350), so the actual streams are unknowable here ("parity unpinned" for the
random numbers themselves).  What *is* pinned is this file: Philox4x32-10
(checked against the Random123 known-answer vectors in tests/) and a set of
float32 transforms written with one IEEE operation per step -- no fused
multiply-add, no libm -- so that the GPU (``__fmul_rn``/``__fadd_rn``) and numpy
produce identical bits, and Poisson photon counts match as integers.

Stream layout (key = 64-bit environment seed, counter = 4 x uint32):
    ctr = (i, n, tag, sub)
    tag 1: turbulence innovation   i = lane // 4, n = extrusion number, sub = layer
    tag 2: WFS pixel noise         i = subap * npix^2 + pixel, n = frame, sub = wfs
    tag 4: actor exploration noise i = element // 4, n = step, sub = agent
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)

TAG_ATMOS = 1
TAG_WFS = 2
TAG_ACTOR = 4

POISSON_SWITCH = np.float32(30.0)
POISSON_MAXK = 200


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10.  All arguments broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(
        *[np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3, k0, k1)])
    c0, c1, c2, c3 = c0.copy(), c1.copy(), c2.copy(), c3.copy()
    k0, k1 = k0.copy(), k1.copy()
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            if r < 9:
                k0 = k0 + W0
                k1 = k1 + W1
    return c0, c1, c2, c3


def seed_key(seed):
    seed = np.asarray(seed, dtype=np.int64).astype(np.uint64)
    return (seed & MASK32).astype(np.uint32), (seed >> np.uint64(32)).astype(np.uint32)


def u01(x):
    """uint32 -> float32 in (0,1): ((x >> 8) + 0.5) * 2^-24  (exact in float32)."""
    return ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0**-24)


def u01_open(x):
    """uint32 -> float32 strictly inside (0,1): ((x >> 9) + 0.5) * 2^-23 (exact in float32).  Used by the Poisson
    CDF inversion: u01's (x >> 8) + 0.5 rounds to 2^24 for the largest words (u == 1.0), which no float32 CDF reaches."""
    return ((x >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0**-23)


F32 = np.float32
_LN2 = F32(0.6931471805599453)
_SQRT2 = F32(1.41421354)


def det_log(x):
    """float32 natural log of positive normal x, one rounding per operation."""
    x = np.asarray(x, dtype=np.float32)
    bits = x.view(np.int32)
    e = ((bits >> 23) & 0xFF) - 127
    m = ((bits & 0x007FFFFF) | 0x3F800000).astype(np.int32).view(np.float32)
    big = m > _SQRT2
    m = np.where(big, m * F32(0.5), m).astype(np.float32)
    e = np.where(big, e + 1, e)
    f = m - F32(1.0)
    s = f / (F32(2.0) + f)
    z = s * s
    p = F32(2.0 / 9.0)
    p = p * z + F32(2.0 / 7.0)
    p = p * z + F32(2.0 / 5.0)
    p = p * z + F32(2.0 / 3.0)
    logm = F32(2.0) * s + (s * z) * p
    return (e.astype(np.float32) * _LN2 + logm).astype(np.float32)


_HALFPI = F32(1.5707963267948966)


def det_sincos2pi(u):
    """(cos(2 pi u), sin(2 pi u)) for u in [0,1), float32, one rounding per operation."""
    u = np.asarray(u, dtype=np.float32)
    t = u * F32(4.0)
    q = np.floor(t + F32(0.5))
    r = t - q
    a = r * _HALFPI
    a2 = a * a
    ps = F32(1.0 / 362880.0)
    ps = ps * a2 + F32(-1.0 / 5040.0)
    ps = ps * a2 + F32(1.0 / 120.0)
    ps = ps * a2 + F32(-1.0 / 6.0)
    s = a + (a * a2) * ps
    pc = F32(-1.0 / 3628800.0)
    pc = pc * a2 + F32(1.0 / 40320.0)
    pc = pc * a2 + F32(-1.0 / 720.0)
    pc = pc * a2 + F32(1.0 / 24.0)
    pc = pc * a2 + F32(-0.5)
    c = F32(1.0) + a2 * pc
    qi = q.astype(np.int32) & 3
    cos = np.where(qi == 0, c, np.where(qi == 1, -s, np.where(qi == 2, -c, s)))
    sin = np.where(qi == 0, s, np.where(qi == 1, c, np.where(qi == 2, -s, -c)))
    return cos.astype(np.float32), sin.astype(np.float32)


_LOG2E = F32(1.4426950408889634)
_LN2_HI = F32(0.693145751953125)       # 0x3f317200, 15 significant bits
_LN2_LO = F32(1.42860682030941723212e-6)


def det_exp(x):
    """float32 exp(x) for -80 <= x <= 0, one rounding per operation."""
    x = np.asarray(x, dtype=np.float32)
    n = np.rint(x * _LOG2E).astype(np.float32)
    r = x - n * _LN2_HI
    r = r - n * _LN2_LO
    p = F32(1.0 / 720.0)
    p = p * r + F32(1.0 / 120.0)
    p = p * r + F32(1.0 / 24.0)
    p = p * r + F32(1.0 / 6.0)
    p = p * r + F32(0.5)
    p = p * r + F32(1.0)
    p = p * r + F32(1.0)
    scale = ((n.astype(np.int32) + 127) << 23).astype(np.int32).view(np.float32)
    return (p * scale).astype(np.float32)


def normal_pair(x0, x1):
    """Box-Muller on two uint32 words -> two float32 standard normals."""
    u1 = u01(x0)
    u2 = u01(x1)
    r = np.sqrt(F32(-2.0) * det_log(u1)).astype(np.float32)
    c, s = det_sincos2pi(u2)
    return (r * c).astype(np.float32), (r * s).astype(np.float32)


def normals4(c0, c1, c2, c3, k0, k1):
    x0, x1, x2, x3 = philox4x32(c0, c1, c2, c3, k0, k1)
    z0, z1 = normal_pair(x0, x1)
    z2, z3 = normal_pair(x2, x3)
    return z0, z1, z2, z3


def atmos_noise(seed, layer, extrusion_index, n):
    """n standard normals for extrusion `extrusion_index` of `layer` in the env seeded `seed`."""
    k0, k1 = seed_key(seed)
    nb = (n + 3) // 4
    blk = np.arange(nb, dtype=np.uint32)
    z = normals4(blk, np.uint32(extrusion_index), np.uint32(TAG_ATMOS), np.uint32(layer), k0, k1)
    return np.stack(z, axis=1).reshape(-1)[:n].astype(np.float32)


def actor_noise(seed, agent, step, n):
    k0, k1 = seed_key(seed)
    nb = (n + 3) // 4
    blk = np.arange(nb, dtype=np.uint32)
    z = normals4(blk, np.uint32(step), np.uint32(TAG_ACTOR), np.uint32(agent), k0, k1)
    return np.stack(z, axis=1).reshape(-1)[:n].astype(np.float32)


def poisson_from_words(lam, x0, x1):
    """Integer Poisson sample for each float32 rate `lam`, driven by Philox words x0, x1.

    lam < 30 : CDF inversion with u = u01_open(x0), p0 = det_exp(-lam), p_k = (p_{k-1} * lam) / k; the search ends
               where the float32 CDF stops growing (k > lam), so the largest words give a tail sample, not the cap
    lam >= 30: floor(lam + sqrt(lam) * z + 0.5), z = first Box-Muller normal of (x0, x1), clamped at 0
    """
    lam = np.asarray(lam, dtype=np.float32)
    out = np.zeros(lam.shape, dtype=np.int32)
    small = (lam < POISSON_SWITCH) & (lam > 0)
    if small.any():
        ls = lam[small]
        us = u01_open(np.asarray(x0, dtype=np.uint32)[small])
        p = det_exp(-ls)
        F = p.copy()
        k = np.zeros(ls.shape, dtype=np.int32)
        active = us > F
        it = 0
        while active.any() and it < POISSON_MAXK:
            it += 1
            kf = F32(it)
            pn = ((p * ls) / kf).astype(np.float32)
            Fn = (F + pn).astype(np.float32)
            k = np.where(active, it, k)
            # the float32 CDF stopped growing beyond the mode: this k is the tail sample (never the loop cap)
            stalled = active & (Fn == F) & (kf > ls)
            p = np.where(active, pn, p)
            F = np.where(active & ~stalled, Fn, F)
            active = active & ~stalled & (us > F)
        out[small] = k
    big = lam >= POISSON_SWITCH
    if big.any():
        z0, _ = normal_pair(x0[big], x1[big])
        lb = lam[big]
        v = lb + np.sqrt(lb).astype(np.float32) * z0
        v = np.floor(v.astype(np.float32) + F32(0.5))
        out[big] = np.maximum(v, 0).astype(np.int32)
    return out


def wfs_pixel_noise(seed, wfs, frame, lam, noise):
    """Photon + read noise on a flat float32 intensity array `lam` (subap-major, pixel-minor).

    noise < 0: returns lam unchanged; noise == 0: Poisson only; noise > 0: Poisson + N(0, noise).
    """
    lam = np.asarray(lam, dtype=np.float32)
    if noise < 0:
        return lam.copy()
    k0, k1 = seed_key(seed)
    idx = np.arange(lam.size, dtype=np.uint32)
    x0, x1, x2, x3 = philox4x32(idx, np.uint32(frame), np.uint32(TAG_WFS), np.uint32(wfs), k0, k1)
    cnt = poisson_from_words(lam.reshape(-1), x0, x1).astype(np.float32)
    if noise > 0:
        zr, _ = normal_pair(x2, x3)
        cnt = (cnt + F32(noise) * zr).astype(np.float32)
    return cnt.reshape(lam.shape)
