"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

numpy restatement, one environment at a time, of the closed-loop AO frame that
the reference runs through the un-vendored COMPASS ``sutraWrap`` objects.  Each
function cites the reference call site it stands for (paths relative to
/root/reference) and, where the arithmetic lives inside sutra and is therefore
NOT visible in the reference ("parity unpinned"), states the convention chosen.
The conventions are pinned only by the self-consistency checks in
tests/test_oracle_kats.py (flat wavefront -> zero slopes, tilt -> slope, poke ->
stamp, delay impulse) and by the integer KAT "88 / 1284 actuators survive
correct_dm" obtained by running the reference's own correct_dm on top of
``slopes_geom`` below (oracle/refharness/gen_golden.py).

All arrays are float32 and all pupil-plane arrays are C-ordered ``[y, x]`` with
x the fast axis, which is the flat index ``x + n*y`` used by the reference's
integer maps (phasemap geom_init.py:673-685, influpos dm_init.py:750-814).
"""
import numpy as np

from . import rng

F32 = np.float32
ARCSEC_PER_RAD = 206264.80624709636


# --------------------------------------------------------------------------------------
# (a-1/a-2) turbulence: iterkolmo extrusion  -- shesha/util/iterkolmo.py:255-288, 442-461
# --------------------------------------------------------------------------------------
def extrude_column(screen, A, B, istencil, amp, noise):
    """One +x extrusion (iterkolmo.py:255-288): new right-hand column, screen shifts left.

    screen [N,N] float32, A [N,S], B [N,N], istencil [S] flat indices (x + N*y),
    amp = r0_layer_px**(-5/6) * 0.5/(2 pi)  [sutra: screens held in microns], noise [N] ~ N(0,1).
    float32 accumulate in float64 then round: the GPU GEMM is compared with rel 1e-4.
    """
    n = screen.shape[0]
    z = screen.reshape(-1)[istencil].astype(np.float64)
    zref = np.float64(screen[0, n - 1])
    z -= zref
    col = A.astype(np.float64) @ z + B.astype(np.float64) @ (noise.astype(np.float64) * np.float64(amp)) + zref
    out = np.empty_like(screen)
    out[:, :n - 1] = screen[:, 1:]
    out[:, n - 1] = col.astype(np.float32)
    return out


def extrude(screen, A, B, istx, amp, noise, axis, sign):
    """Extrusion along x (axis=0) or y (axis=1), towards +/- (sign).

    y is the x case on the transposed screen (isty is istx of the transposed stencil,
    iterkolmo.py:241-244); a negative direction is the positive case on the screen rotated by
    180 degrees (mirrored indices n*n-1-i, iterkolmo.py:246-249) [sutra scheduling, unpinned].
    `istx` is always the un-mirrored +x stencil.
    """
    s = screen
    if axis == 1:
        s = s.T
    if sign < 0:
        s = s[::-1, ::-1]
    s = extrude_column(np.ascontiguousarray(s), A, B, istx, amp, noise)
    if sign < 0:
        s = s[::-1, ::-1]
    if axis == 1:
        s = s.T
    return np.ascontiguousarray(s)


class OracleAtmos:
    """Per-environment turbulence state (sutra Atmos; atmos_init.py:76-132, atmosCompass.py:137-161)."""

    def __init__(self, tab, seed):
        self.tab = tab
        self.nl = tab["nscreens"]
        self.seed = int(seed)
        self.screens = [np.zeros((int(n), int(n)), F32) for n in tab["dim_screens"]]
        self.accx = np.zeros(self.nl, np.float64)
        self.accy = np.zeros(self.nl, np.float64)
        self.next_ext = np.zeros(self.nl, np.int64)

    def amp(self, l):
        return F32(F32(self.tab["r0_layers"][l]) ** F32(-5.0 / 6.0) * F32(0.5 / (2 * np.pi)))

    def _one(self, l, axis, sign):
        n = self.screens[l].shape[0]
        noise = rng.atmos_noise(self.seed, l, int(self.next_ext[l]), n)
        self.next_ext[l] += 1
        self.screens[l] = extrude(self.screens[l], self.tab["A"][l], self.tab["B"][l],
                                  self.tab["istx_pos"][l], self.amp(l), noise, axis, sign)

    def reset(self, seed):
        """reset_turbu (atmosCompass.py:137-145): zero the screen, 2N extrusions (iterkolmo.py:442-461)
        in the layer's own wind direction along x [sutra refresh_screen, unpinned]."""
        self.seed = int(seed)
        for l in range(self.nl):
            n = self.screens[l].shape[0]
            self.screens[l][:] = 0
            self.accx[l] = self.accy[l] = 0.0
            self.next_ext[l] = 0
            sx = -1 if self.tab["deltax"][l] < 0 else 1
            for _ in range(2 * n):
                self._one(l, 0, sx)

    def move(self):
        """move_atmos (atmosCompass.py:158-161): accumulate wind, extrude whole pixels, x then y."""
        for l in range(self.nl):
            self.accx[l] += float(self.tab["deltax"][l])
            self.accy[l] += float(self.tab["deltay"][l])
            nx = int(self.accx[l])
            ny = int(self.accy[l])
            self.accx[l] -= nx
            self.accy[l] -= ny
            for _ in range(abs(nx)):
                self._one(l, 0, 1 if nx > 0 else -1)
            for _ in range(abs(ny)):
                self._one(l, 1, 1 if ny > 0 else -1)


# --------------------------------------------------------------------------------------
# (a-3) raytrace  -- sourceCompass.py:54-85; offsets wfs_init.py:175-204, target_init.py:100-141
# --------------------------------------------------------------------------------------
def raytrace_layer(screen, n, xoff, yoff):
    """Bilinear sample of screen at (x + xoff, y + yoff), x,y in [0,n)  [sutra interpolation]."""
    ix = int(np.floor(xoff))
    iy = int(np.floor(yoff))
    fx = F32(xoff - ix)
    fy = F32(yoff - iy)
    s = screen
    a = s[iy:iy + n, ix:ix + n]
    b = s[iy:iy + n, ix + 1:ix + 1 + n]
    c = s[iy + 1:iy + 1 + n, ix:ix + n]
    d = s[iy + 1:iy + 1 + n, ix + 1:ix + 1 + n]
    w00 = (F32(1) - fx) * (F32(1) - fy)
    w01 = fx * (F32(1) - fy)
    w10 = (F32(1) - fx) * fy
    w11 = fx * fy
    return (w00 * a + w01 * b + w10 * c + w11 * d).astype(F32)


def raytrace_atmos(atm, n, xoffs, yoffs):
    ph = np.zeros((n, n), F32)
    for l in range(atm.nl):
        ph += raytrace_layer(atm.screens[l], n, xoffs[l] + atm.accx[l], yoffs[l] + atm.accy[l])
    return ph


# --------------------------------------------------------------------------------------
# (a-4) deformable mirrors -- dm_init.py:330-509 (pzt), 661-694 (tt), 750-814 (gather tables)
# --------------------------------------------------------------------------------------
def pzt_shape(volts, influ, i1, j1, dim):
    """shape[y,x] = sum_k volts[k] * influ[x-i1[k], y-j1[k], k]; pixels outside [0,dim) dropped
    exactly as comp_dmgeom drops them (dm_init.py:776-780)."""
    ss = influ.shape[0]
    shape = np.zeros((dim, dim), F32)
    for k in np.nonzero(volts)[0]:
        x0, y0 = int(i1[k]), int(j1[k])
        xa, xb = max(x0, 0), min(x0 + ss, dim)
        ya, yb = max(y0, 0), min(y0 + ss, dim)
        if xa >= xb or ya >= yb:
            continue
        st = influ[xa - x0:xb - x0, ya - y0:yb - y0, k].T
        shape[ya:yb, xa:xb] += F32(volts[k]) * st
    return shape


def tt_shape(volts, influ_tt):
    return (F32(volts[0]) * influ_tt[:, :, 0].T + F32(volts[1]) * influ_tt[:, :, 1].T).astype(F32)


def crop(shape, n, off):
    assert float(off).is_integer(), "fractional DM offset not supported (on-axis, alt 0 only)"
    off = int(off)
    return shape[off:off + n, off:off + n]


# --------------------------------------------------------------------------------------
# geometric slopes used only to filter actuators -- imats.py:54-112 -> sutra slopes_geom(0)
# --------------------------------------------------------------------------------------
def slopes_geom(phase, mpupil, origins, pdiam, flux, subapd):
    """Mean phase gradient per subaperture in arcsec [sutra, unpinned]:
    central differences inside the tile (one-sided at the tile edge), masked by the pupil,
    averaged over pdiam^2 and divided by the illuminated fraction; micron/m -> arcsec."""
    nv = len(origins)
    out = np.zeros(2 * nv, F32)
    alpha = 0.206265 / subapd
    for k, (r0, c0) in enumerate(origins):
        t = phase[r0:r0 + pdiam, c0:c0 + pdiam].astype(np.float64)
        m = mpupil[r0:r0 + pdiam, c0:c0 + pdiam]
        gx = np.empty_like(t)
        gx[:, 1:-1] = t[:, 2:] - t[:, :-2]
        gx[:, 0] = t[:, 1] - t[:, 0]
        gx[:, -1] = t[:, -1] - t[:, -2]
        gy = np.empty_like(t)
        gy[1:-1, :] = t[2:, :] - t[:-2, :]
        gy[0, :] = t[1, :] - t[0, :]
        gy[-1, :] = t[-1, :] - t[-2, :]
        out[k] = (gx * m).sum() / pdiam / 2.0 * alpha / flux[k]
        out[nv + k] = (gy * m).sum() / pdiam / 2.0 * alpha / flux[k]
    return out


# --------------------------------------------------------------------------------------
# (a-5) Shack-Hartmann image -- tables geom_init.py:622-811, call wfsCompass.py:334-343
# --------------------------------------------------------------------------------------
def sh_bincube(phase, mpupil, w, *, return_hr=False):
    """Noise-free binned spots, shape [nvalid, npix, npix] (row = y, col = x of the detector).

    amplitude = pupil * exp(i (2 pi / lambda * phase - halfxy)) in the [0:pdiam,0:pdiam] corner of an
    Nfft^2 array, I = |FFT2|^2 (numpy forward sign), bins summed through `binmap` on the flat
    FFT-native index (geom_init.py:731-758), scaled so each subaperture holds
    nphotons * fluxPerSub photons  [sutra normalisation, unpinned].
    """
    nv, pd, nfft, npix = w["nvalid"], w["pdiam"], w["Nfft"], w["npix"]
    assert w["Ntot"] == nfft, "hrmap path (Ntot != Nfft) is outside the hot-path scope"
    k2 = F32(2.0 * np.pi / w["Lambda"])
    amp = np.zeros((nv, nfft, nfft), np.complex64)
    pm = w["phasemap"]  # [pd*pd, nv]
    ph = phase.reshape(-1)[pm].T.reshape(nv, pd, pd)
    mk = mpupil.reshape(-1)[pm].T.reshape(nv, pd, pd)
    arg = (k2 * ph - w["halfxy"][None]).astype(F32)
    amp[:, :pd, :pd] = mk * (np.cos(arg) + 1j * np.sin(arg)).astype(np.complex64)
    hr = np.abs(np.fft.fft2(amp)) ** 2
    hr = hr.astype(F32).reshape(nv, nfft * nfft)
    binned = hr[:, w["binmap"]].sum(axis=1)  # [nv, npix*npix]
    tot = binned.sum(axis=1, keepdims=True)
    scale = (F32(w["nphotons"]) * w["fluxPerSub"].astype(F32))[:, None] / tot
    cube = (binned * scale).astype(F32).reshape(nv, npix, npix)
    if return_hr:
        return cube, hr.reshape(nv, nfft, nfft)
    return cube


def sh_noise(cube, noise, seed, wfs_index, frame):
    """Photon + read noise (wfsCompass.py:297-310): noise<0 none, ==0 Poisson, >0 Poisson+N(0,noise)."""
    return rng.wfs_pixel_noise(seed, wfs_index, frame, cube.reshape(-1), noise).reshape(cube.shape)


def bincube_to_binimg(cube, w):
    """Detector mosaic (rlSupervisor.py:857-874 writes cube[k] at [validsubsx[k], validsubsy[k]])."""
    npix = w["npix"]
    side = w["nxsub"] * npix
    img = np.zeros((side, side), F32)
    for k in range(w["nvalid"]):
        x0, y0 = int(w["validsubsx"][k]), int(w["validsubsy"][k])
        img[y0:y0 + npix, x0:x0 + npix] = cube[k]
    return img


# --------------------------------------------------------------------------------------
# (a-7) centre of gravity -- constants rtc_init.py:207-232
# --------------------------------------------------------------------------------------
def cog(cube, w):
    npix = w["npix"]
    offset = F32(npix // 2 - 0.5)
    scale = F32(w["pixsize"])
    c = cube.astype(np.float64)
    tot = c.sum(axis=(1, 2))
    xs = np.arange(npix, dtype=np.float64)
    sx = (c.sum(axis=1) * xs).sum(axis=1)
    sy = (c.sum(axis=2) * xs).sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        gx = np.where(tot > 0, sx / tot, offset)
        gy = np.where(tot > 0, sy / tot, offset)
    return np.concatenate([(gx - offset) * scale, (gy - offset) * scale]).astype(F32)


# ---------------------------------------------------------------------------------------------------
# (f-1) target image and Strehl -- TargetCompass.comp_tar_image / comp_strehl, targetCompass.py:139-196; the
# arithmetic is sutra's (un-vendored): zero-padded FFT of pupil * exp(i k phi) on the Nfft grid of p_geom._ipupil,
# |.|^2, maximum pixel, optional sub-pixel fit, normalised by the same figure of a flat wavefront.  [S] the fit is
# restated as a three-point Gaussian (parabola through the logarithms) per axis around the brightest pixel.
def psf_image(phase, pupil, lam, nfft):
    """|FFT|^2 of the pupil field, in units of the diffraction-limited peak, zero frequency at [0, 0]."""
    pup = np.asarray(pupil) > 0
    f = pup * np.exp(1j * (2 * np.pi / lam) * np.asarray(phase, np.float64))
    return np.abs(np.fft.fft2(f, s=(nfft, nfft))) ** 2 / float(pup.sum()) ** 2


def fit_peak(img):
    """Brightest pixel refined by a three-point Gaussian fit along each axis (periodic neighbours)."""
    j, i = np.unravel_index(np.argmax(img), img.shape)
    pk = img[j, i]
    lg = np.log(pk)
    for m1, p1 in ((img[j - 1, i], img[(j + 1) % img.shape[0], i]), (img[j, i - 1], img[j, (i + 1) % img.shape[1]])):
        if m1 > 0 and p1 > 0:
            a = 0.5 * (np.log(m1) + np.log(p1)) - np.log(pk)
            b = 0.5 * (np.log(p1) - np.log(m1))
            if a < 0:
                lg += -b * b / (4 * a)
    return float(np.exp(lg))


def psf_peak(phase, pupil, lam, nfft, do_fit=True):
    img = psf_image(phase, pupil, lam, nfft)
    return fit_peak(img) if do_fit else float(img.max())
