"""Deformable-mirror tables (host side, numpy, once per configuration).

Piezo-stack mirror on a square actuator grid with the Rigaut influence function, and a tip-tilt
mirror; restates shesha/init/dm_init.py (_dm_init 96-202, make_pzt_dm 330-509, make_tiptilt_dm
661-694, comp_dmgeom 750-814, correct_dm 817-889) with shesha/util/dm_util.py (dim_dm_support
45-63, dim_dm_patch 66-99, createSquarePattern 102-121, select_actuators 245-297, make_zernike
300-380) and shesha/util/influ_util.py:139-183 (makeRigaut).

Convention (see oracle/aoframe.py): mirror surfaces are C-ordered [y, x]; actuator k adds
``volt * influ[x - i1[k], y - j1[k], k]``.  For the square pattern every stamp is the same array
(positions are half-integers, so the sub-pixel offset is 0.5 for all of them); `stamp_is_shared`
verifies that and the CUDA kernels then carry a single stamp.
"""
import math

import numpy as np

from .geom import ARCSEC2RAD, radial_distance


def dm_support(cent, extent, ssize):
    n1 = max(np.floor(cent - extent / 2), 1)
    n2 = min(np.ceil(cent + extent / 2), ssize)
    return int(n1), int(n2)


def dm_patch_diameter(pupdiam, diam, alt, xpos_wfs, ypos_wfs):
    norms = [np.hypot(x, y) for x, y in zip(xpos_wfs, ypos_wfs)] or [0.0]
    return int(pupdiam + 2 * np.max(norms) * ARCSEC2RAD * np.abs(alt) / (diam / pupdiam))


def rigaut_params(coupling):
    irc = 1.16136 + 2.97422 * coupling + (-13.2381) * coupling ** 2 + 20.4395 * coupling ** 3
    p1 = 4.49469 + (7.25509 + (-32.1948 + 17.9493 * coupling) * coupling) * coupling
    p2 = 2.49456 + (-0.65952 + (8.78886 - 6.23701 * coupling) * coupling) * coupling
    tc = 1.0 / np.abs(irc)
    ccc = (coupling - 1.0 + tc ** p1) / (np.log(tc) * tc ** p2)
    return irc, p1, p2, ccc


def rigaut_size(pitch, coupling):
    irc = rigaut_params(coupling)[0]
    size = int(np.ceil(2 * irc * pitch + 10))
    return size + size % 2


def rigaut(pitch, coupling, x, y):
    """Separable Rigaut influence function at local coordinates x, y (pixels, float32 arrays)."""
    irc, p1, p2, ccc = rigaut_params(coupling)
    u = np.clip(np.abs(x) / (irc * pitch), 1e-8, 2.0)
    v = np.clip(np.abs(y) / (irc * pitch), 1e-8, 2.0)
    f = (1.0 - u ** p1 + ccc * np.log(u) * u ** p2) * (1.0 - v ** p1 + ccc * np.log(v) * v ** p2)
    return f * (u <= 1.0) * (v <= 1.0)


def square_pattern(pitch, nxact):
    ax = (np.arange(nxact) - (nxact - 1.0) / 2.0).astype(np.float32)
    gx, gy = np.meshgrid(ax, ax)                     # gx varies along the fast axis
    return np.float32(np.array([gx.reshape(-1), gy.reshape(-1)]) * pitch)


def select_actuators(xc, yc, nxact, pitch, cobs, margin_in, margin_out, N=None):
    dis = np.sqrt(xc ** 2 + yc ** 2)
    rad_in = (((nxact - 1) / 2) * cobs - margin_in) * pitch
    if N is None:
        if margin_out is None:
            margin_out = 1.44
        rad_out = ((nxact - 1.0) / 2.0 + margin_out) * pitch
        return np.where((dis <= rad_out) * (dis >= rad_in))[0]
    valid = np.where(dis >= rad_in)[0]
    order = np.argsort(dis[valid])
    if N > valid.size:
        return valid
    return np.sort(order[:N])


def gather_tables(p_dm, p_geom):
    """Per-pixel gather lists (influpos / ninflu / influstart) -- comp_dmgeom, dm_init.py:750-814.
    Kept for parity with the reference's integer tables; the CUDA path evaluates the same sum
    from the shared stamp without them."""
    ss = p_dm._influsize
    nact = p_dm._ntotact
    dm_dim = int(p_dm._n2 - p_dm._n1 + 1)
    mp = p_geom._mpupil.shape[0]
    if dm_dim < mp:
        offs = (mp - dm_dim) // 2
    else:
        offs = 0
        mp = dm_dim
    t = np.arange(ss, dtype=np.int32)
    px = t[None, None, :] + (offs + p_dm._i1)[:, None, None] + np.zeros((1, ss, 1), np.int32)
    py = t[None, :, None] + (offs + p_dm._j1)[:, None, None] + np.zeros((1, 1, ss), np.int32)
    flat = px + mp * py
    sentinel = mp * mp + 10
    bad = (px < 0) | (py < 0) | (px > dm_dim - 1) | (py > dm_dim - 1)
    flat = np.where(bad, sentinel, flat).reshape(-1)
    order = np.argsort(flat, kind="quicksort").astype(np.int32)
    sorted_flat = flat[order].astype(np.int32)
    npts = np.zeros(mp * mp, dtype=np.int32)
    uniq, cnt = np.unique(sorted_flat, return_counts=True)
    if (uniq > npts.size - 1).any():
        uniq, cnt = uniq[:-1], cnt[:-1]
    npts[uniq] = cnt
    istart = np.zeros(mp * mp, dtype=np.int32)
    istart[1:] = np.cumsum(npts[:-1])
    p_dm._influpos = order[:int(npts.sum())].astype(np.int32)
    p_dm._ninflu = npts
    p_dm._influstart = istart
    p_dm._i1 = p_dm._i1 + offs
    p_dm._j1 = p_dm._j1 + offs


def make_pzt_dm(p_dm, p_geom, cobs):
    if p_dm.influ_type != "default":
        raise NotImplementedError("only the default (Rigaut) influence function is on the hot path")
    if p_dm.type_pattern not in (None, "square"):
        raise NotImplementedError("only the square actuator pattern is on the hot path")
    pitch = p_dm._pitch
    ss = rigaut_size(pitch, p_dm.coupling)
    p_dm._influsize = ss
    p_dm.type_pattern = "square"
    cub = square_pattern(pitch, p_dm.nact + 4)
    if p_dm.alt > 0:
        cobs = 0
    keep = select_actuators(cub[0], cub[1], p_dm.nact, pitch, cobs, p_dm.margin_in,
                            p_dm.margin_out, p_dm._ntotact)
    p_dm._ntotact = keep.size
    cub = cub + np.float32(p_geom.cent)
    pos = cub[:, keep]
    xpos, ypos = pos[0], pos[1]
    i1 = (xpos - ss / 2 - 0.5 - p_dm._n1).astype(np.int32)
    j1 = (ypos - ss / 2 - 0.5 - p_dm._n1).astype(np.int32)
    p_dm._xpos, p_dm._ypos, p_dm._i1, p_dm._j1 = xpos, ypos, i1, j1
    t = np.arange(ss, dtype=np.float32)
    # local coordinates of the stamp pixels relative to each actuator (float32 as in the reference)
    lx = (i1[:, None].astype(np.float32) + t[None, :] + np.float32(p_dm._n1)) - xpos[:, None]
    ly = (j1[:, None].astype(np.float32) + t[None, :] + np.float32(p_dm._n1)) - ypos[:, None]
    influ = np.empty((ss, ss, keep.size), dtype=np.float32)
    for k in range(keep.size):
        influ[:, :, k] = rigaut(pitch, p_dm.coupling, lx[k][:, None], ly[k][None, :])
    influ = influ * float(p_dm.unitpervolt / np.max(influ))
    p_dm._influ = influ
    gather_tables(p_dm, p_geom)


def zernike_numbers(zn):
    """(radial degree, azimuthal order) of Noll index zn (dm_util.py:383-410)."""
    j = 0
    for n in range(101):
        for m in range(n + 1):
            if (n - m) % 2 == 0:
                j += 1
                if j == zn:
                    return n, m
                if m != 0:
                    j += 1
                    if j == zn:
                        return n, m
    raise ValueError(zn)


def zernike_cube(nzer, size, diameter, xc, yc, ext):
    radius = (diameter + 1.0) / 2.0
    zr = radial_distance(size, xc, yc).astype(np.float32).T / radius
    mask = (zr <= 1).astype(np.float32)
    maskmod = (zr <= 1.2).astype(np.float32)
    zrmod = zr * maskmod
    zr = zr * mask
    x = np.tile(np.linspace(1, size, size).astype(np.float32), (size, 1))
    teta = np.arctan2(x - yc, x.T - xc).astype(np.float32)
    z = np.zeros((size, size, nzer), dtype=np.float32)
    rad = zrmod if ext else zr
    for zn in range(nzer):
        n, m = zernike_numbers(zn + 1)
        for i in range((n - m) // 2 + 1):
            z[:, :, zn] = z[:, :, zn] + (-1.0) ** i * rad ** (n - 2.0 * i) * float(
                math.factorial(n - i)) / float(math.factorial(i) * math.factorial((n + m) // 2 - i)
                                               * math.factorial((n - m) // 2 - i))
        if m == 0:
            z[:, :, zn] = z[:, :, zn] * np.sqrt(n + 1.0)
        elif (zn + 1) % 2 == 1:
            z[:, :, zn] = z[:, :, zn] * np.sqrt(2.0 * (n + 1)) * np.sin(m * teta)
        else:
            z[:, :, zn] = z[:, :, zn] * np.sqrt(2.0 * (n + 1)) * np.cos(m * teta)
    return z * (maskmod if ext else mask)[:, :, None]


def make_tiptilt_dm(p_dm, patch_diam, p_geom, diam):
    dim = max(p_dm._n2 - p_dm._n1 + 1, p_geom._mpupil.shape[0])
    c = p_geom.cent - p_dm._n1 + 1
    influ = zernike_cube(3, dim, patch_diam, c, c, 1)[:, :, 1:]
    current = influ[dim // 2 - 1, dim // 2 - 1, 0] - influ[dim // 2 - 2, dim // 2 - 2, 0]
    fact = p_dm.unitpervolt * diam / p_geom.pupdiam * 4.848 / current
    influ = influ * fact
    p_dm._ntotact = influ.shape[2]
    p_dm._influsize = influ.shape[0]
    p_dm._influ = influ


def dm_init(p_dms, p_tel, p_geom, p_wfss=None):
    """Geometry and influence functions of every mirror, in list order (dm_init.py:56-202)."""
    xw = [w.xpos for w in p_wfss] if p_wfss is not None else [0]
    yw = [w.ypos for w in p_wfss] if p_wfss is not None else [0]
    types = [d.type for d in p_dms]
    if "tt" in types and any(t != "tt" for t in types[types.index("tt"):]):
        raise RuntimeError("TT must be defined at the end of the dms parameters")
    max_extent = 0
    for d in p_dms:
        patch = dm_patch_diameter(p_geom.pupdiam, p_tel.diam, d.alt, xw, yw)
        if d.type == "pzt":
            d._pitch = patch / float(d.nact - 1)
            extent = d._pitch * (d.nact + d.pzt_extent)
            d._n1, d._n2 = dm_support(p_geom.cent, extent, p_geom.ssize)
            make_pzt_dm(d, p_geom, p_tel.cobs)
            max_extent = max(max_extent, d._n2 - d._n1 + 1)
            d._dim_screen = max(d._n2 - d._n1 + 1, p_geom._mpupil.shape[0])
        elif d.type == "tt":
            if d.alt == 0 and max_extent != 0:
                extent = int(max_extent * 1.05)
                extent += extent % 2
            else:
                extent = p_geom.pupdiam + 16
            d._n1, d._n2 = dm_support(p_geom.cent, extent, p_geom.ssize)
            max_extent = max(max_extent, d._n2 - d._n1 + 1)
            make_tiptilt_dm(d, patch, p_geom, p_tel.diam)
            d._dim_screen = d._n2 - d._n1 + 1
        else:
            raise NotImplementedError("DM type %r is outside the hot-path scope" % d.type)


def keep_actuators(p_dm, p_geom, ok):
    """Drop actuators not in `ok` and rebuild the gather tables (correct_dm, dm_init.py:857-871)."""
    offs_i1, offs_j1 = p_dm._i1[ok], p_dm._j1[ok]
    p_dm._ntotact = int(len(ok))
    p_dm._influ = p_dm._influ[:, :, ok]
    p_dm._xpos = p_dm._xpos[ok]
    p_dm._ypos = p_dm._ypos[ok]
    p_dm._i1, p_dm._j1 = offs_i1, offs_j1
    gather_tables(p_dm, p_geom)


def stamp_is_shared(p_dm):
    return bool(np.all(p_dm._influ == p_dm._influ[:, :, :1]))
