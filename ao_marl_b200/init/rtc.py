"""Real-time-controller tables (host side, numpy/scipy, once per configuration).

* geometric interaction matrix + unseen-actuator filtering:
  shesha/ao/imats.py:54-112 (imat_geom) and shesha/init/dm_init.py:817-889 (correct_dm), called
  from shesha/init/rtc_init.py:116-123.  The geometric slope itself is computed inside sutra
  (``slopes_geom``) and is not visible in the reference: the convention used here (mean of the
  central-difference gradient over the illuminated tile, micron/m -> arcsec) is pinned by the
  integer outcome the reference's committed files imply: 88 / 1284 actuators kept.
* centroider constants: shesha/init/rtc_init.py:193-268.
* Btt modal basis and the filtered command matrix: shesha/ao/basis.py:169-200, 229-256, 362-443,
  driven as in shesha/supervisor/rlSupervisor.py:169-194, 215-234.
"""
import numpy as np
import scipy.sparse as sp


def dm_offset_in_wfs(p_dm, p_geom):
    """Offset of the mpupil frame inside a DM support (wfs_init.py:187-204, on-axis, alt 0)."""
    dims = p_dm._n2 - p_dm._n1 + 1
    dim = max(p_geom._mpupil.shape[0], dims)
    off = (dim - p_geom._n) / 2
    if p_dm.alt != 0 or not float(off).is_integer():
        raise NotImplementedError("altitude-conjugated DMs are outside the hot-path scope")
    return int(off)


def atmos_offset_in_wfs(p_atmos, p_geom, p_wfs, layer):
    """wfs_init.py:175-185."""
    if p_wfs.gsalt > 0:
        raise NotImplementedError("laser guide stars are outside the hot-path scope")
    from .geom import ARCSEC2RAD
    base = (p_atmos.dim_screens[layer] - p_geom._n) / 2.0
    xoff = p_wfs.xpos * ARCSEC2RAD * p_atmos.alt[layer] / p_atmos.pupixsize + base
    yoff = p_wfs.ypos * ARCSEC2RAD * p_atmos.alt[layer] / p_atmos.pupixsize + base
    return float(xoff), float(yoff)


def _pzt_poke(p_dm, k, dim):
    ss = p_dm._influsize
    shape = np.zeros((dim, dim), np.float32)
    x0, y0 = int(p_dm._i1[k]), int(p_dm._j1[k])
    xa, xb, ya, yb = max(x0, 0), min(x0 + ss, dim), max(y0, 0), min(y0 + ss, dim)
    shape[ya:yb, xa:xb] = p_dm._influ[xa - x0:xb - x0, ya - y0:yb - y0, k].T
    return shape


def geometric_slopes(phase, p_wfs, p_geom):
    """Vectorised over the regular subaperture tiling; returns [x..., y...] over valid subaps."""
    pd, nx = p_wfs._pdiam, p_wfs.nxsub
    n = pd * nx
    t = phase[2:2 + n, 2:2 + n].astype(np.float64).reshape(nx, pd, nx, pd)
    m = p_geom._mpupil[2:2 + n, 2:2 + n].astype(np.float64).reshape(nx, pd, nx, pd)
    gx = np.empty_like(t)
    gx[..., 1:-1] = t[..., 2:] - t[..., :-2]
    gx[..., 0] = t[..., 1] - t[..., 0]
    gx[..., -1] = t[..., -1] - t[..., -2]
    gy = np.empty_like(t)
    gy[:, 1:-1] = t[:, 2:] - t[:, :-2]
    gy[:, 0] = t[:, 1] - t[:, 0]
    gy[:, -1] = t[:, -1] - t[:, -2]
    alpha = 0.206265 / p_wfs._subapd
    sx = (gx * m).sum(axis=(1, 3)) / pd / 2.0 * alpha        # [tile row, tile col]
    sy = (gy * m).sum(axis=(1, 3)) / pd / 2.0 * alpha
    rows = p_wfs._validsubsy // p_wfs.npix
    cols = p_wfs._validsubsx // p_wfs.npix
    flux = p_wfs._fluxPerSub_list.astype(np.float64)
    return np.concatenate([sx[rows, cols] / flux, sy[rows, cols] / flux]).astype(np.float32)


def imat_geom(p_wfs, p_dms_ctrl, p_geom):
    """Geometric interaction matrix [2*nvalid, nactu] for the DMs of one controller."""
    n = p_geom._n
    cols = []
    for d in p_dms_ctrl:
        off = dm_offset_in_wfs(d, p_geom)
        dim = max(d._n2 - d._n1 + 1, n)
        for k in range(d._ntotact):
            if d.type == "pzt":
                shape = _pzt_poke(d, k, dim) * np.float32(d.push4imat)
            else:
                shape = d._influ[:, :, k].T * np.float32(d.push4imat)
            ph = shape[off:off + n, off:off + n]
            cols.append(geometric_slopes(ph, p_wfs, p_geom) / np.float32(d.push4imat))
    return np.stack(cols, axis=1).astype(np.float32)


def correct_dm(p_dms_ctrl, p_geom, imat):
    """Drop piezo actuators whose geometric response is below thresh * max (dm_init.py:857-871)."""
    from .dm import keep_actuators
    resp = np.sqrt(np.sum(imat ** 2, axis=0))
    ind = 0
    for d in p_dms_ctrl:
        nact = d._ntotact
        if d.type == "pzt":
            r = resp[ind:ind + nact]
            ok = np.where(r > d.thresh * np.max(r))[0]
            keep_actuators(d, p_geom, ok)
        ind += nact


def centroider_constants(p_wfs):
    """COG offset (pixels) and scale (arcsec per pixel) -- rtc_init.py:207-217."""
    return float(p_wfs.npix // 2 - 0.5), float(p_wfs.pixsize)


def influence_matrix(p_dms_ctrl, p_geom):
    """Sparse [pupil pixels, nactu] matrix of every actuator's surface inside the pupil
    (compute_dm_basis / compute_IFsparse, basis.py:117-200).  Pixel order is the C-flatten of the
    DM support restricted to the illuminated pixels of ipupil."""
    blocks = []
    ip = p_geom._ipupil
    for d in p_dms_ctrl:
        dims = d._n2 - d._n1 + 1
        dim = max(dims, p_geom._mpupil.shape[0])
        margin = (ip.shape[0] - dims) // 2
        pup = ip[margin:ip.shape[0] - margin, margin:ip.shape[1] - margin]
        valid = np.nonzero(pup.reshape(-1, order="F") > 0)[0]
        lookup = -np.ones(dim * dim, dtype=np.int64)
        lookup[valid] = np.arange(valid.size)
        rows, colsi, vals = [], [], []
        for k in range(d._ntotact):
            if d.type == "pzt":
                ss = d._influsize
                x0, y0 = int(d._i1[k]), int(d._j1[k])
                xa, xb, ya, yb = max(x0, 0), min(x0 + ss, dim), max(y0, 0), min(y0 + ss, dim)
                st = d._influ[xa - x0:xb - x0, ya - y0:yb - y0, k].T
                yy, xx = np.mgrid[ya:yb, xa:xb]
                flat = (xx + dim * yy).reshape(-1)
                v = st.reshape(-1)
            else:
                flat = np.arange(dim * dim)
                v = d._influ[:, :, k].T.reshape(-1)
            r = lookup[flat]
            sel = (r >= 0) & (v != 0)
            rows.append(r[sel])
            colsi.append(np.full(int(sel.sum()), k))
            vals.append(v[sel])
        blocks.append(sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows),
                                                           np.concatenate(colsi))),
                                    shape=(valid.size, d._ntotact)))
    return sp.hstack(blocks).tocsr()


def compute_btt(IFpzt, IFtt):
    """Btt (modes -> volts, [nactu, nactu-3]) and P (volts -> modes) -- basis.py:362-443."""
    N, n = IFpzt.shape
    delta = (IFpzt.T @ IFpzt).toarray() / N
    Tp = np.ones((IFtt.shape[0], 3))
    Tp[:, :2] = IFtt
    deltaT = IFpzt.T @ Tp / N
    tau = np.linalg.inv(delta) @ deltaT
    G = np.identity(n)
    tdt = tau.T @ delta @ tau
    G -= tau @ np.linalg.inv(tdt) @ tau.T @ delta
    gdg = G.T @ delta @ G
    U, s, _ = np.linalg.svd(gdg)
    U = U[:, :-3]
    s = s[:-3]
    B = (G @ U) / np.sqrt(s)[None, :]
    TT = IFtt.T @ IFtt / N
    Btt = np.zeros((n + 2, n - 1))
    Btt[:n, :n - 3] = B
    with np.errstate(divide="ignore"):
        mini = 1.0 / np.sqrt(np.abs(TT))
    mini[0, 1] = mini[1, 0] = 0
    Btt[n:, -2:] = mini
    Delta = np.zeros((n + 2, n + 2))
    Delta[:-2, :-2] = delta
    Delta[-2:, -2:] = TT
    P = Btt.T @ Delta
    return Btt.astype(np.float32), P.astype(np.float32)


def cmat_with_btt(imat, Btt, nfilt):
    """Least-squares command matrix restricted to the Btt modes minus the `nfilt` last non-TT ones
    (compute_cmat_with_Btt, basis.py:229-256).  Returns float32 [nactu, nslopes]."""
    nm = Btt.shape[1]
    Bf = np.zeros((Btt.shape[0], nm - nfilt))
    Bf[:, :nm - nfilt - 2] = Btt[:, :nm - (nfilt + 2)]
    Bf[:, nm - nfilt - 2:] = Btt[:, nm - 2:]
    Dm = imat.astype(np.float32).dot(Bf)
    Dmp = np.linalg.inv(Dm.T.dot(Dm)).dot(Dm.T)
    return Bf.dot(Dmp).astype(np.float32)
