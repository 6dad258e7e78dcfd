"""Host tables of the geometric controller (ControllerType.GEO).

Reference: ``init_controller_geo`` (shesha/init/rtc_init.py:418-448) hands the pupil indices and the mirrors to
sutra's ``init_proj_sparse``, which keeps the sparse influence matrix IF [nactu][pupil points], forms
``IF . IF^T``, inverts it in double precision, and every frame computes (``next_part_one_geo``,
shesha/supervisor/rlSupervisor.py:989-1013)

    com = -(IF . IF^T)^-1 . IF . (phi - <phi>)          phi = target phase over the pupil, atmosphere only

i.e. the least-squares fit of the mirrors to the turbulent phase -- the "fitting error only" upper bound of the paper's
plots.  sutra is not vendored in the reference, so the formula above is the published algorithm restated; parity
for this row is a property test against ``numpy.linalg.lstsq`` (tests/test_geo.py), not a golden vector.

Here the influence functions are the ones the device kernels evaluate: piezo actuator (gy, gx) of the square lattice is
``outer(S[gy], S[gx])`` with S the shifted separable factor of the shared stamp, the tip-tilt mirror two planes.  The
device computes ``b = IF . (m phi)`` with two separable passes and subtracts the piston term through ``sifn``.
"""
import numpy as np
import scipy.sparse as sp


def lattice_factors(f, grid_n, pitch, start0, off, n):
    """S [grid_n][n]: S[g][x] = f[x + off - (start0 + g pitch)] inside the stamp, 0 elsewhere."""
    S = np.zeros((grid_n, n), dtype=np.float64)
    for g in range(grid_n):
        xs = start0 + g * pitch - off
        d0, d1 = max(0, -xs), min(f.size, n - xs)
        if d1 > d0:
            S[g, xs + d0:xs + d1] = f[d0:d1]
    return S


def influence_rows(t):
    """Sparse IF [nactu][n*n] (masked by the pupil, flat index x + n y) from the tables the GPU kernels use."""
    from ..lib import actuator_lattice, separable_factor
    n = int(t.n)
    m = (np.asarray(t.mpupil) != 0)
    pitch, grid_n, i1_0, j1_0, amap = actuator_lattice(t.p_pzt)
    f = separable_factor(t.p_pzt._influ[:, :, 0]).astype(np.float64)
    Sx = lattice_factors(f, grid_n, pitch, i1_0, t.pzt_off, n)
    Sy = lattice_factors(f, grid_n, pitch, j1_0, t.pzt_off, n)
    nact = int(t.p_pzt._ntotact)
    rows, cols, vals = [], [], []
    for c in np.nonzero(amap >= 0)[0]:
        gy, gx = divmod(int(c), grid_n)
        ys, xs = np.nonzero(Sy[gy])[0], np.nonzero(Sx[gx])[0]
        if ys.size == 0 or xs.size == 0:
            continue
        blk = np.outer(Sy[gy, ys], Sx[gx, xs]) * m[np.ix_(ys, xs)]
        yy, xx = np.nonzero(blk)
        rows.append(np.full(yy.size, amap[c], dtype=np.int64))
        cols.append(ys[yy] * n + xs[xx])
        vals.append(blk[yy, xx])
    planes = np.asarray(t.p_tt._influ, dtype=np.float64).transpose(2, 1, 0)     # [2][y][x] on the tip-tilt support
    o = int(t.tt_off)
    for j in range(2):
        pl = planes[j, o:o + n, o:o + n] * m
        yy, xx = np.nonzero(pl)
        rows.append(np.full(yy.size, nact + j, dtype=np.int64))
        cols.append(yy * n + xx)
        vals.append(pl[yy, xx])
    IF = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                       shape=(nact + 2, n * n))
    return IF


def build_geo(t, rcond=1e-9):
    """Sets t.geo_proj [nactu][nactu] float32 (= -(IF IF^T)^+) and t.geo_sifn [nactu] float32 (row sums of IF over the
    pupil divided by the number of pupil points: the piston term of b).  Returns (geo_proj, geo_sifn).

    The Gram matrix is inverted in float64 like the reference; directions whose eigenvalue is below ``rcond`` times
    the largest one (the tip-tilt mirror is almost inside the span of the piezo stamps) are dropped -- they carry
    no phase to fit."""
    IF = influence_rows(t)
    G = (IF @ IF.T).toarray()
    w, V = np.linalg.eigh(G)
    keep = w > rcond * w[-1]
    Ginv = (V[:, keep] / w[keep]) @ V[:, keep].T
    npup = float((np.asarray(t.mpupil) != 0).sum())
    t.geo_proj = np.ascontiguousarray(-Ginv, dtype=np.float32)
    t.geo_sifn = np.asarray(IF.sum(axis=1)).ravel().astype(np.float64) / npup
    t.geo_sifn = t.geo_sifn.astype(np.float32)
    t.geo_eig = w
    return t.geo_proj, t.geo_sifn
