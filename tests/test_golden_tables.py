"""Host table builders vs golden vectors produced by the REFERENCE's own builders
(oracle/refharness/gen_golden.py).  Integer tables bit-exact, float tables to rounding."""
import hashlib

import numpy as np
import pytest

from ao_marl_b200 import tables
from ao_marl_b200.config import load_config_from_file
from ao_marl_b200.init.atmos import transposed_stencil


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class Golden:
    def __init__(self, npz):
        self.g = npz

    def exact(self, key, mine):
        mine = np.asarray(mine)
        if key in self.g.files:
            ref = self.g[key]
            assert ref.shape == mine.shape, key
            assert np.array_equal(ref, mine.astype(ref.dtype)), key
        else:
            dt = self.g[key + "__sample"].dtype
            assert tuple(self.g[key + "__shape"]) == mine.shape, key
            assert sha(mine.astype(dt)) == str(self.g[key + "__sha256"]), key

    def close(self, key, mine, tol):
        mine = np.asarray(mine, dtype=np.float64)
        if key in self.g.files:
            ref = self.g[key].astype(np.float64)
            assert ref.shape == mine.shape, key
        else:
            assert tuple(self.g[key + "__shape"]) == mine.shape, key
            ref = self.g[key + "__sample"].astype(np.float64)
            flat = mine.reshape(-1)
            mine = flat[::max(1, flat.size // 4096)][:4096]
        assert np.abs(ref - mine).max() <= tol * max(np.abs(ref).max(), 1e-30), key


def check_config(t, G, full_basis):
    cfg = t.config
    g = cfg.p_geom
    for k, v in (("pupdiam", g.pupdiam), ("ssize", g.ssize), ("cent", g.cent), ("p1", g._p1), ("p2", g._p2),
                 ("n", g._n), ("n1", g._n1), ("n2", g._n2)):
        G.exact("geom." + k, v)
    G.exact("geom.spupil", g._spupil.astype(np.uint8))
    G.exact("geom.mpupil", g._mpupil.astype(np.uint8))
    for i, w in enumerate(cfg.p_wfss):
        p = "wfs%d." % i
        for k in ("_pdiam", "_Nfft", "_Ntot", "_nrebin", "_nvalid", "npix", "nxsub"):
            G.exact(p + k.lstrip("_"), getattr(w, k))
        for k in ("pixsize", "_qpixsize", "_nphotons", "_subapd"):
            G.close(p + k.lstrip("_"), getattr(w, k), 1e-12)
        for k in ("isvalid", "validsubsx", "validsubsy", "validpuppixx", "validpuppixy", "phasemap", "binmap",
                  "halfxy"):
            G.exact(p + k, getattr(w, "_" + k))
        G.exact(p + "fluxPerSub_list", w._fluxPerSub_list)
    a = cfg.p_atmos
    G.exact("atmos.dim_screens", a.dim_screens)
    G.exact("atmos.deltax", a._deltax)
    G.exact("atmos.deltay", a._deltay)
    G.close("atmos.r0_layers", a._r0_layers, 1e-12)
    G.exact("atmos.stencil_size", a._stencil_size)
    for l in range(a.nscreens):
        n = int(a.dim_screens[l])
        istx = t.istx[l].astype(np.int64)
        isty = transposed_stencil(istx, n)
        if a._deltax[l] < 0:
            istx = n * n - 1 - istx
        if a._deltay[l] < 0:
            isty = n * n - 1 - isty
        G.exact("atmos.istx%d" % l, istx.astype(np.uint32))
        G.exact("atmos.isty%d" % l, isty.astype(np.uint32))
        # A/B come out of an ill-conditioned pseudo-inverse: LAPACK threading alone moves B B^T by ~1e-3
        # between two runs of the reference itself (same-process runs agree bit for bit)
        G.close("atmos.A%d" % l, t.A[l], 1e-5)
        bbt = np.einsum("ij,ij->i", t.B[l].astype(np.float64), t.B[l].astype(np.float64))
        G.close("atmos.BBt_diag%d" % l, bbt, 5e-3)
    d = t.p_pzt
    for k in ("n1", "n2", "influsize"):
        G.exact("dm0." + k, getattr(d, "_" + k))
    G.close("dm0.pitch", d._pitch, 1e-12)
    G.exact("dm0.ntotact", d._ntotact)          # the KAT: 88 / 1284 actuators survive correct_dm
    for k in ("i1", "j1", "xpos", "ypos", "influpos", "ninflu", "influstart"):
        G.exact("dm0." + k, getattr(d, "_" + k))
    G.close("dm0.stamp", d._influ[:, :, 0], 1e-6)
    for k in ("n1", "n2", "influsize"):
        G.exact("dm1." + k, getattr(t.p_tt, "_" + k))
    G.close("dm1.influ", t.p_tt._influ, 1e-6)
    layers = G.g["wfs0.layers"]
    types = list(G.g["wfs0.layer_types"])
    nl = a.nscreens
    assert np.allclose(layers[:nl, 1], t.wfs_xoff) and np.allclose(layers[:nl, 2], t.wfs_yoff)
    assert types[nl:] == ["pzt", "tt"] and layers[nl, 1] == t.pzt_off and layers[nl + 1, 1] == t.tt_off
    G.close("basis.BttP_diag", np.einsum("ij,ji->i", t.Btt.astype(np.float64), t.P.astype(np.float64)), 1e-4)
    assert np.abs(t.P.astype(np.float64) @ t.Btt.astype(np.float64) - np.eye(t.Btt.shape[1])).max() < 1e-5
    if full_basis:
        Bref, Pref = G.g["basis.Btt"], G.g["basis.P"]
        sgn = np.sign(np.sum(Bref * t.Btt, axis=0))
        assert np.abs(Bref - t.Btt * sgn).max() <= 1e-4 * np.abs(Bref).max()
        assert np.abs(Pref - t.P * sgn[:, None]).max() <= 1e-4 * np.abs(Pref).max()


def test_tables_10x10(static10, golden10):
    assert static10.p_pzt._ntotact == 88 and static10.nslopes == 128
    check_config(static10, Golden(golden10), True)


@pytest.mark.slow
def test_tables_40x40(golden40):
    t = tables.build_static(load_config_from_file("production_sh_40x40_8m_3layers.py"))
    tables.build_basis(t)
    assert t.p_pzt._ntotact == 1284 and t.nslopes == 2400 and t.Btt.shape == (1286, 1283)
    check_config(t, Golden(golden40), False)


def test_reference_format_parameter_file(tmp_path):
    """A file written for the reference (import shesha.config as conf; set_* calls) loads unchanged."""
    p = tmp_path / "my_par.py"
    p.write_text(
        "import shesha.config as conf\nimport numpy as np\nsimul_name='x'\n"
        "p_loop = conf.Param_loop()\np_loop.set_niter(10)\np_loop.set_ittime(0.002)\n"
        "p_geom = conf.Param_geom()\np_geom.set_zenithangle(0.)\n"
        "p_tel = conf.Param_tel()\np_tel.set_diam(2.0)\np_tel.set_cobs(0.12)\n"
        "p_atmos = conf.Param_atmos()\np_atmos.set_r0(0.16)\np_atmos.set_nscreens(1)\np_atmos.set_frac([1.0])\n"
        "p_atmos.set_alt([0.0])\np_atmos.set_windspeed([20.0])\np_atmos.set_winddir([45.])\np_atmos.set_L0([1.e5])\n"
        "p_wfs0 = conf.Param_wfs()\np_wfss=[p_wfs0]\np_wfs0.set_type('sh')\np_wfs0.set_nxsub(10)\n"
        "p_wfs0.set_dms_seen(np.array([0, 1]))\n")
    cfg = load_config_from_file(str(p))
    assert cfg.p_tel.diam == 2.0 and cfg.p_atmos.nscreens == 1 and cfg.p_wfss[0].nxsub == 10
    assert cfg.p_atmos.windspeed.dtype == np.float32 and cfg.p_dms is None
    assert list(cfg.p_wfss[0].get_dms_seen()) == [0, 1]


def test_extrude_against_reference():
    """Row a-1 pinned: the oracle's extrusion (oracle/aoframe.py) against ONE column per direction extruded by the
    reference's own iterkolmo.extrude (shesha/util/iterkolmo.py:255-288) with a fixed noise vector
    (tests/golden/ref_extrude.npz, generated by oracle/refharness/gen_golden.py::collect_extrude), and the
    mirrored / transposed stencil conventions against the reference's AB (iterkolmo.py:241-249)."""
    import os
    from oracle import aoframe
    from ao_marl_b200.init import atmos as atm
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_extrude.npz"))
    n = int(g["ext.n"])
    p, eps, A, B = g["ext.p"], g["ext.eps"].astype(np.float32), g["ext.A"], g["ext.B"]
    istx = g["ext.istx_pp"].astype(np.int64)
    amp = np.float32(float(g["ext.r0"]) ** (-5.0 / 6.0))       # the reference's python works in rad @ 0.5 um, no micron factor
    # the product's builders reproduce the reference's operators and stencil (bit-exact integers)
    assert np.array_equal(atm.stencil_indices(n), istx)
    A2, B2, _ = atm.extrusion_operands(n, float(g["ext.L0"]))
    assert np.abs(A2 - A).max() < 2e-5 * np.abs(A).max()
    assert np.abs(B2 @ B2.T - B.astype(np.float64) @ B.astype(np.float64).T).max() < 1e-4 * np.abs(B).max() ** 2   # SVD sign freedom
    assert np.array_equal(atm.transposed_stencil(istx, n), g["ext.isty_pp"].astype(np.int64))
    # the reference's mirrored lists: negative wind = the +x stencil read on the 180-degree rotated screen,
    # y = the +x stencil read on the transposed screen (what oracle.aoframe.extrude does)
    flat = p.reshape(-1)
    assert np.array_equal(flat[g["ext.istx_np"].astype(np.int64)], p[::-1, ::-1].reshape(-1)[istx])
    assert np.array_equal(flat[g["ext.isty_pp"].astype(np.int64)], p.T.reshape(-1)[istx])
    assert np.array_equal(flat[g["ext.isty_pn"].astype(np.int64)], p.T[::-1, ::-1].reshape(-1)[istx])
    assert np.array_equal(g["ext.istx_nn"], g["ext.istx_np"]) and np.array_equal(g["ext.isty_nn"], g["ext.isty_pn"])
    for key, axis, sign in (("px", 0, 1), ("py", 1, 1), ("nx", 0, -1), ("ny", 1, -1)):
        ref = g["ext.p1_" + key]
        mine = aoframe.extrude(p.copy(), A, B, istx, amp, eps, axis, sign)
        assert mine.shape == ref.shape
        assert np.abs(mine - ref).max() < 2e-6 * np.abs(ref).max(), key      # float32 dot in the reference, float64 in the oracle
        # the untouched part of the screen is shifted bit-exactly
        if key == "px":
            assert np.array_equal(mine[:, :-1], p[:, 1:])
        if key == "ny":
            assert np.array_equal(mine[1:, :], p[:-1, :])
