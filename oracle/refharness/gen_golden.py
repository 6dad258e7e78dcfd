"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the REFERENCE's own builders.

    python -m oracle.refharness.gen_golden            # all configs
    python -m oracle.refharness.gen_golden 10x10      # one

Runs only in the development container (needs /root/reference).  What is executed is the
reference's unmodified Python (scratch copy, see prepare_ref.py):
  tel_init        shesha/init/geom_init.py:49-108   (init_wfs_geom / init_wfs_size / init_sh_geom / geom_init)
  atmos_init      shesha/init/atmos_init.py:47-134  (iterkolmo.AB)
  dm_init         shesha/init/dm_init.py:56-202     (make_pzt_dm, make_tiptilt_dm, comp_dmgeom)
  target_init     shesha/init/target_init.py:47-141
  wfs_init        shesha/init/wfs_init.py:47-206
  imat_geom       shesha/ao/imats.py:54-112         (on the recording fake + oracle slopes_geom)
  correct_dm      shesha/init/dm_init.py:817-889
  compute_IFsparse / compute_btt / compute_cmat_with_Btt   shesha/ao/basis.py:169-256, 362-443
  get_modes_chosen                                src/.../helper_states.py:286-317
  DelayedMDP                                      src/.../delayed_mdp.py
  GaussianPolicy                                  src/.../model_rpc.py:70-158
The 10x10 configuration is stored in full; for 40x40 large arrays are stored as SHA-256 digests
plus strided samples so the fixture stays small.
"""
import hashlib
import os
import sys

import numpy as np

from . import prepare_ref

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden")

CONFIGS = {
    "10x10": "data/par/par4rl/production/production_sh_10x10_2m.py",
    "40x40": "data/par/par4rl/production/production_sh_40x40_8m_3layers.py",
    "40x40_d0_noise": "data/par/par4rl/production/production_sh_40x40_8m_3layers_d0_noise.py",
}


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def run_reference_init(root, parfile):
    from shesha.util.utilities import load_config_from_file
    from shesha.init import geom_init, atmos_init, dm_init, wfs_init, target_init
    from shesha.ao import imats
    from .fake_sutra import Context

    cfg = load_config_from_file(os.path.join(root, parfile))
    ctx = Context()
    tel = geom_init.tel_init(ctx, cfg.p_geom, cfg.p_tel, cfg.p_atmos.r0, cfg.p_loop.ittime,
                             cfg.p_wfss)
    atm = atmos_init.atmos_init(ctx, cfg.p_atmos, cfg.p_tel, cfg.p_geom, cfg.p_loop.ittime,
                                p_wfss=cfg.p_wfss, p_targets=cfg.p_targets)
    dms = dm_init.dm_init(ctx, cfg.p_dms, cfg.p_tel, cfg.p_geom, cfg.p_wfss)
    tar = target_init.target_init(ctx, tel, cfg.p_targets, cfg.p_atmos, cfg.p_tel, cfg.p_geom,
                                  cfg.p_dms, brahma=False)
    wfs = wfs_init.wfs_init(ctx, tel, cfg.p_wfss, cfg.p_tel, cfg.p_geom, cfg.p_dms, cfg.p_atmos)
    return cfg, ctx, tel, atm, dms, tar, wfs, imats, dm_init


def collect(name, root, parfile, full):
    out = {}

    def put(key, arr, big=False):
        arr = np.asarray(arr)
        if big and not full:
            out[key + "__sha256"] = np.array(digest(arr))
            out[key + "__shape"] = np.array(arr.shape)
            flat = arr.reshape(-1)
            out[key + "__sample"] = flat[::max(1, flat.size // 4096)][:4096].copy()
        else:
            out[key] = arr

    cfg, ctx, tel, atm, dms, tar, wfs, imats, dm_init = run_reference_init(root, parfile)
    g = cfg.p_geom
    for k in ("pupdiam", "ssize", "cent", "_p1", "_p2", "_n", "_n1", "_n2"):
        put("geom." + k.lstrip("_"), getattr(g, k))
    put("geom.spupil", g._spupil.astype(np.uint8), big=True)
    put("geom.mpupil", g._mpupil.astype(np.uint8), big=True)

    for i, w in enumerate(cfg.p_wfss):
        p = "wfs%d." % i
        for k in ("_pdiam", "_Nfft", "_Ntot", "_nrebin", "pixsize", "_qpixsize", "_nvalid",
                  "_nphotons", "_subapd", "npix", "nxsub"):
            put(p + k.lstrip("_"), getattr(w, k))
        put(p + "isvalid", w._isvalid)
        put(p + "validsubsx", w._validsubsx)
        put(p + "validsubsy", w._validsubsy)
        put(p + "validpuppixx", w._validpuppixx)
        put(p + "validpuppixy", w._validpuppixy)
        put(p + "phasemap", w._phasemap, big=True)
        put(p + "binmap", w._binmap)
        put(p + "halfxy", w._halfxy)
        put(p + "fluxPerSub_list", wfs.d_wfs[i].rec["fluxPerSub"])
        put(p + "ftkernel_abs_sum", np.abs(w._ftkernel).sum())
        put(p + "layers", np.array([[l[1], l[2], l[3]] for l in wfs.d_wfs[i].d_gs.layers]))
        put(p + "layer_types", np.array([l[0] for l in wfs.d_wfs[i].d_gs.layers]))
    for i in range(len(cfg.p_targets)):
        put("target%d.layers" % i, np.array([[l[1], l[2], l[3]] for l in tar.d_targets[i].layers]))
    put("target.Npts", tar.rec["Npts"])

    a = cfg.p_atmos
    put("atmos.dim_screens", a.dim_screens)
    put("atmos.deltax", a._deltax)
    put("atmos.deltay", a._deltay)
    put("atmos.r0_layers", atm.rec["r0_layers"])
    put("atmos.stencil_size", atm.rec["stencil_size"])
    put("atmos.pupixsize", a.pupixsize)
    for l, rec in atm.layers.items():
        put("atmos.A%d" % l, rec["A"], big=True)
        put("atmos.B%d" % l, rec["B"], big=True)
        put("atmos.BBt_diag%d" % l, np.einsum("ij,ij->i", rec["B"].astype(np.float64),
                                              rec["B"].astype(np.float64)))
        put("atmos.istx%d" % l, rec["istx"])
        put("atmos.isty%d" % l, rec["isty"])

    # DMs of controller 0 before actuator filtering
    ctrl = cfg.p_controllers[0]
    for j, nm in enumerate(ctrl.ndm):
        d = cfg.p_dms[nm]
        p = "dm%d." % j
        put(p + "type", np.array(str(d.type)))
        put(p + "n1", d._n1)
        put(p + "n2", d._n2)
        put(p + "influsize", d._influsize)
        put(p + "ntotact_before", d._ntotact)
        if d.type == "pzt":
            put(p + "pitch", d._pitch)
            put(p + "xpos_before", d._xpos)
            put(p + "ypos_before", d._ypos)
            put(p + "i1_before", d._i1)
            put(p + "j1_before", d._j1)
            put(p + "stamp", d._influ[:, :, 0])
            put(p + "stamps_identical",
                np.array(bool(np.all(d._influ == d._influ[:, :, :1]))))
        else:
            put(p + "influ", d._influ, big=True)
            put(p + "influ_center", d._influ[d._influ.shape[0] // 2 - 2:d._influ.shape[0] // 2 + 2,
                                             d._influ.shape[0] // 2 - 2:d._influ.shape[0] // 2 + 2, :])

    # actuator filtering: reference imat_geom + correct_dm on top of the oracle's slopes_geom
    imat_g = imats.imat_geom(wfs, dms, cfg.p_wfss, cfg.p_dms, ctrl, meth=0)
    put("imat_geom.colnorm", np.sqrt(np.sum(imat_g.astype(np.float64) ** 2, axis=0)))
    dm_init.correct_dm(ctx, dms, cfg.p_dms, ctrl, cfg.p_geom, imat_g)
    d0 = cfg.p_dms[ctrl.ndm[0]]
    put("dm0.ntotact", d0._ntotact)
    put("dm0.i1", d0._i1)
    put("dm0.j1", d0._j1)
    put("dm0.xpos", d0._xpos)
    put("dm0.ypos", d0._ypos)
    put("dm0.influpos", d0._influpos, big=True)
    put("dm0.ninflu", d0._ninflu, big=True)
    put("dm0.influstart", d0._influstart, big=True)

    # Btt basis through the reference's own compute_IFsparse / compute_btt
    from shesha.ao import basis as ref_basis
    g_dms = [dms.d_dms[nm] for nm in ctrl.ndm]
    p_dms = [cfg.p_dms[nm] for nm in ctrl.ndm]

    class _ShapeView:
        """np.array(sutra d_shape) shows the device buffer with F-order dims: transpose here."""

        def __init__(self, dm):
            self._dm = dm

        def reset_shape(self):
            self._dm.reset_shape()

        def comp_oneactu(self, i, v):
            self._dm.comp_oneactu(i, v)

        @property
        def d_shape(self):
            return self._dm.d_shape.T

    IFs = ref_basis.compute_IFsparse([_ShapeView(d) for d in g_dms], p_dms, cfg.p_geom).T
    n = IFs.shape[1]
    IFtt = IFs[:, -2:].copy().toarray()
    IFpzt = IFs[:, :n - 2]
    Btt, P = ref_basis.compute_btt(IFpzt, IFtt)
    put("basis.Btt", Btt, big=True)
    put("basis.P", P, big=True)
    put("basis.BttP_diag", np.einsum("ij,ji->i", Btt.astype(np.float64), P.astype(np.float64)))
    put("basis.PBtt_err", np.abs(P.astype(np.float64) @ Btt.astype(np.float64)
                                 - np.eye(Btt.shape[1])).max())
    put("basis.IF_nnz", IFs.nnz)
    put("basis.IF_colsum", np.asarray(IFs.sum(axis=0)).ravel())
    return out, (cfg, dms, wfs, tel, Btt, P)


def collect_rl(out):
    """Integer index tables and tiny-network outputs from the reference's RL helpers."""
    import types
    import torch
    from src.reinforcement_learning.rpc_training.helper_rpc import helper_states as hs
    from src.reinforcement_learning.environment.delayed_mdp import DelayedMDP
    from src.reinforcement_learning.rpc_training.algorithms_rpc.model_rpc import GaussianPolicy
    hs.debug_modes_chosen = False

    def layout(total_modes, start, end, n_agents_modes, window, tt_windowed, n_filtered, keys):
        # create_agents_dictionary_original (train_rpc.py:265-296)
        per = (end - start) // n_agents_modes
        d = {}
        wid = 1
        for m in range(start, end, per):
            d[wid] = [m, m + per]
            wid += 1
        d[wid] = [total_modes - 2, total_modes]
        ios = {}
        i0 = 0
        for k in keys:
            ios[k] = [i0, i0 + total_modes if window > -1 else i0 + (end - start + 2)]
            i0 = ios[k][1]
        cfg = types.SimpleNamespace(env_rl=dict(window_n_zernike=window, include_tip_tilt=True,
                                                include_tip_tilt_windowed=tt_windowed,
                                                tt_treated_as_mode=False))
        mc = hs.get_modes_chosen(d, ios, cfg, n_filtered, "golden", total_modes,
                                 end - start + 2, start)
        return d, mc

    keys = ["dm_history_2", "dm_history_1", "dm_before_linear", "dm_residual"]
    # 10x10: 2 agents (80 modes + TT), no window  (BASELINE configs[0/1])
    d, mc = layout(87, 0, 80, 1, -1, False, 5, keys)
    for w, v in mc.items():
        out["rl.10x10.modes_chosen_%d" % w] = np.asarray(v, dtype=np.int64)
    # 40x40 "14_8_3l_dir_w20": 43 agents, modes 0..1260 in groups of 30, window 20, 5 filtered
    d, mc = layout(1283, 0, 1260, 42, 20, False, 5, keys)
    for w, v in mc.items():
        out["rl.40x40w20.modes_chosen_%d" % w] = np.asarray(v, dtype=np.int64)
    d, mc = layout(1283, 0, 1260, 14, 20, True, 5, keys)
    for w, v in mc.items():
        out["rl.40x40_14w20tt.modes_chosen_%d" % w] = np.asarray(v, dtype=np.int64)

    # DelayedMDP slot order: feed integers, record what credit_assignment returns
    for delay, modif in ((1, False), (2, False), (0, False), (1, True)):
        m = DelayedMDP(delay, modif)
        rec = []
        for t in range(8):
            if m.check_update_possibility():
                s, a, sn = m.credit_assignment()
                rec.append((t, s, a, sn))
            m.save(100 + t, 200 + t, 300 + t)
        out["rl.delayed_mdp_d%d_m%d" % (delay, int(modif))] = np.array(rec, dtype=np.int64)

    # GaussianPolicy forward on fixed weights
    torch.manual_seed(1234)
    pol = GaussianPolicy(num_inputs=24, hidden_dim=32, num_actions=6, num_layers=2,
                         activation="relu", initialize_last_layer_zero=False,
                         initialize_last_layer_near_zero=False, action_scale=1.0, action_bias=0.0,
                         LOG_SIG_MAX=2.0)
    with torch.no_grad():
        pol.mean_linear.bias.uniform_(-0.5, 0.5)
        pol.log_std_linear.bias.uniform_(-3.0, 3.0)
        x = torch.randn(5, 24)
        mean, log_std = pol.forward(x)
    sd = pol.state_dict()
    for k, v in sd.items():
        out["rl.policy." + k] = v.numpy()
    out["rl.policy.x"] = x.numpy()
    out["rl.policy.mean"] = mean.numpy()
    out["rl.policy.log_std"] = log_std.numpy()


def collect_extrude(out):
    """One extruded column per direction from the reference's own iterkolmo.extrude (shesha/util/iterkolmo.py:255-288)
    with np.random.normal patched to return a fixed vector, on a 32 x 32 screen with the operators of the
    reference's own AB (190-252) for every sign combination of the wind (stencil mirroring 246-249, isty 241-244)."""
    from shesha.util import iterkolmo as itK
    n, L0, r0 = 32, 1.0e5, 7.3
    r = np.random.default_rng(42)
    p = r.standard_normal((n, n)).astype(np.float32).cumsum(axis=0).cumsum(axis=1) / np.float32(40.0)
    eps = r.standard_normal(n)
    out["ext.n"], out["ext.L0"], out["ext.r0"] = np.int64(n), np.float64(L0), np.float64(r0)
    out["ext.p"], out["ext.eps"] = p.astype(np.float32), eps.astype(np.float64)
    real_normal = np.random.normal
    for tag, (dx, dy) in {"pp": (1.0, 1.0), "np": (-1.0, 1.0), "pn": (1.0, -1.0), "nn": (-1.0, -1.0)}.items():
        A, B, istx, isty = itK.AB(n, L0, dx, dy)
        if tag == "pp":
            out["ext.A"], out["ext.B"] = np.ascontiguousarray(A), np.ascontiguousarray(B)
        out["ext.istx_" + tag], out["ext.isty_" + tag] = istx, isty
    np.random.normal = lambda loc, scale, size: eps.copy()
    try:
        A, B = np.asfortranarray(out["ext.A"]), np.asfortranarray(out["ext.B"])
        # +x: the reference function as it stands
        out["ext.p1_px"] = itK.extrude(p.copy(), r0, A, B, out["ext.istx_pp"])
        # the reference's python extrude always appends on the right; the other three directions are the same call
        # on the transposed / 180-degree rotated screen with the +x stencil (what the mirrored index lists encode)
        out["ext.p1_py"] = itK.extrude(np.ascontiguousarray(p.T), r0, A, B, out["ext.istx_pp"]).T.copy()
        out["ext.p1_nx"] = itK.extrude(np.ascontiguousarray(p[::-1, ::-1]), r0, A, B, out["ext.istx_pp"])[::-1, ::-1].copy()
        out["ext.p1_ny"] = itK.extrude(np.ascontiguousarray(p.T[::-1, ::-1]), r0, A, B, out["ext.istx_pp"])[::-1, ::-1].T.copy()
    finally:
        np.random.normal = real_normal


def main(which=None):
    root = prepare_ref.activate(fake_sutra=True)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    cwd = os.getcwd()
    os.chdir(root)
    try:
        for name, par in CONFIGS.items():
            if which and name not in which:
                continue
            out, _ = collect(name, root, par, full=(name == "10x10"))
            np.savez_compressed(os.path.join(GOLDEN_DIR, "ref_tables_%s.npz" % name), **out)
            print("wrote", name, len(out), "entries")
        if not which or "extrude" in which:
            out = {}
            collect_extrude(out)
            np.savez_compressed(os.path.join(GOLDEN_DIR, "ref_extrude.npz"), **out)
            print("wrote extrude", len(out), "entries")
        if not which or "rl" in which:
            out = {}
            collect_rl(out)
            np.savez_compressed(os.path.join(GOLDEN_DIR, "ref_rl.npz"), **out)
            print("wrote rl", len(out), "entries")
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    main(sys.argv[1:] or None)
