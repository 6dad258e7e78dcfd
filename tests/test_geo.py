"""Geometric controller: host tables and CUDA path against the oracle's least-squares fit.

The arithmetic of the reference sits in sutra (not vendored) -- parity is unpinned and anchored on the published
algorithm com = -(IF IF^T)^-1 IF (phi - <phi>) (oracle/geo.py): the residual of a least-squares fit is unique, so the
checks compare the residual phase variance over the pupil (and the commands where the Gram matrix is well conditioned).

Tolerances: host float64 pipeline vs LSQR 1e-6 relative on the residual variance; CUDA float32 pipeline 1e-3 on the
10x10 case and on the 40x40 case (Gram condition number 1e5; measured 8e-5 with the exact-FFMA product -- the tensor-core
GEMM's truncated accumulator gave 16 % there)."""
import numpy as np
import pytest


def smooth_phase(n, seed):
    r = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n] / n
    ph = sum(r.normal() * np.cos(2 * np.pi * (k * xx + l * yy) + r.uniform(0, 6)) / (1 + k * k + l * l)
             for k in range(6) for l in range(6))
    return ph + 3 * xx - 2 * yy + 5


@pytest.fixture(scope="module")
def geo10(static10, oracle_tab10):
    from ao_marl_b200.init import geo
    from oracle import geo as ogeo
    geo.build_geo(static10)
    IFt, idx = ogeo.influence_matrix(oracle_tab10)
    return geo.influence_rows(static10), IFt, idx


def test_influence_rows_match_the_oracle_mirror_model(static10, geo10):
    """Lattice / separable-factor rows of the product == unit commands through the oracle's stamp superposition."""
    IF, IFt, idx = geo10
    a = IF.toarray()[:, idx]
    b = IFt.toarray().T
    assert a.shape == b.shape == (static10.nactu, idx.size)
    assert np.abs(a - b).max() < 5e-6 * np.abs(b).max()
    assert np.abs(IF.toarray()).sum() == pytest.approx(np.abs(a).sum())      # nothing outside the pupil


def test_projector_is_the_least_squares_fit(static10, geo10):
    from oracle import geo as ogeo
    IF, IFt, idx = geo10
    n = static10.n
    m = (static10.mpupil != 0).ravel()
    assert static10.geo_proj.shape == (static10.nactu, static10.nactu)
    for seed in (0, 1):
        phi = smooth_phase(n, seed).ravel()
        com_o, res_o = ogeo.geo_command(IFt, idx, phi.reshape(n, n))
        # the device's flow in float64: raw pupil sums, piston through sifn, one product with the projector
        b = IF @ (m * phi) - (m * phi).sum() * static10.geo_sifn.astype(np.float64)
        com = static10.geo_proj.astype(np.float64) @ b
        res = (phi[idx] - phi[idx].mean()) + IFt @ com
        assert abs(res.var() - res_o.var()) < 1e-6 * res_o.var()
        assert res.var() < 0.05 * phi[idx].var()
        assert np.abs(com - com_o).max() < 1e-3 * np.abs(com_o).max()


def test_geo_tables_are_optional_and_checked(static10):
    """Rtc / target components only route index 1 to the geometric controller when the parameter file has one."""
    from ao_marl_b200.config import load_config_from_file
    cfg = load_config_from_file("production_sh_10x10_2m.py")
    types = [c.type for c in cfg.p_controllers]
    assert types == ["ls", "geo"]
    assert sorted(int(d) for d in cfg.p_targets[1].dms_seen) == sorted(int(d) for d in cfg.p_controllers[1].ndm)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_geo_control_matches_the_oracle_fit(static10, geo10):
    import torch
    from ao_marl_b200.lib import Simulator
    from oracle import geo as ogeo
    IF, IFt, idx = geo10
    sim = Simulator(static10, 3, rl=None)
    try:
        sim.reset(np.array([201, 202, 203], dtype=np.int64))
        for _ in range(3):
            sim.move_atmos()
        phase = sim.raytrace_wfs(atmos=True, dms=False).cpu().numpy().astype(np.float64)
        com = sim.do_control_geo().cpu().numpy().astype(np.float64)
        sim.check_device()
        assert float(sim.rows("GEO_VOLTS", static10.nactu).abs().max()) == 0.0      # not applied yet
        sim.apply_control_geo()
        lam = 1.65
        s_geo = sim.comp_strehl(lam, atmos=True, dms=True, geo=True).cpu().numpy()
        s_main = sim.comp_strehl(lam, atmos=True, dms=True).cpu().numpy()            # main mirrors are flat
        for e in range(3):
            com_o, res_o = ogeo.geo_command(IFt, idx, phase[e])
            phi = phase[e].ravel()[idx]
            res = (phi - phi.mean()) + IFt @ com[e]
            assert abs(res.var() - res_o.var()) < 1e-3 * res_o.var(), (e, res.var(), res_o.var())
            assert np.abs(com[e] - com_o).max() < 2e-3 * np.abs(com_o).max()
            assert abs(s_geo[e, 2] - res.var()) < 1e-3 * res.var() + 1e-7
            assert abs(s_main[e, 2] - phi.var()) < 1e-4 * phi.var()
            assert s_geo[e, 0] > s_main[e, 0]
    finally:
        sim.close()


@pytest.mark.gpu
def test_pupil_sweep_matches_the_per_pixel_kernels(static10, geo10):
    """Staged-row sweep (default) against the plain-load kernels: geometric command and Strehl figures, with mirrors."""
    import torch
    from ao_marl_b200.lib import Simulator
    sim = Simulator(static10, 5, rl=None)
    try:
        sim.reset(np.array([11, 12, 13, 14, 15], dtype=np.int64))
        r = np.random.default_rng(2)
        sim.set_dm_volts(torch.as_tensor((r.standard_normal((5, static10.nactu)) * 3).astype(np.float32), device="cuda"))
        out = {}
        for it in range(2):
            sim.move_atmos()
            for path in ("pixel", "sweep"):
                sim.set_pupil_path(path)
                com = sim.do_control_geo().clone()
                s_atm = sim.comp_strehl(1.65, atmos=True, dms=False, accumulate=False).clone()
                s_all = sim.comp_strehl(1.65, atmos=True, dms=True, accumulate=False).clone()
                s_dm = sim.comp_strehl(1.65, atmos=False, dms=True, accumulate=False).clone()
                out[path] = (com, s_atm, s_all, s_dm)
            sim.check_device()
            a, b = out["pixel"], out["sweep"]
            assert float((a[0] - b[0]).abs().max()) < 2e-4 * float(a[0].abs().max())
            for k in (1, 2, 3):
                assert float((a[k][:, 2] - b[k][:, 2]).abs().max()) < 2e-5 * float(a[k][:, 2].max()), k
                assert float((a[k][:, 0] - b[k][:, 0]).abs().max()) < 1e-5
    finally:
        sim.set_pupil_path("sweep")
        sim.close()


@pytest.mark.gpu
def test_step_with_geo_leaves_the_main_loop_untouched(static10, oracle_imat10, geo10):
    """AOM_OPT_GEO: the fused step runs the projection beside the sensor frame; the LS loop is bit-identical with and
    without it, and the command it leaves equals an explicit aom_do_control_geo on the same screens."""
    import copy
    import torch
    from ao_marl_b200.init import rtc as rtc_b
    from ao_marl_b200.lib import Simulator
    t = copy.copy(static10)
    t.imat = oracle_imat10
    t.cmat = rtc_b.cmat_with_btt(t.imat, t.Btt, 5)
    seeds = np.array([7, 8], dtype=np.int64)
    out = {}
    for on in (False, True):
        sim = Simulator(t, 2, rl=None)
        try:
            sim.step_with_geo(on)
            sim.reset(seeds)
            for _ in range(4):
                sim.step(mode=2)
            out[on] = (sim.rows("COM", t.nactu).clone().cpu(), sim.rows("SLOPES", t.nslopes).clone().cpu())
            if on:
                g = sim.rows("GEO_COM", t.nactu).clone()
                v = sim.rows("GEO_VOLTS", t.nactu).clone()
                assert float(g.abs().max()) > 0 and torch.equal(g, v)
                g2 = sim.do_control_geo().clone()
                assert torch.equal(g, g2)
            sim.check_device()
        finally:
            sim.close()
    assert torch.equal(out[False][0], out[True][0]) and torch.equal(out[False][1], out[True][1])


@pytest.mark.gpu
def test_geo_supervisor_surface():
    """RlSupervisor.next_part_one_geo / rtc index 1 / target index 1 on the 10x10 'geo' layout, E == 1."""
    from ao_marl_b200.env.ao_env import AoEnv
    from ao_marl_b200.env.config_rl import Config
    cfg = Config(parameters_telescope="production_sh_10x10_2m.py", n_zernike_start_end=[0, 80],
                 n_reverse_filtered_from_cmat=5)
    env = AoEnv(cfg, n_env=1, world_size=3, initial_seed=1234)
    sup = env.supervisor
    try:
        assert sup.geo_index == 1
        env.reset()
        for _ in range(3):
            sup.next_part_one(geo=True)
            sup.next_part_two(np.zeros(env.action_size, np.float32), linear_control=True)
        com = sup.rtc.get_command(1)
        assert isinstance(com, np.ndarray) and com.shape == (90,) and np.abs(com).max() > 0
        assert np.array_equal(sup.rtc.get_voltages(1), com)
        sup.target.comp_tar_image(1)
        sup.target.comp_tar_image(0)
        se1, le1, var1, _ = sup.target.get_strehl(1)
        se0, le0, var0, _ = sup.target.get_strehl(0)
        assert 0.0 < var1 < var0            # the fitting-only residual bounds the closed loop from below
        assert se1 > se0
        # the reference's env surface: linear_step / rl_step with geometric_do_control (ao_env.py:873-937)
        env.linear_step(geometric_do_control=True)
        r, done, info, r_geo = env.rl_step(np.zeros(env.action_size, np.float32), geometric_do_control=True)
        assert isinstance(r, float) and isinstance(r_geo, float) and 0.0 < r_geo <= 1.0
        assert len(env.rl_step(np.zeros(env.action_size, np.float32))) == 3
    finally:
        sup.sim.close()


@pytest.mark.gpu
def test_40x40_geo_fit():
    """Full-size lattice (1284 + 2 actuators, 644^2 pupil, 3 layers): the CUDA projection against LSQR on the
    materialised phase."""
    import torch
    from ao_marl_b200 import tables
    from ao_marl_b200.config import load_config_from_file
    from ao_marl_b200.init import geo
    from ao_marl_b200.lib import Simulator
    from oracle import geo as ogeo
    t = tables.build_static(load_config_from_file("production_sh_40x40_8m_3layers.py"))
    geo.build_geo(t)
    tab = t.as_oracle_dict()
    IFt, idx = ogeo.influence_matrix(tab)
    sim = Simulator(t, 2, rl=None)
    try:
        sim.reset(np.array([301, 302], dtype=np.int64))
        sim.move_atmos()
        phase = sim.raytrace_wfs(atmos=True, dms=False).cpu().numpy().astype(np.float64)
        com = sim.do_control_geo().cpu().numpy().astype(np.float64)
        sim.apply_control_geo()
        s_geo = sim.comp_strehl(1.65, atmos=True, dms=True, geo=True).cpu().numpy()
        sim.check_device()
        for e in range(2):
            _, res_o = ogeo.geo_command(IFt, idx, phase[e])
            phi = phase[e].ravel()[idx]
            res = (phi - phi.mean()) + IFt @ com[e]
            assert abs(res.var() - res_o.var()) < 1e-3 * res_o.var(), (res.var(), res_o.var())
            assert res.var() < 0.02 * phi.var()
            assert abs(s_geo[e, 2] - res.var()) < 2e-3 * res.var()
    finally:
        sim.close()


@pytest.mark.gpu
def test_step_with_the_turbulence_advanced_by_the_caller(static10, oracle_imat10):
    """aom_step(mode | AOM_STEP_ATMOS_DONE) after an explicit aom_move_atmos == plain aom_step, bit for bit."""
    import copy
    import torch
    from ao_marl_b200.init import rtc as rtc_b
    from ao_marl_b200.lib import Simulator
    t = copy.copy(static10)
    t.imat = oracle_imat10
    t.cmat = rtc_b.cmat_with_btt(t.imat, t.Btt, 5)
    seeds = np.array([31, 32, 33], dtype=np.int64)
    a, b = Simulator(t, 3, rl=None), Simulator(t, 3, rl=None)
    try:
        a.reset(seeds)
        b.reset(seeds)
        side = torch.cuda.Stream()
        for _ in range(4):
            a.step(mode=2)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                b.move_atmos()
            torch.cuda.current_stream().wait_stream(side)
            b.step(mode=2, atmos_done=True)
        for name, n in (("SLOPES", t.nslopes), ("COM", t.nactu), ("VOLTS", t.nactu)):
            assert torch.equal(a.rows(name, n), b.rows(name, n)), name
        a.check_device()
        b.check_device()
    finally:
        a.close()
        b.close()
