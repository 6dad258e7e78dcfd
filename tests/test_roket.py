"""ROKET error breakdown (reference guardians/roket_generalized_rl.py:171-376, driven as in
src/error_budget/error_budget_multiple_agents.py:306-343).  The arithmetic behind the reference's calls is sutra's, so the
checks are the properties ROKET itself is validated with: contributors that must vanish do, the loop keeps running from
the state it had, and the Strehl rebuilt from the breakdown agrees with the measured long-exposure Strehl."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _env(n_env, noise=None):
    from ao_marl_b200.env.ao_env import AoEnv
    from ao_marl_b200.env.config_rl import Config
    cfg = Config(parameters_telescope="production_sh_10x10_2m.py", n_zernike_start_end=[0, 80], n_reverse_filtered_from_cmat=5)
    return AoEnv(cfg, n_env=n_env, world_size=3, initial_seed=77, roket=True)


def _run(env, n_total, n_pre):
    sup = env.supervisor
    sup.init_config_roket(N_total=n_total, N_preloop=n_pre, agent=None, gamma=1.0)
    env.reset()
    a = np.zeros(env.action_size, np.float32)
    for step in range(n_total):
        env.rl_step(action=a, linear_control=True, apply_control=False, compute_tar_psf=False)
        sup.do_error_breakdown(a)
        env.linear_step()
        if step + 1 == n_pre:
            sup.target.reset_strehl(0)
    return sup


def test_breakdown_leaves_the_loop_untouched_and_vanishing_terms_vanish():
    """Same seeds with and without the breakdown: identical commands; without detector noise the noise contributor is
    zero; on axis the tomographic one is zero; without agents zeta is zero."""
    from ao_marl_b200.env.ao_env import AoEnv
    from ao_marl_b200.env.config_rl import Config
    env = _env(2)
    try:
        sup = _run(env, 30, 10)
        com_roket = sup.sim.rows("COM", sup.nactus).clone()
        assert float(sup.noise_com.abs().max()) == 0.0                   # the 10x10 file has no detector noise
        assert float(sup.tomo_com.abs().max()) == 0.0
        assert float(sup.zeta_contributor.abs().max()) == 0.0
        for name in ("trunc_com", "alias_wfs_com", "H_com", "bp_com", "mod_com"):
            assert float(getattr(sup, name)[10:].abs().max()) > 0.0, name
        assert float(sup.fit[10:].min()) > 0.0
    finally:
        env.sim.close()
    cfg = Config(parameters_telescope="production_sh_10x10_2m.py", n_zernike_start_end=[0, 80], n_reverse_filtered_from_cmat=5)
    ref = AoEnv(cfg, n_env=2, world_size=3, initial_seed=77)
    try:
        ref.reset()
        a = np.zeros(ref.action_size, np.float32)
        for _ in range(30):
            ref.rl_step(action=a, linear_control=True, apply_control=True)
            ref.linear_step()
        com_plain = ref.sim.rows("COM", com_roket.shape[1]).clone()
    finally:
        ref.sim.close()
    assert torch.equal(com_roket, com_plain)


def test_strehl_from_the_breakdown_matches_the_measured_one():
    """ROKET's own validation: exp(-(2 pi / lambda)^2 (sum of contributor variances + fitting)) against the long-exposure
    Strehl of the same frames.  The breakdown neglects cross terms and the Marechal form is approximate: 15 % relative."""
    env = _env(3)
    try:
        sup = _run(env, 700, 200)
        sr2 = sup.strehl_from_breakdown().cpu().numpy()
        le = np.asarray([float(x) for x in sup.target.get_strehl(0)[1].cpu()])
        cov, cor = sup.cov_cor()
        var = torch.diagonal(cov, dim1=1, dim2=2).cpu().numpy()
        meas_var = np.asarray([float(x) for x in sup.target.get_strehl(0)[3].cpu()])
        print("SR measured", le, "SR from breakdown", sr2, "fitting", sup.fit[200:].mean(0).cpu().numpy())
        print("measured mean phase variance [um^2]", meas_var, "breakdown total", var.sum(1) + sup.fit[200:].mean(0).cpu().numpy())
        print("variances [noise, non linearity, aliasing, filtered, bandwidth, tomography]", np.array2string(var, precision=5, max_line_width=200))
        e_cmd = (sup.com[200:700] - sup.mod_com[200:700] - sup.H_com[200:700]) @ sup.P.T       # actual minus ideal command, modal
        v_e = ((e_cmd * e_cmd).mean(0) - e_cmd.mean(0) ** 2)
        print("variance of (actual - ideal command), all modes", v_e.sum(-1).cpu().numpy(), "tip-tilt only", v_e[:, -2:].sum(-1).cpu().numpy())
        # the identity behind ROKET: actual minus commanded-modes command = sum of the loop-filtered contributors
        lhs = (sup.com[200:700] - sup.mod_com[200:700]) @ sup.P.T
        rhs = (sup.noise_com[200:700] + sup.trunc_com[200:700] + sup.alias_wfs_com[200:700] + sup.bp_com[200:700] + sup.tomo_com[200:700]) @ sup.P.T
        d = lhs - rhs
        rel = (d.var(0, unbiased=False).sum(-1) / lhs.var(0, unbiased=False).sum(-1)).cpu().numpy()
        print("unexplained fraction of the command-error variance", rel)
        assert np.all(rel < 0.05), rel
        bp = sup.bp_com[200:700] @ sup.P.T
        v_bp = ((bp * bp).mean(0) - bp.mean(0) ** 2)
        print("bandwidth variance, tip-tilt only", v_bp[:, -2:].sum(-1).cpu().numpy(), "g", sup.g, "delay", sup.delay)
        assert np.all(le > 0.2)
        assert np.all(np.abs(sr2 / le - 1) < 0.15), (sr2, le)
        assert np.all(var[:, 4] > var[:, 0])                              # bandwidth dominates a noise-free 10x10 loop
        assert cor.shape == (3, 6, 6) and float(cor.abs().max()) <= 1.0 + 1e-5
        import tempfile, os
        with tempfile.TemporaryDirectory() as d:
            sup.save(os.path.join(d, "roket.npz"))
            z = np.load(os.path.join(d, "roket.npz"))
            assert z["bandwidth"].shape == (3, sup.nactus, 500) and z["cov"].shape == (3, 6, 6)
    finally:
        env.sim.close()


def test_geometric_slopes_kernel_against_oracle():
    """aom_do_centroids_geom against oracle.aoframe.slopes_geom on the same phase."""
    from oracle import aoframe as af
    from ao_marl_b200 import tables
    from ao_marl_b200.config import load_config_from_file
    from ao_marl_b200.lib import Simulator
    t = tables.build_static(load_config_from_file("production_sh_10x10_2m.py"))
    sim = Simulator(t, 2, rl=None)
    try:
        sim.reset(np.array([5, 6], dtype=np.int64))
        r = np.random.default_rng(1)
        sim.set_dm_volts(torch.as_tensor((r.standard_normal((2, t.nactu)) * 0.5).astype(np.float32), device="cuda"))
        sim.move_atmos()
        sim.do_centroids_geom()
        got = sim.rows("SLOPES", t.nslopes).cpu().numpy()
        ph = sim.raytrace_wfs().cpu().numpy()
        tab = t.as_oracle_dict()
        w = tab["wfs"]
        for e in range(2):
            want = af.slopes_geom(ph[e], tab["mpupil"], w["tile_origin"], w["pdiam"], w["fluxPerSub"], float(t.p_wfs._subapd))
            assert np.abs(got[e] - want).max() < 1e-4 * np.abs(want).max()
        sim.check_device()
    finally:
        sim.close()
