"""Parity of the CUDA path (through the C ABI) against the numpy oracle on identical seeds.

Tolerances: integers (photon counts, ring offsets) bit-exact; float32 slopes / commands / rewards
rel 1e-4 of the oracle's scale (BASELINE.json north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def sim10(static10, torch):
    from ao_marl_b200.lib import Simulator
    sim = Simulator(static10, 4, rl=None)
    yield sim
    sim.close()


GEMM_TOL = {"simt": 1e-5, "tcgen05": 5e-5}     # the tensor core truncates its float32 accumulator


def test_gemm_tn(sim10, torch):
    from ao_marl_b200.lib import LD
    g = torch.Generator(device="cuda").manual_seed(0)
    try:
        for path in ("simt", "tcgen05"):
            sim10.set_gemm_path(path)
            for (M, N, K) in ((1, 5, 7), (4, 648, 1957), (300, 90, 128), (257, 129, 100)):
                A = torch.zeros((M, LD(K)), device="cuda")
                Bm = torch.zeros((N, LD(K)), device="cuda")
                A[:, :K] = torch.randn((M, K), device="cuda", generator=g)
                Bm[:, :K] = torch.randn((N, K), device="cuda", generator=g)
                bias = torch.randn(N, device="cuda", generator=g)
                C = sim10.gemm_tn(A, Bm, bias=bias, relu=True)
                sim10.check_device()
                ref = torch.relu(A.double() @ Bm.double().T + bias.double())
                assert relerr(C[:, :N].cpu().numpy(), ref.cpu().numpy()) < GEMM_TOL[path], (path, M, N, K)
                assert float(C[:, N:].abs().max()) == 0.0 if C.shape[1] > N else True
    finally:
        sim10.set_gemm_path("tcgen05")


def test_gemm_paths(sim10, torch):
    """tcgen05 (3 x TF32 split) GEMM against float64 and against the FFMA kernel, shapes of the 40x40 step."""
    from ao_marl_b200.lib import LD
    g = torch.Generator(device="cuda").manual_seed(1)
    try:
        # (4096, 1286, 200) / (4000, 1283, 96): M large enough for the 144-column tile to be selected (one CTA wave)
        for (M, N, K) in ((256, 648, 1957), (130, 1286, 2400), (512, 1283, 1286), (3, 60, 256), (128, 128, 16),
                          (4096, 1286, 200), (4000, 1283, 96)):
            A = torch.zeros((M, LD(K)), device="cuda")
            Bm = torch.zeros((N, LD(K)), device="cuda")
            A[:, :K] = torch.randn((M, K), device="cuda", generator=g) * 3
            Bm[:, :K] = torch.randn((N, K), device="cuda", generator=g)
            ref = (A.double() @ Bm.double().T).cpu().numpy()
            out = {}
            for path in ("tcgen05", "simt"):
                sim10.set_gemm_path(path)
                out[path] = sim10.gemm_tn(A, Bm)[:, :N].cpu().numpy()
                sim10.check_device()
            assert relerr(out["simt"], ref) < GEMM_TOL["simt"]
            assert relerr(out["tcgen05"], ref) < GEMM_TOL["tcgen05"], (M, N, K)
    finally:
        sim10.set_gemm_path("tcgen05")


def test_pixel_noise_bit_exact(sim10, torch):
    from oracle import rng
    r = np.random.default_rng(3)
    lam = (r.random(200000) ** 3 * 400).astype(np.float32)
    lam[:100] = 0
    for noise in (0.0, 3.0):
        out = sim10.pixel_noise(torch.as_tensor(lam, device="cuda"), noise, 2 ** 40 + 17, 9, 1).cpu().numpy()
        ref = rng.wfs_pixel_noise(2 ** 40 + 17, 1, 9, lam, noise)
        assert np.array_equal(out, ref)


def _oracle_atmos(tab, seed):
    from oracle import aoframe
    a = aoframe.OracleAtmos(tab, seed)
    a.reset(seed)
    return a


def _logical_screen(sim, layer, env):
    n = int(sim.cfg.screen_dim[layer])
    scr = sim.buffer("SCREEN", layer).view(sim.n_env, n, n)[env]
    ox = int(sim.buffer("RING_OX", layer)[env])
    oy = int(sim.buffer("RING_OY", layer)[env])
    return np.roll(scr.cpu().numpy(), (-oy, -ox), axis=(0, 1)), ox, oy


def test_turbulence_extrusion(sim10, oracle_tab10, torch):
    seeds = np.array([1234, 1235, 77, 2 ** 33 + 5], dtype=np.int64)
    sim10.reset(seeds)
    torch.cuda.synchronize()
    oracles = [_oracle_atmos(oracle_tab10, s) for s in seeds[:2]]
    for e, o in enumerate(oracles):
        got, ox, oy = _logical_screen(sim10, 0, e)
        assert relerr(got, o.screens[0]) < RTOL
    for _ in range(7):
        sim10.move_atmos()
        for o in oracles:
            o.move()
    torch.cuda.synchronize()
    for e, o in enumerate(oracles):
        got, ox, oy = _logical_screen(sim10, 0, e)
        assert relerr(got, o.screens[0]) < RTOL
    # environments with different seeds differ
    a, _, _ = _logical_screen(sim10, 0, 0)
    b, _, _ = _logical_screen(sim10, 0, 2)
    assert np.abs(a - b).max() > 0.1 * np.abs(a).max()


def test_phase_and_frame(sim10, oracle_tab10, oracle_imat10, static10, torch):
    """raytrace + mirrors + SH image + COG against the oracle on a turbulent screen with non-zero volts."""
    from oracle import loop
    seeds = np.array([1234, 1235, 77, 99], dtype=np.int64)
    sim10.reset(seeds)
    r = np.random.default_rng(5)
    volts = (r.standard_normal((4, static10.nactu)) * 20).astype(np.float32)
    volts[:, -2:] *= 5
    sim10.set_dm_volts(torch.as_tensor(volts, device="cuda"))
    for _ in range(3):
        sim10.move_atmos()
    envs = []
    for e in range(2):
        o = loop.OracleEnv(oracle_tab10, np.zeros((static10.nactu, static10.nslopes), np.float32), seed=int(seeds[e]))
        o.reset(int(seeds[e]))
        o.volts = volts[e].copy()
        for _ in range(3):
            o.atm.move()
        envs.append(o)
    ph = sim10.raytrace_wfs().cpu().numpy()
    for e, o in enumerate(envs):
        ref = o.wfs_phase()
        assert relerr(ph[e], ref) < 2e-5
        assert relerr(sim10.raytrace_wfs(atmos=False)[e].cpu().numpy(), o.wfs_phase(atmos=False)) < 2e-5
    sim10.comp_wfs_image(keep_image=True, noise=-1.0)
    sim10.do_centroids()
    s = sim10.rows("SLOPES", static10.nslopes).cpu().numpy()
    cube = sim10.buffer("BINCUBE").view(4, static10.p_wfs._nvalid, 16, 16).cpu().numpy()
    for e, o in enumerate(envs):
        sref, cref = o.comp_wfs_image(noise=-1.0, keep=True)
        assert relerr(cube[e], cref) < 1e-4
        assert relerr(s[e], sref) < RTOL


def test_wfs_paths_agree(sim10, static10, torch):
    """Every tensor-core DFT path (tcgen05 default and the round-1 mma.sync kernels; fp16 split, 3 products per stage) against
    the float32 shared-memory FFT on the same frame."""
    seeds = np.array([21, 22, 23, 24], dtype=np.int64)
    sim10.reset(seeds)
    r = np.random.default_rng(9)
    volts = (r.standard_normal((4, static10.nactu)) * 10).astype(np.float32)
    sim10.set_dm_volts(torch.as_tensor(volts, device="cuda"))
    out = {}
    try:
        for path in ("simt", "tensor", "tensor_fast", "tensor_reg", "umma", "umma_fast"):
            sim10.set_wfs_path(path)
            if path in ("umma", "umma_fast"):
                assert sim10.wfs_kernel() == "wfs_frame_umma_kernel", sim10.lib.aom_last_error(sim10._ctx)
            if path in ("tensor", "tensor_fast"):
                assert sim10.wfs_kernel() == "wfs_frame_tma_kernel", sim10.lib.aom_last_error(sim10._ctx)
            sim10.comp_wfs_image(keep_image=True, noise=-1.0)
            sim10.do_centroids()
            out[path] = (sim10.rows("SLOPES", static10.nslopes).cpu().numpy().copy(),
                         sim10.buffer("BINCUBE").cpu().numpy().copy())
    finally:
        sim10.set_wfs_path(sim10.DEFAULT_WFS_PATH)
    assert relerr(out["tensor"][1], out["simt"][1]) < 2e-5
    assert relerr(out["tensor"][0], out["simt"][0]) < 2e-5
    assert relerr(out["tensor_fast"][0], out["simt"][0]) < 5e-4
    for path in ("tensor_reg", "umma"):
        assert relerr(out[path][1], out["simt"][1]) < 2e-5
        assert relerr(out[path][0], out["simt"][0]) < 2e-5
    assert relerr(out["umma_fast"][0], out["simt"][0]) < 5e-4


def test_wfs_staged_kernel_over_the_seam(sim10, static10, torch):
    """The TMA-staged kernels against the float32 FFT kernel while the torus seam sweeps through the pupil
    (tiles that straddle it take the plain-load fill), with and without the image / normalisation path."""
    seeds = np.array([31, 32, 33, 34], dtype=np.int64)
    sim10.reset(seeds)
    r = np.random.default_rng(10)
    volts = (r.standard_normal((4, static10.nactu)) * 10).astype(np.float32)
    sim10.set_dm_volts(torch.as_tensor(volts, device="cuda"))
    n = static10.nslopes
    try:
        for it in range(24):
            for _ in range(3):
                sim10.move_atmos()
            res = {}
            for path in ("simt", "tensor", "umma", "umma_ws"):
                sim10.set_wfs_path(path)
                sim10.comp_wfs_image(keep_image=(it % 2 == 0), noise=-1.0)
                sim10.do_centroids()
                res[path] = sim10.rows("SLOPES", n).cpu().numpy().copy()
            sim10.check_device()
            assert relerr(res["tensor"], res["simt"]) < 2e-5, it
            assert relerr(res["umma"], res["simt"]) < 2e-5, it
            assert np.array_equal(res["umma_ws"], res["umma"]), it      # same arithmetic on specialised warps
    finally:
        sim10.set_wfs_path(sim10.DEFAULT_WFS_PATH)


def test_specialised_warps_on_ragged_batches(static10, torch):
    """wfs_frame_ws_kernel (the default) against the fused tcgen05 kernel, bit for bit, for batch sizes whose last CTA
    holds a number of work items that is not a multiple of the eight subapertures of a tensor-core tile, and for a
    single environment."""
    from ao_marl_b200.lib import Simulator
    r = np.random.default_rng(3)
    for n_env in (1, 3, 5, 13):
        sim = Simulator(static10, n_env, rl=None)
        try:
            sim.reset(np.arange(1, n_env + 1, dtype=np.int64))
            sim.set_dm_volts(torch.as_tensor((r.standard_normal((n_env, static10.nactu)) * 0.3).astype(np.float32), device="cuda"))
            for _ in range(3):
                sim.move_atmos()
            res = {}
            for path in ("umma", "umma_ws"):
                sim.set_wfs_path(path)
                sim.comp_wfs_image(noise=-1.0)          # the same turbulence frame for both
                sim.do_centroids()
                res[path] = sim.rows("SLOPES", static10.nslopes).cpu().numpy().copy()
            assert sim.wfs_kernel() == "wfs_frame_ws_kernel", sim.lib.aom_last_error(sim._ctx)
            sim.check_device()
            assert np.abs(res["umma"]).max() > 1e-3
            assert np.array_equal(res["umma_ws"], res["umma"]), n_env
        finally:
            sim.close()


def test_noisy_frame_counts(sim10, oracle_tab10, static10, torch):
    """Photon noise: the oracle's sampler applied to the GPU's noise-free image reproduces the GPU's
    noisy image exactly (same Philox stream, same integer Poisson draws) -- for every sensor-kernel generation,
    each of which lays the detector pixels out differently over the lanes."""
    from oracle import aoframe
    seeds = np.array([11, 12, 13, 14], dtype=np.int64)
    try:
        for path in ("umma", "tensor", "tensor_reg", "simt"):
            sim10.set_wfs_path(path)
            sim10.reset(seeds)
            sim10.comp_wfs_image(keep_image=True, noise=-1.0)       # frame 0
            clean = sim10.buffer("BINCUBE").clone().view(4, -1).cpu().numpy()
            sim10.reset(seeds)
            sim10.comp_wfs_image(keep_image=True, noise=3.0)        # frame 0 again, same screens
            noisy = sim10.buffer("BINCUBE").view(4, -1).cpu().numpy()
            sim10.check_device()
            for e in range(4):
                ref = aoframe.sh_noise(clean[e], 3.0, int(seeds[e]), static10.wfs_index, 0)
                assert np.array_equal(noisy[e], ref), path
    finally:
        sim10.set_wfs_path(sim10.DEFAULT_WFS_PATH)


def test_imat_matches_oracle(static10, oracle_imat10, torch):
    from ao_marl_b200 import calibration
    D = calibration.measure_imat(static10)
    assert D.shape == oracle_imat10.shape
    assert relerr(D, oracle_imat10) < RTOL


@pytest.fixture(scope="module")
def system10(static10, oracle_imat10, torch):
    """Full 10x10 system (2 agents: 80 modes + tip-tilt) built on the GPU-measured imat."""
    import copy
    from ao_marl_b200 import calibration
    from ao_marl_b200.init import rtc as rtc_b
    from ao_marl_b200.lib import Simulator
    from ao_marl_b200.rl.layout import RLLayout
    t = copy.copy(static10)
    t.imat = calibration.measure_imat(static10)
    t.cmat = rtc_b.cmat_with_btt(t.imat, t.Btt, 5)
    rl = RLLayout(t.Btt.shape[1], dict(parameters_telescope="production_sh_10x10_2m.py", n_zernike_start_end=[0, 80],
                                       n_reverse_filtered_from_cmat=5), None, world_size=3, seed=3)
    import torch as th
    th.manual_seed(20240607)     # reproducible heads: an unseeded draw once gave a 1e-3 action outlier in 1 of ~10 runs
    with th.no_grad():           # non-trivial heads so that actions are not all zero-mean/unit-std
        for p in rl.policies:
            p.mean_linear.weight.normal_(0, 0.05)
            p.log_std_linear.weight.normal_(0, 0.05)
            p.log_std_linear.bias.fill_(-1.0)
    sim = Simulator(t, 3, rl)
    yield sim, t, rl
    sim.close()


def test_closed_loop_against_oracle(system10, oracle_tab10, torch):
    from oracle import loop
    sim, t, rl = system10
    seeds = np.array([1234, 4321, 999], dtype=np.int64)
    sim.reset(seeds)
    envs = [loop.OracleEnv(oracle_tab10, t.cmat, t.Btt, t.P, rl, seed=int(s)) for s in seeds[:2]]
    for o, s in zip(envs, seeds):
        o.reset(int(s))
    # initial linear step (AoEnv.reset ends with one, ao_env.py:354)
    sim.state_begin(); sim.move_atmos(); sim.comp_wfs_image(); sim.do_centroids(); sim.do_control(); sim.state_end()
    states = [o.linear_step() for o in envs]
    st = sim.rows("STATE", rl.state_dim).cpu().numpy()
    for e in range(2):
        assert relerr(st[e], states[e]) < RTOL
    worst = {}
    for it in range(12):
        sim.step(mode=0)
        act = sim.rows("ACTION", rl.action_dim).cpu().numpy()
        rew = sim.buffer("REWARD").view(3, rl.n_agents).cpu().numpy()
        st = sim.rows("STATE", rl.state_dim).cpu().numpy()
        com = sim.rows("COM", t.nactu).cpu().numpy()
        sl = sim.rows("SLOPES", t.nslopes).cpu().numpy()
        for e, o in enumerate(envs):
            a, _ = o.actors(states[e])
            worst["actions"] = max(worst.get("actions", 0), relerr(act[e], a))
            states[e], r = o.env_step(act[e])     # feed the GPU's action so that errors do not compound through tanh
            for k, (x, y) in dict(slopes=(sl[e], o.slopes), commands=(com[e], o.com), rewards=(rew[e], r),
                                  state=(st[e], states[e])).items():
                worst[k] = max(worst.get(k, 0), relerr(x, y))
    print("10x10 closed loop, worst rel errors over 12 steps:", worst)
    for k, v in worst.items():
        assert v < RTOL, (k, worst)


def test_delay_impulse(system10, torch):
    """An action impulse at step t shows in d_err at frame t+2 for delay 1 (train_rpc.py:620-630)."""
    sim, t, rl = system10
    sim.reset(np.array([5, 5, 5], dtype=np.int64))
    assert int(round(t.delay)) == 1
    sim.set_loop(False)                      # integrator frozen: only the injected action moves the mirror
    try:
        errs = []
        for it in range(5):
            a = torch.zeros((3, rl.action_dim), device="cuda")
            if it == 1:
                a[1, 0] = 5.0                # env 1 only
            sim.rows("ACTION", rl.action_dim).copy_(a)
            sim.step(mode=1)
            e = sim.rows("ERR", t.nactu)
            errs.append(float((e[1] - e[0]).abs().max()))
        assert errs[0] == 0.0 and errs[1] == 0.0
        assert errs[2] > 0.0                 # frame t+1 in 0-based counting of frames after the action step
    finally:
        sim.set_loop(True)


def test_delay_zero_impulse(static10, oracle_imat10, torch):
    """Same impulse with a delay-0 controller (the *_d0_noise parameter files): it shows one frame earlier."""
    import copy
    from ao_marl_b200.init import rtc as rtc_b
    from ao_marl_b200.lib import Simulator
    from ao_marl_b200.rl.layout import RLLayout
    t = copy.copy(static10)
    t.delay = 0.0
    t.imat = oracle_imat10
    t.cmat = rtc_b.cmat_with_btt(t.imat, t.Btt, 5)
    rl = RLLayout(t.Btt.shape[1], dict(parameters_telescope="production_sh_10x10_2m.py", n_zernike_start_end=[0, 80],
                                       n_reverse_filtered_from_cmat=5), None, world_size=3, seed=3)
    sim = Simulator(t, 3, rl)
    try:
        assert sim.cfg.delay == 0
        sim.reset(np.array([5, 5, 5], dtype=np.int64))
        sim.set_loop(False)
        errs = []
        for it in range(4):
            a = torch.zeros((3, rl.action_dim), device="cuda")
            if it == 1:
                a[1, 0] = 5.0
            sim.rows("ACTION", rl.action_dim).copy_(a)
            sim.step(mode=1)
            e = sim.rows("ERR", t.nactu)
            errs.append(float((e[1] - e[0]).abs().max()))
        assert errs[0] == 0.0
        assert errs[1] > 0.0                 # the frame of the action step itself
        sim.check_device()
    finally:
        sim.close()


# ------------------------------------------------------------------------------------------------
# Full-size configuration (production_sh_40x40_8m_3layers, 1200 subapertures, 1286 actuators, 3 layers,
# 43 windowed agents): the oracle needs minutes per frame here, so the checks are size-independent properties.
@pytest.fixture(scope="module")
def system40(torch):
    from ao_marl_b200.system import build_system
    sim, t, rl = build_system("production_sh_40x40_8m_3layers.py", 6,
                              env_rl=dict(n_zernike_start_end=[0, 1260], window_n_zernike=20,
                                          include_tip_tilt_windowed=True, n_reverse_filtered_from_cmat=5,
                                          delayed_assignment=2), world_size=44, seed=0)
    yield sim, t, rl
    sim.close()


def test_40x40_flat_and_tilt(system40, torch):
    """Flat wavefront -> zero slopes; a tip-tilt command -> the same slope on every subaperture, linear in the
    command; all on the staged kernel."""
    sim, t, rl = system40
    assert sim.wfs_kernel() == "wfs_frame_ws_kernel", sim.lib.aom_last_error(sim._ctx)
    nv = t.p_wfs._nvalid
    sim.reset(np.arange(6, dtype=np.int64) + 500)
    sim.reset_dm()
    sim.comp_wfs_image(atmos=False, dms=True, noise=-1.0)
    sim.do_centroids()
    s = sim.rows("SLOPES", t.nslopes).cpu().numpy()
    assert np.abs(s).max() < 2e-6
    volts = np.zeros((6, t.nactu), np.float32)
    volts[:, -2] = [50.0, 100.0, 200.0, 0.0, 0.0, 50.0]       # ~0.025 .. 0.1 arcsec on a 0.258 arcsec pixel
    volts[:, -1] = [0.0, 0.0, 0.0, 50.0, 100.0, 50.0]
    sim.set_dm_volts(torch.as_tensor(volts, device="cuda"))
    sim.comp_wfs_image(atmos=False, dms=True, noise=-1.0)
    sim.do_centroids()
    s = sim.rows("SLOPES", t.nslopes).cpu().numpy()
    sx, sy = s[:, :nv], s[:, nv:]
    full = np.asarray(t.p_wfs._fluxPerSub_list) > 0.999          # fully illuminated subapertures
    for e in range(6):
        for comp in (sx[e][full], sy[e][full]):
            assert np.abs(comp - comp.mean()).max() < 6e-2 * max(np.abs(s[e]).max(), 1e-9) + 1e-6
    m = np.hypot(sx[:, full].mean(axis=1), sy[:, full].mean(axis=1))
    assert m[0] > 5e-3                                                            # a real signal
    assert abs(m[1] / m[0] - 2.0) < 1e-2 and abs(m[2] / m[0] - 4.0) < 4e-2      # linear in the command
    assert abs(m[4] / m[3] - 2.0) < 1e-2
    # the two mirror axes are orthogonal on the sensor
    d0 = np.array([sx[0, full].mean(), sy[0, full].mean()])
    d3 = np.array([sx[3, full].mean(), sy[3, full].mean()])
    assert abs(d0 @ d3) < 1e-3 * np.linalg.norm(d0) * np.linalg.norm(d3)


def test_40x40_kernel_generations_agree(system40, torch):
    """Three layers + random mirror shape: staged tensor kernel == register tensor kernel == float32 FFT kernel,
    also after the torus seam has moved into the pupil."""
    sim, t, rl = system40
    sim.reset(np.arange(6, dtype=np.int64) + 900)
    r = np.random.default_rng(3)
    volts = (r.standard_normal((6, t.nactu)) * 0.3).astype(np.float32)
    sim.set_dm_volts(torch.as_tensor(volts, device="cuda"))
    try:
        for it in range(3):
            for _ in range(1 + 40 * it):
                sim.move_atmos()
            res = {}
            for path in ("simt", "tensor", "tensor_reg", "umma", "umma_ws"):
                sim.set_wfs_path(path)
                sim.comp_wfs_image(noise=-1.0)
                sim.do_centroids()
                res[path] = sim.rows("SLOPES", t.nslopes).cpu().numpy().copy()
            sim.check_device()
            assert np.abs(res["simt"]).max() > 1e-3
            assert relerr(res["tensor"], res["simt"]) < 2e-5, it
            assert relerr(res["tensor_reg"], res["simt"]) < 2e-5, it
            assert relerr(res["umma"], res["simt"]) < 2e-5, it
            assert np.array_equal(res["umma_ws"], res["umma"]), it      # same arithmetic on specialised warps
    finally:
        sim.set_wfs_path(sim.DEFAULT_WFS_PATH)


def test_40x40_closed_loop_properties(system40, torch):
    """Integrator-only loop converges (residual slopes far below open loop); the modal projector applied by
    rl_control with a zero action is idempotent; rewards are negative and finite for every agent."""
    sim, t, rl = system40
    sim.reset(np.arange(6, dtype=np.int64) + 1234)
    sim.comp_wfs_image(noise=-1.0)
    sim.do_centroids()
    open_rms = float(sim.rows("SLOPES", t.nslopes).square().mean().sqrt())
    for _ in range(40):
        sim.step(mode=2)
    closed_rms = float(sim.rows("SLOPES", t.nslopes).square().mean().sqrt())
    assert closed_rms < 0.5 * open_rms, (open_rms, closed_rms)
    zero = torch.zeros((6, rl.action_dim), device="cuda")
    sim.rl_control(zero)
    c1 = sim.rows("COM", t.nactu).clone()
    sim.rl_control(zero)
    c2 = sim.rows("COM", t.nactu)
    assert float((c1 - c2).abs().max()) < 1e-4 * float(c1.abs().max())
    sim.step(mode=0)
    rw = sim.buffer("REWARD").view(6, rl.n_agents)
    assert torch.isfinite(rw).all() and float(rw.max()) <= 0.0
    st = sim.rows("STATE", rl.state_dim)
    assert torch.isfinite(st).all()
    sim.check_device()


def test_batched_trainer_learns_on_device(torch):
    """AoEnv + BatchedTrainer + BatchedSAC on the 10x10 system: pooled replay fills with E transitions per step
    after the credit-assignment window, updates run, and the actors uploaded into the simulator reproduce the
    learner's own forward pass (aom_actor_forward == torch policy, eval mode)."""
    from ao_marl_b200.env.ao_env import AoEnv
    from ao_marl_b200.env.config_rl import Config
    from ao_marl_b200.env.trainer import BatchedTrainer
    from ao_marl_b200.rl.sac import BatchedSAC
    cfg = Config(parameters_telescope="production_sh_10x10_2m.py", n_zernike_start_end=[0, 80],
                 n_reverse_filtered_from_cmat=5, delayed_assignment=2)
    cfg.sac.update(batch_size=64, hidden_size_critic=64, num_layers_critic=2)
    E = 8
    env = AoEnv(cfg, n_env=E, world_size=3, initial_seed=1234)
    rl = env.supervisor.rl
    learner = BatchedSAC.from_layout(rl, device="cuda", seed=3, memory_size=4096)
    tr = BatchedTrainer(env, learner, seed=1234, updates_per_episode=6)
    steps = 24
    r_total, stats = tr.train_episode(steps=steps)
    depth = cfg.env_rl["delayed_assignment"] + 1
    assert len(learner.memory) == E * (steps - depth)
    assert stats is not None and all(torch.isfinite(v).all() for v in stats.values())
    assert np.isfinite(r_total) and r_total < 0
    # stored transitions are consistent: next state of a slot differs from its state, rewards negative
    assert float(learner.memory.r[:, :len(learner.memory)].max()) <= 0
    # uploaded actors == learner forward
    env.sim.actor_forward(True)
    a_sim = env.sim.rows("ACTION_MEAN", rl.action_dim).clone()
    s = env.sim.rows("STATE", rl.state_dim)
    with torch.no_grad():
        _, _, mean = learner.policy_sample(tr.states_per_agent(s))
    a_ref = torch.zeros_like(a_sim)
    for a in range(rl.n_agents):
        n = learner.act_dims[a]
        a_ref[:, tr._act[a, :n]] = mean[a, :, :n]
    assert float((a_sim - a_ref).abs().max()) < 2e-4
    env.sim.check_device()
    env.sim.close()


def test_denoiser_in_the_centroid_path(torch):
    """Config 4 (production_sh_40x40_8m_3layers_d0_noise: magnitude 9, 3 e- read noise): the reference's trained
    denoiser, fed from AOM_B_BINCUBE and handed back through aom_set_bincube, brings the centre-of-gravity slopes
    closer to the noise-free ones than the raw noisy image does."""
    from ao_marl_b200 import tables
    from ao_marl_b200.config import load_config_from_file
    from ao_marl_b200.denoiser import Autoencoder
    from ao_marl_b200.lib import Simulator
    t = tables.build_static(load_config_from_file("production_sh_40x40_8m_3layers_d0_noise.py"))
    sim = Simulator(t, 3, rl=None)
    try:
        assert abs(sim.cfg.noise - 3.0) < 1e-6 and 200 < sim.cfg.nphotons < 300
        ae = Autoencoder(dict(type="cnn_single_subaperture", path="autoencoder_M9_rms_3"), device="cuda")
        sim.reset(np.array([5, 6, 7], dtype=np.int64))
        n, nv = t.nslopes, t.p_wfs._nvalid
        sim.comp_wfs_image(noise=-1.0)
        sim.do_centroids()
        clean = sim.rows("SLOPES", n).clone()
        sim.comp_wfs_image(keep_image=True)              # configured noise: photon + 3 e- read noise
        sim.do_centroids()
        raw = sim.rows("SLOPES", n).clone()
        cube = sim.buffer("BINCUBE").view(3, nv, 256)
        den = ae.predict(cube.reshape(3 * nv, 16, 16)).reshape(3, nv, 256).contiguous()
        sim.set_bincube(den)
        sim.do_centroids()
        denoised = sim.rows("SLOPES", n).clone()
        e_raw = float((raw - clean).square().mean().sqrt())
        e_den = float((denoised - clean).square().mean().sqrt())
        assert torch.isfinite(denoised).all()
        assert e_den < 0.7 * e_raw, (e_raw, e_den)
        sim.check_device()
    finally:
        sim.close()


def test_40x40_closed_loop_statistics_against_reference(system40, torch):
    """The only pin against the real COMPASS simulator: the reference authors committed per-feature statistics of
    1000-frame integrator loops (state_normalization/*.pickle, exported to ao_marl_b200/data/normalization).  The
    closed-loop slope rms and the normalised state blocks of the batched simulator must sit in the same
    distribution (different RNG, so a distribution-level gate: SURVEY.md 8(c))."""
    from ao_marl_b200.rl.layout import load_normalization
    sim, t, rl = system40
    norm, _ = load_normalization("production_sh_40x40_8m_3layers.py")
    sim.reset(np.arange(6, dtype=np.int64) + 2000)
    for _ in range(80):
        sim.step(mode=2)
    sl, st = [], []
    for _ in range(160):
        sim.step(mode=2)
        sl.append(sim.rows("SLOPES", t.nslopes).std(dim=1).clone())
        st.append(sim.rows("STATE", rl.state_dim).clone())
    rms = float(torch.stack(sl).mean())
    ref = float(norm["wfs"]["std"].mean())
    assert abs(rms / ref - 1) < 0.3, (rms, ref)
    # state = (v2m . x - mean_ref) / std_ref per mode: with the reference's own statistics every block has to
    # come out roughly standard (rms over frames, environments and modes; TT and high orders included)
    S = torch.stack(st)                                        # [frames, E, state_dim]
    for key in ("dm_before_linear", "dm_residual"):
        a, b = rl.indices_of_state[key]
        blk = S[:, :, a:b]
        spread = float(blk.std(dim=(0, 1)).median())
        assert 0.5 < spread < 2.0, (key, spread)


def test_single_environment_keeps_the_reference_shapes(torch):
    """Drop-in surface with E == 1 (what ao_env.py / train_rpc.py see): numpy float32 arrays of the reference's shapes,
    the reference's error behaviour (rtcCompass.py:471-472 raises ValueError("Dimension mismatch")), and the same step
    sequencing as AoEnv.reset / rl_step / linear_step (ao_env.py:336-354, 871-939)."""
    from ao_marl_b200.env.ao_env import AoEnv
    from ao_marl_b200.env.config_rl import Config
    cfg = Config(parameters_telescope="production_sh_10x10_2m.py", n_zernike_start_end=[0, 80],
                 n_reverse_filtered_from_cmat=5)
    env = AoEnv(cfg, n_env=1, world_size=3, initial_seed=1234)
    sup = env.supervisor
    try:
        s = env.reset()
        assert isinstance(s, np.ndarray) and s.dtype == np.float32 and s.shape == (env.state_size,) == (4 * 82,)
        d = env.linear_step(return_dict=True)
        assert list(d) == ["dm_history_2", "dm_history_1", "dm_before_linear", "dm_residual"]
        assert all(v.shape == (82,) for v in d.values())
        slopes = sup.rtc.get_slopes(0)
        assert isinstance(slopes, np.ndarray) and slopes.shape == (128,) and slopes.dtype == np.float32
        assert sup.rtc.get_command(0).shape == (90,) and sup.rtc.get_err(0).shape == (90,)
        assert sup.rtc.get_voltages(0).shape == (90,)
        assert sup.modes2volts.shape == (90, 87) and sup.volts2modes.shape == (87, 90)
        assert sup.wfs.get_wfs_image(0).shape == (160, 160)
        assert sup.wfs.get_wfs_phase(0).shape == (164, 164)
        assert sup.atmos.get_atmos_layer(0).shape == (168, 168)
        se, le, var, avg = sup.target.get_strehl(0)
        assert all(isinstance(x, float) for x in (se, le, var, avg))
        with pytest.raises(ValueError):
            sup.rtc.set_command(0, np.zeros(91, np.float32))
        a = np.zeros(env.action_size, np.float32)
        r, done, info = env.rl_step(a)
        assert isinstance(r, float) and r <= 0.0 and done is False
        s2 = env.linear_step()
        assert s2.shape == s.shape and np.isfinite(s2).all() and np.abs(s2 - s).max() > 0
        # set_command / get_command round trip in the reference's units
        com = np.linspace(-1, 1, 90).astype(np.float32)
        sup.rtc.set_command(0, com)
        assert np.array_equal(sup.rtc.get_command(0), com)
        g0 = sup.rtc._rtc.d_control[0].gain
        env.set_gain(0.25)
        assert abs(sup.rtc._rtc.d_control[0].gain - 0.25) < 1e-7 and g0 != 0.25
        env.sim.check_device()
    finally:
        env.sim.close()


def test_normalization_producer_against_reference_statistics(static10, oracle_imat10, torch, tmp_path):
    """ao_marl_b200.tools.obtain_normalization (the reference's normalization_loop, batched over the seeds): its
    statistics land in the distribution of the ones the reference authors committed for the same parameter file,
    and its output file loads through the RL layout."""
    import copy
    from ao_marl_b200 import calibration
    from ao_marl_b200.init import rtc as rtc_b
    from ao_marl_b200.lib import Simulator
    from ao_marl_b200.rl.layout import RLLayout, load_normalization
    from ao_marl_b200.tools import obtain_normalization as on
    t = copy.copy(static10)
    t.imat = calibration.measure_imat(static10)
    t.cmat = rtc_b.cmat_with_btt(t.imat, t.Btt, 0)
    sim = Simulator(t, 20, rl=None)
    try:
        norm, zn = on.normalization_loop(sim, t, n_frames=400, first_seed=1, settle=20)
    finally:
        sim.close()
    ref, zn_ref = load_normalization("production_sh_10x10_2m.py")
    assert zn.shape == zn_ref.shape == (87,)
    assert abs(norm["wfs"]["std"].mean() / ref["wfs"]["std"].mean() - 1) < 0.25
    ratio = norm["dm"]["std"] / ref["dm"]["std"]
    assert 0.5 < np.median(ratio) < 2.0, np.median(ratio)
    assert 0.4 < np.median(zn / zn_ref) < 2.5
    out = tmp_path / "norm.npz"
    on.save(str(out), norm, zn)
    z = np.load(out)
    rl = RLLayout(87, dict(parameters_telescope="production_sh_10x10_2m.py", n_zernike_start_end=[0, 80]), None, 3,
                  norm={k: {s: z["%s_%s" % (k, s)] for s in ("mean", "std", "max", "min")}
                        for k in ("dm", "wfs", "dm_residual")}, zn_norm=z["zn_norm"])
    assert rl.state_dim == 4 * 82 and np.isfinite(rl.freedom).all()


def test_strehl_kernel_matches_the_materialised_phase(sim10, static10, torch):
    """aom_comp_strehl (phase evaluated and reduced on the fly) against the variance and the on-axis intensity ratio of
    the materialised pupil phase; long-exposure means over frames; reset."""
    seeds = np.array([41, 42, 43, 44], dtype=np.int64)
    sim10.reset(seeds)
    r = np.random.default_rng(4)
    sim10.set_dm_volts(torch.as_tensor((r.standard_normal((4, static10.nactu)) * 5).astype(np.float32), device="cuda"))
    lam = 1.65
    pup = torch.as_tensor(static10.mpupil, device="cuda") > 0
    se_hist, var_hist = [], []
    for it in range(3):
        sim10.move_atmos()
        s = sim10.comp_strehl(lam).clone()
        ph = sim10.raytrace_wfs()
        var = ph[:, pup].double().var(dim=1, unbiased=False).float()
        k = 2 * np.pi / lam
        pd = ph[:, pup].double() * k
        se = (pd.cos().mean(dim=1) ** 2 + pd.sin().mean(dim=1) ** 2).float()     # on-axis intensity |<exp(i k phi)>|^2
        assert float((se - torch.exp(-var * k * k)).abs().max()) < 0.5               # Marechal only holds for small residuals
        assert float((s[:, 2] - var).abs().max()) < 2e-5 * float(var.max())
        assert float((s[:, 0] - se).abs().max()) < 1e-5
        se_hist.append(se)
        var_hist.append(var)
        assert float((s[:, 1] - torch.stack(se_hist).mean(0)).abs().max()) < 1e-5
        assert float((s[:, 3] - torch.stack(var_hist).mean(0)).abs().max()) < 2e-5 * float(var.max())
    sim10.reset_strehl()
    assert float(sim10.buffer("STREHL").abs().max()) == 0.0
    sim10.check_device()


def test_psf_core_strehl_against_oracle(system10, static10, torch):
    """Peak Strehl (comp_strehl(do_fit=True), targetCompass.py:139-196): the 3 x 3 PSF core summed directly in the pupil
    sweep + the three-point fit, against the oracle's full zero-padded FFT image (oracle.aoframe.psf_image / fit_peak) of
    the same phase; short exposure per frame, long exposure = fit of the accumulated image; both sweep forms; and the
    on-demand image getter.  Closed loop with the integrator so that the peak sits within a pixel of the axis."""
    from oracle import aoframe as af
    from ao_marl_b200.supervisor.components import TargetB200
    sim, t, rl = system10
    g = t.config.p_geom
    pd, off = int(g.pupdiam), (int(g._n) - int(g.pupdiam)) // 2
    sp = np.asarray(g._spupil) > 0
    nfft = sim.psf_nfft
    assert nfft == np.asarray(g._ipupil).shape[0] == 512
    lam = 1.65
    sim.reset(np.array([21, 22, 23], dtype=np.int64))
    sim.reset_strehl()
    for _ in range(25):                                        # let the integrator converge
        sim.move_atmos(); sim.comp_wfs_image(); sim.do_centroids(); sim.do_control(); sim.apply_control()
    le_img = [np.zeros((nfft, nfft)) for _ in range(3)]
    worst = 0.0
    try:
        for it in range(4):
            sim.set_pupil_path("pixel" if it & 1 else "sweep")
            sim.move_atmos(); sim.comp_wfs_image(); sim.do_centroids(); sim.do_control(); sim.apply_control()
            s = sim.comp_strehl(lam, peak=True).cpu().numpy()
            on_axis = sim.comp_strehl(lam, accumulate=False).cpu().numpy()[:, 0]
            ph = sim.raytrace_wfs().cpu().numpy()[:, off:off + pd, off:off + pd]
            for e in range(3):
                img = af.psf_image(ph[e], sp, lam, nfft)
                le_img[e] += img
                se, le = af.fit_peak(img), af.fit_peak(le_img[e] / (it + 1))
                assert np.unravel_index(np.argmax(img), img.shape) == (0, 0)     # closed loop: brightest pixel on axis
                assert se > 0.3 and se >= img[0, 0] and abs(on_axis[e] - img[0, 0]) < 1e-5
                worst = max(worst, abs(s[e, 0] - se), abs(s[e, 1] - le))
                assert abs(s[e, 0] - se) < 2e-5 and abs(s[e, 1] - le) < 2e-5, (it, e, s[e], se, le)
    finally:
        sim.set_pupil_path("sweep")
    print("peak Strehl vs oracle FFT image, worst abs error:", worst)
    # the image getter against the oracle image (centred)
    tar = TargetB200(sim, t.config, t)
    tar.raytrace(0, atm=type("A", (), {"is_enable": True})(), dms=object())
    img = tar.get_tar_image(0, envs=[1])
    ref = np.fft.fftshift(af.psf_image(ph[1], sp, lam, nfft))
    assert img.shape == (nfft, nfft) and np.abs(img - ref).max() < 1e-5
    sim.reset_strehl()
    sim.check_device()


def test_strehl_ordering_in_the_fused_step(system10, torch):
    """AOM_OPT_STREHL inside aom_step: value 1 publishes the phase as traced in the previous next_part_one, i.e. with the
    voltages that were on the mirrors before this step's apply_control (rlSupervisor.py:964-965 then 944-947); value 2
    re-traces after apply_control ("modification_online", rlSupervisor.py:936-940).  Checked (a) against the stand-alone
    call issued right before each step and (b), without atmosphere, through the identity SE2[t] == SE1[t+1]: both see
    the voltages of step t."""
    from ao_marl_b200.lib import Simulator
    sim, t, rl = system10
    seeds = np.array([7, 8, 9], dtype=np.int64)
    g = torch.Generator(device="cuda").manual_seed(3)
    actions = [torch.randn((3, rl.action_dim), device="cuda", generator=g).clamp_(-1, 1) for _ in range(5)]

    def run(s, how, delay0=False):
        s.reset(seeds)
        s.reset_strehl()
        s.step_with_strehl(how == "fused", 1.65, pure_delay_0=delay0)
        s.state_begin(); s.move_atmos(); s.comp_wfs_image(); s.do_centroids(); s.do_control(); s.state_end()
        out = []
        for a in actions:
            if how == "standalone":
                out.append(s.comp_strehl(1.65)[:, :2].clone())
            s.rows("ACTION", rl.action_dim).copy_(a)
            s.step(mode=1)
            if how == "fused":
                out.append(s.buffer("STREHL").view(3, 4)[:, :2].clone())
        s.step_with_strehl(False)
        return torch.stack(out)

    fused1 = run(sim, "fused")
    alone = run(sim, "standalone")
    assert float((fused1 - alone).abs().max()) < 1e-6       # same kernels, same screens and voltages (atomic sum order differs)
    fused2 = run(sim, "fused", delay0=True)
    assert float((fused2[:, :, 0] - fused1[:, :, 0]).abs().max()) > 1e-4
    flat = Simulator(t, 3, rl, atmosphere=False)
    try:
        f1 = run(flat, "fused")[:, :, 0]
        f2 = run(flat, "fused", delay0=True)[:, :, 0]
        assert float((f2[:-1] - f1[1:]).abs().max()) < 1e-6
        assert float((f2 - f1).abs().max()) > 1e-4
        assert float(f1.min()) > 0.0 and float(f1.max()) <= 1.0 + 1e-6
    finally:
        flat.close()
    sim.check_device()


def test_error_behaviour_and_edge_sizes(static10, torch):
    """Edge cases of the C ABI: rejected sizes raise (ValueError for dimension problems, as the reference's
    set_command does, rtcCompass.py:471-472), empty work is a no-op, calls before their prerequisites report
    AOM_ERR_STATE instead of reading unset tables."""
    import ctypes
    from ao_marl_b200.denoiser import Autoencoder
    from ao_marl_b200.lib import Simulator
    with pytest.raises(ValueError):
        Simulator(static10, 0, rl=None)
    with pytest.raises(ValueError):
        Simulator(static10, 65536, rl=None)
    import copy
    t = copy.copy(static10)
    t.geo_proj = None                                        # another module may have built the projector tables
    sim = Simulator(t, 1, rl=None, atmosphere=False)
    try:
        with pytest.raises(ValueError):
            sim.set_table("MPUPIL", np.zeros(7, np.float32))
        with pytest.raises(ValueError):
            sim.set_command(torch.zeros((1, static10.nactu + 1), device="cuda"))
        with pytest.raises(RuntimeError):
            sim.denoise()                                    # no denoiser parameters uploaded
        Autoencoder(dict(type="cnn_single_subaperture", path="autoencoder_M9_rms_3"), device="cuda", sim=sim)
        with pytest.raises(RuntimeError):
            sim.denoise()                                    # no detector cube kept yet
        empty = torch.empty((0, 256), device="cuda")
        assert sim.denoise(empty).shape == (0, 256)
        with pytest.raises(ValueError):
            sim.denoise(torch.zeros(100, device="cuda"))
        with pytest.raises(RuntimeError):
            sim.do_control_geo()                             # projector tables not uploaded
        # no atmosphere: flat wavefront, zero slopes, unit Strehl, zero geometric command
        sim.reset(np.array([1], dtype=np.int64))
        sim.comp_wfs_image(noise=-1.0)
        sim.do_centroids()
        assert float(sim.rows("SLOPES", static10.nslopes).abs().max()) < 2e-6
        s = sim.comp_strehl(1.65)
        assert abs(float(s[0, 0]) - 1.0) < 1e-6 and float(s[0, 2]) == 0.0
        from ao_marl_b200.init import geo
        sim.set_geo(*geo.build_geo(static10))
        assert float(sim.do_control_geo().abs().max()) == 0.0
        sim.check_device()
    finally:
        sim.close()
