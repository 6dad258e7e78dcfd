"""Denoiser module against the reference's own network output (tests/golden/ref_autoencoder.npz, written by
oracle/refharness/export_autoencoder.py from /root/reference)."""
import os

import numpy as np
import pytest
import torch

from ao_marl_b200.denoiser import Autoencoder, load_weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_autoencoder.npz")


def test_denoiser_matches_reference_output():
    g = np.load(GOLDEN)
    ae = Autoencoder(dict(type="cnn_single_subaperture", path="autoencoder_M9_rms_3"), device="cpu")
    y = ae.predict(torch.from_numpy(g["x"])).numpy()
    assert y.shape == g["y"].shape
    assert np.abs(y - g["y"]).max() < 1e-5 * np.abs(g["y"]).max()
    # flat [n, 256] input (the layout of AOM_B_BINCUBE) gives the same result
    y2 = ae.predict(torch.from_numpy(g["x"].reshape(-1, 256))).numpy().reshape(-1, 16, 16)
    assert np.array_equal(y, y2)


def test_shipped_weights_have_the_reference_shapes():
    sd = load_weights("autoencoder_M9_rms_3")
    want = {"encoder1.weight": (16, 1, 3, 3), "encoder2.weight": (32, 16, 3, 3), "encoder3.weight": (64, 32, 3, 3),
            "decoder1.weight": (64, 32, 4, 4), "decoder2.weight": (32, 16, 4, 4), "decoder3.weight": (16, 1, 3, 3)}
    for k, shp in want.items():
        assert tuple(sd[k].shape) == shp
    assert sum(v.numel() for v in sd.values()) == 64449


def test_packed_parameters_follow_the_kernel_layout():
    """pack_weights: [input channel][tap][output channel] per layer, biases after each (csrc/denoise_kernels.cuh)."""
    from ao_marl_b200.denoiser import pack_weights
    sd = load_weights("autoencoder_M9_rms_3")
    p = pack_weights(sd)
    assert p.dtype == np.float32 and p.size == 64452 and p[64449:].sum() == 0
    e2 = sd["encoder2.weight"].numpy()
    off = 144 + 16
    assert p[off + (3 * 9 + 2 * 3 + 1) * 32 + 7] == e2[7, 3, 2, 1]
    d1 = sd["decoder1.weight"].numpy()
    off = 144 + 16 + 4608 + 32 + 18432 + 64
    assert p[off + (5 * 16 + 3 * 4 + 2) * 32 + 9] == d1[5, 9, 3, 2]
    assert p[64448] == sd["decoder3.bias"].numpy()[0]
    with pytest.raises(ValueError):
        pack_weights({k: v for k, v in sd.items() if not k.startswith("encoder3")} | {
            "encoder3.weight": torch.zeros(8, 32, 3, 3), "encoder3.bias": torch.zeros(8)})


@pytest.mark.gpu
def test_fused_kernel_matches_the_reference_module(static10):
    """aom_denoise against the reference module's own output (golden vectors) and against the torch module on raw
    photo-electron spots; spot counts that do not fill the last CTA batch; float32 tolerance 2e-5 of the output scale."""
    from ao_marl_b200.lib import Simulator
    g = np.load(GOLDEN)
    sim = Simulator(static10, 2, rl=None)
    try:
        ae = Autoencoder(dict(type="cnn_single_subaperture", path="autoencoder_M9_rms_3"), device="cuda", sim=sim)
        ref = Autoencoder(dict(type="cnn_single_subaperture", path="autoencoder_M9_rms_3"), device="cpu")
        # both kernels: the tensor-core one (default: four layers as implicit GEMMs on tcgen05, fp16 hi / lo operands) and
        # the float32 FFMA one; same tolerance
        for path in ("tcgen05", "simt"):
            sim.set_denoise_path(path)
            y = ae.predict(torch.from_numpy(g["x"]).cuda()).cpu().numpy()
            sim.check_device()
            assert y.shape == g["y"].shape
            assert np.abs(y - g["y"]).max() < 2e-5 * np.abs(g["y"]).max(), path
            gen = torch.Generator().manual_seed(3)
            for n in (1, 3, 4, 5, 6, 1027):
                for scale in (30.0, 3000.0):                   # faint and bright spots (fp16 operand range)
                    x = torch.poisson(torch.rand((n, 16, 16), generator=gen) * scale, generator=gen) + \
                        torch.randn((n, 16, 16), generator=gen) * 3
                    want = ref.predict(x).numpy()
                    got = ae.predict(x.cuda()).cpu().numpy()
                    assert got.shape == want.shape
                    assert np.abs(got - want).max() < 2e-5 * np.abs(want).max(), (path, n, scale)
                    again = ae.predict(x.cuda()).cpu().numpy()
                    assert np.array_equal(got, again), (path, n)      # run-to-run reproducible
        sim.set_denoise_path("tcgen05")
        sim.check_device()
    finally:
        sim.close()


@pytest.mark.gpu
def test_in_place_denoising_feeds_the_centroider(static10):
    """aom_denoise(NULL, NULL): the frame's detector cube is denoised in place and the next aom_do_centroids reads it --
    same slopes as the explicit predict + set_bincube round trip of the reference's flow."""
    from ao_marl_b200.lib import Simulator
    sim = Simulator(static10, 3, rl=None)
    try:
        ae = Autoencoder(dict(type="cnn_single_subaperture", path="autoencoder_M9_rms_3"), device="cuda", sim=sim)
        sim.reset(np.array([9, 10, 11], dtype=np.int64))
        nv, n = static10.p_wfs._nvalid, static10.nslopes
        sim.comp_wfs_image(keep_image=True, noise=3.0)
        cube = sim.buffer("BINCUBE").view(3, nv, 256).clone()
        den = ae.predict(cube)
        sim.set_bincube(den)
        sim.do_centroids()
        a = sim.rows("SLOPES", n).clone()
        sim.denoise()
        assert torch.equal(sim.buffer("BINCUBE").view(3, nv, 256), den)
        sim.do_centroids()
        b = sim.rows("SLOPES", n).clone()
        assert torch.equal(a, b)
        sim.check_device()
    finally:
        sim.close()


@pytest.mark.gpu
def test_step_with_denoiser_equals_the_manual_sequence(static10, oracle_imat10):
    """AOM_OPT_DENOISE: aom_step = apply_control, move_atmos, frame (image kept), denoise, centroids, control --
    bit-identical to the calls RlSupervisor.next_part_one_integrator makes one by one (rlSupervisor.py:949-987)."""
    import copy
    from ao_marl_b200.init import rtc as rtc_b
    from ao_marl_b200.lib import Simulator
    t = copy.copy(static10)
    t.imat = oracle_imat10
    t.cmat = rtc_b.cmat_with_btt(t.imat, t.Btt, 5)
    seeds = np.array([21, 22], dtype=np.int64)
    sims = [Simulator(t, 2, rl=None) for _ in range(2)]
    try:
        for sim in sims:
            Autoencoder(dict(type="cnn_single_subaperture", path="autoencoder_M9_rms_3"), device="cuda", sim=sim)
            sim.reset(seeds)
        sims[0].step_with_denoiser(True)
        for _ in range(3):
            sims[0].step(mode=2)
            m = sims[1]
            m.apply_control()
            m.move_atmos()
            m.comp_wfs_image(keep_image=True)
            m.denoise()
            m.do_centroids()
            m.do_control()
        for name, n in (("SLOPES", t.nslopes), ("COM", t.nactu)):
            assert torch.equal(sims[0].rows(name, n), sims[1].rows(name, n)), name
        raw = Simulator(t, 2, rl=None)
        raw.reset(seeds)
        for _ in range(3):
            raw.step(mode=2)
        assert not torch.equal(raw.rows("SLOPES", t.nslopes), sims[0].rows("SLOPES", t.nslopes))
        raw.close()
        for sim in sims:
            sim.check_device()
    finally:
        for sim in sims:
            sim.close()
