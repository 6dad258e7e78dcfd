"""Denoiser module against the reference's own network output (tests/golden/ref_autoencoder.npz, written by
oracle/refharness/export_autoencoder.py from /root/reference)."""
import os

import numpy as np
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
