"""TEST/DATA INFRASTRUCTURE -- exports the reference's trained denoiser weights and a golden input/output pair.

    python -m oracle.refharness.export_autoencoder

* ao_marl_b200/data/autoencoder/autoencoder_M9_rms_3.npz : the state_dict of the committed checkpoint
  /root/reference/output/autoencoder/autoencoder_weights/autoencoder_M9_rms_3 (data, 64 449 float32 weights).
* tests/golden/ref_autoencoder.npz : output of the REFERENCE's own module
  (src/autoencoder/autoencoder_models.py:130-197, DenoisingAutoencoderCNN2DSingleSubapeture, imported from
  /root/reference with matplotlib / torchvision stubbed) on a seeded batch of noisy spots, float32 on CPU.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")


def main():
    for name in ("matplotlib", "matplotlib.pyplot", "torchvision", "torchvision.transforms"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    sys.path.insert(0, os.path.join(REF, "src", "autoencoder"))
    import autoencoder_models as ref
    sd = torch.load(os.path.join(REF, "output/autoencoder/autoencoder_weights/autoencoder_M9_rms_3"), map_location="cpu")
    out = os.path.join(ROOT, "ao_marl_b200", "data", "autoencoder", "autoencoder_M9_rms_3.npz")
    np.savez_compressed(out, **{k: v.numpy() for k, v in sd.items()})
    model = ref.DenoisingAutoencoderCNN2DSingleSubapeture()
    model.load_state_dict(sd)
    model.eval()
    rng = np.random.default_rng(7)
    yy, xx = np.mgrid[0:16, 0:16]
    spots = []
    for i in range(12):
        cx, cy = 7.5 + rng.normal(0, 1.2, 2)
        clean = 240.0 * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * 1.6 ** 2)) / (2 * np.pi * 1.6 ** 2)
        noisy = rng.poisson(clean) + rng.normal(0, 3.0, clean.shape)
        spots.append(noisy.astype(np.float32))
    x = np.stack(spots)
    with torch.no_grad():
        y = model(torch.from_numpy(x).view(-1, 1, 16, 16)).numpy().reshape(-1, 16, 16)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_autoencoder.npz"), x=x, y=y)
    print("wrote", out, "and golden pair", x.shape, y.shape, float(np.abs(y).max()))


if __name__ == "__main__":
    main()
