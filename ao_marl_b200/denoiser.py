"""WFS-image denoiser of the `_d0_noise` configurations (BASELINE.json config 4).

Architecture and parameter names follow the reference's `DenoisingAutoencoderCNN2DSingleSubapeture`
(src/autoencoder/autoencoder_models.py:130-197; MSE variant: no batch-norm, no sigmoid) so its checkpoints load
unchanged; `Autoencoder` mirrors the reference wrapper (autoencoder_models.py:199-240) with a device-resident
`predict`: the reference moves every frame GPU -> CPU -> GPU and loops over subapertures in Python
(rlSupervisor.py:857-891), here the whole [E * nvalid, 1, 16, 16] cube of a frame goes through the network in
chunks and the result feeds aom_set_bincube / aom_do_centroids without leaving the device.

The network itself runs on cuDNN through PyTorch (library code; a hand-written tcgen05 implicit-GEMM kernel is
the next step for this row): 3.42 MFLOP per subaperture, 4.1 GFLOP per 40x40 frame.
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "autoencoder")


class DenoisingAutoencoderCNN2DSingleSubapeture(nn.Module):
    def __init__(self, criterion="MSE", batch_norm=False):
        super().__init__()
        if batch_norm:
            raise NotImplementedError("the shipped checkpoints have no batch-norm layers")
        self.encoder1 = nn.Conv2d(1, 16, kernel_size=3, stride=1, padding=1)
        self.encoder2 = nn.Conv2d(16, 32, kernel_size=3, stride=1, padding=1)
        self.encoder3 = nn.Conv2d(32, 64, kernel_size=3, stride=1, padding=1)
        self.decoder1 = nn.ConvTranspose2d(64, 32, kernel_size=4, stride=2, padding=1)
        self.decoder2 = nn.ConvTranspose2d(32, 16, kernel_size=4, stride=2, padding=1)
        self.decoder3 = nn.ConvTranspose2d(16, 1, kernel_size=3, stride=1, padding=1)
        self.criterion = criterion

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.encoder1(x)), 2)
        x = F.max_pool2d(F.relu(self.encoder2(x)), 2)
        x = F.relu(self.encoder3(x))
        x = F.relu(self.decoder1(x))
        x = F.relu(self.decoder2(x))
        x = self.decoder3(x)
        return torch.sigmoid(x) if self.criterion == "BCE" else x


def load_weights(name_or_path):
    """state_dict from a reference checkpoint (torch.save of model.state_dict()) or a shipped .npz export."""
    path = name_or_path
    if not os.path.exists(path):
        path = os.path.join(DATA_DIR, os.path.basename(name_or_path) + ".npz")
    if path.endswith(".npz"):
        z = np.load(path)
        return {k: torch.from_numpy(z[k].copy()) for k in z.files}
    return torch.load(path, map_location="cpu")


class Autoencoder:
    """Reference-shaped wrapper: `Autoencoder(config)` with config.autoencoder = {'type', 'path'}."""

    def __init__(self, config, device=None, chunk=65536):
        ae = config.autoencoder if hasattr(config, "autoencoder") else dict(config)
        self.type = str(ae.get("type", "cnn_single_subaperture")).lower()
        if self.type != "cnn_single_subaperture":
            raise NotImplementedError("only the per-subaperture CNN denoiser is used by the production files")
        self.device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        self.model = DenoisingAutoencoderCNN2DSingleSubapeture()
        if ae.get("path") is not None:
            self.model.load_state_dict(load_weights(ae["path"]))
        self.model.to(self.device).eval()
        self.chunk = int(chunk)

    @torch.no_grad()
    def predict(self, noisy_tensor, only_inference_time=False):
        """[..., 16, 16] (or [..., 256]) spots -> denoised spots of the same shape, on the input's device."""
        x = torch.as_tensor(noisy_tensor, dtype=torch.float32, device=self.device)
        shape = x.shape
        x = x.reshape(-1, 1, 16, 16)
        out = torch.empty_like(x)
        for i in range(0, x.shape[0], self.chunk):
            out[i:i + self.chunk] = self.model(x[i:i + self.chunk])
        return out.reshape(shape)
