"""WFS-image denoiser of the `_d0_noise` configurations (BASELINE.json config 4).

Architecture and parameter names follow the reference's `DenoisingAutoencoderCNN2DSingleSubapeture`
(src/autoencoder/autoencoder_models.py:130-197; MSE variant: no batch-norm, no sigmoid) so its checkpoints load
unchanged; `Autoencoder` mirrors the reference wrapper (autoencoder_models.py:199-240) with a device-resident
`predict`: the reference moves every frame GPU -> CPU -> GPU and loops over subapertures in Python
(rlSupervisor.py:857-891), here the whole [E * nvalid, 1, 16, 16] cube of a frame goes through the network in
chunks and the result feeds aom_set_bincube / aom_do_centroids without leaving the device.

On the product path the network is ONE hand-written fused CUDA kernel (csrc/denoise_kernels.cuh, `aom_denoise`):
`Autoencoder(config, sim=simulator)` packs the checkpoint into the kernel's layout and every `predict` of a CUDA tensor
goes through the library.  The torch module below is the holder of the reference's parameter names (checkpoint
loading, the golden-vector test against the reference's own module on the CPU) and what `predict` uses when no
simulator is attached (host-side tooling); measured on B200 it needs 3.5 ms per environment through cuDNN in float32
(0.36 ms with TF32 allowed) against the fused kernel's figure in DESIGN.md section 4.
3.42 MFLOP per subaperture, 4.1 GFLOP per 40x40 frame and environment.
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


def pack_weights(state_dict):
    """Parameters in the fused kernel's order and layout (csrc/denoise_kernels.cuh: DN_E1W ... DN_D3B): convolution
    weights as [input channel][tap][output channel], biases after each, padded to a multiple of four floats."""
    sd = {k: np.asarray(v.detach().cpu() if hasattr(v, "detach") else v, dtype=np.float32) for k, v in state_dict.items()}
    parts = []
    for name in ("encoder1", "encoder2", "encoder3"):            # Conv2d weight [co][ci][ky][kx]
        w = sd[name + ".weight"]
        parts += [w.transpose(1, 2, 3, 0).reshape(-1), sd[name + ".bias"]]
    for name in ("decoder1", "decoder2"):                        # ConvTranspose2d weight [ci][co][ky][kx]
        w = sd[name + ".weight"]
        parts += [w.transpose(0, 2, 3, 1).reshape(-1), sd[name + ".bias"]]
    parts += [sd["decoder3.weight"].reshape(-1), sd["decoder3.bias"]]      # [16][1][3][3] -> [ci][tap]
    flat = np.concatenate(parts).astype(np.float32)
    expect = 144 + 16 + 4608 + 32 + 18432 + 64 + 32768 + 32 + 8192 + 16 + 144 + 1
    if flat.size != expect:
        raise ValueError("Dimension mismatch: the checkpoint is not the per-subaperture CNN (%d parameters)" % flat.size)
    out = np.zeros((flat.size + 3) & ~3, dtype=np.float32)
    out[:flat.size] = flat
    return out


# ---- tensor-core form of the same network (csrc/denoise_tc.cuh) -------------------------------------------------------
# Geometry shared with the kernel: G spots per pass on a 4 x 4 coarse grid with a shared pad row / column (pitch 5), every
# finer map stored space-to-depth on that grid; activations scaled by SCALE (a power of two: exact) so that fp16 hi / lo
# operands cannot overflow on bright spots.  CT_TAPS[parity] = the two (kernel index, input offset) pairs a stride-2
# 4 x 4 transposed convolution uses for an output row / column of that parity.
DT = dict(G=5, PITCH=5, LEAD=8, NPOS=144, SCALE=2.0 ** -4, CT_TAPS=(((1, 0), (3, -1)), ((0, 1), (2, 0))),
          CHUNKS=(0, 18432, 55296, 92160, 124928, 157696, 190464, 223232, 256000), NPARAM=452)


def _tile_f16(w_nk):
    """[N][K] float -> (hi, lo) fp16 tiles in the UMMA K-major core-matrix order [K / 8][N][8]."""
    w = np.asarray(w_nk, np.float32)
    hi = w.astype(np.float16)
    lo = (w - hi.astype(np.float32)).astype(np.float16)
    n, k = w.shape
    order = lambda t: np.ascontiguousarray(t.reshape(n, k // 8, 8).transpose(1, 0, 2)).reshape(-1)
    return order(hi), order(lo)


def _tile_f16_merged(w_nk):
    """[N][K] float -> one fp16 tile of 2N rows, hi rows then lo rows, in the UMMA K-major order [K / 8][2N][8]: A_hi times
    the whole tile gives A_hi W_hi and A_hi W_lo side by side in one instruction, A_lo times its first N rows A_lo W_hi."""
    w = np.asarray(w_nk, np.float32)
    hi = w.astype(np.float16)
    lo = (w - hi.astype(np.float32)).astype(np.float16)
    both = np.concatenate([hi, lo], axis=0)
    n2, k = both.shape
    return np.ascontiguousarray(both.reshape(n2, k // 8, 8).transpose(1, 0, 2)).reshape(-1)


def pack_weights_tc(state_dict, as_float=False):
    """Weight tiles of the four tensor-core layers in the order the kernel streams them, and the float parameters of the
    two SIMT layers + the (scaled) biases.  as_float: float64 tiles as a dict (profiles/dev/denoise_tc_model.py)."""
    sd = {k: np.asarray(v.detach().cpu() if hasattr(v, "detach") else v, dtype=np.float32) for k, v in state_dict.items()}
    S, ct = DT["SCALE"], DT["CT_TAPS"]
    e2, e3, d1, d2 = sd["encoder2.weight"], sd["encoder3.weight"], sd["decoder1.weight"], sd["decoder2.weight"]
    L2 = [e2[:, :, ky, kx] for ky in range(3) for kx in range(3)]                       # [n = co][k = ci]
    L3 = [e3[:, :, ky, kx] for ky in range(3) for kx in range(3)]
    L4 = [[d1[:, :, ct[py][ty][0], ct[px][tx][0]].T for ty in range(2) for tx in range(2)]
          for py in range(2) for px in range(2)]                                        # ConvTranspose2d: [ci][co] -> [co][ci]
    L5 = [d2[:, :, ky, kx].T for ky in range(4) for kx in range(4)]
    if as_float:
        f = lambda t: np.asarray(t, np.float64)
        return dict(L2=[f(t) for t in L2], L3=[f(t) for t in L3], L4=[[f(t) for t in c] for c in L4], L5=[f(t) for t in L5],
                    b2=f(sd["encoder2.bias"]) * S, b3=f(sd["encoder3.bias"]) * S, b4=f(sd["decoder1.bias"]) * S,
                    b5=f(sd["decoder2.bias"]) * S)
    parts = [_tile_f16_merged(t) for t in L2]                                            # e2: hi | lo side by side
    parts += [_tile_f16(t)[0] for t in L3] + [_tile_f16(t)[1] for t in L3]               # e3: hi chunk, lo chunk
    for c in L4:
        parts += [_tile_f16_merged(t) for t in c]                                        # d1
    parts += [_tile_f16_merged(t) for t in L5]                                           # d2
    blob = np.concatenate(parts)
    assert blob.nbytes == DT["CHUNKS"][-1], blob.nbytes
    prm = np.concatenate([sd["encoder1.weight"].reshape(16, 9).reshape(-1), sd["encoder1.bias"],
                          sd["encoder2.bias"] * S, sd["encoder3.bias"] * S, sd["decoder1.bias"] * S, sd["decoder2.bias"] * S,
                          sd["decoder3.weight"].reshape(16, 9).reshape(-1), sd["decoder3.bias"].reshape(1),
                          np.array([S, 1.0 / S], np.float32)]).astype(np.float32)
    out = np.zeros(DT["NPARAM"], np.float32)
    out[:prm.size] = prm
    return blob.view(np.uint16), out


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

    def __init__(self, config, device=None, chunk=65536, sim=None):
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
        self.sim = None
        if sim is not None:
            self.attach(sim)

    def attach(self, sim):
        """Route `predict` through the fused CUDA kernel of this simulator context (uploads the packed parameters)."""
        sim.set_denoiser(pack_weights(self.model.state_dict()), pack_weights_tc(self.model.state_dict()))
        self.sim = sim

    @torch.no_grad()
    def predict(self, noisy_tensor, only_inference_time=False):
        """[..., 16, 16] (or [..., 256]) spots -> denoised spots of the same shape, on the input's device."""
        if self.sim is not None:
            x = torch.as_tensor(noisy_tensor, dtype=torch.float32, device="cuda")
            return self.sim.denoise(x.contiguous())
        x = torch.as_tensor(noisy_tensor, dtype=torch.float32, device=self.device)
        shape = x.shape
        x = x.reshape(-1, 1, 16, 16)
        out = torch.empty_like(x)
        for i in range(0, x.shape[0], self.chunk):
            out[i:i + self.chunk] = self.model(x[i:i + self.chunk])
        return out.reshape(shape)
