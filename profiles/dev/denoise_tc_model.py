"""numpy model of the tcgen05 denoiser schedule (csrc/denoise_tc.cuh): same activation planes, tap -> (phase, shift)
maps and weight tiles as the kernel, evaluated with plain matrix products, against the torch module on the CPU.
Run: python profiles/dev/denoise_tc_model.py"""
import sys
import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from ao_marl_b200.denoiser import DenoisingAutoencoderCNN2DSingleSubapeture, load_weights, pack_weights_tc, DT  # noqa: E402

G, PITCH, LEAD, NPOS = DT["G"], DT["PITCH"], DT["LEAD"], DT["NPOS"]


def pos(s, Y, X):
    return (1 + 5 * s + Y) * PITCH + 1 + X


def fdiv2(a):
    return a // 2            # floor


def run(x, sd):
    """x [G][16][16] float64 -> out [G][16][16] via the kernel's data flow (float64 arithmetic, no splits)."""
    W = pack_weights_tc(sd, as_float=True)      # dict of float tiles in the kernel's order
    S = DT["SCALE"]
    # ---- L1 (SIMT in the kernel): a1 pooled 8x8x16, stored space-to-depth: A1s[phase][ch][pos]
    w1, b1 = sd["encoder1.weight"].double().numpy(), sd["encoder1.bias"].double().numpy()
    A1s = np.zeros((4, 16, NPOS))
    xp = np.zeros((G, 18, 18)); xp[:, 1:17, 1:17] = x
    for s in range(G):
        for Y in range(4):
            for X in range(4):
                win = xp[s, 4 * Y:4 * Y + 6, 4 * X:4 * X + 6]
                for c in range(16):
                    conv = np.zeros((4, 4))
                    for fy in range(4):
                        for fx in range(4):
                            conv[fy, fx] = b1[c] + (w1[c, 0] * win[fy:fy + 3, fx:fx + 3]).sum()
                    conv = np.maximum(conv, 0)
                    for py in range(2):
                        for px in range(2):
                            A1s[py * 2 + px, c, LEAD + pos(s, Y, X)] = conv[2 * py:2 * py + 2, 2 * px:2 * px + 2].max() * S
    real = np.zeros(128, bool)
    for s in range(G):
        for Y in range(4):
            for X in range(4):
                real[pos(s, Y, X)] = True

    def tile(A, shift):                    # A [K][NPOS] -> [128][K] rows = positions
        return A[:, LEAD + shift:LEAD + shift + 128].T

    # ---- L2: 4 phase accumulators, 9 taps each
    D2 = np.zeros((4, 128, 32))
    for py in range(2):
        for px in range(2):
            for ky in range(3):
                for kx in range(3):
                    dy, dx = ky - 1, kx - 1
                    iph = ((py + dy) & 1) * 2 + ((px + dx) & 1)
                    sh = fdiv2(py + dy) * PITCH + fdiv2(px + dx)
                    D2[py * 2 + px] += tile(A1s[iph], sh) @ W["L2"][ky * 3 + kx].T      # W tile [N][K]
    a2 = np.maximum(D2.max(0) + W["b2"], 0)
    A2 = np.zeros((32, NPOS)); A2[:, LEAD:LEAD + 128][:, real] = a2[real].T
    # ---- L3
    D3 = np.zeros((128, 64))
    for ky in range(3):
        for kx in range(3):
            D3 += tile(A2, (ky - 1) * PITCH + kx - 1) @ W["L3"][ky * 3 + kx].T
    a3 = np.maximum(D3 + W["b3"], 0)
    A3 = np.zeros((64, NPOS)); A3[:, LEAD:LEAD + 128][:, real] = a3[real].T
    # ---- L4: 4 classes x 4 taps
    A4s = np.zeros((4, 32, NPOS))
    for py in range(2):
        for px in range(2):
            D = np.zeros((128, 32))
            for ty in range(2):
                for tx in range(2):
                    ky, dy = DT["CT_TAPS"][py][ty]
                    kx, dx = DT["CT_TAPS"][px][tx]
                    D += tile(A3, dy * PITCH + dx) @ W["L4"][py * 2 + px][ty * 2 + tx].T
            a4 = np.maximum(D + W["b4"], 0)
            A4s[py * 2 + px][:, LEAD:LEAD + 128][:, real] = a4[real].T
    # ---- L5: 16 super-classes, + L6 scatter
    w6, b6 = sd["decoder3.weight"].double().numpy(), float(sd["decoder3.bias"])
    out = np.zeros((G, 18, 18))
    for py in range(2):
        for px in range(2):
            for qy in range(2):
                for qx in range(2):
                    D = np.zeros((128, 16))
                    for ty in range(2):
                        for tx in range(2):
                            ky, dy = DT["CT_TAPS"][qy][ty]
                            kx, dx = DT["CT_TAPS"][qx][tx]
                            iph = ((py + dy) & 1) * 2 + ((px + dx) & 1)
                            sh = fdiv2(py + dy) * PITCH + fdiv2(px + dx)
                            D += tile(A4s[iph], sh) @ W["L5"][ky * 4 + kx].T
                    a5 = np.maximum(D + W["b5"], 0)
                    fy, fx = 2 * py + qy, 2 * px + qx
                    for s in range(G):
                        for Y in range(4):
                            for X in range(4):
                                v = a5[pos(s, Y, X)]
                                for ky in range(3):
                                    for kx in range(3):
                                        out[s, 4 * Y + fy + ky, 4 * X + fx + kx] += (v * w6[:, 0, ky, kx]).sum()
    return out[:, 1:17, 1:17] / S + b6


if __name__ == "__main__":
    sd = load_weights("autoencoder_M9_rms_3")
    m = DenoisingAutoencoderCNN2DSingleSubapeture(); m.load_state_dict(sd); m = m.double()
    z = np.load("tests/golden/ref_autoencoder.npz")
    x = z["x"][:G].astype(np.float64) * 30
    with torch.no_grad():
        ref = m(torch.as_tensor(x)[:, None])[:, 0].numpy()
    got = run(x, sd)
    print("model vs torch: max rel err", np.abs(got - ref).max() / np.abs(ref).max())
