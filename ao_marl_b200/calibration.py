"""Interaction-matrix measurement on the GPU.

Reference: shesha/ao/imats.py:115-173 (imat_init -> sutra ``rtc.do_imat``): every actuator of the
controller's mirrors is pushed by its ``push4imat`` and the noise-free centre-of-gravity slopes are
recorded.  sutra's arithmetic is not in the reference; the convention here is the symmetric
push-pull D[:, k] = (s(+push_k) - s(-push_k)) / (2 push_k), measured through the same fused
Shack-Hartmann kernel the loop uses (one environment per poke, atmosphere off).
"""
import numpy as np


def measure_imat(tables, max_env=8192):
    import torch
    from .lib import Simulator
    pushes = np.concatenate([np.full(tables.p_pzt._ntotact, tables.p_pzt.push4imat, np.float32),
                             np.full(tables.p_tt._ntotact, tables.p_tt.push4imat, np.float32)])
    nactu = len(pushes)
    per = max(1, min(nactu, max_env // 2))
    sim = Simulator(tables, 2 * per, rl=None, atmosphere=False)
    imat = np.zeros((tables.nslopes, nactu), np.float32)
    try:
        for k0 in range(0, nactu, per):
            k1 = min(nactu, k0 + per)
            v = torch.zeros((2 * per, nactu), dtype=torch.float32, device="cuda")
            idx = torch.arange(k0, k1, device="cuda")
            p = torch.as_tensor(pushes[k0:k1], device="cuda")
            v[2 * (idx - k0), idx] = p
            v[2 * (idx - k0) + 1, idx] = -p
            sim.set_dm_volts(v)
            sim.comp_wfs_image(atmos=False, dms=True, noise=-1.0)
            sim.do_centroids()
            s = sim.rows("SLOPES", tables.nslopes)[:2 * (k1 - k0)]
            d = (s[0::2] - s[1::2]) / (2.0 * p[:, None])
            imat[:, k0:k1] = d.T.cpu().numpy()
    finally:
        sim.close()
    return imat
