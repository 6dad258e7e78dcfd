"""Development measurement: batched SAC updates per second for the 43-agent 40x40 layout (config 5 learner side)."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from ao_marl_b200.rl.layout import RLLayout
from ao_marl_b200.rl.sac import BatchedSAC

rl = RLLayout(1283, dict(parameters_telescope="production_sh_40x40_8m_3layers.py", n_zernike_start_end=[0, 1260],
                         window_n_zernike=20, include_tip_tilt_windowed=True, n_reverse_filtered_from_cmat=5,
                         delayed_assignment=2), None, world_size=44, seed=0)
L = BatchedSAC.from_layout(rl, device="cuda", seed=1, memory_size=65536)
A, IN, ACT = L.A, L.IN, L.ACT
g = torch.Generator(device="cuda").manual_seed(0)
n = 8192
L.memory.push(torch.randn(A, n, IN, device="cuda", generator=g), torch.tanh(torch.randn(A, n, ACT, device="cuda", generator=g)),
              -torch.rand(A, n, device="cuda", generator=g), torch.randn(A, n, IN, device="cuda", generator=g),
              torch.ones(A, n, device="cuda"))
for _ in range(5):
    L.update()
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 50
for _ in range(K):
    L.update()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("agents %d, batch %d: %.1f updates/s of all agents (%.2f ms per update, %.0f agent-updates/s)"
      % (A, L.batch_size, K / dt, dt / K * 1e3, K * A / dt))
