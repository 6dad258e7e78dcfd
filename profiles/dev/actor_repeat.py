"""Repeat the batched actor forward on a fixed state and compare the actions bit for bit (race detector)."""
import sys, copy
import numpy as np
import torch
sys.path.insert(0, ".")
from ao_marl_b200 import tables, calibration
from ao_marl_b200.config import load_config_from_file
from ao_marl_b200.init import rtc as rtc_b
from ao_marl_b200.lib import Simulator
from ao_marl_b200.rl.layout import RLLayout

name = sys.argv[1] if len(sys.argv) > 1 else "production_sh_10x10_2m.py"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t = tables.build_static(load_config_from_file(name)); tables.build_basis(t)
if name.startswith("production_sh_10x10"):
    t.imat = calibration.measure_imat(t); t.cmat = rtc_b.cmat_with_btt(t.imat, t.Btt, 5)
    rl = RLLayout(t.Btt.shape[1], dict(parameters_telescope=name, n_zernike_start_end=[0, 80], n_reverse_filtered_from_cmat=5), None, world_size=3, seed=3)
else:
    rl = RLLayout(t.Btt.shape[1], dict(parameters_telescope=name, n_zernike_start_end=[0, 1260], window_n_zernike=20, include_tip_tilt_windowed=True, n_reverse_filtered_from_cmat=5), None, world_size=44, seed=3)
with torch.no_grad():
    for p in rl.policies:
        p.mean_linear.weight.normal_(0, 0.05); p.log_std_linear.weight.normal_(0, 0.05); p.log_std_linear.bias.fill_(-1.0)
sim = Simulator(t, E, rl)
sim.reset(np.arange(1, E + 1, dtype=np.int64))
st = sim.rows("STATE", rl.state_dim)
g = torch.Generator(device="cuda").manual_seed(1)
st.copy_(torch.randn(st.shape, device="cuda", generator=g))
ref = None; bad = 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
for it in range(n):
    sim.actor_forward(eval_mode=True)
    a = sim.rows("ACTION", rl.action_dim).clone()
    if ref is None: ref = a
    elif not torch.equal(a, ref):
        bad += 1
        d = (a - ref).abs()
        print("iteration", it, "differs: max", float(d.max()), "entries", int((d > 0).sum()), flush=True)
sim.check_device()
print("repetitions", n, "mismatching", bad)
