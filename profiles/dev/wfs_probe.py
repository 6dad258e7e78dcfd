"""Development probe: staged Shack-Hartmann kernel vs the float32 FFT kernel on the 10x10 geometry."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ao_marl_b200 import tables
from ao_marl_b200.config import load_config_from_file
from ao_marl_b200.lib import Simulator

mode = sys.argv[1] if len(sys.argv) > 1 else "atmos"
t = tables.build_static(load_config_from_file("production_sh_10x10_2m.py"))
sim = Simulator(t, 4)
sim.reset(np.array([1, 2, 3, 4], dtype=np.int64))
r = np.random.default_rng(0)
sim.set_dm_volts(torch.as_tensor((r.standard_normal((4, t.nactu)) * 10).astype(np.float32), device="cuda"))
print("kernel:", sim.wfs_kernel(), sim.lib.aom_last_error(sim._ctx))
out = {}
paths = ("simt",) + tuple(sys.argv[2:] or ["tensor"])
for path in paths:
    sim.set_wfs_path(path)
    sim.comp_wfs_image(atmos=(mode == "atmos"), dms=True, keep_image=True, noise=-1.0)
    sim.do_centroids()
    torch.cuda.synchronize()
    out[path] = (sim.rows("SLOPES", t.nslopes).cpu().numpy().copy(), sim.buffer("BINCUBE").cpu().numpy().copy())
    print(path, sim.wfs_kernel(), "ok", float(np.abs(out[path][0]).max()))
sim.check_device()
for path in paths[1:]:
    for i in (0, 1):
        a, b = out[path][i], out["simt"][i]
        print(path, "relerr", "slopes" if i == 0 else "cube", float(np.abs(a - b).max() / np.abs(b).max()))
