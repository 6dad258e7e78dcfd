"""Development probe: warp-specialised sensor kernel against the fused tcgen05 kernel (same tables, same screens)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from ao_marl_b200 import tables
from ao_marl_b200.config import load_config_from_file
from ao_marl_b200.lib import Simulator

name = sys.argv[1] if len(sys.argv) > 1 else "production_sh_10x10_2m.py"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 6
t = tables.build_static(load_config_from_file(name))
sim = Simulator(t, E)
sim.reset(np.arange(1, E + 1, dtype=np.int64))
r = np.random.default_rng(0)
sim.set_dm_volts(torch.as_tensor((r.standard_normal((E, t.nactu)) * 0.3).astype(np.float32), device="cuda"))
out = {}
for path in ("umma", "umma_ws"):
    sim.set_wfs_path(path)
    for rep in range(3):
        sim.comp_wfs_image(atmos=True, dms=True, noise=-1.0)
        sim.do_centroids()
    torch.cuda.synchronize()
    sim.check_device()
    out[path] = sim.rows("SLOPES", t.nslopes).cpu().numpy().copy()
    print(path, sim.wfs_kernel(), "max |slope|", float(np.abs(out[path]).max()), flush=True)
a, b = out["umma_ws"], out["umma"]
print("relerr ws vs umma:", float(np.abs(a - b).max() / np.abs(b).max()), "identical:", bool((a == b).all()))
