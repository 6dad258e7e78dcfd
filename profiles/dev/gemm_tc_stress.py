"""Stress of the tcgen05 TF32x3 GEMM (gemm_tc_kernel) for rare precision outliers: many small products against float64."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ao_marl_b200 import tables
from ao_marl_b200.config import load_config_from_file
from ao_marl_b200.lib import Simulator, LD

t = tables.build_static(load_config_from_file("production_sh_10x10_2m.py"))
sim = Simulator(t, 4, rl=None)
sim.set_gemm_path("tcgen05")
g = torch.Generator(device="cuda").manual_seed(0)
worst, bad = 0.0, 0
for (M, N, K) in ((3, 256, 328), (3, 256, 256), (3, 164, 256), (4096, 256, 280), (300, 90, 128)):
    A = torch.zeros((M, LD(K)), device="cuda"); Bm = torch.zeros((N, LD(K)), device="cuda")
    for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 300):
        A[:, :K] = torch.randn((M, K), device="cuda", generator=g)
        Bm[:, :K] = torch.randn((N, K), device="cuda", generator=g)
        C = sim.gemm_tn(A, Bm)[:, :N].double()
        ref = A[:, :K].double() @ Bm[:, :K].double().T
        e = float((C - ref).abs().max() / ref.abs().max())
        worst = max(worst, e)
        if e > 5e-5:
            bad += 1
            print("outlier", (M, N, K), it, e, flush=True)
    print((M, N, K), "worst so far %.3e, outliers %d" % (worst, bad), flush=True)
sim.check_device()
print("OK" if bad == 0 else "OUTLIERS %d" % bad)
