"""GPU probe of the tensor-core denoiser: parity against the float32 kernel and the torch module, then timing.
Run on the box: python profiles/dev/denoise_tc_probe.py [n_time_spots]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from ao_marl_b200 import tables
from ao_marl_b200.config import load_config_from_file
from ao_marl_b200.denoiser import Autoencoder
from ao_marl_b200.lib import Simulator

t = tables.build_static(load_config_from_file("production_sh_10x10_2m.py"))
sim = Simulator(t, 2, rl=None)
ae = Autoencoder(dict(type="cnn_single_subaperture", path="autoencoder_M9_rms_3"), device="cuda", sim=sim)
ref = Autoencoder(dict(type="cnn_single_subaperture", path="autoencoder_M9_rms_3"), device="cpu")
gen = torch.Generator().manual_seed(3)
ok = True
for n in (1, 4, 5, 6, 23, 1027):
    for scale in (30.0, 3000.0):
        x = torch.poisson(torch.rand((n, 16, 16), generator=gen) * scale, generator=gen) + torch.randn((n, 16, 16), generator=gen) * 3
        want = ref.model.double()(x.double()[:, None])[:, 0].detach().numpy() if hasattr(ref, "model") else ref.predict(x).numpy()
        sim.set_denoise_path("tcgen05")
        got = ae.predict(x.cuda()).cpu().numpy()
        sim.set_denoise_path("simt")
        got2 = ae.predict(x.cuda()).cpu().numpy()
        e1 = np.abs(got - want).max() / np.abs(want).max()
        e2 = np.abs(got2 - want).max() / np.abs(want).max()
        print("n %5d scale %6.0f  tcgen05 rel err %.3e   simt rel err %.3e" % (n, scale, e1, e2), flush=True)
        ok &= bool(e1 < 1e-4)
sim.check_device()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1200 * 1024
x = torch.poisson(torch.rand((n, 256), device="cuda") * 30) + torch.randn((n, 256), device="cuda") * 3
y = torch.empty_like(x)
for path in ("tcgen05", "simt"):
    sim.set_denoise_path(path)
    sim.denoise(x, out=y)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        sim.denoise(x, out=y)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print("%s: %.2f ms for %d spots = %.1f ns per spot, %.1f TFLOP/s (3.42 MFLOP per spot)" % (path, ms, n, ms * 1e6 / n, 3.42e6 * n / ms / 1e9))
sim.check_device()
print("OK" if ok else "FAILED")
