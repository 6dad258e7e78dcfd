"""State-normalisation statistics and action bounds from the batched simulator.

Batched counterpart of the reference's `Preprocessor.normalization_loop` / `run_obtain_normalization_and_freedom`
(src/reinforcement_learning/helper_functions/preprocessing/normalization/obtain_normalization.py:139-250, 60-90):
the reference runs 20 seeds x 1000 integrator frames one after the other (hours with the 40x40 file); here the 20
seeds are 20 environments of one context and a frame is one `aom_step(mode=2)`, so the whole loop takes seconds.

Per Btt mode / slope: mean, std, max, min of
    dm          = v2m . rtc.get_command(0)        (obtain_normalization.py:172-176)
    wfs         = rtc.get_slopes(0)               (171)
    dm_residual = v2m . rtc.get_err(0)            (178-182)
and zn_norm = (|max| + |min|) / 2 of the Btt coefficients of the commands (71-73), i.e. the action bound that
`RlSupervisor.load_freedom_parameter_modal_space` divides by `norm_scale_zernike_actions` (rlSupervisor.py:255-282).

    python -m ao_marl_b200.tools.obtain_normalization production_sh_40x40_8m_3layers.py out.npz
"""
import sys

import numpy as np


class RunningStats:
    """mean / std / max / min over the leading dimensions of a stream of [E, n] device tensors."""

    def __init__(self):
        self.n = 0
        self.s = self.s2 = self.mx = self.mn = None

    def add(self, x):
        x = x.double()
        if self.s is None:
            self.s, self.s2 = x.sum(0), (x * x).sum(0)
            self.mx, self.mn = x.max(0).values, x.min(0).values
        else:
            self.s += x.sum(0)
            self.s2 += (x * x).sum(0)
            self.mx = self.mx.max(x.max(0).values)
            self.mn = self.mn.min(x.min(0).values)
        self.n += x.shape[0]

    def result(self):
        mean = self.s / self.n
        var = (self.s2 / self.n - mean * mean).clamp(min=0)
        f = lambda t: t.float().cpu().numpy()
        return dict(mean=f(mean), std=f(var.sqrt()), max=f(self.mx), min=f(self.mn))


def normalization_loop(sim, tables, n_frames=1000, first_seed=1, settle=0):
    """Integrator-only closed loop on every environment of `sim` (seeds first_seed, first_seed + 1, ...).
    Returns ({'dm', 'wfs', 'dm_residual'} -> {'mean', 'std', 'max', 'min'}, zn_norm)."""
    import torch
    E = sim.n_env
    sim.reset(np.arange(E, dtype=np.int64) + int(first_seed))
    P = torch.as_tensor(np.ascontiguousarray(tables.P, dtype=np.float32), device="cuda")      # volts -> modes
    stats = {k: RunningStats() for k in ("dm", "wfs", "dm_residual")}
    for frame in range(settle + n_frames):
        sim.step(mode=2)
        if frame < settle:
            continue
        com = sim.rows("COM", tables.nactu)
        err = sim.rows("ERR", tables.nactu)
        stats["dm"].add(com @ P.T)
        stats["dm_residual"].add(err @ P.T)
        stats["wfs"].add(sim.rows("SLOPES", tables.nslopes))
    norm = {k: v.result() for k, v in stats.items()}
    zn_norm = (np.abs(norm["dm"]["max"]) + np.abs(norm["dm"]["min"])) / 2.0
    return norm, zn_norm.astype(np.float32)


def save(path, norm, zn_norm):
    """Same .npz layout as ao_marl_b200/data/normalization/*.npz (rl/layout.py::load_normalization)."""
    out = {"%s_%s" % (k, s): np.asarray(v[s], np.float32) for k, v in norm.items() for s in ("mean", "std", "max", "min")}
    out["zn_norm"] = np.asarray(zn_norm, np.float32)
    np.savez_compressed(path, **out)


def main(argv):
    from ..lib import Simulator
    from ..system import build_tables
    par, out = argv[0], argv[1]
    n_env = int(argv[2]) if len(argv) > 2 else 20
    n_frames = int(argv[3]) if len(argv) > 3 else 1000
    t = build_tables(par, nfilt=0)
    # the RL layout is not needed for the integrator loop, but aom_step wants the modal tables: a plain context
    # with the command matrix and the basis is enough (state / reward stages are skipped when state_dim == 0)
    sim = Simulator(t, n_env, rl=None)
    norm, zn = normalization_loop(sim, t, n_frames=n_frames)
    save(out, norm, zn)
    print("wrote %s: slopes std %.4f arcsec, first mode std %.4f" % (out, norm["wfs"]["std"].mean(), norm["dm"]["std"][0]))
    sim.close()


if __name__ == "__main__":
    main(sys.argv[1:])
