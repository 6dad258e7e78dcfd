"""End-to-end construction of a batched simulator from a parameter set: host tables ->
interaction matrix on the GPU -> Btt basis -> filtered command matrix -> GPU context.

Same order as RlSupervisor.__init__ (shesha/supervisor/rlSupervisor.py:112-194):
GenericSupervisor._init_components, compute_modes_to_volts_basis("Btt"), obtain_and_set_cmat_filtered.
"""
from . import calibration, tables as tables_mod
from .config import load_config_from_file
from .init import rtc as rtc_b
from .lib import Simulator
from .rl.layout import RLLayout


def build_tables(parameters, nfilt=0, verbose=False):
    config = load_config_from_file(parameters) if isinstance(parameters, str) else parameters
    t = tables_mod.build_static(config, verbose=verbose)
    t.imat = calibration.measure_imat(t)
    tables_mod.build_basis(t)
    t.nfilt = max(int(nfilt), 0)
    t.cmat = rtc_b.cmat_with_btt(t.imat, t.Btt, t.nfilt)
    return t


def build_system(parameters, n_env, env_rl=None, sac=None, world_size=None, seed=0, tables=None, norm=None,
                 zn_norm=None):
    env_rl = dict(env_rl or {})
    if isinstance(parameters, str):
        env_rl.setdefault("parameters_telescope", parameters)
    t = tables if tables is not None else build_tables(parameters, env_rl.get("n_reverse_filtered_from_cmat", 0))
    rl = RLLayout(t.Btt.shape[1], env_rl, sac, world_size, norm=norm, zn_norm=zn_norm, seed=seed)
    return Simulator(t, n_env, rl), t, rl
