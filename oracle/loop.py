"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

One-environment numpy restatement of the closed loop the reference sequences in Python:
  RlSupervisor.reset / next_part_one / next_part_two / rl_control   shesha/supervisor/rlSupervisor.py:236-246, 713-733, 784-818, 900-1051
  AoEnv.linear_step / rl_step / state assembly                      src/.../environment/ao_env.py:470-583, 871-939
  TrainerRPC.env_step / divide_rewards_for_agents / choose_action   src/.../rpc_training/train_rpc.py:402-427, 633-732
  GaussianPolicy.sample(only_choosing_action=True)                  src/.../algorithms_rpc/model_rpc.py:131-144
The sutra stages (unpinned) follow SURVEY.md Appendix A.2 and are implemented in oracle/aoframe.py:
  do_control:    err = -cmat . s ; com += gain * err                         (LS integrator)
  apply_control: volt = com (delay 0) or previous com (delay 1); delay line shifts
  do_imat:       D[:, k] = (s(+push_k) - s(-push_k)) / (2 push_k), no noise
"""
import numpy as np

from . import aoframe as af
from . import rng

F32 = np.float32


def wfs_frame(tab, phase, noise, seed, frame, keep=False):
    w = tab["wfs"]
    cube = af.sh_bincube(phase, tab["mpupil"], w)
    cube = af.sh_noise(cube, noise, seed, tab.get("wfs_index", 0), frame)
    s = af.cog(cube, w)
    return (s, cube) if keep else s


def dm_phase(tab, volts):
    n = tab["n"]
    pz, tt = tab["pzt"], tab["tt"]
    sh = af.pzt_shape(volts[:pz["nact"]], pz["influ"], pz["i1"], pz["j1"], pz["dim"])
    ph = af.crop(sh, n, pz["off"]).copy()
    ph += af.crop(af.tt_shape(volts[pz["nact"]:pz["nact"] + 2], tt["influ"]), n, tt["off"])
    return ph


def measure_imat(tab, push_pzt, push_tt):
    nact = tab["pzt"]["nact"]
    nactu = nact + 2
    D = np.zeros((tab["nslopes"], nactu), F32)
    for k in range(nactu):
        p = push_pzt if k < nact else push_tt
        v = np.zeros(nactu, F32)
        v[k] = p
        sp = wfs_frame(tab, dm_phase(tab, v), -1.0, 0, 0)
        sm = wfs_frame(tab, dm_phase(tab, -v), -1.0, 0, 0)
        D[:, k] = (sp - sm) / F32(2 * p)
    return D


class OracleEnv:
    """One environment; `tab` = StaticTables.as_oracle_dict() (+ wfs_index), plus cmat / Btt / P and an
    optional RL layout (ao_marl_b200.rl.layout.RLLayout supplies only integer tables and constants)."""

    def __init__(self, tab, cmat, Btt=None, P=None, rl=None, seed=1234):
        self.tab = tab
        self.cmat = np.asarray(cmat, F32)
        self.Btt = None if Btt is None else np.asarray(Btt, F32)
        self.P = None if P is None else np.asarray(P, F32)
        self.rl = rl
        self.atm = af.OracleAtmos(tab, seed)
        self.gain = F32(tab["gain"])
        self.delay = int(round(tab["delay"]))
        self.closed = True
        self.seed = int(seed)
        self._clear()

    def _clear(self):
        na, ns = self.tab["nactu"], self.tab["nslopes"]
        self.com = np.zeros(na, F32)
        self.com1 = np.zeros(na, F32)
        self.volts = np.zeros(na, F32)
        self.err = np.zeros(na, F32)
        self.slopes = np.zeros(ns, F32)
        self.frame = 0
        self.step_count = 0
        if self.rl is not None:
            self.hist = [np.zeros(self.rl.state_modes, F32) for _ in range(self.rl.n_hist)]
            self.res_modes = np.zeros(self.P.shape[0], F32)

    # -- supervisor stages -------------------------------------------------------------------------
    def reset(self, seed):
        self.seed = int(seed)
        self.atm.reset(seed)
        self._clear()

    def wfs_phase(self, atmos=True, dms=True):
        n = self.tab["n"]
        ph = np.zeros((n, n), F32)
        if atmos:
            ph += af.raytrace_atmos(self.atm, n, self.tab["wfs_xoff"], self.tab["wfs_yoff"])
        if dms:
            ph += dm_phase(self.tab, self.volts)
        return ph

    def comp_wfs_image(self, atmos=True, dms=True, noise=None, keep=False):
        noise = self.tab["wfs"]["noise"] if noise is None else noise
        out = wfs_frame(self.tab, self.wfs_phase(atmos, dms), noise, self.seed, self.frame, keep=keep)
        self.frame += 1
        self.frame_noise = noise
        if keep:
            self._slopes_frame, self.cube = out
        else:
            self._slopes_frame = out
        return out

    def set_bincube(self, cube):
        """Continue from a given detector cube [nvalid, npix, npix] (the tests hand the GPU's photon counts over so
        that a Poisson draw that sat on a rounding boundary does not fork the two closed loops)."""
        self.cube = np.asarray(cube, F32).reshape(self.tab["wfs"]["nvalid"], self.tab["wfs"]["npix"], -1)
        self._slopes_frame = af.cog(self.cube, self.tab["wfs"])

    def do_centroids(self):
        self.slopes = self._slopes_frame.copy()

    def do_control(self):
        self.err = (-(self.cmat.astype(np.float64) @ self.slopes.astype(np.float64))).astype(F32)
        if self.closed:
            self.com = (self.com + self.gain * self.err).astype(F32)

    def apply_control(self):
        self.volts = (self.com1 if self.delay else self.com).copy()
        self.com1 = self.com.copy()

    def rl_control(self, action):
        rl = self.rl
        a = np.asarray(action, F32) * F32(rl.env_rl["normalization_std_inside_environment"]) + \
            F32(rl.env_rl["normalization_mean_inside_environment"])
        m = (self.P.astype(np.float64) @ self.com.astype(np.float64)).astype(F32)
        m[rl.action_map] += a * rl.freedom[rl.action_map]
        self.com = (self.Btt.astype(np.float64) @ m.astype(np.float64)).astype(F32)

    # -- env-level ---------------------------------------------------------------------------------
    def linear_step(self, cube_hook=None):
        """AoEnv.linear_step: returns the normalised state vector.  cube_hook(own_cube) -> cube to continue from."""
        rl = self.rl
        com_before = self.com.copy()
        self.atm.move()
        if cube_hook is None:
            self.comp_wfs_image()
        else:
            self.comp_wfs_image(keep=True)
            self.set_bincube(cube_hook(self.cube))
        self.do_centroids()
        self.do_control()
        if rl is None:
            return None
        before = (self.P.astype(np.float64) @ com_before.astype(np.float64)).astype(F32)[rl.state_map]
        self.res_modes = (self.P.astype(np.float64) @ self.err.astype(np.float64)).astype(F32)
        res = self.res_modes[rl.state_map]
        dm, rs = rl.norm["dm"], rl.norm["dm_residual"]
        blocks = [((h - dm["mean"]) / dm["std"]).astype(F32) for h in self.hist]
        if rl.n_hist:
            self.hist = self.hist[1:] + [before.copy()]
        blocks.append(((before - dm["mean"]) / dm["std"]).astype(F32))
        blocks.append(((res - rs["mean"]) / rs["std"]).astype(F32))
        return np.concatenate(blocks)

    def rewards(self):
        rl = self.rl
        sq = self.res_modes.astype(np.float64) ** 2
        return np.array([-rl.reward_factor * sq[a0:a1].mean() for a0, a1 in rl.agent_reward], F32)

    def actors(self, state, eval_mode=False):
        """All agents' tanh-Gaussian actions scattered into the global action vector (+ the tanh means)."""
        import torch
        rl = self.rl
        action = np.zeros(rl.action_dim, F32)
        mean_v = np.zeros(rl.action_dim, F32)
        for a, pol in enumerate(rl.policies):
            nin = pol.linear1.weight.shape[1]
            nout = pol.mean_linear.weight.shape[0]
            x = torch.tensor(state[rl.agent_idx[a, :nin]][None, :])
            eps = torch.tensor(rng.actor_noise(self.seed, a, self.step_count, nout)[None, :])
            with torch.no_grad():
                act, _, mean = pol.sample(x, only_choosing_action=True, noise=eps)
            slots = rl.agent_act[a, :nout]
            action[slots] = (mean if eval_mode else act).numpy()[0]
            mean_v[slots] = mean.numpy()[0]
        self.step_count += 1
        return action, mean_v

    def env_step(self, action=None, cube_hook=None):
        """TrainerRPC.env_step: rl half-step + reward + linear half-step.  action=None: integrator only."""
        if action is not None:
            self.rl_control(action)
        self.apply_control()
        r = self.rewards() if (self.rl is not None and self.rl.n_agents) else None
        s = self.linear_step(cube_hook)
        return s, r
