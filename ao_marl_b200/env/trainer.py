"""In-process multi-agent rollout loop: TrainerRPC.episode without RPC
(src/reinforcement_learning/rpc_training/train_rpc.py:503-553, 620-648, 706-757): the master process of the
reference asks one remote SAC worker per agent for an action over torch.distributed.rpc and steps one
simulator; here every agent's actor and every environment advance in one aom_step on the device.
"""
import numpy as np

from ..rl.delayed_mdp import DelayedMDP


class BatchedRollout:
    def __init__(self, env, seed=1234):
        self.env = env
        self.sim = env.sim
        self.rl = env.supervisor.rl
        self.seed = seed
        self.total_step = 0
        self.num_episode = 0
        self.transitions = []       # (s, a, s_next, r) device tensors, per-agent views taken lazily

    def divide_states_for_agents(self, state):
        """{worker id: state[:, modes_chosen]} (train_rpc.py:418-427)."""
        import torch
        return {w: state[:, torch.as_tensor(np.asarray(idx, dtype=np.int64), device=state.device)]
                for w, idx in self.rl.modes_chosen.items()}

    def episode(self, steps=None, eval_mode=False, linear_control=False, on_transition=None):
        e = self.rl.env_rl
        steps = e["max_steps_per_episode"] if steps is None else steps
        self.env.set_sim_seed(self.seed)
        s = self.env.reset().clone() if self.env.n_env > 1 else self.sim.rows("STATE", self.rl.state_dim).clone()
        mdp = DelayedMDP(e["delayed_assignment"], e["modification_online"])
        r_total = 0.0
        for _ in range(steps):
            self.sim.step(mode=2 if linear_control else 0, eval_mode=eval_mode)
            a = self.sim.rows("ACTION", self.rl.action_dim).clone()
            r = self.sim.buffer("REWARD").view(self.env.n_env, self.rl.n_agents).clone()
            s_next = self.sim.rows("STATE", self.rl.state_dim).clone()
            if mdp.check_update_possibility():
                s0, a0, s2 = mdp.credit_assignment()
                if on_transition is not None:
                    on_transition(s0, a0, s2, r)
            mdp.save(s, a, s_next)
            r_total += float(r.sum(dim=1).mean())
            s = s_next
            self.total_step += 1
        self.num_episode += 1
        return r_total
