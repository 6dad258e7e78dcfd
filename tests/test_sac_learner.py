"""Batched SAC learner (ao_marl_b200/rl/sac.py) against a per-agent restatement of the reference's update
(train_rpc.py:987-1133 with QNetwork / GaussianPolicy modules and torch.optim.Adam), the device replay ring, and
the gradient all-reduce over a 2-rank gloo group.  CPU only."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from ao_marl_b200.rl.policy import GaussianPolicy
from ao_marl_b200.rl.sac import BatchedSAC, DeviceReplay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAC = dict(hidden_size_actor=32, num_layers_actor=2, hidden_size_critic=24, num_layers_critic=2, gamma=0.1,
           batch_size=16, lr=3e-3, tau=0.05, alpha=0.2, automatic_entropy_tuning=True, LOG_SIG_MAX=2.0,
           gaussian_std=1.0, gaussian_mu=0.0, initialize_last_layer_0=False, target_update_interval=1)


class QNet(nn.Module):
    """Restatement of the reference QNetwork for the scalar hidden_dim case (model_rpc.py:17-69)."""

    def __init__(self, nin, nact, hidden, layers):
        super().__init__()
        self.Q1_input, self.Q1_output = nn.Linear(nin + nact, hidden), nn.Linear(hidden, 1)
        self.Q2_input, self.Q2_output = nn.Linear(nin + nact, hidden), nn.Linear(hidden, 1)
        self.hidden_Q1 = nn.ModuleList(nn.Linear(hidden, hidden) for _ in range(layers - 1))
        self.hidden_Q2 = nn.ModuleList(nn.Linear(hidden, hidden) for _ in range(layers - 1))

    def forward(self, s, a):
        x = torch.cat([s, a], 1)
        x1, x2 = F.relu(self.Q1_input(x)), F.relu(self.Q2_input(x))
        for h1, h2 in zip(self.hidden_Q1, self.hidden_Q2):
            x1, x2 = F.relu(h1(x1)), F.relu(h2(x2))
        return self.Q1_output(x1), self.Q2_output(x2)


class RefAgent:
    def __init__(self, learner, a):
        nin, nact = learner.in_dims[a], learner.act_dims[a]
        self.policy = GaussianPolicy(nin, nact, hidden_dim=32, num_layers=2, initialize_last_layer_zero=False)
        self.policy.load_state_dict(learner.actor_state_dict(a))
        self.critic, self.critic_target = QNet(nin, nact, 24, 2), QNet(nin, nact, 24, 2)
        self.critic.load_state_dict(learner.critic_state_dict(a))
        self.critic_target.load_state_dict(learner.critic_state_dict(a, target=True))
        self.log_alpha = torch.zeros(1, requires_grad=True)
        self.alpha = 0.2          # reference starts from config alpha until the first temperature update
        self.po, self.co = torch.optim.Adam(self.policy.parameters(), lr=3e-3), torch.optim.Adam(self.critic.parameters(), lr=3e-3)
        self.ao = torch.optim.Adam([self.log_alpha], lr=3e-3)
        self.target_entropy = -float(nact)

    def update(self, s, a, r, s2, m, n0, n1):
        r, m = r[:, None], m[:, None]
        with torch.no_grad():
            a2, lp2, _ = self.policy.sample(s2, noise=n0)
            q1t, q2t = self.critic_target(s2, a2)
            nq = r + m * 0.1 * (torch.min(q1t, q2t) - self.alpha * lp2)
        q1, q2 = self.critic(s, a)
        loss = F.mse_loss(q1, nq) + F.mse_loss(q2, nq)
        self.co.zero_grad(); loss.backward(); self.co.step()
        pi, lp, _ = self.policy.sample(s, noise=n1)
        q1p, q2p = self.critic(s, pi)
        pl = (self.alpha * lp - torch.min(q1p, q2p)).mean()
        self.po.zero_grad(); pl.backward(); self.po.step()
        al = -(self.log_alpha * (lp + self.target_entropy).detach()).mean()
        self.ao.zero_grad(); al.backward(); self.ao.step()
        self.alpha = self.log_alpha.exp().detach()
        with torch.no_grad():
            for t, p in zip(self.critic_target.parameters(), self.critic.parameters()):
                t.copy_(t * (1 - 0.05) + p * 0.05)


def _batch(learner, B, seed):
    g = torch.Generator().manual_seed(seed)
    A, IN, ACT = learner.A, learner.IN, learner.ACT
    s = torch.randn(A, B, IN, generator=g) * learner.in_mask[:, None, :]
    s2 = torch.randn(A, B, IN, generator=g) * learner.in_mask[:, None, :]
    a = torch.tanh(torch.randn(A, B, ACT, generator=g)) * learner.act_mask[:, None, :]
    r = -torch.rand(A, B, generator=g)
    m = torch.ones(A, B)
    n0, n1 = torch.randn(A, B, ACT, generator=g), torch.randn(A, B, ACT, generator=g)
    return (s, a, r, s2, m), (n0, n1)


def test_batched_update_matches_per_agent_reference():
    in_dims, act_dims = [12, 7, 12], [5, 2, 3]
    L = BatchedSAC(in_dims, act_dims, dict(SAC, alpha=0.2), device="cpu", seed=4, memory_size=64)
    # the reference uses config alpha for the very first update and exp(log_alpha) afterwards: start equal
    refs = [RefAgent(L, a) for a in range(L.A)]
    for r in refs:
        r.alpha = 1.0
    for it in range(4):
        batch, noise = _batch(L, 16, 100 + it)
        L.update(batch, noise)
        for a, ref in enumerate(refs):
            s, act, r, s2, m = [x[a] for x in batch]
            nin, nact = in_dims[a], act_dims[a]
            ref.update(s[:, :nin], act[:, :nact], r, s2[:, :nin], m, noise[0][a][:, :nact], noise[1][a][:, :nact])
    for a, ref in enumerate(refs):
        for k, v in L.actor_state_dict(a).items():
            assert torch.allclose(v, ref.policy.state_dict()[k], rtol=2e-4, atol=2e-6), (a, k)
        for k, v in L.critic_state_dict(a).items():
            assert torch.allclose(v, ref.critic.state_dict()[k], rtol=2e-4, atol=2e-6), (a, k)
        for k, v in L.critic_state_dict(a, target=True).items():
            assert torch.allclose(v, ref.critic_target.state_dict()[k], rtol=2e-4, atol=2e-6), (a, k)
        assert abs(float(L.log_alpha[a].detach()) - float(ref.log_alpha.detach())) < 1e-5
    # padding stays inert
    assert float((L.actor.W[0][1][:, 7:]).abs().max()) == 0.0
    assert float((L.head.W[0][1][2:L.ACT]).abs().max()) == 0.0


def test_replay_ring_and_checkpoint_format():
    mem = DeviceReplay(2, 5, 3, 2, "cpu")
    for t in range(4):
        n = 2
        base = torch.full((2, n, 3), float(t))
        mem.push(base, torch.zeros(2, n, 2), torch.full((2, n), float(t)), base + 0.5, torch.ones(2, n))
    assert len(mem) == 5 and mem.position == 3
    # slots 0..2 were overwritten by t = 2 (second half) and t = 3, slots 3, 4 still hold t = 1, 2
    assert mem.r[0].tolist() == [2.0, 3.0, 3.0, 1.0, 2.0]
    s, a, r, s2, m = mem.sample(8, torch.Generator().manual_seed(0))
    assert s.shape == (2, 8, 3) and torch.equal(s2, s + 0.5) and torch.equal(r, s[..., 0])
    L = BatchedSAC([6, 4], [3, 2], SAC, device="cpu", seed=1, memory_size=8)
    ck = L.checkpoint(1, worker_id=2, modes_controlled=[30, 32])
    assert set(ck) == {"worker_id", "models_controlled", "model_state_dict"}
    pol = GaussianPolicy(4, 2, hidden_dim=32, num_layers=2)
    pol.load_state_dict(ck["model_state_dict"])          # reference-format actor loads into the per-agent module
    x = torch.randn(5, 4)
    mean, _ = pol(x)
    xm = torch.zeros(2, 5, 6)
    xm[1, :, :4] = x
    assert torch.allclose(L.policy_forward(xm)[0][1, :, :2], mean, atol=1e-6)
    packed = L.pack_for_sim()
    assert packed["ACTOR_W1"].shape == (2, 32, 16) and packed["ACTOR_WH"].shape == (2, 6, 32)


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
from ao_marl_b200.rl.sac import BatchedSAC
from test_sac_learner import SAC, _batch
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
L = BatchedSAC([9, 5], [4, 2], SAC, device="cpu", seed=7, dist=dist, memory_size=8)
for it in range(3):
    (s, a, r, s2, m), (n0, n1) = _batch(L, 16, 50 + it)
    h = slice(0, 8) if dist.get_rank() == 0 else slice(8, 16)          # each rank sees half of the pooled batch
    L.update((s[:, h], a[:, h], r[:, h], s2[:, h], m[:, h]), (n0[:, h], n1[:, h]))
torch.save({"actor": L.actor_state_dict(0), "critic": L.critic_state_dict(1), "la": L.log_alpha.detach()}, sys.argv[2])
dist.destroy_process_group()
"""


def test_gradient_allreduce_gloo_world2(tmp_path):
    """Two ranks, each on half of a batch, end with identical weights equal to one process on the whole batch."""
    port = 29000 + os.getpid() % 2000
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, port=port))
    outs = [str(tmp_path / ("r%d.pt" % r)) for r in range(2)]
    procs = [subprocess.Popen([sys.executable, str(script), str(r), outs[r]], cwd=ROOT) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    r0, r1 = torch.load(outs[0]), torch.load(outs[1])
    L = BatchedSAC([9, 5], [4, 2], SAC, device="cpu", seed=7, memory_size=8)
    for it in range(3):
        batch, noise = _batch(L, 16, 50 + it)
        L.update(batch, noise)
    for k in r0["actor"]:
        assert torch.equal(r0["actor"][k], r1["actor"][k])                    # replicas stay bit-identical
        assert torch.allclose(r0["actor"][k], L.actor_state_dict(0)[k], rtol=1e-4, atol=1e-6), k
    for k in r0["critic"]:
        assert torch.equal(r0["critic"][k], r1["critic"][k])
        assert torch.allclose(r0["critic"][k], L.critic_state_dict(1)[k], rtol=1e-4, atol=1e-6), k
    assert torch.allclose(r0["la"], L.log_alpha.detach(), atol=1e-6)


NCCL_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
from ao_marl_b200.rl.sac import BatchedSAC
from test_sac_learner import SAC, _batch
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = BatchedSAC([9, 5], [4, 2], SAC, device="cuda", seed=7, dist=dist, memory_size=8)
ref = BatchedSAC([9, 5], [4, 2], SAC, device="cuda", seed=7, memory_size=8)
for it in range(3):
    cpu = BatchedSAC([9, 5], [4, 2], SAC, device="cpu", seed=7, memory_size=8) if it == 0 else cpu
    (s, a, r, s2, m), (n0, n1) = _batch(cpu, 16, 50 + it)
    full = [x.cuda() for x in (s, a, r, s2, m)]
    noise = [x.cuda() for x in (n0, n1)]
    h = slice(0, 8) if rank == 0 else slice(8, 16)
    L.update(tuple(x[:, h] for x in full), tuple(x[:, h] for x in noise))
    ref.update(tuple(full), tuple(noise))
flat = torch.cat([p.detach().reshape(-1) for p in L.actor_params + L.critic_params] + [L.log_alpha.detach()])
gathered = [torch.empty_like(flat) for _ in range(2)]
dist.all_gather(gathered, flat)
flat_ref = torch.cat([p.detach().reshape(-1) for p in ref.actor_params + ref.critic_params] + [ref.log_alpha.detach()])
ok_same = bool(torch.equal(gathered[0], gathered[1]))
err = float((flat - flat_ref).abs().max() / flat_ref.abs().max())
if rank == 0:
    print("NCCL_SAC same=%%s err=%%.3e" %% (ok_same, err))
    assert ok_same and err < 2e-4, (ok_same, err)
dist.destroy_process_group()
"""


@pytest.mark.gpu
def test_gradient_allreduce_nccl_two_gpus(tmp_path):
    """Same check as the gloo test on two real GPUs over NCCL (skipped on a single-GPU box)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    script = tmp_path / "nccl_worker.py"
    script.write_text(NCCL_WORKER % dict(root=ROOT))
    port = 29500 + os.getpid() % 400
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "NCCL_SAC same=True" in r.stdout
