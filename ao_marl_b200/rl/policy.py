"""SAC actor (tanh-Gaussian MLP) and its packing for the batched CUDA forward.

Architecture and initialisation follow the reference's ``GaussianPolicy``
(src/reinforcement_learning/rpc_training/algorithms_rpc/model_rpc.py:70-158): ``num_layers`` hidden
ReLU layers of ``hidden_dim`` units, two heads (mean, log-std clamped to [-20, LOG_SIG_MAX]), Xavier
uniform weights, zero biases, optionally zeroed head weights.  Parameter names are kept
(``linear1``, ``hidden.N``, ``mean_linear``, ``log_std_linear``) so that the reference's saved actors
(train_rpc.py:1140-1161, ``{'worker_id', 'models_controlled', 'model_state_dict'}``) load directly.
The torch module is the fp32 parity reference of the kernel path and the trainable copy; stepping uses
`pack_actors` + aom_actor_forward.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

LOG_SIG_MIN = -20.0
EPS = 1e-5          # model_rpc.py:13 (epsilon)


class GaussianPolicy(nn.Module):
    def __init__(self, num_inputs, num_actions, hidden_dim=256, num_layers=2, activation="relu",
                 initialize_last_layer_zero=True, initialize_last_layer_near_zero=False,
                 action_scale=1.0, action_bias=0.0, LOG_SIG_MAX=2.0):
        super().__init__()
        if activation not in ("relu", "leaky_relu"):
            raise NotImplementedError(activation)
        self.activation = F.relu if activation == "relu" else F.leaky_relu
        self.linear1 = nn.Linear(num_inputs, hidden_dim)
        self.hidden = nn.ModuleList(nn.Linear(hidden_dim, hidden_dim) for _ in range(num_layers - 1))
        self.mean_linear = nn.Linear(hidden_dim, num_actions)
        self.log_std_linear = nn.Linear(hidden_dim, num_actions)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight, gain=1)
                nn.init.constant_(m.bias, 0)
        with torch.no_grad():
            if initialize_last_layer_zero:
                self.mean_linear.weight.zero_()
                self.log_std_linear.weight.zero_()
            elif initialize_last_layer_near_zero:
                nn.init.xavier_uniform_(self.mean_linear.weight, gain=1e-4)
                nn.init.xavier_uniform_(self.log_std_linear.weight, gain=1e-4)
        self.action_scale = torch.tensor(float(action_scale))
        self.action_bias = torch.tensor(float(action_bias))
        self.LOG_SIG_MAX = LOG_SIG_MAX

    def forward(self, state):
        x = self.activation(self.linear1(state))
        for layer in self.hidden:
            x = self.activation(layer(x))
        mean = self.mean_linear(x)
        log_std = torch.clamp(self.log_std_linear(x), min=LOG_SIG_MIN, max=self.LOG_SIG_MAX)
        return mean, log_std

    def sample(self, state, only_choosing_action=False, noise=None):
        """(action, log_prob, tanh-mean).  `noise` lets a caller inject the N(0,1) draw (parity tests)."""
        mean, log_std = self.forward(state)
        std = log_std.exp()
        eps = torch.randn_like(mean) if noise is None else noise
        x_t = mean + std * eps
        y_t = torch.tanh(x_t)
        action = y_t * self.action_scale + self.action_bias
        log_prob = None
        if not only_choosing_action:
            log_prob = -0.5 * eps.pow(2) - log_std - 0.5 * np.log(2 * np.pi)
            log_prob = log_prob - torch.log(self.action_scale * (1 - y_t.pow(2).clamp(0, 1)) + EPS)
            log_prob = log_prob.sum(1, keepdim=True)
        return action, log_prob, torch.tanh(mean) * self.action_scale + self.action_bias

    def to(self, device):
        self.action_scale = self.action_scale.to(device)
        self.action_bias = self.action_bias.to(device)
        return super().to(device)


def _ld(n):
    return (int(n) + 15) & ~15


def pack_actors(policies, actor_in, hidden, actor_out):
    """Stack per-agent parameters into the zero-padded batched layout of aomarl.h (AOM_T_ACTOR_*)."""
    A = len(policies)
    W1 = np.zeros((A, hidden, _ld(actor_in)), np.float32)
    B1 = np.zeros((A, hidden), np.float32)
    W2 = np.zeros((A, hidden, _ld(hidden)), np.float32)
    B2 = np.zeros((A, hidden), np.float32)
    WH = np.zeros((A, 2 * actor_out, _ld(hidden)), np.float32)
    BH = np.zeros((A, 2 * actor_out), np.float32)
    for a, pol in enumerate(policies):
        if len(pol.hidden) != 1 or pol.activation is not F.relu:
            raise NotImplementedError("the batched actor kernel implements 2 hidden ReLU layers")
        sd = {k: v.detach().cpu().numpy() for k, v in pol.state_dict().items()}
        nin = sd["linear1.weight"].shape[1]
        nout = sd["mean_linear.weight"].shape[0]
        if sd["linear1.weight"].shape[0] != hidden or nin > actor_in or nout > actor_out:
            raise ValueError("Dimension mismatch")
        W1[a, :, :nin] = sd["linear1.weight"]
        B1[a] = sd["linear1.bias"]
        W2[a, :, :hidden] = sd["hidden.0.weight"]
        B2[a] = sd["hidden.0.bias"]
        WH[a, :nout, :hidden] = sd["mean_linear.weight"]
        WH[a, actor_out:actor_out + nout, :hidden] = sd["log_std_linear.weight"]
        BH[a, :nout] = sd["mean_linear.bias"]
        BH[a, actor_out:actor_out + nout] = sd["log_std_linear.bias"]
    return dict(ACTOR_W1=W1, ACTOR_B1=B1, ACTOR_W2=W2, ACTOR_B2=B2, ACTOR_WH=WH, ACTOR_BH=BH)
