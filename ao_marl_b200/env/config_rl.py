"""RL-side configuration object with the reference's shape: ``.sac``, ``.env_rl``, ``.autoencoder`` dicts
(src/reinforcement_learning/config/GlobalConfig.py:4-157, defaults of parameters*.cfg + the hard-coded
overrides at GlobalConfig.py:61-63, 86, 98, 102-104, 123-139)."""
from ..rl.layout import DEFAULT_ENV_RL, DEFAULT_SAC


class Config:
    def __init__(self, **env_overrides):
        self.cuda = True
        self.algorithm = "SAC"
        self.sac = dict(DEFAULT_SAC, hidden_size_critic=[256], num_layers_critic=2, policy="Gaussian",
                        target_update_interval=1, updates_per_step=1, memory_size=1000000,
                        updates_per_episode_rpc=1000, l2_norm_policy=-1, sac_reward_scaling=1.0)
        self.env_rl = dict(DEFAULT_ENV_RL, verbose=False, move_atmos=True, create_norm_param=False)
        self.env_rl.update(env_overrides)
        self.autoencoder = dict(path=None, type="cnn_single_subaperture")
