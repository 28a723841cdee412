"""Mirrors sorrel/examples/iowa/world.py:13-23."""
from sorrel_b200 import config as _config
from sorrel_b200.worlds import Gridworld


class GamblingWorld(Gridworld):
    def __init__(self, config, default_entity, num_envs: int = 1, device="cuda", seed: int = 0,
                 env_id_base: int = 0):
        layers = 2
        config = _config.wrap(config)
        super().__init__(config.world.height, config.world.width, layers, default_entity, num_envs=num_envs,
                         device=device, seed=seed, env_id_base=env_id_base)
        self.spawn_prob = config.world.spawn_prob
