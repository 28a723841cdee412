"""`GamblingWorld` of sorrel/examples/iowa/world.py:13-23: two layers, spawn probability from `config.world`."""
from sorrel_b200.examples._common import batch_args, section
from sorrel_b200.worlds import Gridworld


class GamblingWorld(Gridworld):
    LAYERS = 2  # visible Sand floor + the layer walls, decks and agents live on

    def __init__(self, config, default_entity, num_envs: int = 1, device="cuda", seed: int = 0,
                 env_id_base: int = 0):
        w = section(config, "world")
        Gridworld.__init__(self, w.height, w.width, self.LAYERS, default_entity,
                           **batch_args(num_envs, device, seed, env_id_base))
        self.spawn_prob = w.spawn_prob
