"""`TreasurehuntWorld` of sorrel/examples/treasurehunt/world.py:13-29: two layers, item values and the spawn
probability read from `config.world`."""
from sorrel_b200.examples._common import batch_args, section
from sorrel_b200.worlds import Gridworld

ITEM_KEYS = ("gem", "food", "bone")  # config.world.<key>_value


class TreasurehuntWorld(Gridworld):
    LAYERS = 2  # floor (Sand) + the layer walls, items and agents live on

    def __init__(self, config, default_entity, num_envs: int = 1, device="cuda", seed: int = 0,
                 env_id_base: int = 0):
        w = section(config, "world")
        Gridworld.__init__(self, w.height, w.width, self.LAYERS, default_entity,
                           **batch_args(num_envs, device, seed, env_id_base))
        self.spawn_prob = w.spawn_prob
        self.values = {key: w[f"{key}_value"] for key in ITEM_KEYS}
