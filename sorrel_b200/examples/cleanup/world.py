"""Mirrors sorrel/examples/cleanup/world.py:18-51."""
from sorrel_b200 import config as _config
from sorrel_b200.worlds import Gridworld


class CleanupWorld(Gridworld):
    def __init__(self, config, default_entity, num_envs: int = 1, device="cuda", seed: int = 0,
                 env_id_base: int = 0):
        config = _config.wrap(config)
        self.config = config
        self.channels = config.agent.agent.obs.channels
        self.full_mdp = config.env.full_mdp
        self.object_layer = 0
        self.agent_layer = 1
        self.beam_layer = 2
        # world.py:29: set once and never updated in the reference (`measure_pollution` has no caller), so
        # AppleTree's `not world.pollution > world.pollution_threshold` gate is always open
        self.pollution = 0
        super().__init__(config.env.height, config.env.width, config.env.layers, default_entity,
                         num_envs=num_envs, device=device, seed=seed, env_id_base=env_id_base)
        self.mode = config.env.mode
        self.max_turns = config.experiment.max_turns
        self.pollution_threshold = config.env.pollution_threshold
        self.pollution_spawn_chance = config.env.pollution_spawn_chance
        self.apple_spawn_chance = config.env.apple_spawn_chance
        self.initial_apples = config.env.initial_apples
        self.turn = 0

    def measure_pollution(self, env: int = 0) -> float:
        """world.py:41-51 for one device instance."""
        names = [k.__class__.__name__ for k in self._kinds]
        kinds = self.engine.grid_numpy()[env]
        pollution = sum(int((kinds == i).sum()) for i, n in enumerate(names) if n == "Pollution")
        river = sum(int((kinds == i).sum()) for i, n in enumerate(names) if n == "River")
        return pollution / (pollution + river)
