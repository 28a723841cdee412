"""`CleanupWorld` of sorrel/examples/cleanup/world.py:18-51: three layers (objects, agents, beams) and the
pollution / apple parameters of `config.env`."""
from sorrel_b200 import config as _config
from sorrel_b200.examples._common import batch_args
from sorrel_b200.worlds import Gridworld

# attributes copied verbatim from config.env
ENV_KEYS = ("mode", "full_mdp", "pollution_threshold", "pollution_spawn_chance", "apple_spawn_chance", "initial_apples")


class CleanupWorld(Gridworld):
    object_layer, agent_layer, beam_layer = 0, 1, 2

    def __init__(self, config, default_entity, num_envs: int = 1, device="cuda", seed: int = 0,
                 env_id_base: int = 0):
        self.config = config = _config.wrap(config)
        env = config.env
        # world.py:29: set once and never updated in the reference (`measure_pollution` has no caller), so the
        # AppleTree gate `not world.pollution > world.pollution_threshold` is always open
        self.pollution = 0
        Gridworld.__init__(self, env.height, env.width, env.layers, default_entity,
                           **batch_args(num_envs, device, seed, env_id_base))
        for key in ENV_KEYS:
            setattr(self, key, env[key])
        self.channels = config.agent.agent.obs.channels
        self.max_turns = config.experiment.max_turns
        self.turn = 0

    def measure_pollution(self, env: int = 0) -> float:
        """Polluted share of the river cells of ONE device instance (world.py:41-51)."""
        kinds = self.engine.grid_instance(env)
        count = {n: 0 for n in ("Pollution", "River")}
        for k, proto in enumerate(self._kinds):
            name = proto.__class__.__name__
            if name in count:
                count[name] += int((kinds == k).sum())
        return count["Pollution"] / (count["Pollution"] + count["River"])
