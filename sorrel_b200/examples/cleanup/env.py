"""Mirrors sorrel/examples/cleanup/env.py:29-140."""
import numpy as np

from sorrel_b200.action import ActionSpec
from sorrel_b200.environment import Environment
from sorrel_b200.examples.cleanup.agents import CleanupAgent, CleanupObservation
from sorrel_b200.examples.cleanup.entities import Apple, AppleTree, River, Sand, Wall
from sorrel_b200.models import RandomModel

ENTITY_LIST = ["EmptyEntity", "Wall", "River", "Pollution", "AppleTree", "Apple", "CleanBeam", "ZapBeam",
               "CleanupAgent"]


class CleanupEnv(Environment):
    def __init__(self, world, config, model_factory=None) -> None:
        self.model_factory = model_factory
        super().__init__(world, config)

    def setup_agents(self):
        agents = []
        cfg = self.config.agent.agent
        for _ in range(cfg.num):
            observation_spec = CleanupObservation(entity_list=ENTITY_LIST, vision_radius=cfg.obs.vision,
                                                  embedding_size=int(cfg.obs.get("embeddings", 3)))
            action_spec = ActionSpec(["up", "down", "left", "right", "clean", "zap"])
            if self.model_factory is not None:
                model = self.model_factory(observation_spec.input_size, action_spec.n_actions)
            else:  # the reference builds a PyTorchIQN here; the policy network is out of the engine's scope
                model = RandomModel(observation_spec.input_size, action_spec.n_actions,
                                    memory_size=int(self.config.model.iqn.parameters.get("memory_size", 1024)),
                                    n_frames=int(cfg.obs.get("n_frames", 5)))
            agents.append(CleanupAgent(observation_spec, action_spec, model, beam_radius=int(cfg.beam_radius)))
        self.agents = agents

    def populate_environment(self):
        spawn_points, apple_spawn_points = [], []
        world = self.world
        for index in np.ndindex(world.map.shape):
            H, W, L = index
            if H in [0, world.height - 1] or W in [0, world.width - 1]:
                world.add(index, Wall())
            elif L == 0 and world.mode == "APPLE":  # env.py:117-123: an orchard everywhere
                world.add(index, AppleTree())
                if ((H % 2) == 0) and ((W % 2) == 0):
                    spawn_points.append((index[0], index[1], world.agent_layer))
                else:
                    apple_spawn_points.append(index)
            elif L == 0:
                if (H > 0 and H < (world.height // 3)) or (
                        H < ((world.height // 3) * 2 - 1) and W in [world.width // 3, 1 + world.width // 3]):
                    world.add(index, River())
                elif H > (world.height - 1 - (world.height // 3)) and H < (world.height - 1):
                    world.add(index, AppleTree())
                    apple_spawn_points.append(index)
                else:
                    world.add(index, Sand())
                    spawn_points.append((index[0], index[1], world.agent_layer))
        # the reference draws ONE placement of apples and agents here; every device instance draws its own
        self.place_items(Apple(), apple_spawn_points, world.initial_apples)
        self.place_agents(spawn_points)
