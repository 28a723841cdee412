"""Mirrors sorrel/examples/treasurehunt/env.py:24-147."""
import numpy as np

from sorrel_b200.action import ActionSpec
from sorrel_b200.environment import Environment
from sorrel_b200.examples.treasurehunt.agents import TreasurehuntAgent
from sorrel_b200.examples.treasurehunt.entities import Sand, Wall
from sorrel_b200.models import RandomModel
from sorrel_b200.observation import OneHotObservationSpec, RGBObservationSpec


class TreasurehuntEnv(Environment):
    def __init__(self, world, config, model_factory=None) -> None:
        self.model_factory = model_factory
        super().__init__(world, config)

    def setup_agents(self):
        agent_num = int(self.config.model.get("num_agents", 2))
        agents = []
        for _ in range(agent_num):
            entity_list = ["EmptyEntity", "Wall", "Gem", "Bone", "Food", "TreasurehuntAgent"]
            obs_spec_type = self.config.model.get("observation_spec", "onehot")
            if obs_spec_type == "onehot":
                observation_spec = OneHotObservationSpec(
                    entity_list, full_view=False, vision_radius=self.config.model.agent_vision_radius)
            elif obs_spec_type == "rgb":
                observation_spec = RGBObservationSpec(
                    entity_list, full_view=False, vision_radius=self.config.model.agent_vision_radius)
            else:
                raise ValueError(f"Unknown observation spec type: {obs_spec_type}")
            observation_spec.override_input_size((int(np.prod(observation_spec.input_size)),))
            action_spec = ActionSpec(["up", "down", "left", "right"])
            if self.model_factory is not None:
                model = self.model_factory(observation_spec.input_size, action_spec.n_actions)
            else:  # the reference builds an IQN here; the policy network is out of the engine's scope
                model = RandomModel(observation_spec.input_size, action_spec.n_actions,
                                    memory_size=int(self.config.model.get("memory_size", 1024)),
                                    n_frames=int(self.config.model.get("n_frames", 5)))
            agents.append(TreasurehuntAgent(observation_spec, action_spec, model))
        self.agents = agents

    def populate_environment(self):
        valid_spawn_locations = []
        for index in np.ndindex(self.world.map.shape):
            y, x, z = index
            if (y in [0, self.world.height - 1] or x in [0, self.world.width - 1]) and z == 1:
                self.world.add(index, Wall())
            elif z == 0:
                self.world.add(index, Sand())
            elif z == 1:
                valid_spawn_locations.append(index)
        # the reference draws ONE placement here; every device instance draws its own at reset
        self.place_agents(valid_spawn_locations)
