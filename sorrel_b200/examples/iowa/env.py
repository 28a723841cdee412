"""Mirrors sorrel/examples/iowa/env.py:30-137."""
import numpy as np

from sorrel_b200.action import ActionSpec
from sorrel_b200.environment import Environment
from sorrel_b200.examples.iowa.agents import GamblingAgent
from sorrel_b200.examples.iowa.entities import Sand, Wall
from sorrel_b200.models import RandomModel
from sorrel_b200.observation import OneHotObservationSpec

ENTITY_LIST = ["EmptyEntity", "Wall", "Sand", "DeckA", "DeckB", "DeckC", "DeckD", "GamblingAgent"]


class GamblingEnv(Environment):
    def __init__(self, world, config, model_factory=None) -> None:
        self.model_factory = model_factory
        super().__init__(world, config)

    def setup_agents(self):
        agents = []
        for _ in range(2):  # env.py:41
            observation_spec = OneHotObservationSpec(ENTITY_LIST, full_view=False,
                                                     vision_radius=self.config.model.agent_vision_radius)
            observation_spec.override_input_size((int(np.prod(observation_spec.input_size)),))
            action_spec = ActionSpec(["up", "down", "left", "right"])
            if self.model_factory is not None:
                model = self.model_factory(observation_spec.input_size, action_spec.n_actions)
            else:  # the reference builds a PyTorchIQN here; the policy network is out of the engine's scope
                params = self.config.model.get("parameters", {})
                model = RandomModel(observation_spec.input_size, action_spec.n_actions,
                                    memory_size=int(params.get("memory_size", 1024)),
                                    n_frames=int(params.get("n_frames", 5)))
            agents.append(GamblingAgent(observation_spec, action_spec, model))
        self.agents = agents

    def populate_environment(self):
        valid_spawn_locations = []
        for index in np.ndindex(self.world.map.shape):
            y, x, z = index
            if y in [0, self.world.height - 1] or x in [0, self.world.width - 1]:
                self.world.add(index, Wall())  # on BOTH layers (env.py:115-117)
            elif z == 0:
                self.world.add(index, Sand())
            elif z == 1:
                valid_spawn_locations.append(index)
        self.place_agents(valid_spawn_locations)
