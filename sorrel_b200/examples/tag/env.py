"""Mirrors sorrel/examples/tag/env.py:25-117 (base `Gridworld`, `Wall` / `EmptyEntity` of sorrel_b200.entities)."""
import numpy as np

from sorrel_b200.action import ActionSpec
from sorrel_b200.entities import Wall
from sorrel_b200.environment import Environment
from sorrel_b200.examples.tag.agents import TagAgent
from sorrel_b200.models import RandomModel
from sorrel_b200.observation import OneHotObservationSpec


class TagEnv(Environment):
    def __init__(self, world, config, model_factory=None) -> None:
        self.model_factory = model_factory
        super().__init__(world, config)

    def setup_agents(self):
        agents = []
        for _ in range(self.config.agent.num_agents):
            entity_list = ["EmptyEntity", "Wall", "It", "NotIt"]
            observation_spec = OneHotObservationSpec(entity_list, full_view=False,
                                                     vision_radius=self.config.agent.vision_radius)
            input_size = (int(np.prod(observation_spec.input_size) + 1), 1)  # one more input for self.it
            observation_spec.override_input_size(input_size)
            action_spec = ActionSpec(["up", "down", "left", "right"])
            if self.model_factory is not None:
                model = self.model_factory(observation_spec.input_size, action_spec.n_actions)
            else:  # the reference builds a PyTorchIQN here; the policy network is out of the engine's scope
                model = RandomModel(observation_spec.input_size, action_spec.n_actions,
                                    memory_size=int(self.config.model.get("memory_size", 1024)),
                                    n_frames=int(self.config.model.get("n_frames", 5)))
            agents.append(TagAgent(observation_spec, action_spec, model))
        # randomly choose one agent to be "it" (env.py:91-94); every instance starts with this choice
        it_agent_index = int(self.config.agent.get("it_agent", np.random.choice(len(agents))))
        agents[it_agent_index].it = True
        self.agents = agents

    def populate_environment(self):
        valid_spawn_locations = []
        for index in np.ndindex(self.world.map.shape):
            y, x, z = index
            if y in [0, self.world.height - 1] or x in [0, self.world.width - 1]:
                self.world.add(index, Wall())
            else:
                valid_spawn_locations.append(index)
        self.place_agents(valid_spawn_locations)
