"""`TagEnv` of sorrel/examples/tag/env.py:25-117 on the batched `Environment` (base `Gridworld`, the `Wall` /
`EmptyEntity` of sorrel_b200.entities)."""
import numpy as np

from sorrel_b200.action import ActionSpec
from sorrel_b200.entities import Wall
from sorrel_b200.environment import Environment
from sorrel_b200.examples.tag.agents import TagAgent
from sorrel_b200.models import RandomModel
from sorrel_b200.observation import OneHotObservationSpec

ENTITY_LIST = ["EmptyEntity", "Wall", "It", "NotIt"]  # env.py:44
ACTIONS = ["up", "down", "left", "right"]


class TagEnv(Environment):
    def __init__(self, world, config, model_factory=None) -> None:
        self.model_factory = model_factory
        super().__init__(world, config)

    def _make_model(self, input_size, n_actions):
        if self.model_factory is not None:
            return self.model_factory(input_size, n_actions)
        m = self.config.model  # the reference builds an IQN here; the policy network is outside the engine
        return RandomModel(input_size, n_actions, memory_size=int(m.get("memory_size", 1024)),
                           n_frames=int(m.get("n_frames", 5)))

    def setup_agents(self):
        cfg = self.config.agent
        self.agents = []
        for _ in range(cfg.num_agents):
            spec = OneHotObservationSpec(ENTITY_LIST, full_view=False, vision_radius=cfg.vision_radius)
            spec.override_input_size((int(np.prod(spec.input_size) + 1), 1))  # + the agent's own `it` flag
            actions = ActionSpec(ACTIONS)
            self.agents.append(TagAgent(spec, actions, self._make_model(spec.input_size, actions.n_actions)))
        # env.py:91-94 draws the It agent once, at construction; every device instance starts from that choice
        chosen = cfg.get("it_agent")
        self.agents[int(np.random.choice(len(self.agents)) if chosen is None else chosen)].it = True

    def populate_environment(self):
        top, right = self.world.height - 1, self.world.width - 1
        free = []
        for index in np.ndindex(self.world.map.shape):
            if index[0] in (0, top) or index[1] in (0, right):
                self.world.add(index, Wall())
            else:
                free.append(index)
        self.place_agents(free)
