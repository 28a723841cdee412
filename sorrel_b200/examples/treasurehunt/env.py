"""`TreasurehuntEnv` of sorrel/examples/treasurehunt/env.py:24-147 on the batched `Environment`.

Same two hooks as the reference -- `setup_agents()` builds one agent slot per `config.model.num_agents`,
`populate_environment()` lays out walls / sand and names the cells agents may start on -- but the agent placement
itself is drawn per instance on the device (`place_agents`), and the policy is pluggable (`model_factory`)."""
import numpy as np

from sorrel_b200.action import ActionSpec
from sorrel_b200.environment import Environment
from sorrel_b200.examples.treasurehunt.agents import TreasurehuntAgent
from sorrel_b200.examples.treasurehunt.entities import Sand, Wall
from sorrel_b200.models import RandomModel
from sorrel_b200.observation import OneHotObservationSpec, RGBObservationSpec

ENTITY_LIST = ["EmptyEntity", "Wall", "Gem", "Bone", "Food", "TreasurehuntAgent"]  # env.py:43-50: channel order
ACTIONS = ["up", "down", "left", "right"]  # env.py:79
SPEC_TYPES = {"onehot": OneHotObservationSpec, "rgb": RGBObservationSpec}  # config.model.observation_spec


class TreasurehuntEnv(Environment):
    def __init__(self, world, config, model_factory=None) -> None:
        self.model_factory = model_factory  # (input_size, n_actions) -> BaseModel; default: RandomModel
        super().__init__(world, config)

    def _make_model(self, input_size, n_actions):
        if self.model_factory is not None:
            return self.model_factory(input_size, n_actions)
        m = self.config.model  # the reference builds an IQN here; the policy network is outside the engine
        return RandomModel(input_size, n_actions, memory_size=int(m.get("memory_size", 1024)),
                           n_frames=int(m.get("n_frames", 5)))

    def setup_agents(self):
        m = self.config.model
        kind = m.get("observation_spec", "onehot")
        if kind not in SPEC_TYPES:
            raise ValueError(f"Unknown observation spec type: {kind}")
        self.agents = []
        for _ in range(int(m.get("num_agents", 2))):
            spec = SPEC_TYPES[kind](ENTITY_LIST, full_view=False, vision_radius=m.agent_vision_radius)
            spec.override_input_size((int(np.prod(spec.input_size)),))  # the model sees the flattened view
            actions = ActionSpec(ACTIONS)
            self.agents.append(TreasurehuntAgent(spec, actions, self._make_model(spec.input_size, actions.n_actions)))

    def populate_environment(self):
        top, right = self.world.height - 1, self.world.width - 1
        spawn_cells = []
        for y, x, z in np.ndindex(self.world.map.shape):
            on_border = y in (0, top) or x in (0, right)
            if z == 0:
                self.world.add((y, x, z), Sand())  # floor everywhere, also under the walls (env.py:126-132)
            elif on_border:
                self.world.add((y, x, z), Wall())
            else:
                spawn_cells.append((y, x, z))
        self.place_agents(spawn_cells)  # env.py:141-147, drawn per instance at every reset
