"""Mirrors sorrel/examples/tag/agents.py:15-110.  `it` is a per-instance state on the device (`Engine.agent_flag`);
the attribute here is the value every instance starts with."""
from sorrel_b200.agents import MovingAgent


class TagAgent(MovingAgent):
    state_kinds = ("NotIt", "It")  # `kind` follows `it` (agents.py:44-52): the agent looks different to the others
    touch_swap = True  # agents.py:84-99: an It agent next to a NotIt agent after its move hands the role over
    observe_flag = True  # agents.py:55-65: [self.it] is appended to the flattened view

    def __init__(self, observation_spec, action_spec, model, reward_per_turn=10):
        super().__init__(observation_spec, action_spec, model)
        self.reward_per_turn = reward_per_turn
        self._it = False
        self.kind = "NotIt"

    @property
    def it(self):
        return self._it

    @it.setter
    def it(self, value: bool):
        self._it = value
        self.kind = "It" if value else "NotIt"

    @property
    def flag(self) -> bool:
        return self._it

    @property
    def reward_by_state(self):
        """agents.py:101-105: every agent that is not It after its move earns reward_per_turn."""
        return {"NotIt": self.reward_per_turn, "It": 0}

    def reset(self) -> None:
        self.model.reset()

    def is_done(self, world) -> bool:
        return world.is_done
