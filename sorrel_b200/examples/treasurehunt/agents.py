"""Mirrors sorrel/examples/treasurehunt/agents.py:15-69 (pov / get_action / act / is_done come from
the batched `MovingAgent`; the overrides below are the reference's, tensor-valued)."""
from sorrel_b200.agents import MovingAgent


class TreasurehuntAgent(MovingAgent):
    def __init__(self, observation_spec, action_spec, model):
        super().__init__(observation_spec, action_spec, model)

    def reset(self) -> None:
        self.model.reset()

    def is_done(self, world) -> bool:
        return world.is_done
