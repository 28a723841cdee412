"""Mirrors sorrel/examples/iowa/agents.py:15-65 (the per-agent `encounters` tally is host-side bookkeeping of the
reference's logger and is not kept on the device)."""
from sorrel_b200.agents import MovingAgent


class GamblingAgent(MovingAgent):
    def __init__(self, observation_spec, action_spec, model):
        super().__init__(observation_spec, action_spec, model)

    def reset(self) -> None:
        self.model.reset()

    def is_done(self, world) -> bool:
        return world.is_done
