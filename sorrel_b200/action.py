"""ActionSpec with the reference's interface (sorrel/action/action_spec.py:4-47)."""
from __future__ import annotations


class ActionSpec:
    n_actions: int
    actions: dict[int, str]

    def __init__(self, actions: list[str]):
        self.n_actions = len(actions)
        self.actions = {i: v for i, v in enumerate(actions)}
        self._action_to_index = {v: i for i, v in self.actions.items()}

    def get_readable_action(self, action: int) -> str:
        return self.actions[action]  # KeyError for an unknown index, like the reference

    def get_action_index(self, action_str: str) -> int | None:
        return self._action_to_index.get(action_str)
