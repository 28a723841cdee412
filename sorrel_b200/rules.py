"""Declarative transition rules (what `Entity.transition` does in the reference, as data)."""
from __future__ import annotations

from dataclasses import dataclass

from .entities import Entity


@dataclass
class BernoulliSpawn:
    """With probability `prob` per turn the cell becomes one of `choices`, uniformly.

    EmptyEntity.transition of the treasurehunt example (examples/treasurehunt/entities.py:69-85):
    `np.random.random() < world.spawn_prob` then `np.random.choice([Gem, Food, Bone])`.
    Kernel rule id SB_RULE_SPAWN."""

    prob: float
    choices: list[Entity]
