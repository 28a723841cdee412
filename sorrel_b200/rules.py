"""Declarative transition rules (what `Entity.transition` does in the reference, as data)."""
from __future__ import annotations

from dataclasses import dataclass, field

from .entities import Entity


@dataclass
class BernoulliSpawn:
    """With probability `prob` per turn the cell becomes one of `choices`, uniformly.

    EmptyEntity.transition of the treasurehunt example (examples/treasurehunt/entities.py:69-85):
    `np.random.random() < world.spawn_prob` then `np.random.choice([Gem, Food, Bone])`.
    Kernel rule id SB_RULE_SPAWN."""

    prob: float
    choices: list[Entity]


@dataclass
class BernoulliBecome:
    """With probability `prob` per turn the cell becomes `to` (one uniform, no choice draw).

    River.transition / AppleTree.transition of the cleanup example (examples/cleanup/entities.py:50-54,
    82-90).  Kernel rule id SB_RULE_SPAWN_TO."""

    prob: float
    to: Entity


@dataclass
class BecomeIfLayer:
    """The cell becomes `to` when the cell of the same (y, x) on `layer` holds an entity whose kind string
    is in `kinds`.  Pollution.transition (a CleanBeam on the beam layer) and Apple.transition (an agent on
    the agent layer), examples/cleanup/entities.py:64-71,100-105.  Kernel rule id SB_RULE_BECOME_IF."""

    layer: int
    kinds: list[str]
    to: Entity


@dataclass
class ExpireAfter:
    """The entity survives `sweeps` entity sweeps and becomes `to` at the next one.

    Beam.transition (examples/cleanup/agents.py:199-204): `turn_counter` 0 -> 1 at the first sweep, replaced
    by EmptyEntity at the second.  The per-entity counter becomes part of the kind id: the kind table gets
    `sweeps + 1` rows for the class, chained by SB_RULE_BECOME."""

    sweeps: int
    to: Entity


@dataclass
class ValueRedraw:
    """At every sweep the entity's `value` becomes `value_if` with probability `prob`, else `value_else` (one
    uniform).  Deck.transition -> draw (examples/iowa/entities.py:44-69).  The mutable per-entity value becomes part
    of the kind id: the kind table gets two more rows for the class (same appearance, the two values), all three
    chained by SB_RULE_SPAWN_TO with an else branch; a freshly created entity keeps its constructor value until
    its first sweep, like in the reference."""

    prob: float
    value_if: float
    value_else: float


@dataclass
class SpawnBeam:
    """Effect of a non-movement action: place `entity` on every in-map cell of the beam pattern in front of
    the agent (relative to its facing), on the layer above the agent, unless the cell holds one of
    `blocked_by` (CleanupAgent.spawn_beam, examples/cleanup/agents.py:93-145)."""

    entity: Entity
    radius: int
    blocked_by: list[str] = field(default_factory=lambda: ["Wall"])


# MovingAgent.movement (sorrel/agents/agent.py:199-211): the facing each movement action leaves the agent with
MOVE_FACING = {"up": 0, "down": 2, "left": 3, "right": 1}


def beam_pattern(radius: int) -> list[list[tuple[int, int]]]:
    """(dy, dx) cells of the beam per facing 0 N, 1 E, 2 S, 3 W: `radius` cells straight ahead starting one
    step in front, then the same run shifted one step back on the right flank, then on the left flank
    (cleanup/agents.py:111-123)."""
    ahead = [(-1, 0), (0, 1), (1, 0), (0, -1)]
    out = []
    for facing in range(4):
        fy, fx = ahead[facing]
        ry, rx = ahead[(facing + 1) % 4]  # right-hand side of the facing
        cells = [(fy * i, fx * i) for i in range(1, radius + 1)]
        cells += [(ry + fy * i, rx + fx * i) for i in range(radius)]
        cells += [(-ry + fy * i, -rx + fx * i) for i in range(radius)]
        out.append(cells)
    return out
