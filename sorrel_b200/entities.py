"""Entity classes with the reference's attributes (sorrel/entities/entity.py:9-68,
sorrel/entities/basic_entities.py:13-47).

In the reference every grid cell holds a Python `Entity` object and `transition(world)` is called
on each of them every turn.  Here a cell is a uint8 *kind id*; an `Entity` subclass is a row of the
kind table (value, passable, transition rule, appearance) and instances only serve as prototypes
when a world is described (`world.add(location, Wall())`) or read back (`world.observe(location)`).

A kind's per-turn physics is declared, not coded: override `transition_rule(world)` to return a
rule object from `sorrel_b200.rules`.  `transition()` is kept for signature compatibility; the
engine never calls Python per cell.
"""
from __future__ import annotations

from typing import Optional


class Entity:
    _location: Optional[tuple]
    value: float
    passable: bool
    has_transitions: bool
    kind: str

    def __init__(self):
        self._location = None
        self.value = 0  # entity.py:31
        self.passable = False  # entity.py:32-34
        self.has_transitions = False
        self.kind = str(self)

    def __str__(self):
        return str(self.__class__.__name__)

    def __repr__(self):
        return f"{self.__class__.__name__}(value={self.value})"

    @property
    def location(self) -> tuple:
        if self._location is None:  # entity.py:51-52
            raise AttributeError(f"{self.kind} location is None.")
        return self._location

    @location.setter
    def location(self, value: tuple):
        self._location = value

    def transition(self, world) -> None:
        """entity.py:60-68: no-op by default (kept for API compatibility)."""

    def transition_rule(self, world):
        """Declarative replacement of `transition()`: return a rule from sorrel_b200.rules or None."""
        return None

    # identity of a kind-table row: class + the attributes the kernels read
    def kind_key(self):
        return (self.__class__.__name__, str(self.kind), float(self.value), bool(self.passable))


class Wall(Entity):
    """basic_entities.py:13-22"""

    def __init__(self):
        super().__init__()
        self.value = -1


class EmptyEntity(Entity):
    """basic_entities.py:25-34"""

    def __init__(self):
        super().__init__()
        self.passable = True


class Gem(Entity):
    """basic_entities.py:37-47"""

    def __init__(self, value: float | int):
        super().__init__()
        self.passable = True
        self.value = value
