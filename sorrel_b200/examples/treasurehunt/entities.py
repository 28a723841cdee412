"""Mirrors sorrel/examples/treasurehunt/entities.py:14-85; the spawn physics is declared as a rule."""
from sorrel_b200.entities import Entity
from sorrel_b200.rules import BernoulliSpawn


class Wall(Entity):
    def __init__(self):
        super().__init__()
        self.value = -1  # Walls penalize contact


class Sand(Entity):
    def __init__(self):
        super().__init__()
        self.passable = True
        self.kind = "EmptyEntity"  # entities.py:31: looks like empty space to the observation


class Gem(Entity):
    def __init__(self, value):
        super().__init__()
        self.passable = True
        self.value = value


class Food(Gem):
    pass


class Bone(Gem):
    pass


class EmptyEntity(Entity):
    def __init__(self):
        super().__init__()
        self.passable = True
        self.has_transitions = True

    def transition_rule(self, world):
        """entities.py:69-85: with probability world.spawn_prob become Gem / Food / Bone."""
        return BernoulliSpawn(
            prob=world.spawn_prob,
            choices=[Gem(world.values["gem"]), Food(world.values["food"]), Bone(world.values["bone"])],
        )
