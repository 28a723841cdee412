"""Mirrors sorrel/examples/cleanup/entities.py:14-105 and the beams of cleanup/agents.py:186-216; every
`transition()` of the reference is declared as a rule the step kernel evaluates per cell."""
from sorrel_b200.entities import Entity
from sorrel_b200.rules import BecomeIfLayer, BernoulliBecome, ExpireAfter


class EmptyEntity(Entity):
    def __init__(self):
        super().__init__()
        self.passable = True


class Sand(Entity):
    def __init__(self):
        super().__init__()
        self.passable = True
        self.kind = "EmptyEntity"  # entities.py:29-31


class Wall(Entity):
    def __init__(self):
        super().__init__()


class River(Entity):
    def __init__(self):
        super().__init__()
        self.has_transitions = True

    def transition_rule(self, world):
        """entities.py:50-54: `np.random.random() < world.pollution_spawn_chance` -> Pollution."""
        return BernoulliBecome(prob=world.pollution_spawn_chance, to=Pollution())


class Pollution(Entity):
    def __init__(self):
        super().__init__()
        self.has_transitions = True

    def transition_rule(self, world):
        """entities.py:64-71: a CleanBeam on the beam layer of this tile -> River."""
        return BecomeIfLayer(layer=world.beam_layer, kinds=["CleanBeam"], to=River())


class AppleTree(Entity):
    def __init__(self):
        super().__init__()
        self.has_transitions = True

    def transition_rule(self, world):
        """entities.py:82-90: `np.random.random() < world.apple_spawn_chance` -> Apple.  The pollution gate
        (`not world.pollution > world.pollution_threshold`) is constant: see CleanupWorld.pollution."""
        if world.pollution > world.pollution_threshold:
            return None
        return BernoulliBecome(prob=world.apple_spawn_chance, to=Apple())


class Apple(Entity):
    def __init__(self):
        super().__init__()
        self.value = 1
        self.has_transitions = True

    def transition_rule(self, world):
        """entities.py:100-105: an agent on the agent layer of this tile -> AppleTree."""
        return BecomeIfLayer(layer=world.agent_layer, kinds=["CleanupAgent"], to=AppleTree())


class Beam(Entity):
    """cleanup/agents.py:186-204: persists for one full turn, then disappears."""

    def __init__(self):
        super().__init__()
        self.turn_counter = 0
        self.has_transitions = True

    def transition_rule(self, world):
        return ExpireAfter(sweeps=1, to=EmptyEntity())


class CleanBeam(Beam):
    pass


class ZapBeam(Beam):
    def __init__(self):
        super().__init__()
        self.value = -1
