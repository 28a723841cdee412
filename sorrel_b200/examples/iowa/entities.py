"""Mirrors sorrel/examples/iowa/entities.py:14-91; `Deck.transition` / `draw` and `EmptyEntity.transition` are
declared as rules."""
from sorrel_b200.entities import Entity
from sorrel_b200.rules import BernoulliSpawn, ValueRedraw


class Wall(Entity):
    def __init__(self):
        super().__init__()
        self.value = -1  # Walls penalize contact


class Sand(Entity):
    def __init__(self):
        super().__init__()
        self.passable = True


class Deck(Entity):
    # name -> (value without a loss, probability of a loss, the loss): entities.py:49-67
    DRAWS = {"a": (1, 0.5, -2.5), "b": (1, 0.1, -12.5), "c": (0.5, 0.5, -0.5), "d": (0.5, 0.1, -2.5)}

    def __init__(self, name: str):
        super().__init__()
        self.passable = True
        self.value = 0
        self.name = name
        self.has_transitions = True
        self.kind = f"Deck{name.upper()}"  # different decks are different entities to the observation

    def transition_rule(self, world):
        """entities.py:44-73: `p_loss = np.random.random()`; value = base (+ loss if p_loss < threshold) + 0.1."""
        base, p_loss, loss = self.DRAWS[self.name]
        value = base
        win = value + 0.1
        value += loss
        return ValueRedraw(prob=p_loss, value_if=value + 0.1, value_else=win)


class EmptyEntity(Entity):
    def __init__(self):
        super().__init__()
        self.passable = True
        self.has_transitions = True

    def transition_rule(self, world):
        """entities.py:84-91"""
        return BernoulliSpawn(prob=world.spawn_prob, choices=[Deck("a"), Deck("b"), Deck("c"), Deck("d")])
