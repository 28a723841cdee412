"""Mirrors sorrel/examples/cleanup/agents.py:21-183."""
import numpy as np

from sorrel_b200.agents import MovingAgent
from sorrel_b200.examples.cleanup.entities import CleanBeam, ZapBeam
from sorrel_b200.observation import OneHotObservationSpec, positional_embedding
from sorrel_b200.rules import SpawnBeam


class CleanupObservation(OneHotObservationSpec):
    """agents.py:21-62: the flattened one-hot crop followed by the positional code of the agent's cell."""

    def __init__(self, entity_list, full_view: bool = False, vision_radius=None, embedding_size: int = 3):
        super().__init__(entity_list, full_view, vision_radius)
        self.embedding_size = embedding_size
        side = 2 * self.vision_radius + 1
        self.input_size = (1, len(entity_list) * side * side + 4 * self.embedding_size)

    def extras(self, world) -> np.ndarray:
        """[H, W, 4 * embedding_size] float64 table the kernels append to every observation row."""
        scale = (self.embedding_size, self.embedding_size)
        return np.array([[positional_embedding((y, x), world, scale) for x in range(world.width)]
                         for y in range(world.height)])

    def observe(self, world, location=None):
        """agents.py:52-62 for every instance.  location None -> float32 [N, A, C*S*S + 4*embedding]: the
        view of every agent; a (y, x[, z]) location -> the view from that cell, float32 [N, C*S*S + 4*embedding]."""
        if location is None:
            return world.engine.observe()
        return world.observe_from(location)


class CleanupAgent(MovingAgent):
    reward_all_layers = True  # agents.py:160-170: values of all layers at the target location
    reward_count = 2  # `world.total_reward += reward` in act() (agents.py:172) and in transition() (agent.py:172)

    def __init__(self, observation_spec, action_spec, model, beam_radius: int = 3):
        super().__init__(observation_spec, action_spec, model)
        self.direction = 2  # agents.py:74
        self.beam_radius = beam_radius

    def action_effects(self, world):
        """agents.py:93-145, 155-156"""
        return {"clean": SpawnBeam(CleanBeam(), self.beam_radius), "zap": SpawnBeam(ZapBeam(), self.beam_radius)}

    def is_done(self, world) -> bool:
        return world.turn >= world.max_turns  # agents.py:179-180 (world.turn never advances in the reference)
