"""Model plugin interface (sorrel/models/base_model.py:10-111), batched.

A model serves ONE agent slot of every world instance: `take_action(state)` receives the stacked
observations of that agent in all N instances (`[N, n_frames * obs]`, a device tensor) and returns
N actions.  `memory` is that agent slot's view of the shared device replay ring.
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np
import torch


class BaseModel:
    input_size: int | Sequence[int]
    action_space: int
    epsilon: float
    #: True when the engine may evaluate this policy inside the step kernel (no host round trip)
    in_kernel_policy = False

    def __init__(self, input_size, action_space: int, memory_size: int, epsilon: float = 0.0, n_frames: int = 1):
        self.input_size = input_size
        self.action_space = action_space
        self.memory_size = int(memory_size)
        self.n_frames = int(n_frames)
        self.memory = None  # attached by Environment (AgentMemory view of the shared ring)
        self.epsilon = epsilon

    def take_action(self, state) -> torch.Tensor:
        raise NotImplementedError

    def train_step(self) -> np.ndarray:
        return np.array(0.0)

    def reset(self):
        pass

    def set_epsilon(self, new_epsilon: float) -> None:
        self.epsilon = new_epsilon

    def epsilon_decay(self, decay_rate: float) -> None:
        self.epsilon *= 1 - decay_rate

    def start_epoch_action(self, **kwargs):
        pass

    def end_epoch_action(self, **kwargs):
        pass

    def save(self, file_path: str | os.PathLike) -> None:
        pass

    @property
    def model_name(self):
        return self.__class__.__name__


class RandomModel(BaseModel):
    """base_model.py:107-111: uniform random action.  Evaluated in-kernel (SB_POLICY_UNIFORM) when
    every agent uses it; `take_action` is the host-visible equivalent for the split-phase path."""

    in_kernel_policy = True

    def take_action(self, state):
        n = state.shape[0] if hasattr(state, "shape") and len(state.shape) > 1 else 1
        device = state.device if isinstance(state, torch.Tensor) else "cpu"
        return torch.randint(0, self.action_space, (n,), dtype=torch.int32, device=device)


class AgentMemory:
    """One agent slot's window onto the shared device `Buffer` (rows agent, agent+A, agent+2A, ...).

    Offers the reference `Buffer` methods a model uses (`current_state`, `sample`, `__len__`,
    `add_empty`, `clear`, `getidx`) restricted to that slot."""

    def __init__(self, buffer, agent: int, num_agents: int):
        self.buffer, self.agent, self.num_agents = buffer, agent, num_agents
        self.extra_data = buffer.extra_data

    @property
    def capacity(self):
        return self.buffer.capacity

    @property
    def n_frames(self):
        return self.buffer.n_frames

    @property
    def obs_shape(self):
        return self.buffer.obs_shape

    @property
    def idx(self):
        return self.buffer.idx

    @property
    def size(self):
        return self.buffer.size

    def __len__(self):
        return self.buffer.size

    def getidx(self):
        return self.buffer.idx

    @property
    def states(self):
        return self.buffer.states[self.agent::self.num_agents]

    @property
    def actions(self):
        return self.buffer.actions[:, self.agent::self.num_agents]

    @property
    def rewards(self):
        return self.buffer.rewards[:, self.agent::self.num_agents]

    @property
    def dones(self):
        return self.buffer.dones[:, self.agent::self.num_agents]

    def current_state(self) -> torch.Tensor:
        """[N, n_frames-1, *obs_shape] for this slot (buffers.py:143-154)."""
        return self.buffer.current_state()[self.agent::self.num_agents]

    def sample(self, batch_size: int):
        """Buffer.sample over this slot's rows (start indices drawn like buffers.py:109-111)."""
        n_envs = self.buffer.rows // self.num_agents
        envs = np.random.randint(0, n_envs, size=batch_size) if n_envs > 1 else np.zeros(batch_size, dtype=np.int64)
        rows, starts = self.buffer.sample_indices(batch_size, rows=envs * self.num_agents + self.agent)
        return self.buffer.gather(rows, starts)

    def add_empty(self):
        raise RuntimeError("the ring is shared: call Environment.add_empty() once per epoch, not per agent")

    def clear(self):
        self.buffer.clear()
