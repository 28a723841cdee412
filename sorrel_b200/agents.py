"""Agent classes with the reference's interface (sorrel/agents/agent.py:14-225), batched.

An `Agent` object stands for ONE agent slot (list index k in `Environment.agents`) of every world
instance.  `transition(world)` runs the same sub-steps in the same order as the reference
(agent.py:155-173): pov -> get_action -> act -> is_done -> total_reward -> add_memory, each for all
N instances at once through split-phase kernel launches, so agent k still observes the grid after
agents 0..k-1 have moved and list order still decides contested cells.
"""
from __future__ import annotations

import torch

from .entities import Entity


class Agent(Entity):
    def __init__(self, observation_spec, action_spec, model, location=None):
        self.observation_spec = observation_spec
        self.action_spec = action_spec
        self.model = model
        self._location = location
        super().__init__()
        self._location = location
        self.has_transitions = True  # agent.py:54-55
        self.index = 0  # position in Environment.agents (conflict priority)
        self.env = None  # set by Environment

    # -- overridable sub-steps (same names as the reference) -------------------------------------
    def reset(self) -> None:
        self.model.reset()

    def pov(self, world) -> torch.Tensor:
        """This slot's observation in every instance, flattened: float32 [N, C*S*S].

        The kernel writes it straight into the replay-ring slot of the current turn and this
        returns that row (a view; no copy)."""
        return self.env._observe_slot(self.index)

    def get_action(self, state: torch.Tensor) -> torch.Tensor:
        """treasurehunt/agents.py:34-41: stack the previous n_frames-1 observations with `state`."""
        stacked = self.env._stacked_view(self.index) if self.env is not None else None
        if stacked is not None:  # the frames already sit next to each other in the ring: no copy
            return self.model.take_action(stacked)
        prev = self.model.memory.current_state()  # [N, F-1, obs]
        if prev.shape[1] == 0:  # n_frames == 1: nothing to stack (a cat would still copy every row: 630 MB at 2^20)
            return self.model.take_action(state)
        stacked = torch.cat((prev.reshape(prev.shape[0], -1), state), dim=1)
        return self.model.take_action(stacked)

    def act(self, world, action: torch.Tensor) -> torch.Tensor:
        """MovingAgent.act (agent.py:215-225) for every instance: float32 rewards [N]."""
        return self.env._act_slot(self.index, action)

    def is_done(self, world) -> bool:
        return world.is_done

    def add_memory(self, state, action, reward, done) -> None:
        """agent.py:113-136.  The act kernel has already recorded (state, action, reward, done,
        post-move position) into the ring slot; nothing is copied here."""

    def transition(self, world) -> None:
        state = self.pov(world)
        action = self.get_action(state)
        reward = self.act(world, action)
        done = self.is_done(world)
        # world.total_reward += reward happens in the act kernel (agent.py:172)
        self.add_memory(state, action, reward, done)


class MovingAgent(Agent):
    """agent.py:176-225: up/down/left/right decode lives in the kernel (ActionSpec order).

    `direction` (agent.py:185) is the facing every instance's agent starts with; the kernels update the
    per-instance facings when a movement action is taken (`Engine.agent_facing`)."""

    direction = 2  # default: facing down
    #: reward = sum of Entity.value over all layers of the target column (cleanup/agents.py:160-170)
    reward_all_layers = False
    #: times the reward is added to world.total_reward per agent turn (agent.py:172; cleanup adds it twice)
    reward_count = 1

    def action_effects(self, world) -> dict:
        """Non-movement actions: action name -> effect from sorrel_b200.rules (e.g. SpawnBeam)."""
        return {}

    # -- agents with a binary state that changes how they look (examples/tag: `it`) -------------------------
    #: (kind string without the flag, kind string with it), e.g. ("NotIt", "It"); None = one kind
    state_kinds = None
    #: after its move a flagged agent hands the flag to the first unflagged neighbour (TagAgent.act)
    touch_swap = False
    #: {kind string: reward}: reward by the agent's own state after the exchange instead of the target's value
    reward_by_state = None
    #: the flag is appended to the observation (TagAgent.pov)
    observe_flag = False

    @property
    def flag(self) -> bool:
        """Initial value of the flag for this agent slot (every instance starts with the same flags)."""
        return False
