"""`Environment` with the reference's interface (sorrel/environment.py:18-300), batched.

Subclass it exactly like the reference: implement `setup_agents()` (assign `self.agents`) and
`populate_environment()` (edit `self.world.map` through `world.add`, then hand the valid spawn
cells to `self.place_agents(...)` instead of drawing one placement with numpy).  `reset()` and
`take_turn()` then advance ALL `world.num_envs` instances on the GPU:

* every agent uses an in-kernel policy (`RandomModel`)  -> one fused launch per `take_turn()`
  (`take_turns(T)` fuses T turns per launch);
* otherwise -> the split-phase path: `Agent.transition` runs pov / get_action / act per agent
  slot with the policy on the host side of the launch boundary, in the reference's order.
"""
from __future__ import annotations

import os
from abc import abstractmethod
from pathlib import Path

import numpy as np
import torch

from . import config as _config
from .agents import Agent
from .buffers import Buffer
from .models import AgentMemory
from .stats import RewardStats, summarize
from .worlds import Gridworld


class Environment:
    world: Gridworld
    agents: list[Agent]
    stop_if_done: bool

    def __init__(self, world: Gridworld, config, stop_if_done: bool = False) -> None:
        if isinstance(config, (list, tuple)):
            self.config = _config.from_dotlist(config)  # environment.py:44-46
        else:
            self.config = _config.wrap(config)
        self.world = world
        self.turn = 0
        self.world.create_world()
        self.stop_if_done = stop_if_done
        self._spawnable = None
        self._item_groups = []
        self.setup_agents()
        self.populate_environment()
        self._compile()
        self.reset()

    # -- to be provided by the subclass (environment.py:56-70) -----------------------------------
    @abstractmethod
    def setup_agents(self) -> None:
        """Create the agent slots and assign them to self.agents."""

    @abstractmethod
    def populate_environment(self) -> None:
        """Fill self.world.map (walls, floor, ...) and call self.place_agents(valid_locations)."""

    def place_agents(self, valid_spawn_locations) -> None:
        """Batched stand-in for `np.random.choice(len(valid), size=A, replace=False)` + `world.add(loc,
        agent)` (examples/treasurehunt/env.py:141-147): every instance draws its own placement on
        the device at each reset."""
        self._spawnable = np.asarray([tuple(loc) for loc in valid_spawn_locations], dtype=np.int64).reshape(-1, 3)

    def place_items(self, entity, valid_locations, count: int) -> None:
        """Batched stand-in for `np.random.choice(len(points), size=count, replace=False)` + `world.add(loc,
        entity)` (examples/cleanup/env.py:126-132): every instance draws its own `count` distinct cells at
        each reset, before the agents are placed.  Call in the order the reference places the groups."""
        self._item_groups.append((entity, [tuple(loc) for loc in valid_locations], int(count)))

    # -- engine wiring ---------------------------------------------------------------------------
    def _compile(self) -> None:
        if self._spawnable is None:
            raise RuntimeError("populate_environment() must call self.place_agents(valid_spawn_locations)")
        a0 = self.agents[0]
        for k, ag in enumerate(self.agents):
            ag.index, ag.env = k, self
            if type(ag) is not type(a0) or (ag.value, ag.passable) != (a0.value, a0.passable):
                raise ValueError("all agent slots must be of one Agent class (one kind-table row per state)")
        A = len(self.agents)
        flags = [int(bool(ag.flag)) for ag in self.agents] if getattr(a0, "state_kinds", None) else None
        eng = self.world.compile(a0, a0.observation_spec, a0.action_spec, A, self._spawnable, self._item_groups, flags)
        model = a0.model
        cap = int(getattr(model, "memory_size", 0) or self.config.get("engine", {}).get("buffer_capacity", 1024))
        n_frames = int(getattr(model, "n_frames", 1))
        # ring layout: several agents per world -> the rows of one turn contiguous in HBM ("slot"); one agent ->
        # frame stacks contiguous ("agent"); `config.engine.ring_layout` overrides (see buffers.Buffer)
        layout = self.config.get("engine", {}).get("ring_layout", "slot" if A > 1 else "agent")
        self.buffer = Buffer(cap, (eng.spec.obs_size,), n_frames=n_frames, rows=eng.N * A, device=eng.device,
                             layout=layout, positions=(2,))
        for k, ag in enumerate(self.agents):
            ag.model.memory = AgentMemory(self.buffer, k, A)
        self.stats = RewardStats(eng.device)
        self._fusable = all(getattr(ag.model, "in_kernel_policy", False) for ag in self.agents)
        self._reward_slot = torch.zeros((1, eng.N, A), dtype=torch.float32, device=eng.device)
        self._action_slot = torch.zeros((1, eng.N, A), dtype=torch.int32, device=eng.device)
        self._turn_open = False
        self._graphs = {}  # (turns, rng) -> StepGraph, see enable_graphs()
        self._turn_graph = None  # CUDA graph of one whole split-phase turn, see enable_turn_graph()
        self._turn_clock = None
        self._obs_fixed = None

    @property
    def engine(self):
        return self.world.engine

    # -- reference API ---------------------------------------------------------------------------
    def reset(self) -> None:
        """environment.py:72-79"""
        self.turn = 0
        self.world.is_done = False
        self.engine.reset()
        for agent in self.agents:
            agent.reset()

    def add_empty(self) -> None:
        """What every IQN model does at epoch start through `memory.add_empty()` (iqn.py:426-436),
        once for the shared ring."""
        self.buffer.add_empty()

    def take_turn(self) -> None:
        """environment.py:81-93 for every instance."""
        if self._fusable:
            self.take_turns(1)
            return
        if self._turn_graph is not None:  # the whole turn (every agent's phases and policy) as one graph replay
            self._sync_turn_clock()
            graph = self._turn_graph[self.buffer.idx] if isinstance(self._turn_graph, list) else self._turn_graph
            graph.replay()
            self.engine.tick += 1  # (the eager path advances it in the first phase launch of the turn)
            self._after_turn()
            return
        if self._turn_clock is not None:
            self._sync_turn_clock()
        self._take_turn_split()
        self._after_turn()

    def _take_turn_split(self) -> None:
        """One turn through the split-phase launches: Agent.transition per agent slot, in list order."""
        self._turn_open = False
        for agent in self.agents:  # entity sweep happens in agent 0's pov launch
            agent.transition(self.world)
        if self._turn_clock is not None:  # device-resident {tick, slot}: advance them in stream order
            from ._abi import check

            check(self.engine.lib.sb_clock_advance(self._turn_clock.tensor.data_ptr(), 1, 1, self.buffer.capacity,
                                                   torch.cuda.current_stream(self.engine.device).cuda_stream),
                  "sb_clock_advance")

    def _after_turn(self) -> None:
        """Host bookkeeping of one finished turn."""
        self.turn += 1
        self.buffer.advance(1)
        if self._turn_clock is not None:
            held = self._turn_clock.held
            self._turn_clock.held = [(held[0] + 1) & 0xFFFFFFFF, (held[1] + 1) % self.buffer.capacity]

    def _sync_turn_clock(self) -> None:
        """Bring the device clock in line with the host (after reset / add_empty / eager turns)."""
        from ._abi import check

        want = [(self.engine.tick + 1) & 0xFFFFFFFF, self.buffer.idx]
        held = self._turn_clock.held
        if want != held:
            cap = self.buffer.capacity
            check(self.engine.lib.sb_clock_advance(self._turn_clock.tensor.data_ptr(), (want[0] - held[0]) & 0xFFFFFFFF,
                                                   (want[1] - held[1]) % cap, cap,
                                                   torch.cuda.current_stream(self.engine.device).cuda_stream),
                  "sb_clock_advance")
            self._turn_clock.held = want

    def enable_turn_graph(self, per_slot: bool = False) -> None:
        """Live policies (anything but the in-kernel `RandomModel`): capture ONE WHOLE TURN -- for every agent slot in
        list order [sweep +] observe, frame stacking, `model.take_action`, act -- as a CUDA graph, with the tick and
        the ring slot kept in device memory.  `take_turn()` then is one graph replay: an on-GPU policy runs with no
        host work between the phases, exactly in the order of Agent.transition (agents/agent.py:155-173).

        Requirements: every model's `take_action` must be capturable (device tensors in, device tensor out, no host
        synchronisation, static shapes).

        `per_slot=False`: ONE graph.  Observations reach the policy through one fixed tensor (the kernel writes each row
        to its ring slot AND there), frame stacks through `Buffer.current_state()`'s device-clock form (a copy).
        `per_slot=True`: one graph PER RING SLOT (`capacity` captures sharing one memory pool), each with the addresses
        of its slot baked in: the policy reads the observation straight from the ring slot the kernel wrote and the
        frame stack is the usual zero-copy view, so a turn writes every row once.  Worth it for large batches, where the
        second copy of the rows costs more than anything else the graph saves."""
        from .graphs import DeviceClock

        if self._fusable:
            raise RuntimeError("in-kernel policies are fused already: use enable_graphs() / take_turns()")
        eng = self.engine
        A = len(self.agents)
        self._turn_clock = DeviceClock(eng.device)
        if not per_slot:
            self.buffer.device_clock = self._turn_clock
            self._obs_fixed = torch.zeros((1, eng.N, A, eng.spec.obs_size), dtype=torch.float32, device=eng.device)
        # one eager turn in this mode first: CUDA loads modules lazily and that cannot be captured
        self._sync_turn_clock()
        self._take_turn_split()
        self._after_turn()
        torch.cuda.synchronize(eng.device)
        tick = eng.tick
        if not per_slot:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._take_turn_split()
            eng.tick = tick  # capturing ran the host-side tick bookkeeping of a turn that was not executed
            self._turn_graph = graph
            return
        idx, graphs, pool = self.buffer.idx, [], None
        for slot in range(self.buffer.capacity):  # the kernels read tick / slot from the device clock; the HOST-side
            self.buffer.idx = slot                # views (observation row, frame stack) are what gets baked in per slot
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=pool):
                self._take_turn_split()
            pool = graph.pool() if pool is None else pool
            eng.tick = tick
            graphs.append(graph)
        self.buffer.idx = idx
        self._turn_graph = graphs

    def take_turns(self, num_turns: int, rng: str = "philox") -> None:
        """`num_turns` x take_turn() in ONE launch (in-kernel policies only).  `rng`: "philox" draws one random byte
        per cell per turn like the reference's draw-per-cell loop; "sparse" samples the same spawn field sparsely.
        Replays the CUDA graph captured by `enable_graphs(num_turns, ...)` when there is one."""
        if not self._fusable:
            raise RuntimeError("take_turns() needs in-kernel policies; use take_turn() with a host policy")
        graph = self._graphs.get((int(num_turns), rng))
        if graph is not None:
            graph.replay_step()
        else:
            self.engine.step(num_turns, buffer=self.buffer, rng=rng)
        self.turn += num_turns

    def enable_graphs(self, num_turns: int, sample_batch: int = 0, rng: str = "philox") -> None:
        """Capture `take_turns(num_turns, rng)` (+ a `sample_batch`-transition `sample_device`) as CUDA graphs with
        the tick / ring slot kept on the device, so that launch-bound regimes (small batches, one shard of many)
        spend no host time per step.  Call after at least one eager `take_turns` / `sample_device` of the same
        shape (module loading is not capturable) and once the ring is full (its population is baked in)."""
        from .graphs import StepGraph

        self._graphs[(int(num_turns), rng)] = StepGraph(self.engine, self.buffer, int(num_turns), rng=rng,
                                                        sample_batch=int(sample_batch))

    def kernel_launches(self) -> int:
        """Kernels of libsorrel_b200 launched so far by this process: direct launches (`sb_launch_count`) plus the
        library kernels inside replayed CUDA graphs."""
        return int(self.engine.lib.sb_launch_count()) + sum(g.kernels_replayed for g in self._graphs.values())

    def sample_device(self, batch_size: int, rng: str = "philox", generator=None):
        """Buffer.sample for the learner with the index draw on the device (buffers.py:98-124): replays the captured
        graph when `enable_graphs` was given this batch size, else `Buffer.sample_device`."""
        for (turns, g_rng), graph in self._graphs.items():
            if g_rng == rng and graph.B == int(batch_size):
                return graph.replay_sample()
        return self.buffer.sample_device(int(batch_size), generator)

    def begin_epoch(self) -> torch.Tensor:
        """What `run_experiment` does between two epochs (environment.py:128-133,191-199) without a host sync: reduce
        the finished episode's `world.total_reward` over every instance of every rank (device tensor {sum, sum_sq,
        min, max, count}; the ONE collective, 40 bytes per rank), `reset()`, `add_empty()`."""
        stats = self.stats.all_reduce(self.engine.total_reward)
        self.reset()
        self.add_empty()
        return stats

    # -- split-phase plumbing used by Agent.pov / Agent.act ----------------------------------------
    def _observe_slot(self, k: int) -> torch.Tensor:
        eng = self.engine
        extra = {} if self._turn_clock is None else {"obs_out": self._obs_fixed, "clock": self._turn_clock}
        if k == 0 and not self._turn_open:
            eng.phase(sweep=True, obs_agent=0, buffer=self.buffer, new_turn=True, **extra)
            self._turn_open = True
        elif k > 0 and self._pending_obs != k:
            eng.phase(obs_agent=k, buffer=self.buffer, **extra)
        self._pending_obs = None
        if self._obs_fixed is not None:  # fixed address: usable inside a captured graph
            return self._obs_fixed[0, :, k]
        A = len(self.agents)
        return self.buffer.states[k::A, self.buffer.idx].reshape(eng.N, -1)

    _pending_obs = None

    def _stacked_view(self, k: int):
        """The policy input of agent slot k -- the previous n_frames-1 observations followed by the one `pov()` just
        produced (treasurehunt/agents.py:34-41) -- as a ZERO-COPY view of the ring, or None when the frames are not
        adjacent in memory.  In the agent-major ring the slots of a row are contiguous, and the kernel has just written
        the observation into the slot at the write head, so frames [idx-(F-1), idx] of a row are one run of F*obs floats
        (`torch.cat` of the same frames copies N*F*obs floats per turn: 0.3 ms at 262144 x 3 frames).  The view is
        row-strided; the caller must not write to it."""
        buf = self.buffer
        F1 = buf.n_frames - 1
        if F1 == 0 or self._obs_fixed is not None or buf.layout != "agent" or buf.idx < F1:
            return None
        A = len(self.agents)
        return buf.states[k::A, buf.idx - F1: buf.idx + 1].reshape(self.engine.N, -1)

    def _act_slot(self, k: int, action) -> torch.Tensor:
        eng = self.engine
        A = len(self.agents)
        act = torch.as_tensor(action, device=eng.device).to(torch.int32).reshape(-1)
        if act.numel() == 1:
            act = act.expand(eng.N)
        self._action_slot[0, :, k] = act
        nxt = k + 1 if k + 1 < A else -1
        extra = {} if self._turn_clock is None else {"obs_out": self._obs_fixed, "clock": self._turn_clock}
        # act agent k and, in the same launch, observe agent k+1 (it must see k's move)
        eng.phase(act_agent=k, obs_agent=nxt, buffer=self.buffer, actions=self._action_slot,
                  reward_out=self._reward_slot, **extra)
        self._pending_obs = nxt if nxt >= 0 else None
        return self._reward_slot[0, :, k]

    # -- experiment loop ---------------------------------------------------------------------------
    def reward_summary(self) -> dict:
        """mean/std/min/max of world.total_reward over every instance of every rank (one tiny
        all_gather; see stats.py)."""
        return summarize(self.stats.all_reduce(self.engine.total_reward))

    def generate_memories(self, num_games: int = 1000, animate: bool = False,
                          output_dir: Path | None = None) -> list[Path]:
        """environment.py:213-300: play `num_games` games with the existing models and write one
        `memories/agent{i}.npz` per agent slot in the `SavedGames` format (buffers.py:358-379: states, actions,
        rewards, dones, n_frames, idx + extra_data such as positions), games back to back.

        Every device instance plays one game per episode, so ceil(num_games / num_envs) episodes are run and the
        first `num_games` instance-episodes are kept (instance-major within an episode).  The ring must hold one
        episode (`capacity >= max_turns`).  `animate` is accepted for signature compatibility (rendering is
        host-side tooling outside the engine)."""
        exp = self.config.experiment
        T, N, A = int(exp.max_turns), self.engine.N, len(self.agents)
        if self.buffer.capacity < T:
            raise ValueError(f"replay ring capacity {self.buffer.capacity} cannot hold an episode of {T} turns")
        if output_dir is None:
            output_dir = Path(exp.output_dir) if hasattr(exp, "output_dir") else Path("./data/")
        output_dir = Path(output_dir)
        os.makedirs(output_dir / "memories", exist_ok=True)
        fuse = int(self.config.get("engine", {}).get("turns_per_launch", 25))
        keys = ("states", "actions", "rewards", "dones", *self.buffer.extra_data.keys())
        chunks = {k: [[] for _ in range(A)] for k in keys}
        done_games, game = 0, 0
        while done_games < num_games:
            self.reset()
            self.buffer.clear()
            for agent in self.agents:
                agent.model.start_epoch_action(epoch=game)
            while not self.turn >= T:
                if self._fusable:
                    self.take_turns(min(fuse, T - self.turn))
                else:
                    self.take_turn()
            self.world.is_done = True
            take = min(N, num_games - done_games)
            host = {"states": self.buffer.states[:, :T].cpu().numpy(),            # [rows, T, obs]
                    "actions": self.buffer.actions[:T].cpu().numpy(),              # [T, rows]
                    "rewards": self.buffer.rewards[:T].cpu().numpy(),
                    "dones": self.buffer.dones[:T].cpu().numpy()}
            for k, v in self.buffer.extra_data.items():
                host[k] = v[:T].cpu().numpy()                                      # [T, rows, ...]
            for i, agent in enumerate(self.agents):
                agent.model.end_epoch_action(epoch=game)
                rows = np.arange(take) * A + i  # this slot's row in every kept instance
                chunks["states"][i].append(host["states"][rows].reshape(take * T, -1))
                for k in keys[1:]:
                    arr = np.moveaxis(host[k][:, rows], 0, 1)  # [take, T, ...]
                    chunks[k][i].append(arr.reshape(take * T, *arr.shape[2:]))
            done_games += take
            game += 1
        paths = []
        for i, agent in enumerate(self.agents):
            data = {k: np.concatenate(chunks[k][i]) for k in keys}
            path = output_dir / "memories" / f"agent{i}.npz"
            np.savez_compressed(path, n_frames=int(getattr(agent.model, "n_frames", 1)),
                                idx=len(data["actions"]) % max(1, num_games * T), **data)
            paths.append(path)
        self.engine.check_errors()
        return paths

    def run_experiment(self, animate: bool = False, logging: bool = True, logger=None,
                       output_dir: Path | None = None) -> list[dict]:
        """environment.py:108-211 without checkpoints; `animate=True` renders device instance 0 of every
        `record_period`-th epoch to `<output_dir>/gifs/` through `sorrel_b200.visualization.ImageRenderer`:
        epochs x (reset, start_epoch_action, max_turns turns, end_epoch_action, train_step, log,
        epsilon decay).  Returns the per-epoch reward summaries."""
        exp = self.config.experiment
        history = []
        fuse = int(self.config.get("engine", {}).get("turns_per_launch", 25))
        renderer = None
        if animate:  # environment.py:136-142: one device instance is read back per turn of a recorded epoch
            from .visualization import ImageRenderer

            renderer = ImageRenderer(self.__class__.__name__, int(exp.get("record_period", 50)), int(exp.max_turns))
            output_dir = Path(output_dir) if output_dir is not None else Path("./data/")
        for epoch in range(exp.epochs + 1):
            self.reset()
            self.add_empty()
            animate_this_epoch = renderer is not None and epoch % renderer.record_period == 0
            for agent in self.agents:
                agent.model.start_epoch_action(epoch=epoch)
            while not self.turn >= exp.max_turns:
                if animate_this_epoch:
                    renderer.add_image(self.world)
                if self._fusable:
                    self.take_turns(1 if animate_this_epoch else min(fuse, exp.max_turns - self.turn))
                else:
                    self.take_turn()
                if self.world.is_done and self.stop_if_done:
                    break
            self.world.is_done = True
            if animate_this_epoch:
                renderer.save_gif(epoch, output_dir / "gifs")
            total_loss = 0
            for agent in self.agents:
                agent.model.end_epoch_action(epoch=epoch)
                total_loss = agent.model.train_step()
            summary = self.reward_summary()
            summary["epoch"], summary["loss"] = epoch, float(np.asarray(total_loss))
            history.append(summary)
            if logging:
                if logger is not None:
                    logger.record_turn(epoch, total_loss, summary["mean"], self.agents[0].model.epsilon)
                elif int(os.environ.get("RANK", "0")) == 0:
                    print(f"epoch {epoch:4d}  reward mean {summary['mean']:8.3f}  std {summary['std']:7.3f}  "
                          f"min {summary['min']:6.1f}  max {summary['max']:6.1f}  envs {summary['count']}")
            for agent in self.agents:
                if hasattr(self.config.model, "epsilon_decay"):
                    agent.model.epsilon_decay(self.config.model.epsilon_decay)
        self.engine.check_errors()
        return history
