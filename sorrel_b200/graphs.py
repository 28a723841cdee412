"""CUDA-graph replay of the fused step (+ optional Buffer.sample) for launch-bound regimes.

A step launch is identical from turn to turn except for two integers (the global tick and the ring
slot).  `SbStepArgs.d_clock` lets the kernel read them from device memory and `sb_clock_advance`
advances them on the device, so one captured graph = [step kernel, clock advance, (index draw, sample
gather)] can be replayed without any per-step host work: small batches (config 2: 4096 instances) and
multi-GPU strong scaling stop being bound by Python / launch latency.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from ._abi import SbStepArgs, check


class DeviceClock:
    """{tick0, slot0} of the next step launch in device memory + the host's record of what it holds.  Graphs that
    step the same engine / ring share one clock."""

    def __init__(self, device):
        self.tensor = torch.zeros(2, dtype=torch.int32, device=device)
        self.held = [0, 0]


class StepGraph:
    def __init__(self, engine, buffer, turns: int, rng: str = "philox", sample_batch: int = 0,
                 actions: torch.Tensor | None = None, reward_out: torch.Tensor | None = None,
                 clock: DeviceClock | None = None):
        """Capture one step of `turns` fused turns recording into `buffer`: with the in-kernel policy, or -- when
        `actions` (int32 [turns, N, A], a FIXED device tensor the caller refills before every replay) is given -- with
        caller-supplied actions; `reward_out` (float32 [turns, N, A], fixed) receives the step's rewards.

        Every kernel involved must have been launched once before (CUDA loads modules lazily and that
        is not capturable): run one eager `engine.step(...)` / `buffer.gather(...)` first."""
        self.eng, self.buf, self.T, self.rng, self.B = engine, buffer, int(turns), rng, int(sample_batch)
        self._keep = (actions, reward_out)
        #: kernels of libsorrel_b200 launched through graph replays (sb_launch_count only sees direct launches)
        self.kernels_replayed = 0
        dev = engine.device
        self._clock = clock or DeviceClock(dev)
        self.clock = self._clock.tensor
        self._args = SbStepArgs()
        a = self._args
        a.mode, a.num_turns, a.seed = _abi.SB_MODE_FUSED, self.T, engine.seed
        a.rng = _abi.SB_RNG_SPARSE if rng == "sparse" else _abi.SB_RNG_PHILOX
        a.policy = _abi.SB_POLICY_UNIFORM if actions is None else _abi.SB_POLICY_TAPE
        shape = (self.T, engine.N, engine.A)
        a.actions_u8 = 1 if actions is not None and actions.dtype == torch.uint8 else 0
        for t, dtype, name in ((actions, torch.uint8 if a.actions_u8 else torch.int32, "actions"),
                               (reward_out, torch.float32, "reward_out")):
            if t is not None and (tuple(t.shape) != shape or t.dtype != dtype or not t.is_contiguous()
                                  or t.device != engine.grid.device):
                raise ValueError(f"{name}: expected contiguous {dtype} {shape} on {engine.grid.device}")
        a.d_actions = None if actions is None else actions.data_ptr()
        a.d_reward_out = None if reward_out is None else reward_out.data_ptr()
        a.act_agent = a.obs_agent = -1
        a.d_error = engine.error.data_ptr()
        a.d_clock = self.clock.data_ptr()
        a.d_agent_facing = None if engine.agent_facing is None else engine.agent_facing.data_ptr()
        engine.apply_tuning(a)
        self.sample = None
        if self.B:
            F, obs = buffer.n_frames, buffer.obs_size
            self._hi = max(1, buffer.size - F - 1)
            self.rows = torch.zeros(self.B, dtype=torch.int64, device=dev)
            self.starts = torch.zeros(self.B, dtype=torch.int64, device=dev)
            self.sample = (torch.empty((self.B, F * obs), dtype=torch.float32, device=dev),
                           torch.empty((self.B, 1), dtype=torch.int64, device=dev),
                           torch.empty((self.B, 1), dtype=torch.float32, device=dev),
                           torch.empty((self.B, F * obs), dtype=torch.float32, device=dev),
                           torch.empty((self.B, 1), dtype=torch.float32, device=dev),
                           torch.empty((self.B, 1), dtype=torch.float64, device=dev))
        self._sync_clock()
        torch.cuda.synchronize(dev)
        # two graphs (step / sample) so that a caller can put CUDA events between them
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._enqueue_step()
        self.sample_graph = None
        if self.B:
            self.sample_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.sample_graph):
                self._enqueue_sample()

    def _enqueue_step(self):
        eng, buf, lib = self.eng, self.buf, self.eng.lib
        stream = torch.cuda.current_stream(eng.device).cuda_stream
        check(lib.sb_step(C.byref(eng.desc), C.byref(self._args), eng.grid.data_ptr(), eng.agent_pos.data_ptr(),
                          eng.total_reward.data_ptr(), C.byref(buf.desc()), stream), "sb_step(graph)")
        check(lib.sb_clock_advance(self.clock.data_ptr(), self.T, self.T, buf.capacity, stream), "sb_clock_advance")

    def _enqueue_sample(self):
        buf, lib = self.buf, self.eng.lib
        stream = torch.cuda.current_stream(self.eng.device).cuda_stream
        self.rows.random_(0, buf.rows)
        self.starts.random_(0, self._hi)
        s = self.sample
        check(lib.sb_buffer_sample(C.byref(buf.desc()), self.rows.data_ptr(), self.starts.data_ptr(), self.B,
                                   s[0].data_ptr(), s[1].data_ptr(), s[2].data_ptr(), s[3].data_ptr(),
                                   s[4].data_ptr(), s[5].data_ptr(), stream), "sb_buffer_sample(graph)")

    def _sync_clock(self):
        """Bring the device clock in line with the host bookkeeping (after reset / add_empty / eager steps)."""
        want = [(self.eng.tick + 1) & 0xFFFFFFFF, self.buf.idx]
        held = self._clock.held
        if want != held:
            cap = self.buf.capacity
            check(self.eng.lib.sb_clock_advance(self.clock.data_ptr(), (want[0] - held[0]) & 0xFFFFFFFF,
                                                (want[1] - held[1]) % cap, cap,
                                                torch.cuda.current_stream(self.eng.device).cuda_stream),
                  "sb_clock_advance")
            self._clock.held = want

    def replay_step(self):
        """`turns` fused turns for every instance."""
        self._sync_clock()
        self.graph.replay()
        self.kernels_replayed += 2  # step kernel + clock advance
        self.eng.tick += self.T
        self.buf.advance(self.T)
        held = self._clock.held
        self._clock.held = [(held[0] + self.T) & 0xFFFFFFFF, (held[1] + self.T) % self.buf.capacity]

    def replay_sample(self):
        """Draw `sample_batch` (row, start) pairs and gather them; returns the (reused) batch tensors."""
        if max(1, self.buf.size - self.buf.n_frames - 1) != self._hi:
            raise RuntimeError("the ring population changed since capture: capture the StepGraph again")
        self.sample_graph.replay()
        self.kernels_replayed += 1  # buffer_sample_kernel
        return self.sample

    def replay(self):
        """One step (+ sample): returns the sampled batch tensors (overwritten by the next replay) or None."""
        self.replay_step()
        return self.replay_sample() if self.B else None
