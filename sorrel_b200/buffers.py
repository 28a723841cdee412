"""Device-resident batched replay buffer with the reference `Buffer` API.

Mirrors `sorrel.buffers.Buffer` (sorrel/buffers.py:11-201): same constructor arguments, attribute
names, dtypes (states f32, actions i64, rewards/dones f32, extra_data i64), the same `idx`/`size`
bookkeeping including the quirks (`add_empty` advances by n_frames-1 without zeroing, buffers.py:65-69;
`sample` ignores the write head, buffers.py:109-111), but holds ONE ring per agent row for
`rows = num_envs * num_agents` agents that all advance in lock-step:

    states    float32 [rows, capacity, *obs_shape]   one reference Buffer.states per row (see `layout` below)
    actions   int64   [capacity, rows]               actions[:, r] is row r's Buffer.actions
    rewards   float32 [capacity, rows]
    dones     float32 [capacity, rows]
    extra_data["positions"] int64 [capacity, rows, 2]

`layout` picks how `states` sits in HBM; the tensor's SHAPE is [rows, capacity, *obs_shape] either way:

* "agent" (default): contiguous as shaped -- the n_frames stacked frames of a transition are one linear run, which is
  what `sample` and `current_state` like best (single-agent worlds);
* "slot": stored [capacity, rows, obs] and exposed through a permuted view -- the rows ONE TURN writes are contiguous, so
  the step kernel fills whole DRAM pages instead of one 600-byte piece per 19 KB (multi-agent worlds: the treasurehunt
  shape with 8 agents steps 27 % faster, profiles/README.md).  `Environment` picks it when a world has several agents.

`rows=1` reproduces a single reference buffer.  All data movement is done by the CUDA kernels
behind `sb_buffer_add / sb_buffer_sample / sb_buffer_current_state`; the environment step writes
observations straight into `states` (no separate observation tensor, no copy).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Sequence

import numpy as np
import torch

from . import _abi
from ._abi import SbBufferDesc, check, lib


class Buffer:
    def __init__(self, capacity: int, obs_shape: Sequence[int], n_frames: int = 1, rows: int = 1,
                 device: torch.device | str = "cuda", layout: str = "agent", **kwargs):
        self.lib = lib()
        if layout not in ("agent", "slot"):
            raise ValueError("layout must be 'agent' ([rows, capacity, obs]) or 'slot' ([capacity, rows, obs])")
        self.layout = layout
        self.capacity = int(capacity)
        self.obs_shape = tuple(int(x) for x in obs_shape)
        self.obs_size = int(np.prod(self.obs_shape))
        self.n_frames = int(n_frames)
        self.rows = int(rows)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _abi.SorrelB200Error("sorrel_b200.Buffer lives on a CUDA device (no CPU fallback)")
        self.extra_data: dict[str, torch.Tensor] = {}
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)  # device error word of the gathers
        self._alloc()
        for key, value in kwargs.items():  # buffers.py:38-44
            shape = (self.capacity, self.rows, *value) if isinstance(value, tuple) else (self.capacity, self.rows)
            self.extra_data[key] = torch.zeros(shape, dtype=torch.int64, device=self.device)
        self.idx = 0
        self.size = 0
        self._desc = None
        #: graphs.DeviceClock or None.  When set (Environment.enable_turn_graph), `current_state` reads the write head
        #: from device memory and always copies into one fixed tensor, so that it can live inside a captured graph.
        self.device_clock = None
        self._cs_fixed = None

    def _alloc(self):
        dev = self.device
        if self.layout == "slot":  # stored slot-major, exposed in the reference's (row, slot, ...) index order
            self._storage = torch.zeros((self.capacity, self.rows, *self.obs_shape), dtype=torch.float32, device=dev)
            self.states = self._storage.transpose(0, 1)
        else:
            self._storage = self.states = torch.zeros((self.rows, self.capacity, *self.obs_shape), dtype=torch.float32,
                                                      device=dev)
        self.actions = torch.zeros((self.capacity, self.rows), dtype=torch.int64, device=dev)
        self.rewards = torch.zeros((self.capacity, self.rows), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((self.capacity, self.rows), dtype=torch.float32, device=dev)
        self._desc = None

    # -- C ABI view ---------------------------------------------------------------------------
    def desc(self) -> SbBufferDesc:
        if self._desc is None:
            d = SbBufferDesc()
            d.capacity, d.obs_size, d.n_frames, d.rows = self.capacity, self.obs_size, self.n_frames, self.rows
            d.d_states = self._storage.data_ptr()
            d.states_row_stride, d.states_slot_stride = int(self.states.stride(0)), int(self.states.stride(1))
            d.d_actions = self.actions.data_ptr()
            d.d_rewards = self.rewards.data_ptr()
            d.d_dones = self.dones.data_ptr()
            pos = self.extra_data.get("positions")
            d.d_positions = pos.data_ptr() if pos is not None else None
            d.d_error = self.error.data_ptr()
            self._desc = d
        return self._desc

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def advance(self, turns: int) -> None:
        """idx/size bookkeeping of `turns` consecutive add()s (buffers.py:62-63)."""
        self.idx = (self.idx + turns) % self.capacity
        self.size = min(self.size + turns, self.capacity)

    # -- reference API -------------------------------------------------------------------------
    def add(self, obs, action, reward, done, **kwargs) -> None:
        """Buffer.add (buffers.py:46-63) for every row at once.

        obs [rows, *obs_shape] f32, action [rows] i64, reward [rows] f32, done [rows] f32 (device
        tensors; scalars/arrays are converted), positions=[rows, 2] i64."""
        dev = self.device

        def t(x, dtype, shape):
            x = torch.as_tensor(x, dtype=dtype, device=dev)
            return x.expand(shape).contiguous() if tuple(x.shape) != shape else x.contiguous()

        obs = torch.as_tensor(obs, dtype=torch.float32, device=dev).reshape(self.rows, self.obs_size).contiguous()
        action = t(action, torch.int64, (self.rows,))
        reward = t(reward, torch.float32, (self.rows,))
        done = t(done, torch.float32, (self.rows,))
        pos = kwargs.get("positions")
        if pos is not None and "positions" in self.extra_data:
            pos = t(pos, torch.int64, (self.rows, 2))
        else:
            pos = None
        check(self.lib.sb_buffer_add(C.byref(self.desc()), self.idx, obs.data_ptr(), action.data_ptr(),
                                     reward.data_ptr(), done.data_ptr(),
                                     pos.data_ptr() if pos is not None else None, self._stream()),
              "sb_buffer_add")
        for key, value in kwargs.items():  # other extra_data keys: plain slot write
            if key != "positions" and key in self.extra_data:
                self.extra_data[key][self.idx] = torch.as_tensor(value, dtype=torch.int64, device=dev)
        self.advance(1)

    def add_empty(self) -> None:
        """buffers.py:65-69: advance the head by n_frames-1, rows are NOT zeroed."""
        self.idx = (self.idx + self.n_frames - 1) % self.capacity
        self.size = min(self.size + 1, self.capacity)

    def add_from_buffer(self, buffer: "Buffer") -> None:
        """buffers.py:71-96 for every row at once: append the first `buffer.size` slots of `buffer` at this ring's
        head (truncated at the end of the ring, the head is NOT wrapped -- like the reference), extra_data included."""
        assert self.obs_shape == buffer.obs_shape, "Cannot add from a buffer with different state shapes."
        if self.rows != buffer.rows:
            raise ValueError(f"row count mismatch: {self.rows} vs {buffer.rows}")
        n = min(self.capacity - self.idx, buffer.size)
        lo, hi = self.idx, self.idx + n
        self.states[:, lo:hi] = buffer.states[:, :n].to(self.device)
        self.actions[lo:hi] = buffer.actions[:n].to(self.device)
        self.rewards[lo:hi] = buffer.rewards[:n].to(self.device)
        self.dones[lo:hi] = buffer.dones[:n].to(self.device)
        for key, value in buffer.extra_data.items():
            if value.shape[1] != self.rows:
                raise ValueError(f"extra_data[{key!r}]: {value.shape[1]} rows, this ring has {self.rows}")
            if key not in self.extra_data:
                self.extra_data[key] = torch.zeros((self.capacity, *value.shape[1:]), dtype=value.dtype,
                                                   device=self.device)
                self._desc = None  # the cached C view may not know this array yet (positions)
            self.extra_data[key][lo:hi] = value[:n].to(self.device)
        self.idx = self.idx + n

    def sample_indices(self, batch_size: int, rows: torch.Tensor | np.ndarray | None = None):
        """The index draw of Buffer.sample (buffers.py:109-111), batched.

        Start slots come from `np.random.choice(max(1, size - n_frames - 1), batch_size,
        replace=False)` -- the very call (and numpy global stream) the reference uses, so a
        single-row buffer seeded like the reference samples the same transitions.  Rows are
        uniform over agents unless given."""
        starts = np.random.choice(max(1, self.size - self.n_frames - 1), batch_size, replace=False)
        if rows is None:
            rows = np.zeros(batch_size, dtype=np.int64) if self.rows == 1 else \
                np.random.randint(0, self.rows, size=batch_size)
        return np.asarray(rows, dtype=np.int64), starts.astype(np.int64)

    def sample_indices_device(self, batch_size: int, generator: torch.Generator | None = None):
        """Index draw for training loops: (rows, starts) drawn ON THE DEVICE, no host synchronisation (the host draw
        above is kept for seed-for-seed agreement with the reference).  Pairs are drawn independently -- with `rows`
        x `size` transitions to choose from, repeats are as rare as they are harmless; a single-row ring whose
        population is smaller than the batch still raises like the reference's `replace=False` draw."""
        hi = max(1, self.size - self.n_frames - 1)
        if self.rows * hi < batch_size:
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")
        rows = torch.randint(0, self.rows, (batch_size,), device=self.device, generator=generator)
        starts = torch.randint(0, hi, (batch_size,), device=self.device, generator=generator)
        return rows, starts

    def sample_device(self, batch_size: int, generator: torch.Generator | None = None):
        """`sample` with the index draw on the device (see `sample_indices_device`)."""
        return self.gather(*self.sample_indices_device(batch_size, generator))

    def check_errors(self) -> None:
        """Synchronising read of the gathers' device error word; IndexError like numpy's fancy indexing."""
        bits = int(self.error.item())
        if bits:
            self.error.zero_()
            raise _abi.error_bits_to_exception(bits)

    def gather(self, rows, starts):
        """Buffer.sample's gather (buffers.py:112-124) for (row, start) pairs; device tensors in
        the reference's shapes/dtypes: states (B, F*obs) f32, actions (B,1) i64, rewards (B,1) f32,
        next_states (B, F*obs) f32, dones (B,1) f32, valid (B,1) f64."""
        dev = self.device
        rows = torch.as_tensor(rows, dtype=torch.int64, device=dev).contiguous()
        starts = torch.as_tensor(starts, dtype=torch.int64, device=dev).contiguous()
        B = int(rows.numel())
        F = self.n_frames
        states = torch.empty((B, F * self.obs_size), dtype=torch.float32, device=dev)
        next_states = torch.empty((B, F * self.obs_size), dtype=torch.float32, device=dev)
        actions = torch.empty((B, 1), dtype=torch.int64, device=dev)
        rewards = torch.empty((B, 1), dtype=torch.float32, device=dev)
        dones = torch.empty((B, 1), dtype=torch.float32, device=dev)
        valid = torch.empty((B, 1), dtype=torch.float64, device=dev)
        check(self.lib.sb_buffer_sample(C.byref(self.desc()), rows.data_ptr(), starts.data_ptr(), B,
                                        states.data_ptr(), actions.data_ptr(), rewards.data_ptr(),
                                        next_states.data_ptr(), dones.data_ptr(), valid.data_ptr(),
                                        self._stream()), "sb_buffer_sample")
        return states, actions, rewards, next_states, dones, valid

    def sample(self, batch_size: int):
        """Buffer.sample (buffers.py:98-124).  Raises ValueError when batch_size exceeds the
        population, like np.random.choice(replace=False) does in the reference."""
        rows, starts = self.sample_indices(batch_size)
        return self.gather(rows, starts)

    def gather_transformer(self, rows, starts):
        """TransformerBuffer.sample_transformer's gather (buffers.py:332-355): like `gather` but the
        actions keep their time axis.  Returns (states (B,F*obs) f32, actions (B,F) i64, next_actions
        (B,F) f32, next_states (B,F*obs) f32, dones (B,1) f32, valid (B,1) f64, agent_ids (B,) i64 =
        the ring row of every sample, the batched form of `combine`'s agent ids)."""
        dev = self.device
        rows = torch.as_tensor(rows, dtype=torch.int64, device=dev).contiguous()
        starts = torch.as_tensor(starts, dtype=torch.int64, device=dev).contiguous()
        B, F = int(rows.numel()), self.n_frames
        states = torch.empty((B, F * self.obs_size), dtype=torch.float32, device=dev)
        next_states = torch.empty((B, F * self.obs_size), dtype=torch.float32, device=dev)
        actions = torch.empty((B, F), dtype=torch.int64, device=dev)
        next_actions = torch.empty((B, F), dtype=torch.float32, device=dev)
        dones = torch.empty((B, 1), dtype=torch.float32, device=dev)
        valid = torch.empty((B, 1), dtype=torch.float64, device=dev)
        check(self.lib.sb_buffer_sample_transformer(C.byref(self.desc()), rows.data_ptr(), starts.data_ptr(), B,
                                                    states.data_ptr(), actions.data_ptr(), next_actions.data_ptr(),
                                                    next_states.data_ptr(), dones.data_ptr(), valid.data_ptr(),
                                                    self._stream()), "sb_buffer_sample_transformer")
        return states, actions, next_actions, next_states, dones, valid, rows

    def sample_transformer(self, batch_size: int):
        """TransformerBuffer.sample_transformer (buffers.py:319-355) over all rows."""
        rows, starts = self.sample_indices(batch_size)
        return self.gather_transformer(rows, starts)

    def combined(self) -> dict[str, np.ndarray]:
        """Host arrays equal to `TransformerBuffer.combine([one reference buffer per row])`
        (buffers.py:273-317): the first `size` slots of every row back to back, `agent_ids` = row
        (the reference allocates agent_ids but never fills them, so there they are all 0)."""
        n = self.size
        out = {
            "states": self.states[:, :n].reshape(self.rows * n, *self.obs_shape).cpu().numpy(),
            "actions": self.actions[:n].t().reshape(-1).cpu().numpy(),
            "rewards": self.rewards[:n].t().reshape(-1).cpu().numpy(),
            "dones": self.dones[:n].t().reshape(-1).cpu().numpy(),
            "agent_ids": np.repeat(np.arange(self.rows, dtype=np.int64), n),
        }
        for k, v in self.extra_data.items():
            out[k] = v[:n].transpose(0, 1).reshape(self.rows * n, *v.shape[2:]).cpu().numpy()
        return out

    def clear(self) -> None:
        """buffers.py:126-133 (extra_data is left untouched, as in the reference)."""
        self._alloc()
        self.idx = 0
        self.size = 0

    def getidx(self) -> int:
        return self.idx

    def current_state(self) -> torch.Tensor:
        """buffers.py:143-154: the last n_frames-1 observations before the write head, per row:
        float32 [rows, n_frames-1, *obs_shape].  A zero-copy view unless the window wraps."""
        F1 = self.n_frames - 1
        if self.device_clock is not None:
            if self._cs_fixed is None:
                self._cs_fixed = torch.zeros((self.rows, F1, *self.obs_shape), dtype=torch.float32, device=self.device)
            if F1 == 0:  # no stacking: zero rows, like states[idx:idx]
                return self._cs_fixed
            check(self.lib.sb_buffer_current_state_dev(C.byref(self.desc()), self.device_clock.tensor.data_ptr(),
                                                       self._cs_fixed.data_ptr(), self._stream()),
                  "sb_buffer_current_state_dev")
            return self._cs_fixed
        if self.idx >= F1:
            return self.states[:, self.idx - F1: self.idx]
        out = torch.empty((self.rows, F1, *self.obs_shape), dtype=torch.float32, device=self.device)
        check(self.lib.sb_buffer_current_state(C.byref(self.desc()), self.idx, out.data_ptr(), self._stream()),
              "sb_buffer_current_state")
        return out

    def __repr__(self):
        return f"Buffer(capacity={self.capacity}, obs_shape={self.obs_shape}, rows={self.rows})"

    __str__ = __repr__

    def __len__(self):
        return self.size

    def __getitem__(self, idx):
        return (self.states[:, idx], self.actions[idx], self.rewards[idx], self.dones[idx])

    def row(self, r: int) -> dict[str, np.ndarray]:
        """Row r as host arrays laid out exactly like one reference Buffer."""
        out = {
            "states": self.states[r].cpu().numpy(),
            "actions": self.actions[:, r].cpu().numpy(),
            "rewards": self.rewards[:, r].cpu().numpy(),
            "dones": self.dones[:, r].cpu().numpy(),
        }
        for k, v in self.extra_data.items():
            out[k] = v[:, r].cpu().numpy()
        return out

    def save(self, output_file: str | Path, row: int = 0, with_extra: bool = False) -> None:
        """Buffer.save (buffers.py:168-178): same .npz keys, one row.  `with_extra=True` adds the
        extra_data arrays like SavedGames.save / TransformerBuffer.save (buffers.py:225-243, 361-379)."""
        r = self.row(row)
        extra = {k: r[k] for k in self.extra_data} if with_extra else {}
        np.savez_compressed(Path(output_file), states=r["states"], actions=r["actions"], rewards=r["rewards"],
                            dones=r["dones"], n_frames=self.n_frames, idx=self.idx, **extra)

    @classmethod
    def load(cls, input_file: str | Path, device: torch.device | str = "cuda") -> "Buffer":
        """Buffer.load (buffers.py:180-201) into a single-row device buffer."""
        with np.load(Path(input_file)) as data:
            states, actions = data["states"], data["actions"]
            out = cls(capacity=len(actions), obs_shape=states.shape[1:], n_frames=int(data["n_frames"]),
                      rows=1, device=device)
            out.states[0].copy_(torch.from_numpy(states))
            out.actions[:, 0].copy_(torch.from_numpy(actions))
            out.rewards[:, 0].copy_(torch.from_numpy(data["rewards"]))
            out.dones[:, 0].copy_(torch.from_numpy(data["dones"]))
            out.idx = int(data["idx"])
            out.size = len(states)
        return out
