"""ctypes wrapper around oracle/_build/liboracle.so (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import restatement as R

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
RNG_NUMPY, RNG_PHILOX, RNG_SPARSE = 0, 1, 2


class OrWorld(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32), ("L", C.c_int32), ("A", C.c_int32), ("K", C.c_int32),
        ("C", C.c_int32), ("v", C.c_int32),
        ("n_dyn", C.c_int32), ("dyn_layers", C.c_int32 * 8),
        ("agent_layer", C.c_int32), ("fill_kind", C.c_int32), ("agent_kind", C.c_int32),
        ("default_kind", C.c_int32),
        ("n_actions", C.c_int32), ("dy", C.c_int32 * 8), ("dx", C.c_int32 * 8),
        ("n_choices", C.c_int32), ("choices", C.c_int32 * 8),
        ("n_spawnable", C.c_int32), ("obs_transform", C.c_int32),
        ("p", C.c_double),
        ("value", C.c_double * 32), ("passable", C.c_uint8 * 32), ("rule", C.c_uint8 * 32),
        ("appearance", C.c_void_p), ("tmpl", C.c_void_p), ("spawnable", C.c_void_p),
        ("rule_prob", C.c_double * 32), ("rule_to", C.c_uint8 * 32), ("rule_else", C.c_uint8 * 32),
        ("rule_layer", C.c_int8 * 32),
        ("rule_mask", C.c_uint32 * 32),
        ("uses_facing", C.c_int32), ("initial_facing", C.c_int32),
        ("action_facing", C.c_int8 * 8), ("action_beam_kind", C.c_uint8 * 8),
        ("beam_layer", C.c_int32), ("n_beam", C.c_int32),
        ("beam_dy", (C.c_int8 * 16) * 4), ("beam_dx", (C.c_int8 * 16) * 4),
        ("beam_block_mask", C.c_uint32),
        ("reward_all_layers", C.c_int32), ("reward_count", C.c_int32),
        ("obs_extra", C.c_int32), ("extra", C.c_void_p),
        ("n_groups", C.c_int32), ("group_kind", C.c_int32 * 4), ("group_layer", C.c_int32 * 4),
        ("group_count", C.c_int32 * 4), ("group_n_cells", C.c_int32 * 4), ("group_cells", C.c_void_p * 4),
        ("agent_kind_alt", C.c_int32), ("touch_swap", C.c_int32), ("reward_own", C.c_int32), ("obs_flag", C.c_int32),
        ("initial_flags", C.c_int32 * 64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-C", HERE])
        L = C.CDLL(LIB)
        L.or_create.restype = C.c_void_p
        L.or_create.argtypes = [C.POINTER(OrWorld), C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                C.c_int, C.c_int]
        for name in ("or_destroy", "or_reset", "or_add_empty"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.or_step.argtypes = [C.c_void_p, C.c_int]
        L.or_step.restype = None
        L.or_set_tick.argtypes = [C.c_void_p, C.c_uint32]
        L.or_set_sparse_cdf.argtypes = [C.c_void_p, C.c_void_p]
        L.or_set_env_ids.argtypes = [C.c_void_p, C.c_void_p]
        for name in ("or_grid", "or_pos", "or_facing", "or_flags", "or_total", "or_states", "or_actions", "or_rewards", "or_dones",
                     "or_positions", "or_tape_u", "or_tape_c", "or_tape_a", "or_tape_placement"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_void_p
        L.or_idx.argtypes = L.or_size.argtypes = [C.c_void_p]
        L.or_max_threads.restype = C.c_int
        L.or_np_probe.argtypes = [C.c_uint32, C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_int,
                                  C.c_int, C.c_void_p]
        L.or_philox_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class COracle:
    """N instances of an oracle WorldSpec stepped by the C restatement (pthreads over instances)."""

    def __init__(self, spec: R.WorldSpec, num_envs: int, *, capacity: int = 0, n_frames: int = 1,
                 rng: int = RNG_PHILOX, seed: int = 0, env_id_base: int = 0, threads: int = 1,
                 record: bool = False, env_ids=None):
        """`env_ids` (optional, length num_envs): global instance id of every instance -- replays a SAMPLE of a larger
        counter-based batch (numpy mode: instance n is seeded seed + n regardless)."""
        self.L = lib()
        self.spec, self.N, self.capacity = spec, int(num_envs), int(capacity)
        w = OrWorld()
        w.H, w.W, w.L, w.A = spec.height, spec.width, spec.layers, spec.num_agents
        w.K, w.C, w.v = len(spec.kind_names), spec.channels, spec.vision
        w.n_dyn = len(spec.dynamic_layers)
        for i, z in enumerate(spec.dynamic_layers):
            w.dyn_layers[i] = z
        w.agent_layer, w.fill_kind = spec.agent_layer, spec.fill_kind
        w.agent_kind, w.default_kind = spec.agent_kind, spec.default_kind
        w.n_actions = spec.num_actions
        for i, (dy, dx) in enumerate(spec.action_deltas[: spec.num_actions]):
            w.dy[i], w.dx[i] = dy, dx
        w.n_choices = len(spec.spawn_choices)
        for i, k in enumerate(spec.spawn_choices):
            w.choices[i] = k
        w.p = spec.spawn_prob
        w.obs_transform = int(getattr(spec, "obs_transform", 0))
        for k in range(w.K):
            w.value[k] = float(spec.kind_value[k])
            w.passable[k] = 1 if spec.kind_passable[k] else 0
            w.rule[k] = int(spec.kind_rule[k])
        self._app = np.ascontiguousarray(spec.appearance, dtype=np.float64)
        self._tmpl = np.ascontiguousarray(spec.template, dtype=np.uint8)
        sp = np.asarray(spec.spawnable)
        self._spawn = np.ascontiguousarray(sp[:, 0] * spec.width + sp[:, 1], dtype=np.int32)
        w.n_spawnable = len(self._spawn)
        w.appearance, w.tmpl, w.spawnable = self._app.ctypes.data, self._tmpl.ctypes.data, self._spawn.ctypes.data
        # ---- column rules / facings / beams / positional code / item groups (examples/cleanup) ----
        for k in range(w.K):
            w.rule_prob[k] = float(spec.kind_rule_prob[k]) if spec.kind_rule_prob else 0.0
            w.rule_to[k] = int(spec.kind_rule_to[k]) if spec.kind_rule_to else R.NO_KIND
            w.rule_else[k] = int(spec.kind_rule_else[k]) if spec.kind_rule_else else R.NO_KIND
            w.rule_layer[k] = int(spec.kind_rule_layer[k]) if spec.kind_rule_layer else -1
            w.rule_mask[k] = int(spec.kind_rule_mask[k]) if spec.kind_rule_mask else 0
        w.uses_facing = 1 if (spec.action_facing is not None or spec.action_beam_kind is not None) else 0
        w.initial_facing = int(spec.initial_facing)
        for i in range(8):
            w.action_facing[i] = int(spec.action_facing[i]) if spec.action_facing and i < len(spec.action_facing) else -1
            w.action_beam_kind[i] = int(spec.action_beam_kind[i]) \
                if spec.action_beam_kind and i < len(spec.action_beam_kind) else R.NO_KIND
        w.beam_layer = int(spec.beam_layer)
        if spec.beam_offsets is not None:
            w.n_beam = int(spec.beam_offsets.shape[1])
            for f in range(4):
                for b in range(w.n_beam):
                    w.beam_dy[f][b], w.beam_dx[f][b] = int(spec.beam_offsets[f, b, 0]), int(spec.beam_offsets[f, b, 1])
        w.beam_block_mask = int(spec.beam_block_mask)
        w.reward_all_layers = 1 if spec.reward_all_layers else 0
        w.reward_count = int(spec.reward_count)
        self._extra = None
        if spec.obs_extra is not None:
            self._extra = np.ascontiguousarray(spec.obs_extra, dtype=np.float64).reshape(spec.height * spec.width, -1)
            w.obs_extra, w.extra = int(self._extra.shape[1]), self._extra.ctypes.data
        self._groups = []
        w.n_groups = len(spec.item_groups)
        for gi, (kind, layer, count, cells) in enumerate(spec.item_groups):
            flat = np.ascontiguousarray(np.asarray(cells)[:, 0] * spec.width + np.asarray(cells)[:, 1], dtype=np.int32)
            self._groups.append(flat)
            w.group_kind[gi], w.group_layer[gi], w.group_count[gi] = int(kind), int(layer), int(count)
            w.group_n_cells[gi], w.group_cells[gi] = len(flat), flat.ctypes.data
        # ---- agents with a second kind (examples/tag) ----
        w.agent_kind_alt = int(spec.agent_kind_alt)
        w.touch_swap, w.obs_flag = int(bool(spec.touch_swap)), int(bool(spec.obs_flag))
        w.reward_own = 1 if spec.reward_own is not None else 0
        if spec.reward_own is not None:  # the target's value is never read in this mode
            for k in range(w.K):
                w.value[k] = float(spec.reward_own[k])
        if spec.initial_flags is not None:
            for a, f in enumerate(spec.initial_flags):
                w.initial_flags[a] = int(f)
        self._w = w
        self.threads = threads
        self.h = self.L.or_create(C.byref(w), self.N, self.capacity, n_frames, rng, seed, env_id_base, threads,
                                  1 if record else 0)
        self.plane_stride = R.plane_stride(spec.height, spec.width)
        self._T = 0
        if rng == RNG_SPARSE:
            cdf = R.binomial_cdf_u64(len(spec.dynamic_layers) * self.plane_stride, float(spec.spawn_prob))
            self._cdf = np.array(cdf, dtype=np.uint64)
            self.L.or_set_sparse_cdf(self.h, self._cdf.ctypes.data)
        if env_ids is not None:
            ids = np.ascontiguousarray(env_ids, dtype=np.uint64)
            assert ids.shape == (self.N,)
            self.L.or_set_env_ids(self.h, ids.ctypes.data)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.or_destroy(self.h)
            self.h = None

    def reset(self):
        self.L.or_reset(self.h)

    def step(self, T: int = 1):
        self.L.or_step(self.h, T)
        self._T = T

    def add_empty(self):
        self.L.or_add_empty(self.h)

    def set_tick(self, tick: int):
        self.L.or_set_tick(self.h, tick)

    # -- views into the C state (valid until the next call / destruction) -----------------------
    @property
    def grid(self):
        s = self.spec
        return _view(self.L.or_grid(self.h), (self.N, s.height, s.width, s.layers), np.uint8)

    @property
    def pos(self):
        return _view(self.L.or_pos(self.h), (self.N, self.spec.num_agents), np.int32)

    @property
    def facing(self):
        return _view(self.L.or_facing(self.h), (self.N, self.spec.num_agents), np.int32)

    @property
    def flags(self):
        return _view(self.L.or_flags(self.h), (self.N, self.spec.num_agents), np.int32)

    @property
    def total_reward(self):
        return _view(self.L.or_total(self.h), (self.N,), np.float64)

    @property
    def idx(self):
        return self.L.or_idx(self.h)

    @property
    def size(self):
        return self.L.or_size(self.h)

    def ring(self):
        rows, cap, obs = self.N * self.spec.num_agents, self.capacity, self.spec.obs_size
        return {
            "states": _view(self.L.or_states(self.h), (rows, cap, obs), np.float32),
            "actions": _view(self.L.or_actions(self.h), (cap, rows), np.int64),
            "rewards": _view(self.L.or_rewards(self.h), (cap, rows), np.float32),
            "dones": _view(self.L.or_dones(self.h), (cap, rows), np.float32),
            "positions": _view(self.L.or_positions(self.h), (cap, rows, 2), np.int64),
        }

    def tapes(self):
        """Dense replay tapes of the LAST step() call, kernel layout [T, N, n_dyn*plane_stride]."""
        T, es = self._T, len(self.spec.dynamic_layers) * self.plane_stride
        return {
            "spawn_u": _view(self.L.or_tape_u(self.h), (T, self.N, es), np.float64),
            "spawn_c": _view(self.L.or_tape_c(self.h), (T, self.N, es), np.uint8),
            "actions": _view(self.L.or_tape_a(self.h), (T, self.N, self.spec.num_agents), np.int32),
            "placement": _view(self.L.or_tape_placement(self.h), (self.N, self.spec.num_agents), np.int32),
        }


def max_threads() -> int:
    return int(lib().or_max_threads())
