"""Low-level driver of the CUDA kernels: device-resident state of N world instances.

This is the layer directly above the C ABI.  It owns the PyTorch tensors (device memory is
PyTorch's job, see include/sorrel_b200.h "Conventions"), fills `SbWorldDesc` from a `WorldSpec`
and exposes one method per entry point.  The reference-shaped classes (`Gridworld`,
`Environment`, `Buffer`, ...) in the sibling modules are thin layers over this one.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _abi
from ._abi import SbStepArgs, SbWorldDesc, check, lib

# MovingAgent decode (sorrel/agents/agent.py:187-213): ["up","down","left","right"] -> (dy, dx)
MOVE_DELTAS = {"up": (-1, 0), "down": (1, 0), "left": (0, -1), "right": (0, 1)}


@dataclass
class WorldSpec:
    """A world family: kind table (one row per entity class), template and agent setup.

    `template[H, W, L]` is the map right after `populate_environment()` without the agents;
    `spawnable` lists the (y, x) cells on `agent_layer` an agent may be placed on, in the order
    the reference enumerates them (np.ndindex, examples/treasurehunt/env.py:122-137).
    """

    height: int
    width: int
    layers: int
    kind_names: list[str]
    kind_value: list[float]
    kind_passable: list[bool]
    kind_rule: list[int]
    appearance: np.ndarray  # [K, C] float64
    template: np.ndarray  # [H, W, L] uint8
    spawnable: np.ndarray  # [n, 2] (y, x)
    agent_kind: int
    agent_layer: int
    default_kind: int
    fill_kind: int
    vision: int
    num_agents: int
    spawn_prob: float = 0.0
    spawn_choices: list[int] = field(default_factory=list)
    action_deltas: list[tuple[int, int]] = field(
        default_factory=lambda: [MOVE_DELTAS[a] for a in ("up", "down", "left", "right")])
    dynamic_layers: list[int] | None = None  # default: every layer holding a non-static kind
    obs_transform: int = 0  # _abi.SB_OBS_SUM (one-hot / raw sums) or _abi.SB_OBS_RGB
    # ---- column rules, facings + beams, all-layer rewards, positional code (examples/cleanup) ----
    kind_rule_prob: list[float] | None = None  # SB_RULE_SPAWN_TO probability per kind
    kind_rule_to: list[int] | None = None  # resulting kind of SPAWN_TO / BECOME_IF / BECOME
    kind_rule_else: list[int] | None = None  # SPAWN_TO: kind when the draw fails (SB_NO_KIND = unchanged)
    kind_rule_layer: list[int] | None = None  # BECOME_IF: layer of the same column that is tested
    kind_rule_mask: list[int] | None = None  # BECOME_IF: bit j set <=> kind j satisfies the test
    action_facing: list[int] | None = None  # facing 0 N, 1 E, 2 S, 3 W an action sets; -1 unchanged
    action_beam_kind: list[int] | None = None  # kind an action's beam places; SB_NO_KIND = none
    beam_layer: int = -1
    beam_offsets: np.ndarray | None = None  # [4, n, 2] (dy, dx) per facing
    beam_block_mask: int = 0
    reward_all_layers: bool = False
    reward_count: int = 1
    initial_facing: int = 2  # MovingAgent.direction default (agents/agent.py:185)
    obs_extra: np.ndarray | None = None  # [H, W, E] appended to the flattened crop (positional code)
    item_groups: list = field(default_factory=list)  # (kind, layer, count, cells[n, 2]) placed at reset, in order
    # ---- agents with a binary state shown as a second kind (examples/tag) ----
    agent_kind_alt: int = -1
    initial_flags: np.ndarray | None = None  # [A] 0/1 at construction (every instance starts with the same flags)
    touch_swap: bool = False
    reward_own: list[float] | None = None  # reward by the agent's own kind after the exchange
    obs_flag: bool = False
    toroidal: bool = False  # engine option (not a reference behaviour): crops and moves wrap around the map edges

    @property
    def channels(self) -> int:
        return int(self.appearance.shape[1])

    @property
    def obs_side(self) -> int:
        return 2 * self.vision + 1

    @property
    def onehot_size(self) -> int:
        return self.channels * self.obs_side ** 2

    @property
    def obs_size(self) -> int:
        return self.onehot_size + (0 if self.obs_extra is None else int(self.obs_extra.shape[2])) + \
            (1 if self.obs_flag else 0)

    @property
    def uses_facing(self) -> bool:
        return self.action_facing is not None or self.action_beam_kind is not None

    def resolved_dynamic_layers(self) -> list[int]:
        if self.dynamic_layers is not None:
            return sorted(self.dynamic_layers)
        dyn = {self.agent_layer}
        if self.beam_layer >= 0:
            dyn.add(self.beam_layer)
        for z in range(self.layers):
            kinds = np.unique(self.template[:, :, z])
            if any(self.kind_rule[k] != _abi.SB_RULE_NONE for k in kinds):
                dyn.add(z)
            # a static layer that is VISIBLE (iowa's Sand floor) is stored per instance as well: the one-hot fast
            # path reads stored planes only; an invisible one (treasurehunt's Sand) stays in the shared template
            elif any(np.any(self.appearance[k] != 0) for k in kinds):
                dyn.add(z)
        return sorted(dyn)


def spawn_threshold_53(p: float) -> int:
    """ceil(p * 2^53): numpy's `random_sample() < p` is `k < this` for the 53-bit integer k."""
    import math

    if not p > 0.0:
        return 0
    if p >= 1.0:
        return 1 << 53
    return int(math.ceil(math.ldexp(p, 53)))


def binomial_cdf_u64(n_cells: int, p: float, kmax: int = _abi.SB_SPARSE_KMAX) -> list[int] | None:
    """SbWorldDesc.spawn_cdf: floor(2^64 * P(X <= k)), X ~ Binomial(n_cells, P/2^53), exact rational
    arithmetic.  None when the table cannot hold (almost) all of the mass (dense spawning) -- then
    only the per-cell mode SB_RNG_PHILOX is available."""
    from fractions import Fraction

    P = spawn_threshold_53(p)
    if P == 0 or P >= 1 << 53:
        return None
    q = Fraction(P, 1 << 53)
    if n_cells * q > 16:
        return None
    out, acc, term = [], Fraction(0), (1 - q) ** n_cells
    for k in range(kmax):
        acc += term
        out.append(min((acc.numerator << 64) // acc.denominator, (1 << 64) - 1))
        term = term * (n_cells - k) / (k + 1) * q / (1 - q)
    if (1 << 64) - out[-1] > 1 << 8:  # tail beyond kmax must be negligible (< 2^-56)
        return None
    return out


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def build_desc(spec: WorldSpec, num_envs: int, env_id_base: int, dyn_layers: list[int], env_stride: int,
               h_appearance: np.ndarray, d_appearance: int, d_template: int, static_layers_invisible: bool,
               obs_extra_width: int = 0, d_obs_extra: int = 0, d_agent_flag: int = 0) -> tuple[SbWorldDesc, bool]:
    """Fill `SbWorldDesc` from a `WorldSpec` (pure host code: the pointers are passed as integers, so the
    descriptor logic and `sb_validate` are testable without a GPU).  Returns (descriptor, sparse mode available)."""
    H, W, L = spec.height, spec.width, spec.layers
    d = SbWorldDesc()
    d.abi_version = _abi.SB_ABI_VERSION
    d.num_envs, d.num_agents = num_envs, int(spec.num_agents)
    d.height, d.width, d.num_layers = H, W, L
    d.num_dyn_layers = len(dyn_layers)
    for i, z in enumerate(dyn_layers):
        d.dyn_layers[i] = z
    d.agent_layer = spec.agent_layer
    d.num_kinds = len(spec.kind_names)
    d.num_channels = spec.channels
    d.vision = spec.vision
    d.fill_kind, d.agent_kind, d.default_kind = spec.fill_kind, spec.agent_kind, spec.default_kind
    d.num_actions = len(spec.action_deltas)
    for i, (dy, dx) in enumerate(spec.action_deltas[: _abi.SB_MAX_ACTIONS]):
        d.action_dy[i], d.action_dx[i] = dy, dx
    d.num_spawn_choices = len(spec.spawn_choices)
    for i, k in enumerate(spec.spawn_choices[: _abi.SB_MAX_CHOICES]):
        d.spawn_choice_kinds[i] = k
    d.spawn_prob = float(spec.spawn_prob)
    d.env_id_base = int(env_id_base)
    for k in range(min(len(spec.kind_names), _abi.SB_MAX_KINDS)):
        d.kind_value[k] = float(spec.kind_value[k])
        d.kind_passable[k] = 1 if spec.kind_passable[k] else 0
        d.kind_rule[k] = int(spec.kind_rule[k])
    d.d_appearance = int(d_appearance)
    d.h_appearance = h_appearance.ctypes.data
    d.d_template = int(d_template)
    d.static_layers_invisible = 1 if static_layers_invisible else 0
    d.obs_transform = int(spec.obs_transform)
    K = min(len(spec.kind_names), _abi.SB_MAX_KINDS)
    for k in range(K):
        d.kind_rule_prob[k] = float(spec.kind_rule_prob[k]) if spec.kind_rule_prob else 0.0
        d.kind_rule_to[k] = int(spec.kind_rule_to[k]) if spec.kind_rule_to else _abi.SB_NO_KIND
        d.kind_rule_else[k] = int(spec.kind_rule_else[k]) if spec.kind_rule_else else _abi.SB_NO_KIND
        d.kind_value64[k] = float(spec.kind_value[k])
        if spec.reward_own is not None:  # the target's value is never read in this mode: the table holds own rewards
            d.kind_value[k] = d.kind_value64[k] = float(spec.reward_own[k])
        d.kind_rule_layer[k] = int(spec.kind_rule_layer[k]) if spec.kind_rule_layer else -1
        d.kind_rule_mask[k] = int(spec.kind_rule_mask[k]) if spec.kind_rule_mask else 0
    # values that float32 cannot hold (iowa: 1.1, -1.4, ...): total_reward must accumulate the Python floats
    d.use_value64 = 1 if any(float(np.float32(v)) != float(v) for v in spec.kind_value) else 0
    d.agent_kind_alt = int(spec.agent_kind_alt)
    d.touch_swap = 1 if spec.touch_swap else 0
    d.reward_own = 1 if spec.reward_own is not None else 0
    d.obs_flag = 1 if spec.obs_flag else 0
    d.toroidal = 1 if spec.toroidal else 0
    d.d_agent_flag = int(d_agent_flag) if d_agent_flag else None
    d.uses_facing = 1 if spec.uses_facing else 0
    for i in range(_abi.SB_MAX_ACTIONS):
        d.action_facing[i] = int(spec.action_facing[i]) if spec.action_facing and i < len(spec.action_facing) else -1
        d.action_beam_kind[i] = int(spec.action_beam_kind[i]) \
            if spec.action_beam_kind and i < len(spec.action_beam_kind) else _abi.SB_NO_KIND
    d.beam_layer = int(spec.beam_layer)
    if spec.beam_offsets is not None:
        bo = np.asarray(spec.beam_offsets)
        if bo.shape[0] != 4 or bo.shape[1] > _abi.SB_MAX_BEAM_CELLS:
            raise ValueError("beam_offsets must be [4, n <= SB_MAX_BEAM_CELLS, 2]")
        d.num_beam_cells = int(bo.shape[1])
        for f in range(4):
            for b in range(bo.shape[1]):
                d.beam_dy[f][b], d.beam_dx[f][b] = int(bo[f, b, 0]), int(bo[f, b, 1])
    d.beam_block_mask = int(spec.beam_block_mask)
    d.reward_all_layers = 1 if spec.reward_all_layers else 0
    d.reward_count = int(spec.reward_count)
    if obs_extra_width:
        d.obs_extra = int(obs_extra_width)
        d.d_obs_extra = int(d_obs_extra)
    cdf = binomial_cdf_u64(env_stride, spec.spawn_prob) if env_stride < 65536 else None
    if cdf is not None:
        d.spawn_cdf_len = len(cdf)
        for i, v in enumerate(cdf):
            d.spawn_cdf[i] = v
    return d, cdf is not None


class Engine:
    """N instances of one `WorldSpec` on one CUDA device (one rank's shard)."""

    def __init__(self, spec: WorldSpec, num_envs: int, device: torch.device | str = "cuda",
                 env_id_base: int = 0, seed: int = 0, tuning: dict | None = None):
        """`tuning` pins the launch variant (SbStepArgs.tune_*): {"lpe": 4|8|16|32, "staged": 1|-1, "par": -1};
        anything left out is chosen by the library.  Tests use it to check a specific kernel variant."""
        self.lib = lib()  # raises if the CUDA library is missing
        self.tuning = dict(tuning or {})
        unknown = set(self.tuning) - {"lpe", "staged", "par"}
        if unknown:
            raise ValueError(f"unknown tuning keys {sorted(unknown)}")
        self.spec = spec
        self.N = int(num_envs)
        self.A = int(spec.num_agents)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _abi.SorrelB200Error("sorrel_b200 runs on CUDA devices only (no CPU fallback)")
        self.seed = int(seed)
        self.tick = 0  # global turn counter (Philox counter word 1); first turn is tick 1
        self.dyn_layers = spec.resolved_dynamic_layers()
        self.plane_stride = int(self.lib.sb_plane_stride(spec.height, spec.width))
        self.n_planes = len(self.dyn_layers)
        self.env_stride = self.n_planes * self.plane_stride
        H, W, L = spec.height, spec.width, spec.layers

        # ---- shared tables ------------------------------------------------------------------
        self.h_appearance = np.ascontiguousarray(spec.appearance, dtype=np.float64)
        self.d_appearance = torch.from_numpy(self.h_appearance).to(self.device)
        tmpl = np.full((L, self.plane_stride), _abi.SB_PAD_KIND, dtype=np.uint8)
        tmpl[:, : H * W] = np.moveaxis(spec.template, 2, 0).reshape(L, H * W)
        self.d_template = torch.from_numpy(tmpl).to(self.device)
        spawn_cells = (np.asarray(spec.spawnable)[:, 0] * W + np.asarray(spec.spawnable)[:, 1]).astype(np.int32)
        self.n_spawnable = int(spawn_cells.size)
        self.d_spawnable = torch.from_numpy(spawn_cells).to(self.device)
        static_kinds = {int(k) for z in range(L) if z not in self.dyn_layers
                        for k in np.unique(spec.template[:, :, z])}
        invisible = all(not self.h_appearance[k].any() for k in static_kinds)

        # ---- per-instance state ---------------------------------------------------------------
        self.grid = torch.empty((self.N, self.n_planes, self.plane_stride), dtype=torch.uint8, device=self.device)
        self.agent_pos = torch.zeros((self.N, self.A), dtype=torch.int32, device=self.device)
        self.total_reward = torch.zeros(self.N, dtype=torch.float64, device=self.device)
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)
        # MovingAgent.direction of every agent: set at construction, NOT reset by Environment.reset
        self.agent_facing = torch.full((self.N, self.A), int(spec.initial_facing), dtype=torch.int32,
                                       device=self.device) if spec.uses_facing else None
        # flags of agents with a second kind (tag's `it`): like the agent attribute they survive reset
        self.agent_flag = None
        if spec.agent_kind_alt >= 0:
            f0 = np.zeros(self.A, dtype=np.int32) if spec.initial_flags is None else np.asarray(spec.initial_flags, dtype=np.int32)
            self.agent_flag = torch.from_numpy(np.broadcast_to(f0, (self.N, self.A)).copy()).to(self.device)
        self.d_obs_extra = None
        if spec.obs_extra is not None:  # f64 -> f32 exactly like `Buffer.states[idx] = obs` (buffers.py:56)
            ex = np.ascontiguousarray(spec.obs_extra, dtype=np.float64).reshape(H * W, -1).astype(np.float32)
            self.d_obs_extra = torch.from_numpy(ex).to(self.device)
        self.item_groups = []  # (kind, layer, count, device cells)
        for kind, layer, count, cells in spec.item_groups:
            cells = np.asarray(cells)
            flat = (cells[:, 0] * W + cells[:, 1]).astype(np.int32)
            self.item_groups.append((int(kind), int(layer), int(count), torch.from_numpy(flat).to(self.device)))

        self.desc, self.sparse_available = build_desc(
            spec, self.N, env_id_base, self.dyn_layers, self.env_stride, self.h_appearance,
            self.d_appearance.data_ptr(), self.d_template.data_ptr(), invisible,
            0 if self.d_obs_extra is None else int(self.d_obs_extra.shape[1]),
            0 if self.d_obs_extra is None else self.d_obs_extra.data_ptr(),
            0 if self.agent_flag is None else self.agent_flag.data_ptr())
        d = self.desc
        self.desc = d
        check(self.lib.sb_validate(C.byref(d)), "world description")

    # ------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def reset(self, placement: torch.Tensor | None = None, item_placement=None) -> None:
        """Environment.reset (environment.py:72-79): template copy, item groups, agent placement.

        placement: int32 [N, A] indices into the spawnable list (a replayed
        `np.random.choice(len(valid), size=A, replace=False)`), or None to draw in-kernel.
        item_placement: one int32 [N, count] tensor per item group (replayed), or None."""
        if placement is not None:
            placement = placement.to(device=self.device, dtype=torch.int32).contiguous()
            assert placement.shape == (self.N, self.A)
        groups = (_abi.SbItemGroup * max(1, len(self.item_groups)))()
        keep = []
        for g, (kind, layer, count, cells) in enumerate(self.item_groups):
            groups[g].kind, groups[g].layer, groups[g].count, groups[g].n_cells = kind, layer, count, int(cells.numel())
            groups[g].d_cells = cells.data_ptr()
            if item_placement is not None:
                ip = item_placement[g].to(device=self.device, dtype=torch.int32).contiguous()
                assert ip.shape == (self.N, count)
                keep.append(ip)
                groups[g].d_placement = ip.data_ptr()
        check(self.lib.sb_reset_ex(C.byref(self.desc), self.grid.data_ptr(), self.agent_pos.data_ptr(),
                                   self.total_reward.data_ptr(), self.d_spawnable.data_ptr(), self.n_spawnable,
                                   _ptr(placement), groups, len(self.item_groups), self.seed, self.tick + 1,
                                   self._stream()), "sb_reset")
        self._keep = (placement, keep)  # the launch is asynchronous: keep the tapes alive until the next call

    def step(self, num_turns: int = 1, *, buffer=None, actions: torch.Tensor | None = None,
             spawn_u: torch.Tensor | None = None, spawn_c: torch.Tensor | None = None,
             obs_out: torch.Tensor | None = None, reward_out: torch.Tensor | None = None,
             action_out: torch.Tensor | None = None, drew_out: torch.Tensor | None = None,
             rng: str = "philox") -> None:
        """`num_turns` fused Environment.take_turn()s (environment.py:81-93) for every instance.

        actions given -> replayed / host policy; None -> in-kernel uniform policy (RandomModel).
        spawn_u/spawn_c given -> replayed numpy draws; None -> Philox: rng="philox" draws one byte
        per cell (dense), rng="sparse" samples the same Bernoulli field sparsely (SB_RNG_SPARSE).
        With a `buffer` (buffers.Buffer) every agent-turn is recorded into its ring."""
        T = int(num_turns)
        a = SbStepArgs()
        a.mode = _abi.SB_MODE_FUSED
        a.num_turns = T
        a.tick0 = self.tick + 1
        a.seed = self.seed
        a.rng = _abi.SB_RNG_TAPE if spawn_u is not None else \
            (_abi.SB_RNG_SPARSE if rng == "sparse" else _abi.SB_RNG_PHILOX)
        a.policy = _abi.SB_POLICY_TAPE if actions is not None else _abi.SB_POLICY_UNIFORM
        a.act_agent = a.obs_agent = -1
        self._fill_ptrs(a, T, actions, spawn_u, spawn_c, obs_out, reward_out, action_out, drew_out)
        bdesc = None
        if buffer is not None:
            a.slot0 = buffer.idx
            bdesc = C.byref(buffer.desc())
        check(self.lib.sb_step(C.byref(self.desc), C.byref(a), self.grid.data_ptr(), self.agent_pos.data_ptr(),
                               self.total_reward.data_ptr(), bdesc, self._stream()), "sb_step")
        self.tick += T
        if buffer is not None:
            buffer.advance(T)

    def phase(self, *, sweep: bool = False, act_agent: int = -1, obs_agent: int = -1, buffer=None,
              actions: torch.Tensor | None = None, spawn_u=None, spawn_c=None, obs_out=None, reward_out=None,
              new_turn: bool = False, rng: str = "philox", clock=None) -> None:
        """One split-phase launch: [sweep] -> [act agent k] -> [observe agent j] (SB_MODE_SPLIT).

        Lets a live policy run between `pov()` and `act()` of each agent exactly like
        Agent.transition (agents/agent.py:155-173).  `new_turn` advances the tick first."""
        if new_turn:
            self.tick += 1
        a = SbStepArgs()
        a.mode = _abi.SB_MODE_SPLIT
        a.num_turns = 1
        a.tick0 = self.tick
        a.seed = self.seed
        a.rng = _abi.SB_RNG_TAPE if spawn_u is not None else \
            (_abi.SB_RNG_SPARSE if rng == "sparse" else _abi.SB_RNG_PHILOX)
        a.policy = _abi.SB_POLICY_TAPE if actions is not None else _abi.SB_POLICY_UNIFORM
        a.do_sweep = 1 if sweep else 0
        a.act_agent, a.obs_agent = act_agent, obs_agent
        self._fill_ptrs(a, 1, actions, spawn_u, spawn_c, obs_out, reward_out, None, None)
        if clock is not None:
            a.d_clock = clock.tensor.data_ptr()
        bdesc = None
        if buffer is not None:
            a.slot0 = buffer.idx
            bdesc = C.byref(buffer.desc())
        check(self.lib.sb_step(C.byref(self.desc), C.byref(a), self.grid.data_ptr(), self.agent_pos.data_ptr(),
                               self.total_reward.data_ptr(), bdesc, self._stream()), "sb_step(split)")

    def _fill_ptrs(self, a, T, actions, spawn_u, spawn_c, obs_out, reward_out, action_out, drew_out):
        N, A, ES = self.N, self.A, self.env_stride

        def chk(t, shape, dtype, name):
            if t is None:
                return None
            if t.device != self.grid.device or t.dtype != dtype or not t.is_contiguous() or tuple(t.shape) != shape:
                raise ValueError(f"{name}: expected contiguous {dtype} {shape} on {self.grid.device}, "
                                 f"got {t.dtype} {tuple(t.shape)} on {t.device}")
            return t.data_ptr()

        # action tapes: int32, or uint8 (an ActionSpec has at most 8 actions: a quarter of the bytes a host policy ships)
        a.actions_u8 = 1 if actions is not None and actions.dtype == torch.uint8 else 0
        a.d_actions = chk(actions, (T, N, A), torch.uint8 if a.actions_u8 else torch.int32, "actions")
        a.d_spawn_u = chk(spawn_u, (T, N, ES), torch.float64, "spawn_u")
        a.d_spawn_c = chk(spawn_c, (T, N, ES), torch.uint8, "spawn_c")
        a.d_obs_out = chk(obs_out, (T, N, A, self.spec.obs_size), torch.float32, "obs_out")
        a.d_reward_out = chk(reward_out, (T, N, A), torch.float32, "reward_out")
        a.d_action_out = chk(action_out, (T, N, A), torch.int32, "action_out")
        a.d_drew_out = chk(drew_out, (T, N, ES), torch.uint8, "drew_out")
        a.d_error = self.error.data_ptr()
        a.d_agent_facing = _ptr(self.agent_facing)
        self.apply_tuning(a)

    def apply_tuning(self, a: SbStepArgs) -> None:
        a.tune_lpe = int(self.tuning.get("lpe", 0) or 0)
        a.tune_staged = int(self.tuning.get("staged", 0) or 0)
        a.tune_par = int(self.tuning.get("par", 0) or 0)

    def step_plan(self, num_turns: int = 1, rng: str = "philox", replay: bool = False) -> dict:
        """The launch `step(...)` would make (kernel variant, lanes per instance, shared memory, grid)."""
        a = SbStepArgs()
        a.mode, a.num_turns, a.seed = _abi.SB_MODE_FUSED, int(num_turns), self.seed
        a.rng = _abi.SB_RNG_TAPE if replay else (_abi.SB_RNG_SPARSE if rng == "sparse" else _abi.SB_RNG_PHILOX)
        a.policy = _abi.SB_POLICY_UNIFORM
        a.act_agent = a.obs_agent = -1
        self.apply_tuning(a)
        plan = _abi.SbStepPlan()
        check(self.lib.sb_step_plan(C.byref(self.desc), C.byref(a), C.byref(plan)), "sb_step_plan")
        return plan.as_dict()

    def observe(self, out: torch.Tensor | None = None) -> torch.Tensor:
        """ObservationSpec.observe for every (instance, agent): float32 [N, A, C, S, S], or [N, A, obs_size] when
        rows carry more than the crop (positional code, state flag)."""
        s = self.spec
        if out is None:
            shape = (self.N, self.A, s.channels, s.obs_side, s.obs_side) if s.obs_size == s.onehot_size \
                else (self.N, self.A, s.obs_size)
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        if out.numel() != self.N * self.A * s.obs_size or out.dtype != torch.float32 or not out.is_contiguous() \
                or out.device != self.grid.device:
            raise ValueError(f"observe(out=...): expected a contiguous float32 tensor of "
                             f"{self.N * self.A * s.obs_size} elements on {self.grid.device}, got "
                             f"{tuple(out.shape)} {out.dtype} on {out.device}")
        check(self.lib.sb_observe(C.byref(self.desc), self.grid.data_ptr(), self.agent_pos.data_ptr(),
                                  out.data_ptr(), self._stream()), "sb_observe")
        return out

    def check_errors(self) -> None:
        """Synchronising read of the device error word; raises the reference's exception type."""
        bits = int(self.error.item())
        if bits:
            self.error.zero_()
            raise _abi.error_bits_to_exception(bits)

    # ------------------------------------------------------------------------------------------
    def grid_numpy(self) -> np.ndarray:
        """Read back all layers as uint8 [N, H, W, L] (static layers filled from the template)."""
        s = self.spec
        H, W, L = s.height, s.width, s.layers
        out = np.broadcast_to(s.template[None], (self.N, H, W, L)).copy()
        g = self.grid.cpu().numpy()
        for pl, z in enumerate(self.dyn_layers):
            out[:, :, :, z] = g[:, pl, : H * W].reshape(self.N, H, W)
        return out

    def grid_instance(self, env: int) -> np.ndarray:
        """One instance as uint8 [H, W, L] (rendering / debugging: reads back a single instance, not the batch)."""
        s = self.spec
        H, W = s.height, s.width
        out = s.template.copy()
        g = self.grid[env].cpu().numpy()
        for pl, z in enumerate(self.dyn_layers):
            out[:, :, z] = g[pl, : H * W].reshape(H, W)
        return out

    def agent_yx(self) -> np.ndarray:
        """int64 [N, A, 2] agent locations (y, x)."""
        p = self.agent_pos.cpu().numpy().astype(np.int64)
        return np.stack([p // self.spec.width, p % self.spec.width], axis=-1)
