"""Shared helpers for the test-suite: golden fixtures, oracle runs, oracle-spec -> engine-spec."""
from __future__ import annotations

import glob
import os

import numpy as np

from oracle import restatement as R

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(max_cells: int = 64 * 64, with_large: bool = False):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        name = os.path.basename(f)[:-4]
        if name.startswith(("observe_", "transformer_", "cleanup_", "iowa_", "tag_")):
            continue
        if name.startswith("large") and not with_large:
            continue
        out.append(name)
    return out


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def cleanup_cases():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "cleanup_*.npz")))


def cleanup_spec_from_golden(g) -> R.WorldSpec:
    return R.cleanup_spec(int(g["H"]), int(g["W"]), int(g["A"]), int(g["v"]), int(g["initial_apples"]),
                          float(g["apple_p"]), float(g["pollution_p"]), int(g["beam_radius"]),
                          mode=str(g["mode"]) if "mode" in g else "DEFAULT")


def iowa_cases():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "iowa_*.npz")))


def iowa_spec_from_golden(g) -> R.WorldSpec:
    return R.iowa_spec(int(g["H"]), int(g["W"]), int(g["A"]), int(g["v"]), float(g["p"]))


def tag_cases():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "tag_*.npz")))


def tag_spec_from_golden(g) -> R.WorldSpec:
    return R.tag_spec(int(g["H"]), int(g["W"]), int(g["A"]), int(g["v"]), 10, int(np.argmax(g["initial_flags"])))


def spec_from_golden(g) -> R.WorldSpec:
    obs = str(g["obs_spec"]) if "obs_spec" in g else "onehot"
    return R.treasurehunt_spec(int(g["H"]), int(g["W"]), int(g["A"]), int(g["v"]), float(g["p"]), obs=obs)


def engine_spec(o: R.WorldSpec):
    """oracle WorldSpec -> sorrel_b200.engine.WorldSpec (field-for-field)."""
    from sorrel_b200.engine import WorldSpec

    return WorldSpec(
        height=o.height, width=o.width, layers=o.layers, kind_names=list(o.kind_names),
        kind_value=list(o.kind_value), kind_passable=list(o.kind_passable), kind_rule=list(o.kind_rule),
        appearance=np.array(o.appearance, dtype=np.float64), template=o.template.copy(),
        spawnable=np.asarray(o.spawnable)[:, :2].copy(), agent_kind=o.agent_kind, agent_layer=o.agent_layer,
        default_kind=o.default_kind, fill_kind=o.fill_kind, vision=o.vision, num_agents=o.num_agents,
        spawn_prob=o.spawn_prob, spawn_choices=list(o.spawn_choices),
        dynamic_layers=list(o.dynamic_layers) if o.dynamic_layers else None,
        obs_transform=int(getattr(o, "obs_transform", 0)),
        action_deltas=[tuple(d) for d in o.action_deltas],
        kind_rule_prob=o.kind_rule_prob, kind_rule_to=o.kind_rule_to, kind_rule_else=o.kind_rule_else,
        kind_rule_layer=o.kind_rule_layer,
        kind_rule_mask=o.kind_rule_mask, action_facing=o.action_facing, action_beam_kind=o.action_beam_kind,
        beam_layer=o.beam_layer, beam_offsets=o.beam_offsets, beam_block_mask=o.beam_block_mask,
        reward_all_layers=o.reward_all_layers, reward_count=o.reward_count, initial_facing=o.initial_facing,
        obs_extra=o.obs_extra, item_groups=list(o.item_groups),
        agent_kind_alt=o.agent_kind_alt, initial_flags=o.initial_flags, touch_swap=o.touch_swap,
        reward_own=o.reward_own, obs_flag=o.obs_flag, toroidal=bool(getattr(o, "toroidal", False)),
    )


def plane_stride(h: int, w: int) -> int:
    return (h * w + 15) // 16 * 16


def run_oracle_episode(spec: R.WorldSpec, draws, T: int, env_id: int = 0, capacity: int | None = None,
                       n_frames: int = 1, tick0: int = 0):
    """reset + T turns; returns the env (with tapes recorded) and per-turn grids."""
    env = R.OracleEnv(spec, draws, env_id=env_id, capacity=capacity or T, n_frames=n_frames, record_tapes=True)
    env.tick = tick0
    env.reset()
    grids = []
    for _ in range(T):
        env.take_turn()
        grids.append(env.grid.copy())
    return env, np.array(grids)


def pad_tapes(spec: R.WorldSpec, tapes: dict):
    """dense [T, Ldyn*H*W] tapes -> the kernel layout [T, Ldyn*plane_stride] (NaN / 255 padding)."""
    HW = spec.height * spec.width
    ps = plane_stride(spec.height, spec.width)
    nd = len(spec.dynamic_layers)
    T = tapes["actions"].shape[0]
    u = np.full((T, nd, ps), np.nan, dtype=np.float64)
    c = np.full((T, nd, ps), 255, dtype=np.uint8)
    u[:, :, :HW] = tapes["spawn_u"].reshape(T, nd, HW)
    c[:, :, :HW] = tapes["spawn_c"].reshape(T, nd, HW)
    return u.reshape(T, nd * ps), c.reshape(T, nd * ps)
