"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage:  python -m oracle.gen_golden

Protocol: build `TreasurehuntWorld` + `TreasurehuntEnv` from /root/reference,
swap each agent's IQN for the reference's own `RandomModel` (models/base_model.py:107-111) so only
env + observation + buffer are exercised, call `set_seed(seed)` (utils/helpers.py:22-32)
immediately before `env.reset()`, then run `take_turn()` and dump everything the engine must
reproduce: per-turn kind-id grids, per-agent buffer arrays, rewards, positions, draw counts.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import shim  # noqa: E402

shim.install()

from sorrel.buffers import Buffer  # noqa: E402
from sorrel.examples.treasurehunt.entities import EmptyEntity  # noqa: E402
from sorrel.examples.treasurehunt.env import TreasurehuntEnv  # noqa: E402
from sorrel.examples.treasurehunt.world import TreasurehuntWorld  # noqa: E402
from sorrel.models.base_model import RandomModel  # noqa: E402
from sorrel.utils.helpers import set_seed  # noqa: E402

from oracle.recorder import DrawLog, record  # noqa: E402

CLASS_IDS = {"EmptyEntity": 0, "Wall": 1, "Gem": 2, "Bone": 3, "Food": 4, "TreasurehuntAgent": 5, "Sand": 6}
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def grid_ids(world) -> np.ndarray:
    H, W, L = world.map.shape
    g = np.zeros((H, W, L), dtype=np.uint8)
    for idx, e in np.ndenumerate(world.map):
        g[idx] = CLASS_IDS[type(e).__name__]
    return g


def build(H, W, A, v, p, T, obs="onehot"):
    config = {
        "experiment": {"epochs": 1, "max_turns": T, "record_period": 50},
        "model": {"agent_vision_radius": v, "epsilon_decay": 0.0005, "observation_spec": obs,
                  "num_agents": A},
        "world": {"height": H, "width": W, "gem_value": 10, "food_value": 5, "bone_value": -10,
                  "spawn_prob": p},
    }
    world = TreasurehuntWorld(config=config, default_entity=EmptyEntity())
    env = TreasurehuntEnv(world, config)
    return world, env


def run_case(name, seed, H, W, A, v, p, T, epochs=1, capacity=None, n_frames=1, sample_batch=0,
             per_turn_grids=True, obs="onehot"):
    world, env = build(H, W, A, v, p, T, obs)
    cap = capacity or T * epochs
    for ag in env.agents:
        m = RandomModel(ag.observation_spec.input_size, 4, cap)
        # same construction the example uses for IQN's memory (env.py:100-102) + frame stacking
        m.memory = Buffer(capacity=cap, obs_shape=tuple(ag.observation_spec.input_size), n_frames=n_frames)
        m.memory.extra_data["positions"] = np.zeros((cap, 2), dtype=np.int64)
        ag.model = m
    log = DrawLog()
    grids, starts, totals, cur_states, n_draws = [], [], [], [], []
    samples = {}
    with record(log):
        set_seed(seed)
        for ep in range(epochs):
            env.reset()
            starts.append([ag.location[:2] for ag in env.agents])
            if n_frames > 1:
                for ag in env.agents:  # iRainbowModel.start_epoch_action iqn.py:426-436
                    ag.model.memory.add_empty()
            for t in range(T):
                env.take_turn()
                if per_turn_grids:
                    grids.append(grid_ids(world))
                n_draws.append(len(log))
                if n_frames > 1:
                    cur_states.append(np.stack([ag.model.memory.current_state() for ag in env.agents]))
            world.is_done = True  # environment.py:171
            totals.append(world.total_reward)
            if not per_turn_grids:
                grids.append(grid_ids(world))
        if sample_batch:
            outs = [ag.model.memory.sample(sample_batch) for ag in env.agents]
            sample_idx = [c[2] for c in log.calls if c[0] == "choice" and c[1][1] == sample_batch]
            samples = {
                "sample_idx": np.stack(sample_idx[-len(env.agents):]).astype(np.int64),
                "sample_states": np.stack([o[0] for o in outs]),
                "sample_actions": np.stack([o[1] for o in outs]),
                "sample_rewards": np.stack([o[2] for o in outs]),
                "sample_next_states": np.stack([o[3] for o in outs]),
                "sample_dones": np.stack([o[4] for o in outs]),
                "sample_valid": np.stack([o[5] for o in outs]),
            }
    mems = [ag.model.memory for ag in env.agents]
    h = hashlib.sha256()
    for b in mems:
        for arr in (b.states, b.actions, b.rewards, b.dones, b.extra_data["positions"]):
            h.update(arr.tobytes())
    data = dict(
        seed=seed, H=H, W=W, A=A, v=v, p=p, T=T, epochs=epochs, capacity=cap, n_frames=n_frames,
        starts=np.array(starts, dtype=np.int64),
        grids=np.array(grids, dtype=np.uint8),
        total_reward=np.array(totals, dtype=np.float64),
        states=np.stack([b.states for b in mems]),
        actions=np.stack([b.actions for b in mems]),
        rewards=np.stack([b.rewards for b in mems]),
        dones=np.stack([b.dones for b in mems]),
        positions=np.stack([b.extra_data["positions"] for b in mems]),
        buf_idx=np.array([b.idx for b in mems]), buf_size=np.array([b.size for b in mems]),
        n_draws=np.array(n_draws, dtype=np.int64),
        n_uniform=np.int64(sum(1 for c in log.calls if c[0] == "random")),
        n_spawn=np.int64(sum(1 for c in log.calls if c[0] == "choice_obj")),
        obs_spec=obs,
        buf_sha=h.hexdigest()[:16],
        grid_sha=hashlib.sha256(grids[-1][:, :, 1].tobytes()).hexdigest()[:16],
        **samples,
    )
    if cur_states:
        data["current_states"] = np.array(cur_states, dtype=np.float32)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **data)
    print(f"{name}: total_reward={totals} buf_sha={data['buf_sha']} grid_sha={data['grid_sha']} "
          f"draws={len(log)} -> {os.path.getsize(path)} B")


def transformer_case():
    """TransformerBuffer.combine + sample_transformer (buffers.py:273-355) on a 2-agent episode."""
    from sorrel.buffers import TransformerBuffer

    T, cap, F = 40, 64, 4
    world, env = build(11, 11, 2, 2, 0.05, T)
    for ag in env.agents:
        m = RandomModel(ag.observation_spec.input_size, 4, cap)
        m.memory = TransformerBuffer(capacity=cap, obs_shape=tuple(ag.observation_spec.input_size), n_frames=F)
        m.memory.extra_data["positions"] = np.zeros((cap, 2), dtype=np.int64)
        ag.model = m
    log = DrawLog()
    with record(log):
        set_seed(21)
        env.reset()
        for _ in range(T):
            env.take_turn()
        comb = TransformerBuffer.combine([ag.model.memory for ag in env.agents])
        out = comb.sample_transformer(24)
        idx = [c[2] for c in log.calls if c[0] == "choice" and c[1][1] == 24][-1]
    mems = [ag.model.memory for ag in env.agents]
    np.savez_compressed(
        os.path.join(OUT, "transformer_s21_a2_11.npz"), seed=21, H=11, W=11, A=2, v=2, p=0.05, T=T, capacity=cap,
        n_frames=F, states=np.stack([b.states for b in mems]), actions=np.stack([b.actions for b in mems]),
        comb_states=comb.states, comb_actions=comb.actions, comb_rewards=comb.rewards, comb_dones=comb.dones,
        comb_positions=comb.extra_data["positions"], comb_agent_ids=comb.extra_data["agent_ids"],
        comb_idx=comb.idx, comb_size=comb.size, sample_idx=np.asarray(idx, dtype=np.int64),
        t_states=out[0], t_actions=out[1], t_next_actions=out[2], t_next_states=out[3], t_dones=out[4],
        t_valid=out[5], t_agent_ids=out[6])
    print("transformer_s21_a2_11: sampled", out[0].shape, out[1].dtype, out[2].dtype, out[6][:6])


def observe_cases(rgb_only=False):
    """Direct OneHotObservationSpec.observe probes on random multi-layer worlds."""
    from sorrel.entities import Entity
    from sorrel.observation.observation_spec import OneHotObservationSpec
    from sorrel.worlds import Gridworld

    names = ["EmptyEntity", "Wall", "Gem", "Bone"]

    def mk(kind):
        e = Entity()
        e.kind = kind
        return e

    rng = np.random.RandomState(1234)
    cases = []
    for _ in range(0 if rgb_only else 60):
        H, W, L = rng.randint(5, 12), rng.randint(5, 12), rng.randint(1, 4)
        v = rng.randint(1, min((H - 1) // 2, (W - 1) // 2) + 1)
        world = Gridworld(H, W, L, mk("EmptyEntity"))
        kinds = rng.randint(0, 4, size=(H, W, L)).astype(np.uint8)
        for idx in np.ndindex(H, W, L):
            world.add(idx, mk(names[kinds[idx]]))
        spec = OneHotObservationSpec(names, full_view=False, vision_radius=int(v))
        for _ in range(5):
            y, x = rng.randint(0, H), rng.randint(0, W)
            obs = spec.observe(world, (y, x, 0))
            cases.append((kinds, int(v), int(y), int(x), obs))
    # RGBObservationSpec on the same kind of worlds (clip(sum, 0, 255) / 255, observation_spec.py:483)
    from sorrel.observation.observation_spec import RGBObservationSpec

    rgb_cases = []
    rng = np.random.RandomState(4321)  # own stream: independent of the one-hot probes above
    for _ in range(30):
        H, W, L = rng.randint(5, 12), rng.randint(5, 12), rng.randint(1, 4)
        v = rng.randint(1, min((H - 1) // 2, (W - 1) // 2) + 1)
        world = Gridworld(H, W, L, mk("EmptyEntity"))
        kinds = rng.randint(0, 4, size=(H, W, L)).astype(np.uint8)
        for idx in np.ndindex(H, W, L):
            world.add(idx, mk(names[kinds[idx]]))
        spec = RGBObservationSpec(names, full_view=False, vision_radius=int(v))
        table = np.stack([spec.entity_map[n] for n in names]).astype(np.float64)
        for _ in range(4):
            y, x = rng.randint(0, H), rng.randint(0, W)
            rgb_cases.append((kinds, int(v), int(y), int(x), spec.observe(world, (y, x, 0)), table))
    np.savez_compressed(
        os.path.join(OUT, "observe_rgb_cases.npz"), n=len(rgb_cases),
        **{f"kinds_{i}": c[0] for i, c in enumerate(rgb_cases)},
        **{f"meta_{i}": np.array(c[1:4]) for i, c in enumerate(rgb_cases)},
        **{f"obs_{i}": c[4] for i, c in enumerate(rgb_cases)},
        **{f"table_{i}": c[5] for i, c in enumerate(rgb_cases)},
    )
    print(f"observe_rgb_cases: {len(rgb_cases)} cases")
    if rgb_only:
        return
    np.savez_compressed(
        os.path.join(OUT, "observe_cases.npz"),
        n=len(cases),
        **{f"kinds_{i}": c[0] for i, c in enumerate(cases)},
        **{f"meta_{i}": np.array(c[1:4]) for i, c in enumerate(cases)},
        **{f"obs_{i}": c[4] for i, c in enumerate(cases)},
    )
    print(f"observe_cases: {len(cases)} cases")


CLEANUP_IDS = {"EmptyEntity": 0, "Wall": 1, "River": 2, "Pollution": 3, "AppleTree": 4, "Apple": 5, "CleanBeam": 6,
               "ZapBeam": 7, "CleanupAgent": 8, "Sand": 9}


def cleanup_grid_ids(world) -> np.ndarray:
    """Kind ids of a CleanupWorld; a beam whose turn_counter is 1 (swept once) gets id 10 / 11."""
    H, W, L = world.map.shape
    g = np.zeros((H, W, L), dtype=np.uint8)
    for idx, e in np.ndenumerate(world.map):
        k = CLEANUP_IDS[type(e).__name__]
        if k in (6, 7) and e.turn_counter >= 1:
            k += 4
        g[idx] = k
    return g


def run_cleanup_case(name, seed, H, W, A, v, T, epochs=1, initial_apples=20, apple_p=0.002, pollution_p=0.009,
                     beam_radius=3, n_frames=1, mode="DEFAULT"):
    """examples/cleanup with the reference's RandomModel (6 actions) instead of IQN; everything else unmodified."""
    from omegaconf import OmegaConf
    from sorrel.examples.cleanup.entities import EmptyEntity as CEmpty
    from sorrel.examples.cleanup.env import CleanupEnv
    from sorrel.examples.cleanup.world import CleanupWorld

    cfg = OmegaConf.create({
        "experiment": {"name": "cleanup", "epochs": epochs, "max_turns": T, "epsilon_decay": 0.0001,
                       "record_period": 50},
        "env": {"height": H, "width": W, "layers": 3, "default_object": "EmptyEntity", "full_mdp": False,
                "pollution_threshold": 0.5, "initial_apples": initial_apples, "apple_spawn_chance": apple_p,
                "pollution_spawn_chance": pollution_p, "mode": mode},
        "model": {"iqn": {"type": "iRainbowModel", "parameters": {
            "layer_size": 16, "n_step": 3, "batch_size": 8, "memory_size": 64, "LR": 0.00025, "TAU": .001,
            "GAMMA": 0.99, "n_quantiles": 4, "sync_freq": 200, "model_update_freq": 4, "epsilon": 0.8,
            "device": "cpu"}}, "epsilon_decay": 0.0001},
        "agent": {"agent": {"num": A, "model": "iqn", "beam_radius": beam_radius, "rotation": False,
                            "obs": {"channels": 8, "vision": v, "n_frames": n_frames, "embeddings": 3}}},
        "root": "x", "log": False})
    world = CleanupWorld(config=cfg, default_entity=CEmpty())
    env = CleanupEnv(world, cfg)
    cap = T * epochs + (n_frames - 1) * epochs
    for ag in env.agents:
        m = RandomModel(ag.observation_spec.input_size, 6, cap)
        m.memory = Buffer(capacity=cap, obs_shape=(ag.observation_spec.input_size[1],), n_frames=n_frames)
        m.memory.extra_data["positions"] = np.zeros((cap, 2), dtype=np.int64)
        ag.model = m
    log = DrawLog()
    grids, starts, totals, facings = [], [], [], []
    with record(log):
        set_seed(seed)
        for ep in range(epochs):
            env.reset()
            starts.append([ag.location[:2] for ag in env.agents])
            if n_frames > 1:
                for ag in env.agents:
                    ag.model.memory.add_empty()
            grids.append(cleanup_grid_ids(world))  # state right after reset (apples + agents placed)
            for t in range(T):
                env.take_turn()
                grids.append(cleanup_grid_ids(world))
                facings.append([ag.direction for ag in env.agents])
            world.is_done = True
            totals.append(world.total_reward)
    mems = [ag.model.memory for ag in env.agents]
    data = dict(
        world="cleanup", mode=mode, seed=seed, H=H, W=W, A=A, v=v, T=T, epochs=epochs, capacity=cap, n_frames=n_frames,
        initial_apples=initial_apples, apple_p=apple_p, pollution_p=pollution_p, beam_radius=beam_radius,
        starts=np.array(starts, dtype=np.int64), grids=np.array(grids, dtype=np.uint8),
        facings=np.array(facings, dtype=np.int64), total_reward=np.array(totals, dtype=np.float64),
        states=np.stack([b.states for b in mems]), actions=np.stack([b.actions for b in mems]),
        rewards=np.stack([b.rewards for b in mems]), dones=np.stack([b.dones for b in mems]),
        positions=np.stack([b.extra_data["positions"] for b in mems]),
        buf_idx=np.array([b.idx for b in mems]), buf_size=np.array([b.size for b in mems]),
        n_uniform=np.int64(sum(1 for c in log.calls if c[0] == "random")),
    )
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **data)
    acts = np.stack([b.actions for b in mems])
    print(f"{name}: total_reward={totals} uniforms={int(data['n_uniform'])} beams={(acts >= 4).sum()} "
          f"rewards!=0: {(np.stack([b.rewards for b in mems]) != 0).sum()} -> {os.path.getsize(path)} B")


def cleanup_cases():
    run_cleanup_case("cleanup_s3_a10_21x31", 3, 21, 31, 10, 5, 60)  # the example's own shape (configs/config.yaml)
    run_cleanup_case("cleanup_s5_a4_15x16", 5, 15, 16, 4, 3, 80, initial_apples=8, apple_p=0.05, pollution_p=0.1,
                     beam_radius=2)  # dense dynamics: many spawns, cleaned pollution, eaten apples
    run_cleanup_case("cleanup_s8_a6_12x19", 8, 12, 19, 6, 2, 40, epochs=2, initial_apples=5, apple_p=0.02,
                     pollution_p=0.05, n_frames=3)  # two episodes: facings persist across reset, ring add_empty
    if len(sys.argv) > 2 and sys.argv[2] != "apple":
        return
    run_cleanup_case("cleanup_apple_s9_a5_11x14", 9, 11, 14, 5, 3, 40, initial_apples=12, apple_p=0.03,
                     beam_radius=2, mode="APPLE")  # env.mode == "APPLE": orchard everywhere, no river


IOWA_IDS = {"EmptyEntity": 0, "Wall": 1, "Sand": 2, "GamblingAgent": 7}
IOWA_DECK_VALUES = {"a": (1.1, -1.4), "b": (1.1, -11.4), "c": (0.6, 0.1), "d": (0.6, -1.9)}


def iowa_grid_ids(world) -> np.ndarray:
    """Kind ids of a GamblingWorld; a Deck's mutable value selects one of its three rows (fresh / win / loss)."""
    H, W, L = world.map.shape
    g = np.zeros((H, W, L), dtype=np.uint8)
    for idx, e in np.ndenumerate(world.map):
        name = type(e).__name__
        if name == "Deck":
            i = "abcd".index(e.name)
            win, lose = IOWA_DECK_VALUES[e.name]
            if e.value == 0:
                g[idx] = 3 + i
            elif e.value == win:
                g[idx] = 8 + 2 * i
            else:
                assert e.value == lose, (e.name, e.value)
                g[idx] = 9 + 2 * i
        else:
            g[idx] = IOWA_IDS[name]
    return g


def run_iowa_case(name, seed, H, W, v, p, T, epochs=1):
    """examples/iowa with the reference's RandomModel instead of IQN (2 agents, as iowa/env.py:41 hard-codes)."""
    from sorrel.examples.iowa.entities import EmptyEntity as IEmpty
    from sorrel.examples.iowa.env import GamblingEnv
    from sorrel.examples.iowa.world import GamblingWorld

    config = {"experiment": {"epochs": epochs, "max_turns": T, "record_period": 50},
              "model": {"agent_vision_radius": v, "epsilon_decay": 0.0005,
                        "parameters": {"layer_size": 16, "epsilon": 0.5, "device": "cpu", "n_frames": 1}},
              "world": {"height": H, "width": W, "spawn_prob": p}}
    world = GamblingWorld(config=config, default_entity=IEmpty())
    env = GamblingEnv(world, config)
    cap = T * epochs
    for ag in env.agents:
        m = RandomModel(ag.observation_spec.input_size, 4, cap)
        m.memory = Buffer(capacity=cap, obs_shape=tuple(ag.observation_spec.input_size), n_frames=1)
        m.memory.extra_data["positions"] = np.zeros((cap, 2), dtype=np.int64)
        ag.model = m
    log = DrawLog()
    grids, starts, totals = [], [], []
    with record(log):
        set_seed(seed)
        for ep in range(epochs):
            env.reset()
            starts.append([ag.location[:2] for ag in env.agents])
            for t in range(T):
                env.take_turn()
                grids.append(iowa_grid_ids(world))
            world.is_done = True
            totals.append(world.total_reward)
    mems = [ag.model.memory for ag in env.agents]
    data = dict(
        world="iowa", seed=seed, H=H, W=W, A=len(env.agents), v=v, p=p, T=T, epochs=epochs, capacity=cap, n_frames=1,
        starts=np.array(starts, dtype=np.int64), grids=np.array(grids, dtype=np.uint8),
        total_reward=np.array(totals, dtype=np.float64),
        states=np.stack([b.states for b in mems]), actions=np.stack([b.actions for b in mems]),
        rewards=np.stack([b.rewards for b in mems]), dones=np.stack([b.dones for b in mems]),
        positions=np.stack([b.extra_data["positions"] for b in mems]),
        buf_idx=np.array([b.idx for b in mems]), buf_size=np.array([b.size for b in mems]),
        n_uniform=np.int64(sum(1 for c in log.calls if c[0] == "random")),
    )
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **data)
    r = np.stack([b.rewards for b in mems])
    print(f"{name}: total_reward={totals} uniforms={int(data['n_uniform'])} distinct rewards={np.unique(r)} "
          f"-> {os.path.getsize(path)} B")


def iowa_cases():
    run_iowa_case("iowa_s4_20", 4, 20, 20, 2, 0.01, 100)  # the example's own shape (iowa/main.py)
    run_iowa_case("iowa_s6_9x12", 6, 9, 12, 2, 0.08, 60, epochs=2)  # dense decks: every value state is visited


def run_tag_case(name, seed, H, W, A, v, T, epochs=1):
    """examples/tag (base Gridworld, 1 layer) with the reference's RandomModel instead of IQN."""
    from sorrel.entities.basic_entities import EmptyEntity as BEmpty
    from sorrel.examples.tag.env import TagEnv
    from sorrel.worlds import Gridworld

    config = {"experiment": {"epochs": epochs, "max_turns": T, "record_period": 50},
              "model": {"epsilon_decay": 0.0001}, "agent": {"num_agents": A, "vision_radius": v},
              "world": {"height": H, "width": W, "layers": 1}}
    np.random.seed(seed + 1000)  # the It agent is drawn in the constructor (tag/env.py:91-94)
    world = Gridworld(**config["world"], default_entity=BEmpty())
    env = TagEnv(world, config)
    cap = T * epochs
    ids = {"EmptyEntity": 0, "Wall": 1, "It": 2, "NotIt": 3}
    initial_flags = np.array([int(ag.it) for ag in env.agents], dtype=np.int64)
    for ag in env.agents:
        n_in = int(ag.observation_spec.input_size[0])
        m = RandomModel((n_in,), 4, cap)
        m.memory = Buffer(capacity=cap, obs_shape=(n_in,), n_frames=1)
        m.memory.extra_data["positions"] = np.zeros((cap, 2), dtype=np.int64)
        ag.model = m

    def grid_ids(w):
        g = np.zeros(w.map.shape, dtype=np.uint8)
        for idx, e in np.ndenumerate(w.map):
            g[idx] = ids[e.kind]
        return g

    log = DrawLog()
    grids, starts, totals, flags = [], [], [], []
    with record(log):
        set_seed(seed)
        for ep in range(epochs):
            env.reset()
            starts.append([ag.location[:2] for ag in env.agents])
            grids.append(grid_ids(world))
            for t in range(T):
                env.take_turn()
                grids.append(grid_ids(world))
                flags.append([int(ag.it) for ag in env.agents])
            world.is_done = True
            totals.append(world.total_reward)
    mems = [ag.model.memory for ag in env.agents]
    data = dict(
        world="tag", seed=seed, H=H, W=W, A=A, v=v, T=T, epochs=epochs, capacity=cap, n_frames=1,
        initial_flags=initial_flags, starts=np.array(starts, dtype=np.int64), grids=np.array(grids, dtype=np.uint8),
        flags=np.array(flags, dtype=np.int64), total_reward=np.array(totals, dtype=np.float64),
        states=np.stack([b.states for b in mems]), actions=np.stack([b.actions for b in mems]),
        rewards=np.stack([b.rewards for b in mems]), dones=np.stack([b.dones for b in mems]),
        positions=np.stack([b.extra_data["positions"] for b in mems]),
    )
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **data)
    fl = np.array(flags)
    print(f"{name}: total_reward={totals} it-agent changes={(np.diff(fl.argmax(axis=1)) != 0).sum()} "
          f"-> {os.path.getsize(path)} B")


def tag_cases():
    run_tag_case("tag_s2_a5_11", 2, 11, 11, 5, 4, 60)  # the example's own shape (tag/main.py), 3 x its episode length
    run_tag_case("tag_s3_a7_8x9", 3, 8, 9, 7, 3, 40, epochs=2)  # crowded: the role changes hands often


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "transformer":
        transformer_case()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "tag":
        tag_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "iowa":
        iowa_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "cleanup":
        cleanup_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "rgb":  # only the fixtures added after the first generation
        run_case("rgb_s13_a2_11", 13, 11, 11, 2, 2, 0.05, 40, obs="rgb")
        observe_cases(rgb_only=True)
        return
    run_case("kat1_s0_a1_21", 0, 21, 21, 1, 2, 0.005, 100)
    run_case("kat2_s0_a8_21", 0, 21, 21, 8, 2, 0.005, 100)
    run_case("kat3_s7_a2_11", 7, 11, 11, 2, 2, 0.05, 50)
    for s in range(1, 7):
        run_case(f"c1_s{s}_a1_21", s, 21, 21, 1, 2, 0.005, 100)
    run_case("dense_s3_a3_7", 3, 7, 7, 3, 3, 0.2, 40)  # v = (H-1)//2: heavy out-of-bounds fill
    run_case("crowd_s5_a12_8", 5, 8, 8, 12, 1, 0.1, 60)  # 12 agents on 36 cells: many conflicts
    run_case("mid_s2_a8_32", 2, 32, 32, 8, 4, 0.005, 30)
    run_case("rect_s9_a4_9x15", 9, 9, 15, 4, 3, 0.03, 50)
    # ring wrap + add_empty + current_state + sample: 3 epochs, capacity 64 < 3*(40+4), F=5
    run_case("ring_s11_a2_11", 11, 11, 11, 2, 2, 0.05, 40, epochs=3, capacity=64, n_frames=5,
             sample_batch=16)
    run_case("ring_s12_a1_21", 12, 21, 21, 1, 2, 0.005, 100, epochs=2, capacity=256, n_frames=5,
             sample_batch=64)
    run_case("large_s4_a64_128", 4, 128, 128, 64, 4, 0.005, 2, per_turn_grids=False)
    # the example's DEFAULT observation spec (examples/treasurehunt/main.py:24)
    run_case("rgb_s13_a2_11", 13, 11, 11, 2, 2, 0.05, 40, obs="rgb")
    transformer_case()
    observe_cases()
    cleanup_cases()
    iowa_cases()
    tag_cases()


if __name__ == "__main__":
    main()
