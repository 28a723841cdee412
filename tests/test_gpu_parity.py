"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Bit-exact everywhere (integer / byte / index work; observations are small integers held in f32).
Nothing here reads /root/reference: golden results were generated from it ahead of time
(oracle/gen_golden.py) and the oracle restatement regenerates the numpy draw tapes from seeds.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import restatement as R
import sb_testutil as util

pytestmark = pytest.mark.gpu

SMALL_CASES = util.golden_cases()


def _engine(spec, n, device, **kw):
    from sorrel_b200.engine import Engine

    return Engine(util.engine_spec(spec), n, device=device, **kw)


def _buffer(spec, n, capacity, device, n_frames=1, layout="agent"):
    from sorrel_b200.buffers import Buffer

    return Buffer(capacity, (spec.obs_size,), n_frames=n_frames, rows=n * spec.num_agents, device=device,
                  layout=layout, positions=(2,))


def _tapes_to_device(spec, env_tapes, device):
    """list (per env) of oracle tapes -> stacked device tensors [T, N, ...]."""
    us, cs = zip(*(util.pad_tapes(spec, t) for t in env_tapes))
    spawn_u = torch.from_numpy(np.stack(us, axis=1)).to(device)
    spawn_c = torch.from_numpy(np.stack(cs, axis=1)).to(device)
    actions = torch.from_numpy(np.stack([t["actions"] for t in env_tapes], axis=1)).to(device)
    placement = torch.from_numpy(np.stack([t["placement"] for t in env_tapes]).astype(np.int32)).to(device)
    return spawn_u, spawn_c, actions, placement


def _compare_buffers(buf, oracle_envs, A):
    for n, env in enumerate(oracle_envs):
        for a in range(A):
            row = buf.row(n * A + a)
            ob = env.buffers[a]
            assert row["states"].dtype == np.float32 and row["actions"].dtype == np.int64
            np.testing.assert_array_equal(row["states"], ob.states, err_msg=f"states env {n} agent {a}")
            np.testing.assert_array_equal(row["actions"], ob.actions)
            np.testing.assert_array_equal(row["rewards"], ob.rewards)
            np.testing.assert_array_equal(row["dones"], ob.dones)
            np.testing.assert_array_equal(row["positions"], ob.positions)


# ------------------------------------------------------------------------------------------------
# 1. replayed numpy draws: golden fixtures (generated from the unmodified reference)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", SMALL_CASES)
def test_golden_replay_per_turn(name, cuda_device):
    """One launch per turn; every turn's grid, and all buffer arrays, equal the reference's."""
    g = util.load_golden(name)
    if int(g["epochs"]) != 1:
        pytest.skip("multi-epoch fixture: covered by test_golden_ring")
    spec = util.spec_from_golden(g)
    T, A = int(g["T"]), int(g["A"])
    env, grids = util.run_oracle_episode(spec, R.NumpyDraws(int(g["seed"])), T)
    np.testing.assert_array_equal(grids, g["grids"])  # oracle == reference (also a CPU test)
    spawn_u, spawn_c, actions, placement = _tapes_to_device(spec, [env.dense_tapes()], cuda_device)
    eng = _engine(spec, 1, cuda_device)
    buf = _buffer(spec, 1, T, cuda_device)
    eng.reset(placement)
    np.testing.assert_array_equal(eng.agent_yx()[0], g["starts"][0])
    drew = torch.zeros((1, 1, eng.env_stride), dtype=torch.uint8, device=cuda_device)
    for t in range(T):
        eng.step(1, buffer=buf, actions=actions[t:t + 1].contiguous(), spawn_u=spawn_u[t:t + 1].contiguous(),
                 spawn_c=spawn_c[t:t + 1].contiguous(), drew_out=drew)
        np.testing.assert_array_equal(eng.grid_numpy()[0], g["grids"][t], err_msg=f"turn {t}")
        np.testing.assert_array_equal(drew.cpu().numpy()[0, 0] == 1, ~np.isnan(spawn_u[t, 0].cpu().numpy()))
    eng.check_errors()
    assert float(eng.total_reward[0]) == float(g["total_reward"][0])
    for a in range(A):
        row = buf.row(a)
        for key in ("states", "actions", "rewards", "dones", "positions"):
            assert row[key].dtype == g[key].dtype
            np.testing.assert_array_equal(row[key], g[key][a], err_msg=key)


@pytest.mark.parametrize("name", SMALL_CASES)
def test_golden_replay_fused(name, cuda_device):
    """All T turns in ONE launch (grid resident in shared memory): same final state and buffers."""
    g = util.load_golden(name)
    if int(g["epochs"]) != 1:
        pytest.skip("multi-epoch fixture: covered by test_golden_ring")
    spec = util.spec_from_golden(g)
    T, A = int(g["T"]), int(g["A"])
    env, _ = util.run_oracle_episode(spec, R.NumpyDraws(int(g["seed"])), T)
    spawn_u, spawn_c, actions, placement = _tapes_to_device(spec, [env.dense_tapes()], cuda_device)
    eng = _engine(spec, 1, cuda_device)
    buf = _buffer(spec, 1, T, cuda_device)
    eng.reset(placement)
    rew = torch.zeros((T, 1, A), dtype=torch.float32, device=cuda_device)
    eng.step(T, buffer=buf, actions=actions, spawn_u=spawn_u, spawn_c=spawn_c, reward_out=rew)
    eng.check_errors()
    np.testing.assert_array_equal(eng.grid_numpy()[0], g["grids"][-1])
    assert float(eng.total_reward[0]) == float(g["total_reward"][0])
    for a in range(A):
        row = buf.row(a)
        for key in ("states", "actions", "rewards", "dones", "positions"):
            np.testing.assert_array_equal(row[key], g[key][a], err_msg=key)
        np.testing.assert_array_equal(rew[:, 0, a].cpu().numpy(), g["rewards"][a])


def test_golden_large_map(cuda_device):
    """128x128, 64 agents, 9x9 window (config 4 shape) against the reference's own two turns."""
    g = util.load_golden("large_s4_a64_128")
    spec = util.spec_from_golden(g)
    T, A = int(g["T"]), int(g["A"])
    env, _ = util.run_oracle_episode(spec, R.NumpyDraws(int(g["seed"])), T)
    spawn_u, spawn_c, actions, placement = _tapes_to_device(spec, [env.dense_tapes()], cuda_device)
    eng = _engine(spec, 1, cuda_device)
    buf = _buffer(spec, 1, T, cuda_device)
    eng.reset(placement)
    eng.step(T, buffer=buf, actions=actions, spawn_u=spawn_u, spawn_c=spawn_c)
    eng.check_errors()
    np.testing.assert_array_equal(eng.grid_numpy()[0], g["grids"][-1])
    for a in range(A):
        row = buf.row(a)
        for key in ("states", "actions", "rewards", "dones", "positions"):
            np.testing.assert_array_equal(row[key], g[key][a], err_msg=f"{key} agent {a}")


@pytest.mark.parametrize("name", ["ring_s11_a2_11", "ring_s12_a1_21"])
def test_golden_ring(name, cuda_device):
    """Multi-epoch fixture: ring wrap, add_empty, current_state and sample against the reference."""
    g = util.load_golden(name)
    spec = util.spec_from_golden(g)
    T, A, E, F, cap = int(g["T"]), int(g["A"]), int(g["epochs"]), int(g["n_frames"]), int(g["capacity"])
    draws = R.NumpyDraws(int(g["seed"]))
    oenv = R.OracleEnv(spec, draws, capacity=cap, n_frames=F, record_tapes=True)
    eng = _engine(spec, 1, cuda_device)
    buf = _buffer(spec, 1, cap, cuda_device, n_frames=F)
    ci = 0
    for ep in range(E):
        oenv.tapes = {"spawn_u": [], "spawn_c": [], "actions": [], "placement": None}
        oenv.reset()
        for b in oenv.buffers:
            b.add_empty()
        for _ in range(T):
            oenv.take_turn()
        spawn_u, spawn_c, actions, placement = _tapes_to_device(spec, [oenv.dense_tapes()], cuda_device)
        eng.reset(placement)
        buf.add_empty()
        # split the episode into launches of uneven length to exercise slot wrap inside a launch
        t = 0
        for chunk in (1, 7, T - 8):
            sl = slice(t, t + chunk)
            eng.step(chunk, buffer=buf, actions=actions[sl].contiguous(), spawn_u=spawn_u[sl].contiguous(),
                     spawn_c=spawn_c[sl].contiguous())
            t += chunk
            cs = buf.current_state().cpu().numpy()  # [rows, F-1, obs]
            np.testing.assert_array_equal(cs, g["current_states"][ci + t - 1])
        ci += T
        assert float(eng.total_reward[0]) == float(g["total_reward"][ep])
    eng.check_errors()
    assert buf.idx == int(g["buf_idx"][0]) and buf.size == int(g["buf_size"][0])
    for a in range(A):
        row = buf.row(a)
        for key in ("states", "actions", "rewards", "dones", "positions"):
            np.testing.assert_array_equal(row[key], g[key][a], err_msg=key)
        # Buffer.sample gather on the reference's own drawn indices
        idx = g["sample_idx"][a]
        out = buf.gather(np.full(len(idx), a, dtype=np.int64), idx)
        for o, key in zip(out, ("sample_states", "sample_actions", "sample_rewards", "sample_next_states",
                                "sample_dones", "sample_valid")):
            ref = g[key][a]
            got = o.cpu().numpy()
            assert got.dtype == ref.dtype and got.shape == ref.shape, key
            np.testing.assert_array_equal(got, ref, err_msg=key)


# ------------------------------------------------------------------------------------------------
# 2. many instances at once, replayed numpy streams (seed = instance id)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("A,H,v,p,T,N,lpe", [(1, 21, 2, 0.005, 100, 96, None), (8, 21, 2, 0.005, 40, 40, None),
                                             (3, 9, 3, 0.15, 30, 33, None), (1, 21, 2, 0.005, 60, 31, 4),
                                             (8, 21, 2, 0.005, 25, 18, 8), (2, 21, 2, 0.01, 25, 12, 16),
                                             (1, 21, 2, 0.005, 40, 35, 2)])
def test_batched_replay(A, H, v, p, T, N, lpe, cuda_device):
    tuning = {"lpe": lpe} if lpe else {}
    spec = R.treasurehunt_spec(H, H, A, v, p)
    envs, tapes, finals = [], [], []
    for n in range(N):
        env, grids = util.run_oracle_episode(spec, R.NumpyDraws(1000 + n), T, env_id=n)
        envs.append(env)
        tapes.append(env.dense_tapes())
        finals.append(grids[-1])
    spawn_u, spawn_c, actions, placement = _tapes_to_device(spec, tapes, cuda_device)
    eng = _engine(spec, N, cuda_device, tuning=tuning)
    buf = _buffer(spec, N, T, cuda_device)
    eng.reset(placement)
    # two launches: T1 + T2 turns
    T1 = T // 3
    eng.step(T1, buffer=buf, actions=actions[:T1].contiguous(), spawn_u=spawn_u[:T1].contiguous(),
             spawn_c=spawn_c[:T1].contiguous())
    eng.step(T - T1, buffer=buf, actions=actions[T1:].contiguous(), spawn_u=spawn_u[T1:].contiguous(),
             spawn_c=spawn_c[T1:].contiguous())
    eng.check_errors()
    np.testing.assert_array_equal(eng.grid_numpy(), np.stack(finals))
    np.testing.assert_array_equal(eng.total_reward.cpu().numpy(), np.array([e.total_reward for e in envs]))
    np.testing.assert_array_equal(eng.agent_yx(), np.stack([e.pos for e in envs]))
    _compare_buffers(buf, envs, A)


# ------------------------------------------------------------------------------------------------
# 3. counter-based mode (what the benchmark runs): kernels vs oracle with the same Philox stream
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("A,H,W,v,p,T,N,lpe", [
    (1, 21, 21, 2, 0.005, 60, 70, None), (8, 21, 21, 2, 0.005, 25, 37, None), (2, 11, 14, 3, 0.3, 20, 19, None),
    (4, 12, 12, 1, 0.9, 6, 9, None), (1, 21, 21, 2, 0.005, 30, 20, 8), (1, 21, 21, 2, 0.005, 30, 20, 32),
    (3, 40, 40, 4, 0.01, 10, 6, None),
    # the group widths large batches run with (small test batches are widened to 32 lanes by pick_lpe):
    (1, 21, 21, 2, 0.005, 50, 41, 4),   # the headline variant step_kernel<4, SIMD, NP1, STAGED>
    (8, 21, 21, 2, 0.01, 20, 19, 4), (8, 21, 21, 2, 0.01, 20, 19, 8), (2, 30, 30, 3, 0.01, 12, 9, 16),
    (3, 11, 14, 3, 0.3, 15, 13, 4),     # generic sweep (p too large for the byte filter) at 4 lanes
    (1, 21, 21, 2, 0.005, 50, 70, 2), (3, 17, 19, 3, 0.02, 20, 37, 2)])  # two lanes per instance
def test_philox_mode(A, H, W, v, p, T, N, lpe, cuda_device):
    tuning = {"lpe": lpe} if lpe else {}
    seed, base = 0x1234ABCD5678, 5
    spec = R.treasurehunt_spec(H, W, A, v, p)
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed, tuning=tuning)
    buf = _buffer(spec, N, T, cuda_device)
    eng.reset()
    eng.step(T, buffer=buf)
    eng.check_errors()
    envs = []
    for n in range(N):
        env, _ = util.run_oracle_episode(spec, R.PhiloxDraws(seed), T, env_id=base + n)
        envs.append(env)
    np.testing.assert_array_equal(eng.grid_numpy(), np.stack([e.grid for e in envs]))
    np.testing.assert_array_equal(eng.total_reward.cpu().numpy(), np.array([e.total_reward for e in envs]))
    _compare_buffers(buf, envs, A)


@pytest.mark.parametrize("A,H,W,v,p,T,N,lpe", [
    (1, 21, 21, 2, 0.005, 60, 70, None), (8, 21, 21, 2, 0.005, 25, 37, None), (2, 11, 14, 3, 0.03, 30, 19, None),
    (1, 21, 21, 2, 0.005, 30, 20, 8), (1, 21, 21, 2, 0.02, 30, 12, 32), (3, 40, 40, 4, 0.002, 10, 6, None),
    (1, 21, 21, 2, 0.005, 40, 33, 4), (2, 30, 30, 3, 0.004, 12, 9, 16), (1, 21, 21, 2, 0.005, 40, 67, 2)])
def test_sparse_mode(A, H, W, v, p, T, N, lpe, cuda_device):
    """SB_RNG_SPARSE (Binomial count + distinct uniform cells): kernels == oracle with the same stream."""
    tuning = {"lpe": lpe} if lpe else {}
    seed, base = 0xFEEDC0DE77, 9
    spec = R.treasurehunt_spec(H, W, A, v, p)
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed, tuning=tuning)
    assert eng.sparse_available
    buf = _buffer(spec, N, T, cuda_device)
    eng.reset()
    eng.step(T // 2, buffer=buf, rng="sparse")
    eng.step(T - T // 2, buffer=buf, rng="sparse")
    eng.check_errors()
    envs = []
    for n in range(N):
        env, _ = util.run_oracle_episode(spec, R.SparseDraws(seed), T, env_id=base + n)
        envs.append(env)
    np.testing.assert_array_equal(eng.grid_numpy(), np.stack([e.grid for e in envs]))
    np.testing.assert_array_equal(eng.total_reward.cpu().numpy(), np.array([e.total_reward for e in envs]))
    _compare_buffers(buf, envs, A)


def test_sparse_and_dense_modes_have_the_same_statistics(cuda_device):
    """Both counter-based modes sample iid Bernoulli(p) per empty cell + uniform kind choice: compare
    spawn counts per kind over 2^16 instances x 1 turn from identical start states."""
    spec = R.treasurehunt_spec(21, 21, 1, 2, 0.005)
    N = 1 << 16
    counts = {}
    for mode in ("philox", "sparse"):
        eng = _engine(spec, N, cuda_device, seed=2025)
        eng.reset()
        eng.step(1, rng=mode)
        g = eng.grid[:, 0, : 441]
        counts[mode] = np.array([int((g == k).sum()) for k in (2, 3, 4)], dtype=np.float64)
    expect = N * 360 * spec.spawn_prob / 3  # 361 interior cells minus the agent, three kinds
    for mode, c in counts.items():
        assert np.all(np.abs(c - expect) < 5 * np.sqrt(expect)), (mode, c, expect)
    with pytest.raises(ValueError):  # dense spawning has no sparse table
        e2 = _engine(R.treasurehunt_spec(9, 9, 1, 2, 0.4), 4, cuda_device)
        assert not e2.sparse_available
        e2.reset()
        e2.step(1, rng="sparse")


def test_cuda_graph_replay_equals_eager(cuda_device):
    """StepGraph (device-resident tick/slot, sb_clock_advance) replays == eager launches, across resets,
    add_empty and ring wrap."""
    from sorrel_b200.graphs import StepGraph

    spec = R.treasurehunt_spec(21, 21, 2, 2, 0.01)
    N, T, cap = 48, 5, 16
    eager, eb = _engine(spec, N, cuda_device, seed=5), None
    eb = _buffer(spec, N, cap, cuda_device, n_frames=3)
    graph_e = _engine(spec, N, cuda_device, seed=5)
    gb = _buffer(spec, N, cap, cuda_device, n_frames=3)
    for e, b in ((eager, eb), (graph_e, gb)):
        e.reset()
        b.add_empty()
        for _ in range(4):  # fills the ring (size == capacity) and loads the kernels before capture
            e.step(T, buffer=b)
    b.gather(np.zeros(4, dtype=np.int64), np.zeros(4, dtype=np.int64))
    sg = StepGraph(graph_e, gb, T, sample_batch=8)
    for it in range(9):
        if it in (3, 6):  # episode boundary on both
            for e, b in ((eager, eb), (graph_e, gb)):
                e.reset()
                b.add_empty()
        eager.step(T, buffer=eb)
        out = sg.replay()
        assert out[0].shape == (8, 3 * spec.obs_size)
    eager.check_errors()
    graph_e.check_errors()
    assert graph_e.tick == eager.tick and gb.idx == eb.idx and gb.size == eb.size
    np.testing.assert_array_equal(graph_e.grid_numpy(), eager.grid_numpy())
    assert torch.equal(gb.states, eb.states) and torch.equal(gb.actions, eb.actions)
    assert torch.equal(gb.rewards, eb.rewards) and torch.equal(graph_e.total_reward, eager.total_reward)
    # the captured sample is a valid gather of the ring
    rows, starts = sg.rows.cpu().numpy(), sg.starts.cpu().numpy()
    ref = gb.gather(rows, starts)
    for o, r in zip(out, ref):
        assert torch.equal(o, r)


def test_shard_invariance(cuda_device):
    """N instances on one engine == the concatenation of two shards (rank offset env_id_base)."""
    spec = R.treasurehunt_spec(21, 21, 2, 2, 0.01)
    N, T = 64, 30
    full = _engine(spec, N, cuda_device, seed=7)
    fb = _buffer(spec, N, T, cuda_device)
    full.reset()
    full.step(T, buffer=fb)
    parts = []
    for r in range(2):
        e = _engine(spec, N // 2, cuda_device, seed=7, env_id_base=r * N // 2)
        b = _buffer(spec, N // 2, T, cuda_device)
        e.reset()
        e.step(T, buffer=b)
        parts.append((e, b))
    np.testing.assert_array_equal(full.grid_numpy(), np.concatenate([e.grid_numpy() for e, _ in parts]))
    assert torch.equal(fb.states, torch.cat([b.states for _, b in parts]))
    assert torch.equal(fb.rewards, torch.cat([b.rewards for _, b in parts], dim=1))


def test_split_phase_equals_fused(cuda_device):
    """Live-policy path: sweep+observe / act per agent (SB_MODE_SPLIT) == the fused launch."""
    spec = R.treasurehunt_spec(21, 21, 3, 2, 0.02)
    N, T, A = 50, 12, 3
    fused = _engine(spec, N, cuda_device, seed=3)
    fb = _buffer(spec, N, T, cuda_device)
    fused.reset()
    acts = torch.zeros((T, N, A), dtype=torch.int32, device=cuda_device)
    fused.step(T, buffer=fb, action_out=acts)
    split = _engine(spec, N, cuda_device, seed=3)
    sb = _buffer(spec, N, T, cuda_device)
    split.reset()
    for t in range(T):
        split.phase(sweep=True, obs_agent=0, buffer=sb, new_turn=True)
        for a in range(A):
            nxt = a + 1 if a + 1 < A else -1
            split.phase(act_agent=a, obs_agent=nxt, buffer=sb, actions=acts[t:t + 1].contiguous())
        sb.advance(1)
    split.check_errors()
    np.testing.assert_array_equal(split.grid_numpy(), fused.grid_numpy())
    assert torch.equal(sb.states, fb.states) and torch.equal(sb.rewards, fb.rewards)
    assert torch.equal(sb.actions, fb.actions) and torch.equal(sb.extra_data["positions"], fb.extra_data["positions"])
    assert torch.equal(split.total_reward, fused.total_reward)


# ------------------------------------------------------------------------------------------------
# 4. observation kernel alone
# ------------------------------------------------------------------------------------------------
def test_observe_golden_cases(cuda_device):
    """sb_observe against OneHotObservationSpec.observe outputs captured from the reference on
    random 1-3 layer worlds (every layer dynamic, sums over layers, heavy out-of-bounds fill)."""
    from sorrel_b200.engine import Engine, WorldSpec

    g = util.load_golden("observe_cases")
    app = np.zeros((4, 4))
    for k in range(1, 4):
        app[k, k] = 1.0
    for i in range(int(g["n"])):
        kinds = g[f"kinds_{i}"]
        v, y, x = (int(t) for t in g[f"meta_{i}"])
        H, W, L = kinds.shape
        spec = WorldSpec(height=H, width=W, layers=L, kind_names=["EmptyEntity", "Wall", "Gem", "Bone"],
                         kind_value=[0, 0, 0, 0], kind_passable=[True] * 4, kind_rule=[0] * 4, appearance=app,
                         template=kinds, spawnable=np.array([[0, 0]]), agent_kind=3, agent_layer=0,
                         default_kind=0, fill_kind=1, vision=v, num_agents=1, dynamic_layers=list(range(L)))
        eng = Engine(spec, 1, device=cuda_device)
        eng.grid.copy_(eng.d_template[None])
        eng.agent_pos[0, 0] = y * W + x
        obs = eng.observe().cpu().numpy()[0, 0]
        np.testing.assert_array_equal(obs.astype(np.float64), g[f"obs_{i}"], err_msg=f"case {i}")


def test_observe_rgb_golden_cases(cuda_device):
    """sb_observe with the RGB transform against RGBObservationSpec.observe outputs from the reference
    (float32 of the reference's float64, i.e. what Buffer.states would hold)."""
    from sorrel_b200.engine import Engine, WorldSpec

    g = util.load_golden("observe_rgb_cases")
    for i in range(int(g["n"])):
        kinds = g[f"kinds_{i}"]
        v, y, x = (int(t) for t in g[f"meta_{i}"])
        H, W, L = kinds.shape
        spec = WorldSpec(height=H, width=W, layers=L, kind_names=["EmptyEntity", "Wall", "Gem", "Bone"],
                         kind_value=[0, 0, 0, 0], kind_passable=[True] * 4, kind_rule=[0] * 4,
                         appearance=g[f"table_{i}"], template=kinds, spawnable=np.array([[0, 0]]), agent_kind=3,
                         agent_layer=0, default_kind=0, fill_kind=1, vision=v, num_agents=1,
                         dynamic_layers=list(range(L)), obs_transform=1)
        eng = Engine(spec, 1, device=cuda_device)
        eng.grid.copy_(eng.d_template[None])
        eng.agent_pos[0, 0] = y * W + x
        obs = eng.observe().cpu().numpy()[0, 0]
        np.testing.assert_array_equal(obs, g[f"obs_{i}"].astype(np.float32), err_msg=f"case {i}")


@pytest.mark.parametrize("general", [False, True])
def test_observe_random(general, cuda_device):
    """Fast one-hot path and the general table path (non one-hot appearance: f64 sum over layers,
    static layer visible) against the oracle's closed form."""
    rng = np.random.RandomState(5)
    spec = R.treasurehunt_spec(13, 17, 5, 3, 0.0)
    if general:
        spec.appearance = rng.randint(0, 256, size=spec.appearance.shape).astype(np.float64)
    N = 24
    eng = _engine(spec, N, cuda_device)
    eng.reset()
    # scatter random items on the dynamic layer (not over walls / agents)
    g = eng.grid.cpu().numpy()
    mask = (g[:, 0, : 13 * 17] == 0) & (rng.rand(N, 13 * 17) < 0.4)
    g[:, 0, : 13 * 17][mask] = rng.randint(2, 5, size=int(mask.sum()))
    eng.grid.copy_(torch.from_numpy(g))
    obs = eng.observe().cpu().numpy()
    full = eng.grid_numpy()
    yx = eng.agent_yx()
    for n in range(N):
        for a in range(5):
            ref = R.observe(spec, full[n], int(yx[n, a, 0]), int(yx[n, a, 1])).astype(np.float32)
            np.testing.assert_array_equal(obs[n, a], ref)


@pytest.mark.parametrize("H,W,v,A,N", [(9, 9, 4, 3, 5), (16, 23, 3, 6, 9), (21, 21, 5, 4, 37), (33, 18, 6, 7, 11),
                                       (40, 40, 7, 9, 6), (21, 21, 2, 8, 40), (128, 128, 4, 64, 3),
                                       (21, 21, 2, 3, 5), (12, 12, 1, 3, 7), (15, 14, 2, 1, 9),
                                       (40, 44, 9, 5, 7), (45, 41, 10, 3, 4), (36, 36, 8, 8, 5)])
def test_standalone_observe_all_window_widths(H, W, v, A, N, cuda_device):
    """`sb_observe` (the bulk-async kernel for small rows -- whole and short last tiles: odd pair counts --, the
    software-pipelined row-wise kernel for 7..16-wide windows, row segments on 32 lanes for wider ones) on worlds
    without walls, with agents parked on corners, edges and the interior of the FIRST and LAST instance as well (the
    row gather may touch bytes outside an instance's planes: guarded at the ends of the allocation), several launches
    back to back (the staging rows must come back zeroed)."""
    rng = np.random.RandomState(H * 100 + W)
    spec = R.treasurehunt_spec(H, W, A, v, 0.0)
    eng = _engine(spec, N, cuda_device)
    eng.reset()
    g = eng.grid.cpu().numpy()
    g[:, 0, : H * W] = rng.choice([0, 0, 1, 2, 3, 4], size=(N, H * W)).astype(np.uint8)  # no border walls
    cells = np.stack([rng.permutation(H * W)[:A] for _ in range(N)]).astype(np.int32)
    corners = [0, W - 1, (H - 1) * W, H * W - 1, W // 2, (H // 2) * W, (H // 2) * W + W - 1, (H - 1) * W + W // 2]
    for n in (0, N - 1):
        for a in range(A):
            cells[n, a] = corners[(a + n) % len(corners)] if a < len(corners) else cells[n, a]
        cells[n] = np.where(np.bincount(cells[n], minlength=H * W)[cells[n]] > 1, rng.permutation(H * W)[:A], cells[n])
    for n in range(N):
        g[n, 0, cells[n]] = 5
    eng.grid.copy_(torch.from_numpy(g))
    eng.agent_pos.copy_(torch.from_numpy(cells))
    full = eng.grid_numpy()
    for _ in range(3):
        obs = eng.observe().cpu().numpy()
        for n in list(range(min(N, 3))) + [N - 1]:
            for a in range(A):
                y, x = divmod(int(cells[n, a]), W)
                np.testing.assert_array_equal(obs[n, a], R.observe(spec, full[n], y, x).astype(np.float32),
                                              err_msg=f"instance {n} agent {a} at ({y}, {x})")


@pytest.mark.parametrize("v,A,N", [(1, 3, 5), (2, 5, 3), (1, 1, 2), (2, 2, 34)])
def test_standalone_observe_odd_row_size(v, A, N, cuda_device):
    """Seven observation channels: rows of 7 * S * S floats (odd), so the bulk-async kernel ships FOUR rows per tile
    (the smallest multiple of 16 bytes); pair counts that leave a short last tile of one, two or three rows."""
    import dataclasses

    rng = np.random.RandomState(17 * v + A)
    H, W = 11, 13
    base = R.treasurehunt_spec(H, W, A, v, 0.0)
    app = np.zeros((7, 7), dtype=np.float64)
    for k in range(1, 6):
        app[k, k + 1] = 1.0  # channel 0 and 1 stay empty, kind k shows on channel k + 1
    spec = dataclasses.replace(base, appearance=app)
    assert spec.obs_size % 2 == 1
    eng = _engine(spec, N, cuda_device)
    eng.reset()
    g = eng.grid.cpu().numpy()
    g[:, 0, : H * W] = rng.choice([0, 0, 1, 2, 3, 4], size=(N, H * W)).astype(np.uint8)
    cells = np.stack([rng.permutation(H * W)[:A] for _ in range(N)]).astype(np.int32)
    for n in range(N):
        g[n, 0, cells[n]] = 5
    eng.grid.copy_(torch.from_numpy(g))
    eng.agent_pos.copy_(torch.from_numpy(cells))
    full = eng.grid_numpy()
    for _ in range(3):  # several launches: the tiles must come back zeroed
        obs = eng.observe().cpu().numpy()
        for n in range(N):
            for a in range(A):
                y, x = divmod(int(cells[n, a]), W)
                np.testing.assert_array_equal(obs[n, a], R.observe(spec, full[n], y, x).astype(np.float32),
                                              err_msg=f"instance {n} agent {a} at ({y}, {x})")


# ------------------------------------------------------------------------------------------------
# 5. replay ring primitives, reward statistics, reset
# ------------------------------------------------------------------------------------------------
def test_buffer_add_matches_reference_semantics(cuda_device):
    from sorrel_b200.buffers import Buffer

    rows, cap, obs, F = 6, 10, 14, 4
    rng = np.random.RandomState(0)
    buf = Buffer(cap, (obs,), n_frames=F, rows=rows, device=cuda_device, positions=(2,))
    refs = [R.OracleBuffer(cap, obs, F) for _ in range(rows)]
    for step in range(27):
        if step % 9 == 0:
            buf.add_empty()
            for r in refs:
                r.add_empty()
        o = rng.rand(rows, obs).astype(np.float32)
        a = rng.randint(0, 4, rows)
        rw = rng.randn(rows).astype(np.float32)
        dn = (rng.rand(rows) < 0.3).astype(np.float32)
        pos = rng.randint(0, 21, (rows, 2))
        buf.add(torch.from_numpy(o).to(cuda_device), a, rw, dn, positions=pos)
        for r in range(rows):
            refs[r].add(o[r], a[r], rw[r], dn[r], tuple(pos[r]))
        cs = buf.current_state().cpu().numpy()
        for r in range(rows):
            np.testing.assert_array_equal(cs[r], refs[r].current_state())
    assert buf.idx == refs[0].idx and len(buf) == refs[0].size
    for r in range(rows):
        row = buf.row(r)
        np.testing.assert_array_equal(row["states"], refs[r].states)
        np.testing.assert_array_equal(row["positions"], refs[r].positions)
    # gather with dones present -> `valid` path
    starts = np.array([0, 1, 2, 3, 4], dtype=np.int64)
    for r in range(rows):
        got = buf.gather(np.full(5, r), starts)
        ref = refs[r].gather(starts)
        for o_, r_ in zip(got, ref):
            np.testing.assert_array_equal(o_.cpu().numpy(), r_)
    with pytest.raises(ValueError):  # np.random.choice(replace=False) over-sampling, buffers.py:109
        buf.sample(10_000)


def test_transformer_buffer_golden(cuda_device):
    """TransformerBuffer.combine + sample_transformer outputs captured from the reference."""
    from sorrel_b200.buffers import Buffer

    g = util.load_golden("transformer_s21_a2_11")
    A, cap, F, T = 2, int(g["capacity"]), int(g["n_frames"]), int(g["T"])
    spec = R.treasurehunt_spec(11, 11, A, 2, 0.05)
    env, _ = util.run_oracle_episode(spec, R.NumpyDraws(21), T)
    spawn_u, spawn_c, actions, placement = _tapes_to_device(spec, [env.dense_tapes()], cuda_device)
    eng = _engine(spec, 1, cuda_device)
    buf = Buffer(cap, (spec.obs_size,), n_frames=F, rows=A, device=cuda_device, positions=(2,))
    eng.reset(placement)
    eng.step(T, buffer=buf, actions=actions, spawn_u=spawn_u, spawn_c=spawn_c)
    comb = buf.combined()
    for key in ("states", "actions", "rewards", "dones", "positions"):
        ref = g["comb_" + key]
        assert comb[key].dtype == ref.dtype, key
        np.testing.assert_array_equal(comb[key], ref, err_msg=key)
    assert len(comb["actions"]) == int(g["comb_size"])
    # The reference's combine() allocates agent_ids but never fills them (buffers.py:298-313 copies only the
    # keys the source buffers hold), so its ids are all 0; the engine returns the ids the docstring promises.
    assert not g["comb_agent_ids"].any() and not g["t_agent_ids"].any()
    np.testing.assert_array_equal(comb["agent_ids"], np.repeat(np.arange(A), T))
    # sample_transformer on the reference's drawn indices of the COMBINED buffer -> (row, start) pairs
    idx = g["sample_idx"]
    rows, starts = idx // T, idx % T
    keep = starts + F + 1 <= T  # windows that stay inside one agent's segment map 1:1 onto a ring row
    out = buf.gather_transformer(rows[keep], starts[keep])
    for o, key in zip(out[:6], ("t_states", "t_actions", "t_next_actions", "t_next_states", "t_dones", "t_valid")):
        got, ref = o.cpu().numpy(), g[key][keep]
        assert got.dtype == ref.dtype and got.shape == ref.shape, key
        np.testing.assert_array_equal(got, ref, err_msg=key)
    np.testing.assert_array_equal(out[6].cpu().numpy(), rows[keep])
    assert int(keep.sum()) >= 12


def test_reward_stats(cuda_device):
    import ctypes as C

    from sorrel_b200._abi import check, lib

    L = lib()
    for n in (1, 31, 4097, 300_001):
        x = torch.from_numpy(np.random.RandomState(n).randint(-50, 80, n).astype(np.float64)).to(cuda_device)
        out = torch.zeros(5, dtype=torch.float64, device=cuda_device)
        scratch = torch.zeros(L.sb_stats_scratch_bytes(), dtype=torch.uint8, device=cuda_device)
        for _ in range(2):  # the ticket must re-arm itself
            check(L.sb_reward_stats(x.data_ptr(), n, out.data_ptr(), scratch.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream))
            xs = x.cpu().numpy()
            np.testing.assert_array_equal(out.cpu().numpy(),
                                          np.array([xs.sum(), (xs * xs).sum(), xs.min(), xs.max(), n]))
    # the partials of several ranks (what the all-gather delivers) folded by sb_stats_combine
    from sorrel_b200.stats import combine_partials

    parts = torch.tensor([[10.0, 60.0, -3.0, 7.0, 4.0], [-2.5, 9.25, -2.5, 0.0, 2.0], [0.0, 0.0, 0.0, 0.0, 1.0]],
                         dtype=torch.float64, device=cuda_device)
    out = torch.zeros(5, dtype=torch.float64, device=cuda_device)
    check(L.sb_stats_combine(parts.data_ptr(), 3, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    assert torch.equal(out, combine_partials(parts)) and out.tolist() == [7.5, 69.25, -3.0, 7.0, 7.0]


def test_reset_philox_placement(cuda_device):
    spec = R.treasurehunt_spec(9, 9, 12, 2, 0.0)
    N, seed = 40, 99
    eng = _engine(spec, N, cuda_device, seed=seed, env_id_base=3)
    for episode_tick in (0, 17):
        eng.tick = episode_tick
        eng.reset()
        yx = eng.agent_yx()
        for n in range(N):
            picks = R.PhiloxDraws(seed).placement(len(spec.spawnable), 12, 3 + n, episode_tick + 1)
            np.testing.assert_array_equal(yx[n], spec.spawnable[picks][:, :2])
        g = eng.grid_numpy()
        assert ((g[:, :, :, 1] == spec.agent_kind).sum(axis=(1, 2)) == 12).all()
        assert (eng.total_reward == 0).all()


def test_errors_and_validation(cuda_device):
    spec = R.treasurehunt_spec(7, 7, 1, 2, 0.0)
    spec.vision = 4  # > (H-1)//2: the reference silently returns a wrong-shaped crop
    with pytest.raises(ValueError):
        _engine(spec, 1, cuda_device)
    # an unwalled world lets the agent step off the grid -> IndexError like the reference
    spec = R.treasurehunt_spec(7, 7, 1, 2, 0.0)
    spec.template[:, :, 1] = 0
    spec.spawnable = np.array([[0, 0, 1]])
    eng = _engine(spec, 1, cuda_device)
    eng.reset(torch.zeros((1, 1), dtype=torch.int32))
    acts = torch.zeros((1, 1, 1), dtype=torch.int32, device=cuda_device)  # "up" from row 0
    eng.step(1, actions=acts)
    with pytest.raises(IndexError):
        eng.check_errors()
    acts[:] = 7  # not an action
    eng.step(1, actions=acts)
    with pytest.raises(KeyError):
        eng.check_errors()


# ------------------------------------------------------------------------------------------------
# 6. full-size property tests (sizes the oracle cannot replay in seconds)
# ------------------------------------------------------------------------------------------------
def test_full_size_invariants(cuda_device):
    """Config-2/5 shape at 262144 instances: size-independent invariants of the rule set."""
    spec = R.treasurehunt_spec(21, 21, 1, 2, 0.005)
    N, T, cap = 1 << 18, 50, 8
    eng = _engine(spec, N, cuda_device, seed=11)
    buf = _buffer(spec, N, cap, cuda_device)
    eng.reset()
    rew = torch.zeros((T, N, 1), dtype=torch.float32, device=cuda_device)
    eng.step(T, buffer=buf, reward_out=rew)
    eng.check_errors()
    g = eng.grid[:, 0, : 21 * 21].view(N, 21, 21)
    border = torch.ones((21, 21), dtype=torch.bool, device=cuda_device)
    border[1:-1, 1:-1] = False
    assert (g[:, border] == 1).all() and (g[:, ~border] != 1).all()  # walls never move or appear
    assert ((g == 5).sum(dim=(1, 2)) == 1).all()  # exactly one agent per instance
    pos = eng.agent_pos[:, 0].long()
    assert (g.reshape(N, -1).gather(1, pos[:, None]) == 5).all()  # ... at its recorded position
    assert torch.equal(eng.total_reward, rew.double().sum(dim=(0, 2)))  # reward bookkeeping
    # spawn rate: items only appear through Bernoulli(p) on empty cells
    items = ((g >= 2) & (g <= 4)).sum().item()
    expect = N * T * 0.005 * 330  # ~330 empty cells on average over the episode start
    assert 0.8 * expect < items < 1.2 * expect
    # every observation row is a valid one-hot crop: 25 cells, each channel-sum <= 1, centre = agent
    st = buf.states.view(N, cap, 6, 5, 5)
    assert (st[:, :, 0] == 0).all() and (st.sum(dim=2) <= 1).all()
    assert (st[:, :, 5, 2, 2] == 1).all()


# ------------------------------------------------------------------------------------------------
# agent-parallel path of the warp-per-instance variant (large maps, many agents)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("A,H,W,v,p,T,N,staged", [
    (16, 64, 64, 3, 0.01, 30, 6, None),   # staged planes, ~10 % of the 4-agent groups conflict -> both paths
    (24, 40, 40, 2, 0.02, 25, 5, None),   # crowded: many fall-backs to the sequential path
    (10, 48, 48, 4, 0.01, 20, 4, None),   # A % 4 != 0: sequential tail
    (64, 128, 128, 4, 0.005, 12, 3, None),  # BASELINE config 4 shape (planes used in place in global memory)
    (16, 64, 64, 3, 0.01, 20, 4, 0),      # the same world with unstaged planes
    (40, 16, 16, 0, 0.02, 25, 6, None),   # v = 0, crowded: agents 2 apart may target the same cell (radius max(v+1, 2))
    (40, 16, 16, 0, 0.02, 25, 6, 0),
])
def test_agent_parallel_path_matches_oracle(A, H, W, v, p, T, N, staged, cuda_device):
    """Groups of 4 consecutive agents that are pairwise more than v+1 apart are observed and moved at once
    (8 lanes each); anything else takes the sequential path.  Both must reproduce the strictly sequential
    oracle (environment.py:92) bit for bit: grids, positions, every buffer array, total reward."""
    from oracle import coracle

    tuning = {"lpe": 32}
    if staged is not None:  # 0: planes used in place in global memory (one warp per instance)
        tuning["staged"] = 1 if staged else -1
        if not staged:
            tuning["lpe"] = 32
    seed, base = 0xC0FFEE1234, 17
    spec = R.treasurehunt_spec(H, W, A, v, p)
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed, tuning=tuning)
    buf = _buffer(spec, N, T, cuda_device)
    obs_out = torch.zeros((T - T // 2, N, A, spec.obs_size), dtype=torch.float32, device=cuda_device)
    eng.reset()
    eng.step(T // 2, buffer=buf)
    eng.step(T - T // 2, buffer=buf, obs_out=obs_out)  # second destination of the rows as well
    eng.check_errors()
    co = coracle.COracle(spec, N, capacity=T, rng=coracle.RNG_PHILOX, seed=seed, env_id_base=base, threads=4)
    co.reset()
    co.step(T)
    np.testing.assert_array_equal(eng.grid_numpy(), co.grid)
    np.testing.assert_array_equal(eng.agent_pos.cpu().numpy(), co.pos)
    np.testing.assert_array_equal(eng.total_reward.cpu().numpy(), co.total_reward)
    ring = co.ring()
    np.testing.assert_array_equal(buf.states.cpu().numpy(), ring["states"])
    np.testing.assert_array_equal(buf.actions.cpu().numpy(), ring["actions"])
    np.testing.assert_array_equal(buf.rewards.cpu().numpy(), ring["rewards"])
    np.testing.assert_array_equal(buf.extra_data["positions"].cpu().numpy(), ring["positions"])
    got = obs_out.cpu().numpy().reshape(T - T // 2, N * A, -1)
    np.testing.assert_array_equal(got, np.moveaxis(ring["states"][:, T // 2:], 0, 1))


@pytest.mark.parametrize("A,H,W,v,p,T,N", [
    (64, 128, 128, 4, 0.005, 10, 5),   # BASELINE config 4 shape
    (16, 64, 64, 3, 0.01, 30, 7),      # exactly one slot per agent
    (24, 40, 40, 2, 0.02, 25, 6),      # crowded: long dependency chains, many levels
    (37, 30, 33, 1, 0.02, 20, 5),      # A neither a multiple of 16 nor of 32, 3x3 windows
    (40, 16, 16, 0, 0.02, 25, 6),      # v = 0, very crowded
    (64, 24, 24, 5, 0.01, 12, 4),      # 64 agents within sight of each other: nearly sequential
])
def test_block_per_instance_kernel_matches_oracle(A, H, W, v, p, T, N, cuda_device):
    """The block-per-instance kernel (plane staged in shared memory, 16 eight-lane slots, agents released level by level
    as their conflicting predecessors finish) must reproduce the strictly sequential oracle (environment.py:92) bit for
    bit in every array: grids, positions, rewards, total reward, every ring array; tape-driven actions as well."""
    from oracle import coracle

    seed, base = 0xB10C + A, 23
    spec = R.treasurehunt_spec(H, W, A, v, p)
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed, tuning={"lpe": 128})
    assert eng.step_plan(T)["lanes_per_instance"] == 128
    buf = _buffer(spec, N, T, cuda_device)
    rew = torch.zeros((T - T // 2, N, A), dtype=torch.float32, device=cuda_device)
    eng.reset()
    eng.step(T // 2, buffer=buf)
    eng.step(T - T // 2, buffer=buf, reward_out=rew)
    eng.check_errors()
    co = coracle.COracle(spec, N, capacity=T, rng=coracle.RNG_PHILOX, seed=seed, env_id_base=base, threads=4,
                         record=True)
    co.reset()
    co.step(T)
    np.testing.assert_array_equal(eng.grid_numpy(), co.grid)
    np.testing.assert_array_equal(eng.agent_pos.cpu().numpy(), co.pos)
    np.testing.assert_array_equal(eng.total_reward.cpu().numpy(), co.total_reward)
    ring = co.ring()
    np.testing.assert_array_equal(buf.states.cpu().numpy(), ring["states"])
    np.testing.assert_array_equal(buf.actions.cpu().numpy(), ring["actions"])
    np.testing.assert_array_equal(buf.rewards.cpu().numpy(), ring["rewards"])
    np.testing.assert_array_equal(buf.extra_data["positions"].cpu().numpy(), ring["positions"])
    np.testing.assert_array_equal(rew.cpu().numpy().reshape(T - T // 2, N * A), ring["rewards"][T // 2:])
    # the same turns driven by an action tape (policy TAPE) from the same start state
    eng2 = _engine(spec, N, cuda_device, env_id_base=base, seed=seed, tuning={"lpe": 128})
    buf2 = _buffer(spec, N, T, cuda_device, layout="slot")
    eng2.reset()
    eng2.step(T, buffer=buf2, actions=torch.from_numpy(co.tapes()["actions"].copy()).to(cuda_device))
    eng2.check_errors()
    np.testing.assert_array_equal(eng2.grid_numpy(), co.grid)
    np.testing.assert_array_equal(buf2.states.cpu().numpy(), ring["states"])


# ------------------------------------------------------------------------------------------------
# engine option: toroidal worlds (crops and moves wrap around the map edges)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lpe,staged,obs,H,W,v", [
    (None, None, "onehot", 12, 15, 3), (4, None, "onehot", 12, 15, 3), (8, None, "onehot", 12, 15, 3),
    (32, None, "onehot", 12, 15, 3), (32, 0, "onehot", 12, 15, 3),   # planes used in place in global memory
    (4, None, "onehot", 9, 11, 4),                                     # v = (H-1)/2: the window spans the whole column
    (None, None, "rgb", 12, 15, 3), (4, None, "rgb", 12, 15, 3)])      # the general (table-sum) path
def test_toroidal_option_matches_oracle(lpe, staged, obs, H, W, v, cuda_device):
    """NOT a reference behaviour (the reference pads crops with the fill kind and indexes the map with plain numpy
    indices): an explicit toroidal mode.  Windows that reach over an edge show the opposite side, an
    agent that steps over an edge re-enters on the opposite side; compared with the oracle restatement's own
    definition of the same option.  One-hot worlds take the FAST observation path (bf16 staging row, 128-bit stores,
    folded window indices) -- the library reports it -- RGB worlds the general one; the standalone observation kernel
    agrees with the fused step."""
    tuning = {"lpe": lpe} if lpe else {}
    if staged is not None:
        tuning["staged"] = 1 if staged else -1
    spec = R.torus_spec(H, W, 3, v, 0.02, obs=obs)
    N, T, seed = 23, 40, 8
    eng = _engine(spec, N, cuda_device, seed=seed, tuning=tuning)
    assert eng.step_plan(T)["fast_onehot"] == (obs == "onehot")
    buf = _buffer(spec, N, T, cuda_device)
    eng.reset()
    eng.step(T, buffer=buf)
    eng.check_errors()  # nothing is "off the grid" on a torus
    crossed = 0
    for n in range(0, N, 4):
        o, _ = util.run_oracle_episode(spec, R.PhiloxDraws(seed), T, env_id=n)
        np.testing.assert_array_equal(eng.grid_numpy()[n], o.grid)
        np.testing.assert_array_equal(eng.agent_yx()[n], o.pos)
        for a in range(3):
            np.testing.assert_array_equal(buf.row(n * 3 + a)["states"], o.buffers[a].states)
            np.testing.assert_array_equal(buf.row(n * 3 + a)["rewards"], o.buffers[a].rewards)
            p = o.buffers[a].positions
            crossed += int((np.abs(np.diff(p[:, 0])) > 1).sum() + (np.abs(np.diff(p[:, 1])) > 1).sum())
    assert crossed > 0  # some agent did cross a seam
    obs_now = eng.observe().reshape(N, 3, -1).cpu().numpy()
    full = eng.grid_numpy()
    for n in (0, 7):
        for a in range(3):
            y, x = eng.agent_yx()[n, a]
            np.testing.assert_array_equal(obs_now[n, a],
                                          R.observe(spec, full[n], y, x).reshape(-1).astype(np.float32))


@pytest.mark.parametrize("lpe", [None, 4, 8])
def test_general_observation_path_batch(lpe, cuda_device):
    """RGBObservationSpec (general path: table sums in f64, clip / 255) on many instances and several agents at the
    group widths large batches use -- the path where every lane reads the grid itself right before lane 0 moves the
    agent (a missing barrier there was found with the toroidal test)."""
    tuning = {"lpe": lpe} if lpe else {}
    spec = R.treasurehunt_spec(13, 16, 5, 3, 0.04, obs="rgb")
    N, T, seed = 48, 25, 77
    eng = _engine(spec, N, cuda_device, seed=seed, tuning=tuning)
    buf = _buffer(spec, N, T, cuda_device)
    eng.reset()
    eng.step(T, buffer=buf)
    eng.check_errors()
    for n in range(0, N, 5):
        o, _ = util.run_oracle_episode(spec, R.PhiloxDraws(seed), T, env_id=n)
        np.testing.assert_array_equal(eng.grid_numpy()[n], o.grid)
        for a in range(5):
            np.testing.assert_array_equal(buf.row(n * 5 + a)["states"], o.buffers[a].states)


# ------------------------------------------------------------------------------------------------
# randomized differential test against the C oracle
# ------------------------------------------------------------------------------------------------
def _fuzz_cases(n=36, seed=20261018):
    rng = np.random.RandomState(seed)
    cases = []
    for i in range(n):
        H, W = int(rng.randint(5, 41)), int(rng.randint(5, 41))
        v = int(rng.randint(1, min(4, (min(H, W) - 1) // 2) + 1))
        A = int(rng.randint(1, min(12, (H - 2) * (W - 2) // 2) + 1))
        p = float(rng.choice([0.001, 0.005, 0.02, 0.1, 0.3, 0.6]))
        lpe = [None, 4, 8, 16, 32][int(rng.randint(0, 5))]
        staged = None if rng.rand() < 0.7 else 0
        mode = "sparse" if (rng.rand() < 0.3 and p <= 0.02) else "philox"
        obs = "rgb" if rng.rand() < 0.2 else "onehot"
        cases.append((i, H, W, A, v, p, lpe, staged, mode, obs, int(rng.randint(3, 25)), int(rng.randint(2, 30))))
    return cases


@pytest.mark.parametrize("case", _fuzz_cases(), ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}a{c[3]}v{c[4]}")
def test_fuzz_against_c_oracle(case, cuda_device):
    """Random map sizes, agent counts, vision radii, spawn probabilities, group widths, staged / unstaged planes,
    dense / sparse sampling, one-hot / RGB views: the fused step in two launches must equal the C restatement on the
    same counter-based stream in every array."""
    from oracle import coracle

    _, H, W, A, v, p, lpe, staged, mode, obs, T, N = case
    tuning = {"lpe": lpe} if lpe else {}
    if staged is not None:  # 0: planes used in place in global memory (one warp per instance)
        tuning["staged"] = 1 if staged else -1
        if not staged:
            tuning["lpe"] = 32
    seed, base = 1000 + case[0], 3 * case[0]
    spec = R.treasurehunt_spec(H, W, A, v, p, obs=obs)
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed, tuning=tuning)
    if mode == "sparse" and not eng.sparse_available:
        mode = "philox"
    buf = _buffer(spec, N, T, cuda_device)
    eng.reset()
    eng.step(T // 2, buffer=buf, rng=mode)
    eng.step(T - T // 2, buffer=buf, rng=mode)
    eng.check_errors()
    co = coracle.COracle(spec, N, capacity=T, seed=seed, env_id_base=base, threads=4,
                         rng=coracle.RNG_SPARSE if mode == "sparse" else coracle.RNG_PHILOX)
    co.reset()
    co.step(T)
    np.testing.assert_array_equal(eng.grid_numpy(), co.grid)
    np.testing.assert_array_equal(eng.agent_pos.cpu().numpy(), co.pos)
    np.testing.assert_array_equal(eng.total_reward.cpu().numpy(), co.total_reward)
    ring = co.ring()
    np.testing.assert_array_equal(buf.states.cpu().numpy(), ring["states"])
    np.testing.assert_array_equal(buf.actions.cpu().numpy(), ring["actions"])
    np.testing.assert_array_equal(buf.rewards.cpu().numpy(), ring["rewards"])
    np.testing.assert_array_equal(buf.extra_data["positions"].cpu().numpy(), ring["positions"])


def _many_agent_fuzz_cases(n=14, seed=20261019):
    rng = np.random.RandomState(seed)
    cases = []
    for i in range(n):
        H, W = int(rng.randint(12, 73)), int(rng.randint(12, 73))
        v = int(rng.randint(0, min(5, (min(H, W) - 1) // 2) + 1))
        A = int(rng.randint(16, min(64, (H - 2) * (W - 2) // 3) + 1))
        p = float(rng.choice([0.002, 0.01, 0.05]))
        variant = ["block", "block", "warp_par", "warp_inplace"][int(rng.randint(0, 4))]
        cases.append((i, H, W, A, v, p, variant, int(rng.randint(4, 16)), int(rng.randint(2, 8))))
    return cases


@pytest.mark.parametrize("case", _many_agent_fuzz_cases(), ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}a{c[3]}v{c[4]}-{c[6]}")
def test_fuzz_many_agents_against_c_oracle(case, cuda_device):
    """Random many-agent worlds (16..64 agents, crowded to sparse, vision 0..5) through the three tilings that run agents
    concurrently -- block per instance (dependency levels, 16 slots), warp per instance with four mutually distant agents
    at once (planes staged, or used in place in global memory) -- against the strictly sequential C oracle, every
    array, both ring layouts."""
    from oracle import coracle

    i, H, W, A, v, p, variant, T, N = case
    tuning = {"block": {"lpe": 128}, "warp_par": {"lpe": 32}, "warp_inplace": {"lpe": 32, "staged": -1}}[variant]
    if variant == "warp_par" and H * W > 4096:
        tuning = {"lpe": 32, "staged": -1}  # (planes above 4 KB are not staged by the per-warp tilings)
    seed, base = 7000 + i, 11 * i
    spec = R.treasurehunt_spec(H, W, A, v, p)
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed, tuning=tuning)
    buf = _buffer(spec, N, T, cuda_device, layout="slot" if i % 2 else "agent")
    eng.reset()
    eng.step(T // 2, buffer=buf)
    eng.step(T - T // 2, buffer=buf)
    eng.check_errors()
    co = coracle.COracle(spec, N, capacity=T, rng=coracle.RNG_PHILOX, seed=seed, env_id_base=base, threads=4)
    co.reset()
    co.step(T)
    np.testing.assert_array_equal(eng.grid_numpy(), co.grid)
    np.testing.assert_array_equal(eng.agent_pos.cpu().numpy(), co.pos)
    np.testing.assert_array_equal(eng.total_reward.cpu().numpy(), co.total_reward)
    ring = co.ring()
    np.testing.assert_array_equal(buf.states.cpu().numpy(), ring["states"])
    np.testing.assert_array_equal(buf.actions.cpu().numpy(), ring["actions"])
    np.testing.assert_array_equal(buf.rewards.cpu().numpy(), ring["rewards"])
    np.testing.assert_array_equal(buf.extra_data["positions"].cpu().numpy(), ring["positions"])
