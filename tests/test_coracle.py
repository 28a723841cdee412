"""CPU: the C restatement (oracle/oracle.c) against numpy's legacy stream, the golden fixtures and
the Python restatement."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

import sb_testutil as util
from oracle import coracle
from oracle import restatement as R


def test_numpy_legacy_stream_reimplementation():
    """MT19937 seeding, random_sample, randint (masked rejection) and permutation vs numpy itself."""
    L = coracle.lib()
    for seed in list(range(40)) + [2**31 - 1, 2**32 - 1, 123456789]:
        for bound, perm_n in ((3, 361), (4, 15876), (7, 49), (1, 2), (200, 1000)):
            u = np.zeros(500)
            r = np.zeros(300, dtype=np.uint32)
            k = min(perm_n, 64)
            p = np.zeros(k, dtype=np.int32)
            L.or_np_probe(seed, 500, u.ctypes.data, 300, bound, r.ctypes.data, perm_n, k, p.ctypes.data)
            rs = np.random.RandomState(seed)
            np.testing.assert_array_equal(u, [rs.random_sample() for _ in range(500)])
            np.testing.assert_array_equal(r, [rs.randint(0, bound) for _ in range(300)])
            np.testing.assert_array_equal(p, rs.permutation(perm_n)[:k])


def test_philox_matches_python():
    L = coracle.lib()
    rng = np.random.RandomState(0)
    for _ in range(200):
        ctr = rng.randint(0, 2**32, 4, dtype=np.uint64).astype(np.uint32)
        key = rng.randint(0, 2**32, 2, dtype=np.uint64).astype(np.uint32)
        out = np.zeros(4, dtype=np.uint32)
        L.or_philox_probe(ctr.ctypes.data, key.ctypes.data, out.ctypes.data)
        assert tuple(int(x) for x in out) == R.philox4x32_10(ctr, key)


@pytest.mark.parametrize("name", util.golden_cases(with_large=True))
def test_c_oracle_reproduces_reference(name):
    g = util.load_golden(name)
    spec = util.spec_from_golden(g)
    T, E, F, A = int(g["T"]), int(g["epochs"]), int(g["n_frames"]), int(g["A"])
    co = coracle.COracle(spec, 1, capacity=int(g["capacity"]), n_frames=F, rng=coracle.RNG_NUMPY,
                         seed=int(g["seed"]))
    per_turn = g["grids"].shape[0] == T * E
    gi = 0
    for ep in range(E):
        co.reset()
        yx = np.stack([co.pos[0] // spec.width, co.pos[0] % spec.width], axis=-1)
        np.testing.assert_array_equal(yx, g["starts"][ep])
        if F > 1:
            co.add_empty()
        if per_turn:
            for t in range(T):
                co.step(1)
                np.testing.assert_array_equal(co.grid[0], g["grids"][gi], err_msg=f"epoch {ep} turn {t}")
                gi += 1
        else:
            co.step(T)
            np.testing.assert_array_equal(co.grid[0], g["grids"][ep])
        assert co.total_reward[0] == float(g["total_reward"][ep])
    ring = co.ring()
    for a in range(A):
        np.testing.assert_array_equal(ring["states"][a], g["states"][a])
        np.testing.assert_array_equal(ring["actions"][:, a], g["actions"][a])
        np.testing.assert_array_equal(ring["rewards"][:, a], g["rewards"][a])
        np.testing.assert_array_equal(ring["dones"][:, a], g["dones"][a])
        np.testing.assert_array_equal(ring["positions"][:, a], g["positions"][a])
    assert co.idx == int(g["buf_idx"][0]) and co.size == int(g["buf_size"][0])


@pytest.mark.parametrize("A,H,W,v,p,T,N", [(1, 21, 21, 2, 0.005, 40, 12), (5, 9, 12, 2, 0.3, 15, 7),
                                           (2, 8, 8, 1, 1.0, 5, 3)])
def test_c_oracle_philox_equals_python(A, H, W, v, p, T, N):
    """Counter-based mode: C (multi-threaded) == Python restatement, including the tapes layout."""
    spec = R.treasurehunt_spec(H, W, A, v, p)
    seed, base = 0xDEADBEEF1234, 11
    co = coracle.COracle(spec, N, capacity=T, rng=coracle.RNG_PHILOX, seed=seed, env_id_base=base, threads=4)
    co.reset()
    co.step(T // 2)
    co.step(T - T // 2)
    ring = co.ring()
    for n in range(N):
        env, _ = util.run_oracle_episode(spec, R.PhiloxDraws(seed), T, env_id=base + n)
        np.testing.assert_array_equal(co.grid[n], env.grid)
        assert co.total_reward[n] == env.total_reward
        for a in range(A):
            np.testing.assert_array_equal(ring["states"][n * A + a], env.buffers[a].states)
            np.testing.assert_array_equal(ring["positions"][:, n * A + a], env.buffers[a].positions)


@pytest.mark.parametrize("A,H,W,v,p,T,N", [(1, 21, 21, 2, 0.005, 40, 10), (8, 21, 21, 2, 0.005, 20, 6),
                                           (2, 11, 14, 3, 0.03, 30, 7), (3, 40, 40, 4, 0.002, 10, 4),
                                           (1, 21, 21, 2, 0.02, 30, 5)])
def test_c_oracle_sparse_equals_python(A, H, W, v, p, T, N):
    """SB_RNG_SPARSE in both restatements: Binomial count by inverse CDF + distinct uniform cells (same stream)."""
    spec = R.treasurehunt_spec(H, W, A, v, p)
    seed, base = 0xFEEDC0DE77, 9
    co = coracle.COracle(spec, N, capacity=T, rng=coracle.RNG_SPARSE, seed=seed, env_id_base=base, threads=3)
    co.reset()
    co.step(T // 3)
    co.step(T - T // 3)
    ring = co.ring()
    spawned = 0
    for n in range(N):
        env, _ = util.run_oracle_episode(spec, R.SparseDraws(seed), T, env_id=base + n)
        np.testing.assert_array_equal(co.grid[n], env.grid)
        assert co.total_reward[n] == env.total_reward
        spawned += int(np.isin(env.grid[:, :, 1], spec.spawn_choices).sum())
        for a in range(A):
            np.testing.assert_array_equal(ring["states"][n * A + a], env.buffers[a].states)
            np.testing.assert_array_equal(ring["rewards"][:, n * A + a], env.buffers[a].rewards)
    assert spawned > 0


def test_c_oracle_env_id_list_replays_a_sample():
    """`env_ids`: instance n of a sampled oracle == instance env_ids[n] of the full batch (counter-based modes)."""
    spec = R.treasurehunt_spec(21, 21, 2, 2, 0.01)
    T, N, seed = 12, 40, 77
    for mode in (coracle.RNG_PHILOX, coracle.RNG_SPARSE):
        full = coracle.COracle(spec, N, capacity=T, rng=mode, seed=seed, env_id_base=100, threads=2)
        ids = np.array([100, 107, 139, 121], dtype=np.uint64)
        part = coracle.COracle(spec, len(ids), capacity=T, rng=mode, seed=seed, env_ids=ids)
        for co in (full, part):
            co.reset()
            co.step(T)
        sel = (ids - 100).astype(int)
        np.testing.assert_array_equal(part.grid, full.grid[sel])
        np.testing.assert_array_equal(part.total_reward, full.total_reward[sel])
        fr, pr = full.ring(), part.ring()
        rows = (sel[:, None] * 2 + np.arange(2)[None]).reshape(-1)
        np.testing.assert_array_equal(pr["states"], fr["states"][rows])
        np.testing.assert_array_equal(pr["positions"], fr["positions"][:, rows])


def test_c_oracle_tapes_equal_python_tapes():
    spec = R.treasurehunt_spec(11, 13, 3, 2, 0.05)
    T, N = 20, 5
    co = coracle.COracle(spec, N, capacity=T, rng=coracle.RNG_NUMPY, seed=500, record=True, threads=2)
    co.reset()
    co.step(T)
    ct = co.tapes()
    for n in range(N):
        env, _ = util.run_oracle_episode(spec, R.NumpyDraws(500 + n), T)
        u, c = util.pad_tapes(spec, env.dense_tapes())
        np.testing.assert_array_equal(ct["spawn_u"][:, n], u)
        np.testing.assert_array_equal(ct["spawn_c"][:, n], c)
        np.testing.assert_array_equal(ct["actions"][:, n], env.dense_tapes()["actions"])
        np.testing.assert_array_equal(ct["placement"][n], env.dense_tapes()["placement"])


@pytest.mark.parametrize("name", util.cleanup_cases())
def test_c_oracle_reproduces_reference_cleanup(name):
    """The C restatement on the cleanup rule set (column rules, beams, facings, all-layer rewards, positional
    code, apples placed before agents) against the fixtures of the unmodified reference."""
    g = util.load_golden(name)
    spec = util.cleanup_spec_from_golden(g)
    T, E, F, A = int(g["T"]), int(g["epochs"]), int(g["n_frames"]), int(g["A"])
    co = coracle.COracle(spec, 1, capacity=int(g["capacity"]), n_frames=F, rng=coracle.RNG_NUMPY,
                         seed=int(g["seed"]))
    gi = 0
    for ep in range(E):
        co.reset()
        if F > 1:
            co.add_empty()
        yx = np.stack([co.pos[0] // spec.width, co.pos[0] % spec.width], axis=-1)
        np.testing.assert_array_equal(yx, g["starts"][ep])
        np.testing.assert_array_equal(co.grid[0], g["grids"][gi], err_msg="after reset")
        gi += 1
        for t in range(T):
            co.step(1)
            np.testing.assert_array_equal(co.grid[0], g["grids"][gi], err_msg=f"epoch {ep} turn {t}")
            np.testing.assert_array_equal(co.facing[0], g["facings"][ep * T + t])
            gi += 1
        assert co.total_reward[0] == float(g["total_reward"][ep])
    ring = co.ring()
    for a in range(A):
        np.testing.assert_array_equal(ring["states"][a], g["states"][a])
        np.testing.assert_array_equal(ring["actions"][:, a], g["actions"][a])
        np.testing.assert_array_equal(ring["rewards"][:, a], g["rewards"][a])
        np.testing.assert_array_equal(ring["positions"][:, a], g["positions"][a])


def test_c_oracle_philox_equals_python_cleanup():
    """Counter-based mode on the cleanup rule set: C (multi-threaded) == Python restatement."""
    spec = R.cleanup_spec(13, 17, 5, 3, initial_apples=6, apple_spawn_chance=0.03, pollution_spawn_chance=0.06,
                          beam_radius=2)
    N, T, seed, base = 9, 16, 77, 500
    co = coracle.COracle(spec, N, capacity=2 * T, rng=coracle.RNG_PHILOX, seed=seed, env_id_base=base, threads=3)
    pys = [R.OracleEnv(spec, R.PhiloxDraws(seed), env_id=base + n, capacity=2 * T) for n in range(N)]
    for ep in range(2):
        co.reset()
        co.step(T)
        for n, o in enumerate(pys):
            o.reset()
            for _ in range(T):
                o.take_turn()
            np.testing.assert_array_equal(co.grid[n], o.grid, err_msg=f"episode {ep} instance {n}")
            np.testing.assert_array_equal(co.facing[n], o.facing)
            assert co.total_reward[n] == o.total_reward
    ring = co.ring()
    for n, o in enumerate(pys):
        for a in range(spec.num_agents):
            row = n * spec.num_agents + a
            np.testing.assert_array_equal(ring["states"][row], o.buffers[a].states)
            np.testing.assert_array_equal(ring["rewards"][:, row], o.buffers[a].rewards)
            np.testing.assert_array_equal(ring["actions"][:, row], o.buffers[a].actions)


@pytest.mark.parametrize("name", util.iowa_cases())
def test_c_oracle_reproduces_reference_iowa(name):
    """The C restatement on the iowa rule set (value states, float rewards, float64 total) vs the reference fixtures."""
    g = util.load_golden(name)
    spec = util.iowa_spec_from_golden(g)
    T, E, A = int(g["T"]), int(g["epochs"]), int(g["A"])
    co = coracle.COracle(spec, 1, capacity=int(g["capacity"]), rng=coracle.RNG_NUMPY, seed=int(g["seed"]))
    gi = 0
    for ep in range(E):
        co.reset()
        for t in range(T):
            co.step(1)
            np.testing.assert_array_equal(co.grid[0], g["grids"][gi], err_msg=f"epoch {ep} turn {t}")
            gi += 1
        assert co.total_reward[0] == float(g["total_reward"][ep])
    ring = co.ring()
    for a in range(A):
        np.testing.assert_array_equal(ring["states"][a], g["states"][a])
        np.testing.assert_array_equal(ring["actions"][:, a], g["actions"][a])
        np.testing.assert_array_equal(ring["rewards"][:, a], g["rewards"][a])
        np.testing.assert_array_equal(ring["positions"][:, a], g["positions"][a])


def test_c_oracle_philox_equals_python_iowa():
    spec = R.iowa_spec(10, 13, 3, 2, 0.06)
    N, T, seed = 7, 25, 31
    co = coracle.COracle(spec, N, capacity=T, rng=coracle.RNG_PHILOX, seed=seed, env_id_base=40, threads=2)
    co.reset()
    co.step(T)
    ring = co.ring()
    for n in range(N):
        o = R.OracleEnv(spec, R.PhiloxDraws(seed), env_id=40 + n, capacity=T)
        o.reset()
        for _ in range(T):
            o.take_turn()
        np.testing.assert_array_equal(co.grid[n], o.grid)
        assert co.total_reward[n] == o.total_reward
        for a in range(3):
            np.testing.assert_array_equal(ring["rewards"][:, n * 3 + a], o.buffers[a].rewards)
            np.testing.assert_array_equal(ring["states"][n * 3 + a], o.buffers[a].states)


@pytest.mark.parametrize("name", util.tag_cases())
def test_c_oracle_reproduces_reference_tag(name):
    g = util.load_golden(name)
    spec = util.tag_spec_from_golden(g)
    T, E, A = int(g["T"]), int(g["epochs"]), int(g["A"])
    co = coracle.COracle(spec, 1, capacity=int(g["capacity"]), rng=coracle.RNG_NUMPY, seed=int(g["seed"]))
    gi = 0
    for ep in range(E):
        co.reset()
        np.testing.assert_array_equal(co.grid[0], g["grids"][gi], err_msg="after reset")
        gi += 1
        for t in range(T):
            co.step(1)
            np.testing.assert_array_equal(co.grid[0], g["grids"][gi], err_msg=f"epoch {ep} turn {t}")
            np.testing.assert_array_equal(co.flags[0], g["flags"][ep * T + t])
            gi += 1
        assert co.total_reward[0] == float(g["total_reward"][ep])
    ring = co.ring()
    for a in range(A):
        np.testing.assert_array_equal(ring["states"][a], g["states"][a])
        np.testing.assert_array_equal(ring["actions"][:, a], g["actions"][a])
        np.testing.assert_array_equal(ring["rewards"][:, a], g["rewards"][a])
        np.testing.assert_array_equal(ring["positions"][:, a], g["positions"][a])
