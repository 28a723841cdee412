"""CPU: the oracle restatement against the fixtures generated from the unmodified reference."""
from __future__ import annotations

import numpy as np
import pytest

from oracle import restatement as R
import sb_testutil as util


@pytest.mark.parametrize("name", util.golden_cases(with_large=True))
def test_restatement_reproduces_reference(name):
    g = util.load_golden(name)
    spec = util.spec_from_golden(g)
    T, E, F = int(g["T"]), int(g["epochs"]), int(g["n_frames"])
    env = R.OracleEnv(spec, R.NumpyDraws(int(g["seed"])), capacity=int(g["capacity"]), n_frames=F)
    per_turn = g["grids"].shape[0] == T * E
    gi = 0
    for ep in range(E):
        env.reset()
        np.testing.assert_array_equal(env.pos, g["starts"][ep])
        if F > 1:
            for b in env.buffers:
                b.add_empty()
        for t in range(T):
            env.take_turn()
            if per_turn:
                np.testing.assert_array_equal(env.grid, g["grids"][gi], err_msg=f"epoch {ep} turn {t}")
                gi += 1
            if F > 1:
                cs = np.stack([b.current_state() for b in env.buffers])
                np.testing.assert_array_equal(cs, g["current_states"][ep * T + t])
        if not per_turn:
            np.testing.assert_array_equal(env.grid, g["grids"][ep])
        assert env.total_reward == float(g["total_reward"][ep])
    for a, b in enumerate(env.buffers):
        for key, arr in (("states", b.states), ("actions", b.actions), ("rewards", b.rewards),
                         ("dones", b.dones), ("positions", b.positions)):
            assert arr.dtype == g[key].dtype, key
            np.testing.assert_array_equal(arr, g[key][a], err_msg=key)
        assert b.idx == int(g["buf_idx"][a]) and b.size == int(g["buf_size"][a])
    if "sample_idx" in g:
        for a, b in enumerate(env.buffers):
            idx = b.sample_indices(g["sample_idx"].shape[1], env.draws.rs)
            np.testing.assert_array_equal(idx, g["sample_idx"][a])
            for o, key in zip(b.gather(idx), ("sample_states", "sample_actions", "sample_rewards",
                                              "sample_next_states", "sample_dones", "sample_valid")):
                assert o.dtype == g[key].dtype and o.shape == g[key][a].shape
                np.testing.assert_array_equal(o, g[key][a])


def test_survey_known_answers():
    """SURVEY.md §8c KAT hashes were taken with set_seed *before* construction (the constructor's
    own placement draw precedes reset's); reproduce that protocol: one extra placement first."""
    import hashlib

    for seed, A, H, v, p, T, total, buf_sha, grid_sha in [
        (0, 1, 21, 2, 0.005, 100, 36.0, "9853cd136a7cd606", "92f6882502c9dcfb"),
        (0, 8, 21, 2, 0.005, 100, 75.0, "5945c54e2c481aaa", "81d4e963d247f69e"),
        (7, 2, 11, 2, 0.05, 50, 30.0, "0704132958714fb2", "27f1c7c75978b829"),
    ]:
        spec = R.treasurehunt_spec(H, H, A, v, p)
        env = R.OracleEnv(spec, R.NumpyDraws(seed), capacity=T)
        env.reset()  # Environment.__init__ -> populate_environment (environment.py:54)
        env.reset()  # env.reset() of the harness
        for _ in range(T):
            env.take_turn()
        assert env.total_reward == total
        h = hashlib.sha256()
        for b in env.buffers:
            for arr in (b.states, b.actions, b.rewards, b.dones, b.positions):
                h.update(arr.tobytes())
        assert h.hexdigest()[:16] == buf_sha
        assert hashlib.sha256(np.ascontiguousarray(env.grid[:, :, 1]).tobytes()).hexdigest()[:16] == grid_sha


def test_observe_closed_form_against_reference_outputs():
    g = util.load_golden("observe_cases")
    app = np.zeros((4, 4))
    for k in range(1, 4):
        app[k, k] = 1.0
    for i in range(int(g["n"])):
        kinds = g[f"kinds_{i}"]
        v, y, x = (int(t) for t in g[f"meta_{i}"])
        H, W, L = kinds.shape
        spec = R.WorldSpec(H, W, L, ["E", "W", "G", "B"], [0] * 4, [True] * 4, [0] * 4, app, kinds,
                           np.zeros((0, 3)), 3, 0, 0, 1, v, 0, 0.0, [])
        o = R.observe(spec, kinds, y, x)
        assert o.dtype == np.float64 and o.shape == g[f"obs_{i}"].shape
        np.testing.assert_array_equal(o, g[f"obs_{i}"])


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kats = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
            ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
            ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
             (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for c, k, exp in kats:
        assert R.philox4x32_10(c, k) == exp


def test_tape_replay_equals_numpy_run():
    """TapeDraws (what the kernels consume) replays a NumpyDraws run exactly."""
    spec = R.treasurehunt_spec(11, 11, 3, 2, 0.05)
    env, grids = util.run_oracle_episode(spec, R.NumpyDraws(42), 30)
    t = env.dense_tapes()
    env2 = R.OracleEnv(spec, R.TapeDraws(t["spawn_u"], t["spawn_c"], t["actions"], t["placement"]), capacity=30)
    env2.reset()
    for i in range(30):
        env2.take_turn()
        np.testing.assert_array_equal(env2.grid, grids[i])
    for a in range(3):
        np.testing.assert_array_equal(env2.buffers[a].states, env.buffers[a].states)


def test_spawn_threshold_is_exact():
    for p in (0.005, 0.05, 0.5, 1 / 3, 1e-9, 0.999999):
        thr = R.spawn_threshold(p)
        assert (thr - 1) / 2**53 < p <= thr / 2**53 or thr / 2**53 == p


def test_sparse_sampler_matches_dense_law():
    """SparseDraws and PhiloxDraws sample the same law: spawn counts per turn ~ Binomial(#empty, p)."""
    spec = R.treasurehunt_spec(21, 21, 1, 2, 0.005)
    n_env = 150  # the dense Python sampler evaluates Philox per cell: keep it small (the GPU test uses 2^16)
    tot = {"dense": 0, "sparse": 0}
    for name, cls in (("dense", R.PhiloxDraws), ("sparse", R.SparseDraws)):
        for e in range(n_env):
            env = R.OracleEnv(spec, cls(99), env_id=e)
            env.reset()
            before = int(((env.grid[:, :, 1] >= 2) & (env.grid[:, :, 1] <= 4)).sum())
            env.take_turn()
            tot[name] += int(((env.grid[:, :, 1] >= 2) & (env.grid[:, :, 1] <= 4)).sum()) - before
    expect = n_env * 360 * 0.005
    for v in tot.values():
        assert abs(v - expect) < 4 * expect ** 0.5 + 1, tot
    cdf = R.binomial_cdf_u64(448, 0.005)
    assert cdf == sorted(cdf) and cdf[-1] == (1 << 64) - 1
    # P(K = 0) = (1 - q)^448 exactly
    from fractions import Fraction
    q = Fraction(R.spawn_threshold(0.005), 1 << 53)
    assert cdf[0] == int((1 - q) ** 448 * (1 << 64))


def test_rgb_observe_against_reference_outputs():
    """RGBObservationSpec.observe outputs captured from the reference (sum over layers, clip, /255)."""
    g = util.load_golden("observe_rgb_cases")
    for i in range(int(g["n"])):
        kinds = g[f"kinds_{i}"]
        v, y, x = (int(t) for t in g[f"meta_{i}"])
        H, W, L = kinds.shape
        spec = R.WorldSpec(H, W, L, ["E", "W", "G", "B"], [0] * 4, [True] * 4, [0] * 4, g[f"table_{i}"], kinds,
                           np.zeros((0, 3)), 3, 0, 0, 1, v, 0, 0.0, [], obs_transform=1)
        o = R.observe(spec, kinds, y, x)
        np.testing.assert_array_equal(o, g[f"obs_{i}"])


@pytest.mark.parametrize("name", util.cleanup_cases())
def test_restatement_reproduces_reference_cleanup(name):
    """examples/cleanup: column rules (River/Pollution/AppleTree/Apple), beams with a
    one-turn lifetime, facings, all-layer rewards, one-hot crop + positional code -- every turn's grid
    (beam ages included), facings, total reward and all buffer arrays equal the reference's."""
    g = util.load_golden(name)
    spec = util.cleanup_spec_from_golden(g)
    T, E, F = int(g["T"]), int(g["epochs"]), int(g["n_frames"])
    env = R.OracleEnv(spec, R.NumpyDraws(int(g["seed"])), capacity=int(g["capacity"]), n_frames=F)
    gi = 0
    for ep in range(E):
        env.reset()
        if F > 1:
            for b in env.buffers:
                b.add_empty()
        np.testing.assert_array_equal(env.pos, g["starts"][ep])
        np.testing.assert_array_equal(env.grid, g["grids"][gi], err_msg=f"epoch {ep} after reset")
        gi += 1
        for t in range(T):
            env.take_turn()
            np.testing.assert_array_equal(env.grid, g["grids"][gi], err_msg=f"epoch {ep} turn {t}")
            np.testing.assert_array_equal(env.facing, g["facings"][ep * T + t])
            gi += 1
        assert env.total_reward == float(g["total_reward"][ep])
    for a, b in enumerate(env.buffers):
        for key, arr in (("states", b.states), ("actions", b.actions), ("rewards", b.rewards),
                         ("dones", b.dones), ("positions", b.positions)):
            assert arr.dtype == g[key].dtype, key
            np.testing.assert_array_equal(arr, g[key][a], err_msg=key)
        assert b.idx == int(g["buf_idx"][a]) and b.size == int(g["buf_size"][a])


@pytest.mark.parametrize("name", util.iowa_cases())
def test_restatement_reproduces_reference_iowa(name):
    """examples/iowa (per-entity mutable value): decks redraw their value at every sweep (three
    kind rows per deck), a deck spawned this turn is worth 0 until its first sweep, Sand is visible, walls on both
    layers; rewards are non-integer Python floats, so `total_reward` is checked in float64 and the buffer in float32."""
    g = util.load_golden(name)
    spec = util.iowa_spec_from_golden(g)
    T, E = int(g["T"]), int(g["epochs"])
    env = R.OracleEnv(spec, R.NumpyDraws(int(g["seed"])), capacity=int(g["capacity"]))
    gi = 0
    for ep in range(E):
        env.reset()
        np.testing.assert_array_equal(env.pos, g["starts"][ep])
        for t in range(T):
            env.take_turn()
            np.testing.assert_array_equal(env.grid, g["grids"][gi], err_msg=f"epoch {ep} turn {t}")
            gi += 1
        assert env.total_reward == float(g["total_reward"][ep])
    for a, b in enumerate(env.buffers):
        for key, arr in (("states", b.states), ("actions", b.actions), ("rewards", b.rewards),
                         ("dones", b.dones), ("positions", b.positions)):
            assert arr.dtype == g[key].dtype, key
            np.testing.assert_array_equal(arr, g[key][a], err_msg=key)


@pytest.mark.parametrize("name", util.tag_cases())
def test_restatement_reproduces_reference_tag(name):
    """examples/tag (agent kind flips): the It agent shows as another kind, hands the role to the
    first adjacent NotIt agent after its move, rewards follow the agent's own state, the flag is appended to the
    observation and survives reset."""
    g = util.load_golden(name)
    spec = util.tag_spec_from_golden(g)
    T, E = int(g["T"]), int(g["epochs"])
    env = R.OracleEnv(spec, R.NumpyDraws(int(g["seed"])), capacity=int(g["capacity"]))
    gi = 0
    for ep in range(E):
        env.reset()
        np.testing.assert_array_equal(env.pos, g["starts"][ep])
        np.testing.assert_array_equal(env.grid, g["grids"][gi], err_msg="after reset")
        gi += 1
        for t in range(T):
            env.take_turn()
            np.testing.assert_array_equal(env.grid, g["grids"][gi], err_msg=f"epoch {ep} turn {t}")
            np.testing.assert_array_equal(env.flags, g["flags"][ep * T + t])
            gi += 1
        assert env.total_reward == float(g["total_reward"][ep])
    for a, b in enumerate(env.buffers):
        for key, arr in (("states", b.states), ("actions", b.actions), ("rewards", b.rewards),
                         ("dones", b.dones), ("positions", b.positions)):
            assert arr.dtype == g[key].dtype, key
            np.testing.assert_array_equal(arr, g[key][a], err_msg=key)
