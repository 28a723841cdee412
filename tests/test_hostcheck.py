"""CPU: the kernels' per-lane device functions (sb_core.cuh), compiled for the host with g++, against the
oracle.  This unit-tests the arithmetic the GPU executes (Philox, byte-SIMD spawn filter, lazy 53-bit
Bernoulli, action decode / move, general observation sum) without a GPU.  The library built here is
test-only and is never loaded by sorrel_b200 (it is not a fallback: it contains no kernel)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import restatement as R

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hc") / "libhostcheck.so")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "hostcheck", "hostcheck.cpp")])
    L = C.CDLL(out)
    L.hc_filter_chunk.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    L.hc_filter_chunk.restype = C.c_uint32
    L.hc_filter_bound.argtypes = [C.c_double]
    L.hc_filter_bound.restype = C.c_uint32
    L.hc_resolve_candidate.argtypes = [C.c_double, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
    L.hc_action.argtypes = [C.c_uint32] * 4 + [C.c_uint64]
    L.hc_action.restype = C.c_uint32
    L.hc_agent_act.restype = C.c_float
    L.hc_agent_act.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_int, C.c_int,
                               C.c_void_p, C.c_int, C.c_void_p]
    L.hc_obs_general.restype = C.c_double
    L.hc_obs_general.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p] + \
                                [C.c_int] * 9
    return L


def test_philox_known_answers_and_round_key_variant(hc):
    rng = np.random.RandomState(1)
    kats = [((0, 0, 0, 0), (0, 0)), ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2),
            ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0))]
    kats += [(tuple(int(x) for x in rng.randint(0, 2**32, 4, dtype=np.uint64)),
              tuple(int(x) for x in rng.randint(0, 2**32, 2, dtype=np.uint64))) for _ in range(300)]
    for ctr, key in kats:
        c, k = np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32)
        o1, o2 = np.zeros(4, dtype=np.uint32), np.zeros(4, dtype=np.uint32)
        hc.hc_philox(c.ctypes.data, k.ctypes.data, o1.ctypes.data, o2.ctypes.data)
        ref = R.philox4x32_10(ctr, key)
        assert tuple(int(x) for x in o1) == ref  # plain Philox4x32-10 (Random123 known answers above)
        # the engine's keyed stream: fixed key, the seed words folded into counter words 1 and 3
        seeded = R.philox4x32_10((ctr[0], ctr[1] ^ key[0], ctr[2], ctr[3] ^ key[1]), R.PHILOX_FIXED_KEY)
        assert tuple(int(x) for x in o2) == seeded
        d = R.PhiloxDraws(key[0] | (key[1] << 32))
        assert d.px(*ctr) == seeded


@pytest.mark.parametrize("p", [0.005, 0.05, 0.2, 0.49, 1e-4, 0.00390625])
def test_byte_simd_filter_is_exact_superset(hc, p):
    """Filter mask == {cell is the spawning kind and r8 < n2}: exactly the cells that pass level 1 of the two-level
    Bernoulli draw, n2 = the power of two R.level1_bound gives (>= the largest accepted top byte + 1)."""
    rng = np.random.RandomState(int(p * 1e6))
    P = R.spawn_threshold(p)
    n2 = hc.hc_filter_bound(p)
    assert n2 == R.level1_bound(P) and n2 & (n2 - 1) == 0 and n2 >= ((P - 1) >> 45) + 1
    assert (P << 19) // n2 <= 1 << 64  # the level-2 bound fits: P(hit) = (n2 / 256) * (T / 2^64) = P / 2^53
    for spawn_kind in (0, 3, 7):
        for _ in range(400):
            kinds = np.where(rng.rand(16) < 0.5, spawn_kind, rng.randint(0, 256, 16)).astype(np.uint8)
            rnd = np.where(rng.rand(16) < 0.4, rng.randint(0, n2 + 2, 16), rng.randint(0, 256, 16)).clip(0, 255)
            rnd = rnd.astype(np.uint8)
            got = hc.hc_filter_chunk(kinds.ctypes.data, rnd.ctypes.data, spawn_kind, p)
            ref = sum(1 << b for b in range(16) if kinds[b] == spawn_kind and rnd[b] < n2)
            assert got == ref


def test_lazy_bernoulli_resolution_matches_oracle(hc):
    """sb_resolve_candidate == the oracle's PhiloxDraws.spawn for the same (seed, env, tick, cell)."""
    rng = np.random.RandomState(7)
    for _ in range(4000):
        p = float(rng.choice([0.005, 0.05, 0.3, 0.9, 1.0, 1 / 256, 1 / 128, 3 / 256]))
        n = int(rng.randint(1, 9))
        env, tick, cell = int(rng.randint(0, 2**20)), int(rng.randint(1, 2**16)), int(rng.randint(0, 4096))
        seed = int(rng.randint(0, 2**62))
        got = hc.hc_resolve_candidate(p, n, env, tick, cell, seed)
        _, c = R.PhiloxDraws(seed).spawn(p, n, env, tick, cell, cell)
        assert got == c


def test_action_draw_matches_oracle(hc):
    rng = np.random.RandomState(3)
    for _ in range(2000):
        env, tick, a, n = int(rng.randint(0, 2**24)), int(rng.randint(1, 2**20)), int(rng.randint(0, 64)), int(rng.randint(1, 9))
        seed = int(rng.randint(0, 2**62))
        assert hc.hc_action(env, tick, a, n, seed) == R.PhiloxDraws(seed).action(n, env, tick, a)


def test_agent_act_matches_reference_move_semantics(hc):
    """Walls block (-1), items are collected (value, cell -> agent, old cell -> empty), agents block (0),
    off-grid sets the error bit (gridworld.py:95-122, agents/agent.py:215-225)."""
    H = W = 6
    value = np.array([0, -1, 10, -10, 5, 0], dtype=np.float32)
    passable = 0b011101
    rng = np.random.RandomState(0)
    for _ in range(500):
        grid = np.full((H * W + 15) // 16 * 16, 0xFF, dtype=np.uint8)
        cells = rng.choice([0, 1, 2, 3, 4, 5], size=H * W, p=[0.4, 0.15, 0.15, 0.1, 0.1, 0.1]).astype(np.uint8)
        y, x, action = int(rng.randint(0, H)), int(rng.randint(0, W)), int(rng.randint(0, 5))
        cells[y * W + x] = 5
        grid[: H * W] = cells
        ref = cells.reshape(H, W).copy()
        yx = np.array([(y << 16) | x], dtype=np.uint32)
        err = np.zeros(1, dtype=np.uint32)
        r = hc.hc_agent_act(grid.ctypes.data, H, W, 6, value.ctypes.data, passable, 5, 0, yx.ctypes.data, action,
                            err.ctypes.data)
        if action >= 4:
            assert err[0] & 4
            continue
        dy, dx = R.ACTION_DELTAS[action]
        ny, nx = y + dy, x + dx
        if not (0 <= ny < H and 0 <= nx < W):
            assert err[0] & 1 and r == 0.0
            continue
        t = ref[ny, nx]
        assert r == value[t] and err[0] == 0
        if (passable >> t) & 1:
            ref[ny, nx], ref[y, x] = 5, 0
            assert int(yx[0]) == (ny << 16) | nx
        else:
            assert int(yx[0]) == (y << 16) | x
        np.testing.assert_array_equal(grid[: H * W].reshape(H, W), ref)


def test_general_observation_sum_matches_oracle(hc):
    rng = np.random.RandomState(11)
    for _ in range(40):
        H, W, L, K, Cn = int(rng.randint(5, 10)), int(rng.randint(5, 10)), int(rng.randint(1, 4)), 5, 4
        v = int(rng.randint(1, min((H - 1) // 2, (W - 1) // 2) + 1))
        app = rng.randint(0, 256, (K, Cn)).astype(np.float64)
        kinds = rng.randint(0, K, (H, W, L)).astype(np.uint8)
        dyn = sorted(rng.choice(L, size=int(rng.randint(1, L + 1)), replace=False).tolist())
        ps = (H * W + 15) // 16 * 16
        planes = np.full((len(dyn), ps), 0xFF, dtype=np.uint8)
        tmpl = np.full((L, ps), 0xFF, dtype=np.uint8)
        pol = np.full(8, -1, dtype=np.int32)
        for z in range(L):
            tmpl[z, : H * W] = kinds[:, :, z].reshape(-1)
            if z in dyn:
                pol[z] = dyn.index(z)
                planes[pol[z], : H * W] = kinds[:, :, z].reshape(-1)
                tmpl[z, : H * W] = 0  # dynamic layers must come from the planes, not the template
        spec = R.WorldSpec(H, W, L, ["k"] * K, [0] * K, [True] * K, [0] * K, app, kinds, np.zeros((0, 3)), 0, 0, 0, 1, v,
                           0, 0.0, [])
        y, x = int(rng.randint(0, H)), int(rng.randint(0, W))
        ref = R.observe(spec, kinds, y, x)
        S = 2 * v + 1
        for c in range(Cn):
            for i in range(S):
                for j in range(S):
                    got = hc.hc_obs_general(planes.ctypes.data, tmpl.ctypes.data, app.ctypes.data, H, W, L,
                                            pol.ctypes.data, K, Cn, v, 1, y, x, c, i, j)
                    assert got == ref[c, i, j]


def test_bernoulli_become_matches_oracle(hc):
    """SB_RULE_SPAWN_TO (River -> Pollution, AppleTree -> Apple): sb_bernoulli_hit == PhiloxDraws.bernoulli."""
    hc.hc_bernoulli_hit.argtypes = [C.c_double, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
    rng = np.random.RandomState(11)
    hits = 0
    for _ in range(6000):
        p = float(rng.choice([0.002, 0.009, 0.05, 0.5, 1.0, 0.0, 1 / 256, 2 / 256, 255 / 256]))
        env, tick, cell = int(rng.randint(0, 2**20)), int(rng.randint(1, 2**16)), int(rng.randint(0, 4096))
        seed = int(rng.randint(0, 2**62))
        got = hc.hc_bernoulli_hit(p, env, tick, cell, seed)
        _, want = R.PhiloxDraws(seed).bernoulli(p, env, tick, cell, cell)
        assert bool(got) == bool(want)
        hits += got
    assert 1000 < hits < 3000  # both outcomes are exercised


def test_column_rules(hc):
    """SB_RULE_BECOME_IF reads the same column on another plane; SB_RULE_BECOME is unconditional."""
    hc.hc_column_rule.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32]
    rng = np.random.RandomState(5)
    stride = 32
    for _ in range(500):
        planes = rng.randint(0, 12, size=2 * stride).astype(np.uint8)
        off = int(rng.randint(0, stride))
        mask = int(rng.randint(0, 1 << 12))
        to = int(rng.randint(0, 12))
        other = int(planes[stride + off])
        got = hc.hc_column_rule(planes.ctypes.data, stride, off, R.RULE_BECOME_IF, to, 1, mask)
        assert got == (to if (mask >> other) & 1 else -1)
        assert hc.hc_column_rule(planes.ctypes.data, stride, off, R.RULE_BECOME, to, 1, mask) == to
        assert hc.hc_column_rule(planes.ctypes.data, stride, off, R.RULE_NONE, to, 1, mask) == -1


def test_agent_act_all_layer_reward(hc):
    """cleanup/agents.py:158-176: reward = sum of values over all layers at the target, whether or not the move
    succeeds; non-movement actions read the agent's own column."""
    hc.hc_agent_act_all_layers.restype = C.c_float
    hc.hc_agent_act_all_layers.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32,
                                           C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    H, W, L, K = 5, 7, 3, 6
    value = np.array([0, 0, 1, -1, 3, 0], dtype=np.float32)  # kinds: 0 empty, 1 wall, 2 apple, 3 zap, 4 gem, 5 agent
    passable = 0b000001
    stride = (H * W + 15) // 16 * 16
    rng = np.random.RandomState(2)
    for _ in range(400):
        planes = np.full(L * stride, 0xFF, dtype=np.uint8)
        img = rng.choice([0, 1, 2, 3, 4], size=(L, H, W), p=[0.5, 0.1, 0.15, 0.15, 0.1]).astype(np.uint8)
        img[1] = rng.choice([0, 1, 5], size=(H, W), p=[0.7, 0.15, 0.15])  # agent layer: empty / wall / other agents
        y, x = int(rng.randint(1, H - 1)), int(rng.randint(1, W - 1))
        img[1, y, x] = 5
        for z in range(L):
            planes[z * stride: z * stride + H * W] = img[z].reshape(-1)
        action = int(rng.randint(0, 6))
        dy, dx = ((-1, 0), (1, 0), (0, -1), (0, 1), (0, 0), (0, 0))[action]
        ny, nx = y + dy, x + dx
        want = float(sum(value[img[z, ny, nx]] for z in range(L)))
        moved = img[1, ny, nx] == 0
        yx = np.array([(y << 16) | x], dtype=np.uint32)
        err = np.zeros(1, dtype=np.uint32)
        r = hc.hc_agent_act_all_layers(planes.ctypes.data, H, W, L, K, value.ctypes.data, passable, 1, 5, 0,
                                       yx.ctypes.data, action, err.ctypes.data)
        assert r == want and err[0] == 0
        assert (int(yx[0]) >> 16, int(yx[0]) & 0xFFFF) == ((ny, nx) if moved else (y, x))
        after = planes[stride: stride + H * W].reshape(H, W)
        if moved:
            assert after[ny, nx] == 5 and after[y, x] == 0
        else:
            assert after[y, x] == 5


def test_touch_swap_order_and_conditions(hc):
    """TagAgent.act (tag/agents.py:84-99): only an "It" agent hands over, to the FIRST "NotIt" neighbour in the order
    up, right, down, left (Location.adjacent, location.py:144); nothing else changes."""
    hc.hc_touch_swap.argtypes = [C.c_void_p] + [C.c_int] * 6
    H, W, IT, NOT = 6, 7, 2, 3
    rng = np.random.RandomState(4)
    stride = (H * W + 15) // 16 * 16
    for _ in range(600):
        plane = np.full(stride, 0xFF, dtype=np.uint8)
        cells = rng.choice([0, 1, NOT, IT], size=(H, W), p=[0.45, 0.1, 0.35, 0.1]).astype(np.uint8)
        y, x = int(rng.randint(0, H)), int(rng.randint(0, W))
        cells[y, x] = IT if rng.rand() < 0.7 else NOT
        plane[: H * W] = cells.reshape(-1)
        want = cells.copy()
        if cells[y, x] == IT:
            for oy, ox in ((-1, 0), (0, 1), (1, 0), (0, -1)):
                qy, qx = y + oy, x + ox
                if 0 <= qy < H and 0 <= qx < W and cells[qy, qx] == NOT:
                    want[y, x], want[qy, qx] = NOT, IT
                    break
        got = hc.hc_touch_swap(plane.ctypes.data, H, W, y, x, NOT, IT)
        np.testing.assert_array_equal(plane[: H * W].reshape(H, W), want)
        assert got == want[y, x]
