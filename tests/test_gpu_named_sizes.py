"""GPU parity at the sizes BASELINE.json's configs name (VERDICT r01, row ★h): bit-exact against the C oracle.

* config 2  4096 instances x 1 agent x 100 turns, REPLAYED numpy draws (`COracle(RNG_NUMPY)` tapes), every array;
* config 3  16384 instances x 8 agents, counter-based stream, every array, ring wrapped;
* config 5  2^20 instances with the bench's 32-deep ring (20 GB: byte offsets far beyond 2^32), bench cadence
            (reset, add_empty, 25 + 25 fused turns at the 4-lane tiling): every 4099th instance and the whole last
            block, dense and sparse samplers;
* config 4  1024 instances of 128x128 maps with 64 agents and 9x9 windows, every array.

Reference semantics: Environment.take_turn sorrel/environment.py:81-93, Buffer.add sorrel/buffers.py:46-63.
The oracle is the checker only; everything under test goes through the C ABI.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import coracle
from oracle import restatement as R
import sb_testutil as util

pytestmark = pytest.mark.gpu


def _engine(spec, n, device, **kw):
    from sorrel_b200.engine import Engine

    return Engine(util.engine_spec(spec), n, device=device, **kw)


def _buffer(spec, n, capacity, device, n_frames=1, layout="agent"):
    from sorrel_b200.buffers import Buffer

    return Buffer(capacity, (spec.obs_size,), n_frames=n_frames, rows=n * spec.num_agents, device=device,
                  layout=layout, positions=(2,))


def _compare_all(eng, buf, co):
    np.testing.assert_array_equal(eng.grid_numpy(), co.grid)
    np.testing.assert_array_equal(eng.agent_pos.cpu().numpy(), co.pos)
    np.testing.assert_array_equal(eng.total_reward.cpu().numpy(), co.total_reward)
    ring = co.ring()
    np.testing.assert_array_equal(buf.actions.cpu().numpy(), ring["actions"])
    np.testing.assert_array_equal(buf.rewards.cpu().numpy(), ring["rewards"])
    np.testing.assert_array_equal(buf.dones.cpu().numpy(), ring["dones"])
    np.testing.assert_array_equal(buf.extra_data["positions"].cpu().numpy(), ring["positions"])
    rows = ring["states"].shape[0]
    for lo in range(0, rows, 1 << 15):  # chunked: the whole ring never sits on the host twice
        hi = min(rows, lo + (1 << 15))
        np.testing.assert_array_equal(buf.states[lo:hi].cpu().numpy(), ring["states"][lo:hi],
                                      err_msg=f"states rows {lo}:{hi}")
    assert (buf.idx, buf.size) == (co.idx, co.size)


def test_config2_4096_envs_numpy_replay(cuda_device):
    """BASELINE configs[1]: "4096 envs x 1 agent ... bit-exact vs reference replayed RNG": 100 turns of recorded
    numpy-legacy draws (one MT19937 stream per instance, seeded like `set_seed(seed + n)`), four 25-turn launches."""
    spec = R.treasurehunt_spec(21, 21, 1, 2, 0.005)
    N, T, chunk = 4096, 100, 25
    co = coracle.COracle(spec, N, capacity=T, rng=coracle.RNG_NUMPY, seed=31000, threads=coracle.max_threads(),
                         record=True)
    co.reset()
    eng = _engine(spec, N, cuda_device)
    buf = _buffer(spec, N, T, cuda_device)
    placed = False
    drew_total = 0
    for _ in range(T // chunk):
        co.step(chunk)
        tp = co.tapes()
        if not placed:
            eng.reset(torch.from_numpy(tp["placement"].copy()))
            placed = True
        drew = torch.zeros((chunk, N, eng.env_stride), dtype=torch.uint8, device=cuda_device)
        eng.step(chunk, buffer=buf, actions=torch.from_numpy(tp["actions"].copy()).to(cuda_device),
                 spawn_u=torch.from_numpy(tp["spawn_u"]).to(cuda_device),
                 spawn_c=torch.from_numpy(tp["spawn_c"]).to(cuda_device), drew_out=drew)
        # stream alignment: the cells that consumed a uniform are exactly the tape's non-NaN cells
        assert torch.equal(drew.bool().cpu(), torch.from_numpy(~np.isnan(tp["spawn_u"])))
        drew_total += int(drew.sum())
    eng.check_errors()  # SB_ERR_TAPE_ALIGN would be raised here
    assert drew_total > N * T * 200
    _compare_all(eng, buf, co)


@pytest.mark.parametrize("layout", ["slot", "agent"])
def test_config3_16384_envs_8_agents(layout, cuda_device):
    """BASELINE configs[2]: 16384 instances x 8 agents (movement conflicts, list-order priority), counter-based
    stream, 25 + 15 turns into a 16-deep ring (wraps), the tiling the library picks at this size; with the ring layout
    `Environment` picks for multi-agent worlds (slot-major) and with the agent-major one."""
    spec = R.treasurehunt_spec(21, 21, 8, 2, 0.005)
    N, cap, seed, base = 16384, 16, 0xC3C3C3, 1 << 20
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed)
    assert eng.step_plan(25)["lanes_per_instance"] == 8
    buf = _buffer(spec, N, cap, cuda_device, n_frames=5, layout=layout)
    co = coracle.COracle(spec, N, capacity=cap, n_frames=5, rng=coracle.RNG_PHILOX, seed=seed, env_id_base=base,
                         threads=coracle.max_threads())
    eng.reset()
    buf.add_empty()
    co.reset()
    co.add_empty()
    for turns in (25, 15):
        eng.step(turns, buffer=buf)
        co.step(turns)
    eng.check_errors()
    _compare_all(eng, buf, co)
    # conflicts did happen: some agent was blocked by another agent (reward 0, position unchanged on an attempt
    # to enter an occupied cell is indistinguishable here, so count instances where two agents ended adjacent)
    pos = eng.agent_yx()
    d = np.abs(pos[:, :, None, :] - pos[:, None, :, :]).sum(-1)
    assert ((d == 1).sum(axis=(1, 2)) > 0).mean() > 0.2


@pytest.mark.parametrize("mode", ["philox", "sparse"])
def test_config5_one_million_envs_sampled(mode, cuda_device):
    """BASELINE configs[4] exactly as bench.py runs it: 2^20 instances, ring capacity 32 (20.1 GB: the last rows start
    ~20 GB into the tensor), F = 5, reset + add_empty, then 25 + 25 fused turns (the second launch wraps the ring) with
    the 4-lane tiling.  Every 4099th instance plus the whole last block are replayed by the C oracle (same global
    instance ids) and compared in every array."""
    spec = R.treasurehunt_spec(21, 21, 1, 2, 0.005)
    N, cap, F, seed = 1 << 20, 32, 5, 0
    free, _ = torch.cuda.mem_get_info(cuda_device)
    if free < 26 << 30:
        pytest.skip("needs ~22 GB of free device memory")
    eng = _engine(spec, N, cuda_device, seed=seed)
    plan = eng.step_plan(25, rng=mode)
    assert plan["lanes_per_instance"] == 4 and plan["planes_staged"] and plan["fast_onehot"]
    buf = _buffer(spec, N, cap, cuda_device, n_frames=F)
    assert buf.states.numel() * 4 > 1 << 34  # offsets beyond 2^32 (and 2^34) bytes are exercised
    ids = np.unique(np.concatenate([np.arange(0, N, 4099), np.arange(N - 32, N)])).astype(np.uint64)
    co = coracle.COracle(spec, len(ids), capacity=cap, n_frames=F, seed=seed, env_ids=ids,
                         rng=coracle.RNG_SPARSE if mode == "sparse" else coracle.RNG_PHILOX, threads=4)
    for _ in range(2):  # two episodes of two launches: reset at tick 0 and tick 50, like the bench's boundary
        eng.reset()
        buf.add_empty()
        co.reset()
        co.add_empty()
        for turns in (25, 25):
            eng.step(turns, buffer=buf, rng=mode)
            co.step(turns)
    eng.check_errors()
    sel = torch.from_numpy(ids.astype(np.int64)).to(cuda_device)
    H = W = 21
    grid = eng.grid[sel][:, 0, : H * W].reshape(-1, H, W).cpu().numpy()
    np.testing.assert_array_equal(grid, co.grid[:, :, :, 1])
    np.testing.assert_array_equal(eng.agent_pos[sel].cpu().numpy(), co.pos)
    np.testing.assert_array_equal(eng.total_reward[sel].cpu().numpy(), co.total_reward)
    ring = co.ring()
    np.testing.assert_array_equal(buf.states[sel].cpu().numpy(), ring["states"])
    np.testing.assert_array_equal(buf.actions[:, sel].cpu().numpy(), ring["actions"])
    np.testing.assert_array_equal(buf.rewards[:, sel].cpu().numpy(), ring["rewards"])
    np.testing.assert_array_equal(buf.dones[:, sel].cpu().numpy(), ring["dones"])
    np.testing.assert_array_equal(buf.extra_data["positions"][:, sel].cpu().numpy(), ring["positions"])
    assert (buf.idx, buf.size) == (co.idx, co.size)
    # Buffer.sample gather on rows at the far end of the 20 GB ring (buffers.py:112-124)
    rows = sel[-16:]
    starts = torch.arange(16, device=cuda_device) % (cap - F - 1)
    s, a, r, ns, d, v = buf.gather(rows, starts)
    buf.check_errors()
    st = ring["states"][-16:]
    for b in range(16):
        s0 = int(starts[b])
        np.testing.assert_array_equal(s[b].cpu().numpy(), st[b, s0:s0 + F].reshape(-1))
        np.testing.assert_array_equal(ns[b].cpu().numpy(), st[b, s0 + 1:s0 + F + 1].reshape(-1))
        assert int(a[b]) == ring["actions"][s0 + F - 1, -16 + b] and float(r[b]) == ring["rewards"][s0 + F - 1, -16 + b]
    del buf, eng
    torch.cuda.empty_cache()


def test_config4_large_map_64_agents(cuda_device):
    """BASELINE configs[3]: 128x128 maps, 64 agents, 9x9 one-hot windows, 1024 instances, every array.  The library's
    choice at this shape (block per instance: plane staged in shared memory, up to 16 independent agents at once) is
    what is checked."""
    spec = R.treasurehunt_spec(128, 128, 64, 4, 0.005)
    N, cap, seed, base = 1024, 4, 0x4C4C, 77
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed)
    plan = eng.step_plan(6)
    assert plan["lanes_per_instance"] == 128 and plan["planes_staged"]  # block per instance
    buf = _buffer(spec, N, cap, cuda_device, layout="slot")
    co = coracle.COracle(spec, N, capacity=cap, rng=coracle.RNG_PHILOX, seed=seed, env_id_base=base,
                         threads=coracle.max_threads())
    eng.reset()
    co.reset()
    for turns in (6, 4):
        eng.step(turns, buffer=buf)
        co.step(turns)
    eng.check_errors()
    _compare_all(eng, buf, co)


def test_sparse_sampler_full_batch_distribution(cuda_device):
    """SB_RNG_SPARSE at 2^20 instances, one turn from identical start states (entities.py:72-75: iid Bernoulli(p) per
    EmptyEntity cell + uniform kind): the number of spawns per instance follows Binomial(360, p) (chi-square on the
    histogram), every interior cell spawns with the same marginal probability, kinds are uniform; and the dense
    per-cell sampler passes the same checks."""
    spec = R.treasurehunt_spec(21, 21, 1, 2, 0.005)
    N = 1 << 20
    from math import comb

    P = R.spawn_threshold(0.005) / float(1 << 53)
    for mode in ("sparse", "philox"):
        eng = _engine(spec, N, cuda_device, seed=9001)
        eng.reset()
        eng.step(1, rng=mode)
        g = eng.grid[:, 0, : 441].view(N, 21, 21)
        item = (g >= 2) & (g <= 4)
        counts = item.sum(dim=(1, 2))
        hist = torch.bincount(counts, minlength=12).double().cpu().numpy()
        n_cells = 360  # 19x19 interior minus the agent's cell
        expect = np.array([comb(n_cells, k) * P ** k * (1 - P) ** (n_cells - k) for k in range(len(hist))]) * N
        keep = expect > 20
        chi2 = float((((hist - expect) ** 2) / np.maximum(expect, 1e-9))[keep].sum())
        assert chi2 < 60, (mode, chi2, hist[:8], expect[:8])  # dof ~ 7: P(chi2 > 60) < 1e-9
        per_cell = item[:, 1:-1, 1:-1].sum(dim=0).double().cpu().numpy()  # spawns per interior cell
        mean = N * P * (360.0 / 361.0)  # a cell holds the agent with probability 1/361
        z = (per_cell - mean) / np.sqrt(mean)
        assert np.abs(z).max() < 6.0, (mode, float(np.abs(z).max()))
        kinds = np.array([int((g == k).sum()) for k in (2, 3, 4)], dtype=np.float64)
        assert np.all(np.abs(kinds - kinds.sum() / 3) < 6 * np.sqrt(kinds.sum() / 3)), (mode, kinds)
        del eng
