"""GPU parity for the cleanup rule set: column rules, beams with a one-turn lifetime,
facings, all-layer rewards and the one-hot crop + positional code, through the C ABI.

Replay mode is compared with fixtures generated from the unmodified reference (oracle/gen_golden.py
cleanup); the counter-based mode with the oracle restatement running the same Philox stream."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import restatement as R
import sb_testutil as util

pytestmark = pytest.mark.gpu


def _engine(spec, n, device, **kw):
    from sorrel_b200.engine import Engine

    return Engine(util.engine_spec(spec), n, device=device, **kw)


def _buffer(spec, n, capacity, device, n_frames=1):
    from sorrel_b200.buffers import Buffer

    return Buffer(capacity, (spec.obs_size,), n_frames=n_frames, rows=n * spec.num_agents, device=device,
                  positions=(2,))


def _episode_tapes(spec, tapes, device):
    u, c = util.pad_tapes(spec, tapes)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(device)  # noqa: E731
    return (dev(u[:, None]), dev(c[:, None]), dev(tapes["actions"][:, None]),
            dev(tapes["placement"].astype(np.int32)[None]),
            [dev(ip.astype(np.int32)[None]) for ip in tapes["item_placement"]])


@pytest.mark.parametrize("name", util.cleanup_cases())
@pytest.mark.parametrize("fused", [False, True])
def test_cleanup_golden_replay(name, fused, cuda_device):
    """Replayed numpy draws: every turn's grid (beam ages included), facings, total reward and all buffer
    arrays equal the reference's, one launch per turn and with the whole episode fused into one launch."""
    g = util.load_golden(name)
    spec = util.cleanup_spec_from_golden(g)
    T, E, A, F = int(g["T"]), int(g["epochs"]), int(g["A"]), int(g["n_frames"])
    oenv = R.OracleEnv(spec, R.NumpyDraws(int(g["seed"])), capacity=int(g["capacity"]), n_frames=F, record_tapes=True)
    eng = _engine(spec, 1, cuda_device)
    buf = _buffer(spec, 1, int(g["capacity"]), cuda_device, n_frames=F)
    gi = 0
    for ep in range(E):
        oenv.tapes.update({"spawn_u": [], "spawn_c": [], "actions": []})
        oenv.reset()
        for _ in range(T):
            oenv.take_turn()
        spawn_u, spawn_c, actions, placement, item_pl = _episode_tapes(spec, oenv.dense_tapes(), cuda_device)
        eng.reset(placement, item_placement=item_pl)
        if F > 1:
            buf.add_empty()
        np.testing.assert_array_equal(eng.agent_yx()[0], g["starts"][ep])
        np.testing.assert_array_equal(eng.grid_numpy()[0], g["grids"][gi], err_msg="after reset")
        gi += 1
        if fused:
            eng.step(T, buffer=buf, actions=actions, spawn_u=spawn_u, spawn_c=spawn_c)
            gi += T
            np.testing.assert_array_equal(eng.grid_numpy()[0], g["grids"][gi - 1])
        else:
            drew = torch.zeros((1, 1, eng.env_stride), dtype=torch.uint8, device=cuda_device)
            for t in range(T):
                eng.step(1, buffer=buf, actions=actions[t:t + 1].contiguous(), spawn_u=spawn_u[t:t + 1].contiguous(),
                         spawn_c=spawn_c[t:t + 1].contiguous(), drew_out=drew)
                np.testing.assert_array_equal(eng.grid_numpy()[0], g["grids"][gi], err_msg=f"epoch {ep} turn {t}")
                np.testing.assert_array_equal(eng.agent_facing.cpu().numpy()[0], g["facings"][ep * T + t])
                np.testing.assert_array_equal(drew.cpu().numpy()[0, 0] == 1, ~np.isnan(spawn_u[t, 0].cpu().numpy()))
                gi += 1
        eng.check_errors()
        assert float(eng.total_reward[0]) == float(g["total_reward"][ep])
        np.testing.assert_array_equal(eng.agent_facing.cpu().numpy()[0], g["facings"][(ep + 1) * T - 1])
    for a in range(A):
        row = buf.row(a)
        for key in ("states", "actions", "rewards", "dones", "positions"):
            assert row[key].dtype == g[key].dtype
            np.testing.assert_array_equal(row[key], g[key][a], err_msg=f"{key} agent {a}")


@pytest.mark.parametrize("lpe", [None, 16, 8])
def test_cleanup_philox_batch(lpe, cuda_device):
    """Counter-based mode (in-kernel item / agent placement, Bernoulli draws, uniform policy): many instances
    against the oracle restatement on the same stream, including a second episode (facings persist)."""
    tuning = {"lpe": lpe} if lpe else {}
    spec = R.cleanup_spec(13, 17, 5, 3, initial_apples=6, apple_spawn_chance=0.03, pollution_spawn_chance=0.06,
                          beam_radius=2)
    N, T, seed, base = 40, 24, 99, 1000
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed, tuning=tuning)
    buf = _buffer(spec, N, 2 * T, cuda_device)
    oenvs = [R.OracleEnv(spec, R.PhiloxDraws(seed), env_id=base + n, capacity=2 * T) for n in range(0, N, 9)]
    for ep in range(2):
        eng.reset()
        for o in oenvs:
            o.reset()
        if ep == 0:
            grids = eng.grid_numpy()
            for i, n in enumerate(range(0, N, 9)):
                np.testing.assert_array_equal(grids[n], oenvs[i].grid, err_msg=f"reset, instance {n}")
        eng.step(T // 2, buffer=buf)  # two launches per episode: state carried through global memory
        eng.step(T - T // 2, buffer=buf)
        for o in oenvs:
            for _ in range(T):
                o.take_turn()
        eng.check_errors()
        grids, tot = eng.grid_numpy(), eng.total_reward.cpu().numpy()
        for i, n in enumerate(range(0, N, 9)):
            np.testing.assert_array_equal(grids[n], oenvs[i].grid, err_msg=f"episode {ep}, instance {n}")
            assert tot[n] == oenvs[i].total_reward
            np.testing.assert_array_equal(eng.agent_facing.cpu().numpy()[n], oenvs[i].facing)
    for i, n in enumerate(range(0, N, 9)):
        for a in range(spec.num_agents):
            row, ob = buf.row(n * spec.num_agents + a), oenvs[i].buffers[a]
            np.testing.assert_array_equal(row["states"], ob.states)
            np.testing.assert_array_equal(row["actions"], ob.actions)
            np.testing.assert_array_equal(row["rewards"], ob.rewards)
            np.testing.assert_array_equal(row["positions"], ob.positions)


def test_cleanup_observe_standalone(cuda_device):
    """sb_observe on a cleanup world == the rows the fused step wrote (crop + positional code)."""
    spec = R.cleanup_spec(21, 31, 10, 5)
    N = 6
    eng = _engine(spec, N, cuda_device, seed=5)
    buf = _buffer(spec, N, 4, cuda_device)
    eng.reset()
    eng.step(3, buffer=buf)
    obs = eng.observe()  # state after 3 turns
    assert tuple(obs.shape) == (N, spec.num_agents, spec.obs_size)
    oenv = R.OracleEnv(spec, R.PhiloxDraws(5), env_id=2, capacity=4)
    oenv.reset()
    for _ in range(3):
        oenv.take_turn()
    for a in range(spec.num_agents):
        y, x = oenv.pos[a]
        want = np.concatenate((R.observe(spec, oenv.grid, y, x).reshape(-1), spec.obs_extra[y, x])).astype(np.float32)
        np.testing.assert_array_equal(obs[2, a].cpu().numpy(), want)
    eng.check_errors()


def test_cleanup_full_size_invariants(cuda_device):
    """The example's own shape at 16 384 instances x 10 agents x 40 turns (sizes the oracle cannot replay in a test):
    size-independent invariants of the rule set."""
    spec = R.cleanup_spec()
    H, W, A = spec.height, spec.width, spec.num_agents
    N, T, cap = 1 << 14, 40, 4
    eng = _engine(spec, N, cuda_device, seed=21)
    buf = _buffer(spec, N, cap, cuda_device)
    eng.reset()
    rew = torch.zeros((T, N, A), dtype=torch.float32, device=cuda_device)
    act = torch.zeros((T, N, A), dtype=torch.int32, device=cuda_device)
    eng.step(T, buffer=buf, reward_out=rew, action_out=act)
    eng.check_errors()
    g = eng.grid[:, :, : H * W].view(N, 3, H, W)
    tmpl = torch.from_numpy(np.moveaxis(spec.template, 2, 0).copy()).to(cuda_device)  # [3, H, W]
    obj, agents, beams = g[:, 0], g[:, 1], g[:, 2]
    # walls never move or appear, on any layer
    assert ((g == 1) == (tmpl == 1)[None]).all()
    # object layer: River <-> Pollution on river cells, AppleTree <-> Apple on orchard cells, Sand fixed
    river, orchard, sand = (tmpl[0] == 2)[None], (tmpl[0] == 4)[None], (tmpl[0] == 9)[None]
    assert (((obj == 2) | (obj == 3)) == river).all() and (((obj == 4) | (obj == 5)) == orchard).all()
    assert ((obj == 9) == sand).all()
    # agent layer: exactly A agents per instance, at their recorded positions; nothing else but walls / empty
    assert ((agents == 8).sum(dim=(1, 2)) == A).all()
    assert (agents.reshape(N, -1).gather(1, eng.agent_pos.long()) == 8).all()
    assert ((agents == 0) | (agents == 1) | (agents == 8)).all()
    # beam layer: empty, wall, or a beam of either age; a beam lives at most two sweeps
    assert ((beams == 0) | (beams == 1) | (beams == 6) | (beams == 7) | (beams == 10) | (beams == 11)).all()
    # an Apple never lies under an agent after the sweep... but it may right after a move: only check the rule's
    # direction: pollution spawns at Bernoulli(0.009) per river cell and turn and is only removed by CleanBeams
    n_river = int(river.sum())
    frac = float((obj == 3).sum()) / (N * n_river)
    assert 0.1 < frac < 1 - (1 - 0.009) ** T + 0.02
    # rewards: +1 per Apple tile entered, -1 per ZapBeam over the target tile, integers in [-1, 1] here; cleanup
    # adds every reward to world.total_reward twice
    assert ((rew == -1) | (rew == 0) | (rew == 1)).all()
    assert torch.equal(eng.total_reward, 2.0 * rew.double().sum(dim=(0, 2)))
    assert ((act >= 0) & (act < 6)).all() and float((act >= 4).float().mean()) > 0.3  # both beams are used
    # facings are what the last movement action left (or the initial 2)
    moved = act < 4
    last = torch.where(moved, torch.arange(T, device=cuda_device)[:, None, None], -1).max(dim=0).values
    table = torch.tensor([0, 2, 3, 1], device=cuda_device)
    want = torch.where(last >= 0, table[act.long().gather(0, last.clamp(min=0)[None])[0]], 2)
    assert torch.equal(eng.agent_facing.long(), want)
    # observation rows: crop entries are small non-negative integers (layers summed), own channel set at the centre,
    # and the last 12 floats are the positional code of the cell the agent observed from
    st = buf.states.view(N * A, cap, -1)
    crop = st[:, :, : spec.onehot_size].view(N * A, cap, 9, 11, 11)
    assert ((crop == 0) | (crop == 1) | (crop == 2) | (crop == 3)).all() and (crop[:, :, 0] == 0).all()
    assert (crop[:, :, 8, 5, 5] == 1).all()
    codes = torch.from_numpy(spec.obs_extra.reshape(H * W, -1).astype(np.float32)).to(cuda_device)
    got = st[:, :, spec.onehot_size:]
    # the row in slot (T-1) % cap was observed from the position BEFORE the last move; check membership instead:
    # every code row equals some cell's code and that cell is within 1 step of the final position
    final = eng.agent_pos.reshape(-1).long()
    slot = (T - 1) % cap
    cand = torch.stack([final, final - 1, final + 1, final - W, final + W], dim=1).clamp(0, H * W - 1)
    match = (codes[cand] == got[:, slot][:, None, :]).all(dim=2).any(dim=1)
    assert match.all()


# ------------------------------------------------------------------------------------------------
# examples/iowa: value states (SPAWN_TO with an else branch), float rewards, float64 total_reward
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", util.iowa_cases())
@pytest.mark.parametrize("fused", [False, True])
def test_iowa_golden_replay(name, fused, cuda_device):
    g = util.load_golden(name)
    spec = util.iowa_spec_from_golden(g)
    T, E, A = int(g["T"]), int(g["epochs"]), int(g["A"])
    oenv = R.OracleEnv(spec, R.NumpyDraws(int(g["seed"])), capacity=int(g["capacity"]), record_tapes=True)
    eng = _engine(spec, 1, cuda_device)
    buf = _buffer(spec, 1, int(g["capacity"]), cuda_device)
    gi = 0
    for ep in range(E):
        oenv.tapes.update({"spawn_u": [], "spawn_c": [], "actions": []})
        oenv.reset()
        for _ in range(T):
            oenv.take_turn()
        spawn_u, spawn_c, actions, placement, _ = _episode_tapes(spec, oenv.dense_tapes(), cuda_device)
        eng.reset(placement)
        np.testing.assert_array_equal(eng.agent_yx()[0], g["starts"][ep])
        if fused:
            eng.step(T, buffer=buf, actions=actions, spawn_u=spawn_u, spawn_c=spawn_c)
            gi += T
            np.testing.assert_array_equal(eng.grid_numpy()[0], g["grids"][gi - 1])
        else:
            for t in range(T):
                eng.step(1, buffer=buf, actions=actions[t:t + 1].contiguous(), spawn_u=spawn_u[t:t + 1].contiguous(),
                         spawn_c=spawn_c[t:t + 1].contiguous())
                np.testing.assert_array_equal(eng.grid_numpy()[0], g["grids"][gi], err_msg=f"epoch {ep} turn {t}")
                gi += 1
        eng.check_errors()
        assert float(eng.total_reward[0]) == float(g["total_reward"][ep])  # float64 sum of 1.1, -1.4, 0.6, ...
    for a in range(A):
        row = buf.row(a)
        for key in ("states", "actions", "rewards", "dones", "positions"):
            assert row[key].dtype == g[key].dtype
            np.testing.assert_array_equal(row[key], g[key][a], err_msg=f"{key} agent {a}")


@pytest.mark.parametrize("lpe", [None, 8])  # 8 lanes is what pick_lpe gives this world at large batch sizes
def test_iowa_philox_batch(lpe, cuda_device):
    from oracle import coracle

    tuning = {"lpe": lpe} if lpe else {}

    spec = R.iowa_spec(20, 20, 2, 2, 0.03)
    N, T, seed, base = 300, 40, 5, 77
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed, tuning=tuning)
    buf = _buffer(spec, N, T, cuda_device)
    eng.reset()
    eng.step(T // 2, buffer=buf)
    eng.step(T - T // 2, buffer=buf)
    eng.check_errors()
    co = coracle.COracle(spec, N, capacity=T, rng=coracle.RNG_PHILOX, seed=seed, env_id_base=base, threads=4)
    co.reset()
    co.step(T)
    np.testing.assert_array_equal(eng.grid_numpy(), co.grid)
    np.testing.assert_array_equal(eng.total_reward.cpu().numpy(), co.total_reward)
    ring = co.ring()
    np.testing.assert_array_equal(buf.states.cpu().numpy(), ring["states"])
    np.testing.assert_array_equal(buf.rewards.cpu().numpy(), ring["rewards"])
    np.testing.assert_array_equal(buf.actions.cpu().numpy(), ring["actions"])
    assert len(np.unique(ring["rewards"])) >= 7  # walls, fresh decks and the win / loss states of every deck


# ------------------------------------------------------------------------------------------------
# examples/tag: agents with a second kind ("It"), role hand-over to a neighbour, own-state rewards, flag in the row
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", util.tag_cases())
@pytest.mark.parametrize("fused", [False, True])
def test_tag_golden_replay(name, fused, cuda_device):
    g = util.load_golden(name)
    spec = util.tag_spec_from_golden(g)
    T, E, A = int(g["T"]), int(g["epochs"]), int(g["A"])
    oenv = R.OracleEnv(spec, R.NumpyDraws(int(g["seed"])), capacity=int(g["capacity"]), record_tapes=True)
    eng = _engine(spec, 1, cuda_device)
    buf = _buffer(spec, 1, int(g["capacity"]), cuda_device)
    gi = 0
    for ep in range(E):
        oenv.tapes.update({"spawn_u": [], "spawn_c": [], "actions": []})
        oenv.reset()
        for _ in range(T):
            oenv.take_turn()
        spawn_u, spawn_c, actions, placement, _ = _episode_tapes(spec, oenv.dense_tapes(), cuda_device)
        eng.reset(placement)  # the flags of the previous episode decide who is drawn as "It"
        np.testing.assert_array_equal(eng.grid_numpy()[0], g["grids"][gi], err_msg="after reset")
        gi += 1
        if fused:
            eng.step(T, buffer=buf, actions=actions, spawn_u=spawn_u, spawn_c=spawn_c)
            gi += T
            np.testing.assert_array_equal(eng.grid_numpy()[0], g["grids"][gi - 1])
        else:
            for t in range(T):
                eng.step(1, buffer=buf, actions=actions[t:t + 1].contiguous(), spawn_u=spawn_u[t:t + 1].contiguous(),
                         spawn_c=spawn_c[t:t + 1].contiguous())
                np.testing.assert_array_equal(eng.grid_numpy()[0], g["grids"][gi], err_msg=f"epoch {ep} turn {t}")
                np.testing.assert_array_equal(eng.agent_flag.cpu().numpy()[0], g["flags"][ep * T + t])
                gi += 1
        eng.check_errors()
        assert float(eng.total_reward[0]) == float(g["total_reward"][ep])
        np.testing.assert_array_equal(eng.agent_flag.cpu().numpy()[0], g["flags"][(ep + 1) * T - 1])
    for a in range(A):
        row = buf.row(a)
        for key in ("states", "actions", "rewards", "dones", "positions"):
            assert row[key].dtype == g[key].dtype
            np.testing.assert_array_equal(row[key], g[key][a], err_msg=f"{key} agent {a}")


@pytest.mark.parametrize("lpe", [None, 8, 4])
def test_tag_philox_batch(lpe, cuda_device):
    from oracle import coracle

    tuning = {"lpe": lpe} if lpe else {}

    spec = R.tag_spec(9, 10, 6, 3, 10, it_agent=2)
    N, T, seed, base = 200, 30, 13, 9
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed, tuning=tuning)
    buf = _buffer(spec, N, 2 * T, cuda_device)
    co = coracle.COracle(spec, N, capacity=2 * T, rng=coracle.RNG_PHILOX, seed=seed, env_id_base=base, threads=4)
    for ep in range(2):  # the second episode starts from the flags the first one left
        eng.reset()
        eng.step(T, buffer=buf)
        co.reset()
        co.step(T)
        eng.check_errors()
        np.testing.assert_array_equal(eng.grid_numpy(), co.grid)
        np.testing.assert_array_equal(eng.agent_flag.cpu().numpy(), co.flags)
        np.testing.assert_array_equal(eng.total_reward.cpu().numpy(), co.total_reward)
    ring = co.ring()
    np.testing.assert_array_equal(buf.states.cpu().numpy(), ring["states"])
    np.testing.assert_array_equal(buf.rewards.cpu().numpy(), ring["rewards"])
    assert (eng.agent_flag.sum(dim=1) == 1).all()  # exactly one "It" per instance, always
    obs = eng.observe()
    assert tuple(obs.shape) == (N, 6, spec.obs_size)
    np.testing.assert_array_equal(obs[:, :, -1].cpu().numpy(), co.flags.astype(np.float32))


# ------------------------------------------------------------------------------------------------
# randomized differential test of the rule-set extensions against the C oracle
# ------------------------------------------------------------------------------------------------
def _fuzz_rule_cases(n=8, seed=77):
    rng = np.random.RandomState(seed)
    out = []
    for i in range(n):
        H, W = int(rng.randint(9, 26)), int(rng.randint(9, 34))
        v = int(rng.randint(1, min(5, (min(H, W) - 1) // 2) + 1))
        lpe = [None, 8, 16, 32][int(rng.randint(0, 4))]
        out.append(("cleanup", i, H, W, int(rng.randint(1, 9)), v, lpe, int(rng.randint(1, 4))))
        out.append(("iowa", i, H, W, int(rng.randint(1, 7)), min(v, 3), lpe, 0))
        out.append(("tag", i, min(H, 16), min(W, 16), int(rng.randint(2, 9)), min(v, (min(H, W, 16) - 1) // 2), lpe, 0))
    return out


@pytest.mark.parametrize("case", _fuzz_rule_cases(), ids=lambda c: f"{c[0]}{c[1]}-{c[2]}x{c[3]}a{c[4]}v{c[5]}")
def test_fuzz_rule_sets_against_c_oracle(case, cuda_device):
    from oracle import coracle

    kind, i, H, W, A, v, lpe, radius = case
    tuning = {"lpe": lpe} if lpe else {}
    if kind == "cleanup":
        spec = R.cleanup_spec(H, W, A, v, initial_apples=3, apple_spawn_chance=0.02, pollution_spawn_chance=0.05,
                              beam_radius=radius)
    elif kind == "iowa":
        spec = R.iowa_spec(H, W, A, v, 0.04)
    else:
        spec = R.tag_spec(H, W, A, v, 10, it_agent=i % A)
    N, T, seed, base = 17, 18, 500 + i, 11 * i
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed, tuning=tuning)
    buf = _buffer(spec, N, T, cuda_device)
    co = coracle.COracle(spec, N, capacity=T, rng=coracle.RNG_PHILOX, seed=seed, env_id_base=base, threads=4)
    eng.reset()
    eng.step(T // 3, buffer=buf)
    eng.step(T - T // 3, buffer=buf)
    eng.check_errors()
    co.reset()
    co.step(T)
    np.testing.assert_array_equal(eng.grid_numpy(), co.grid)
    np.testing.assert_array_equal(eng.total_reward.cpu().numpy(), co.total_reward)
    ring = co.ring()
    np.testing.assert_array_equal(buf.states.cpu().numpy(), ring["states"])
    np.testing.assert_array_equal(buf.actions.cpu().numpy(), ring["actions"])
    np.testing.assert_array_equal(buf.rewards.cpu().numpy(), ring["rewards"])
    if kind == "cleanup":
        np.testing.assert_array_equal(eng.agent_facing.cpu().numpy(), co.facing)
    if kind == "tag":
        np.testing.assert_array_equal(eng.agent_flag.cpu().numpy(), co.flags)


# ------------------------------------------------------------------------------------------------
# column gather of several staged planes (sb_step.cu: scatter_cols_multi): every window width, every group width
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lpe", [4, 8, 16, 32])
@pytest.mark.parametrize("v", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
def test_multi_plane_window_gather_all_widths(v, lpe, cuda_device):
    """Cleanup worlds (three stored planes, 12 kinds: both channel tables; walls on every layer: entries counted
    three times; beams and agents on top of terrain) at every window width a descriptor accepts and every group
    width: all rows of the ring against the C oracle.  Agents start anywhere, so windows hang over every edge."""
    from oracle import coracle

    H, W = 23, 27
    spec = R.cleanup_spec(H, W, 4, v, initial_apples=5, apple_spawn_chance=0.03, pollution_spawn_chance=0.05,
                          beam_radius=2)
    N, T, seed, base = 9, 6, 40 + v, 3 * v
    eng = _engine(spec, N, cuda_device, env_id_base=base, seed=seed, tuning={"lpe": lpe})
    buf = _buffer(spec, N, T, cuda_device)
    try:
        eng.step_plan(T)
    except ValueError:  # 32 instances of this size per block do not fit in shared memory
        assert lpe == 4 and v >= 8
        pytest.skip("tiling pinned by the test does not fit this shape")
    co = coracle.COracle(spec, N, capacity=T, rng=coracle.RNG_PHILOX, seed=seed, env_id_base=base, threads=4)
    eng.reset()
    eng.step(T, buffer=buf)
    eng.check_errors()
    co.reset()
    co.step(T)
    np.testing.assert_array_equal(eng.grid_numpy(), co.grid)
    ring = co.ring()
    np.testing.assert_array_equal(buf.states.cpu().numpy(), ring["states"])
    np.testing.assert_array_equal(buf.rewards.cpu().numpy(), ring["rewards"])


@pytest.mark.parametrize("lpe", [None, 4, 8, 32])
@pytest.mark.parametrize("v", [2, 4])
def test_two_plane_world_with_few_kinds(v, lpe, cuda_device):
    """A treasurehunt world whose floor layer is stored per instance as well: two planes, 7 kinds (one channel
    table), byte-SIMD sweep + column gather, against the C oracle of the same two-plane description (the draw
    stream is keyed by the stored-plane cell index, so it differs from the one-plane layout's)."""
    import dataclasses
    from oracle import coracle

    spec1 = R.treasurehunt_spec(15, 18, 3, v, 0.02)
    spec2 = dataclasses.replace(spec1, dynamic_layers=[0, 1])
    N, T, seed = 21, 12, 7
    eng = _engine(spec2, N, cuda_device, seed=seed, tuning={"lpe": lpe} if lpe else {})
    plan = eng.step_plan(T)
    assert not plan["fast_onehot"] and plan["planes_staged"], plan
    buf = _buffer(spec2, N, T, cuda_device)
    co = coracle.COracle(spec2, N, capacity=T, rng=coracle.RNG_PHILOX, seed=seed, threads=4)
    eng.reset()
    eng.step(T, buffer=buf)
    eng.check_errors()
    co.reset()
    co.step(T)
    np.testing.assert_array_equal(eng.grid_numpy(), co.grid)
    ring = co.ring()
    np.testing.assert_array_equal(buf.states.cpu().numpy(), ring["states"])
    np.testing.assert_array_equal(buf.rewards.cpu().numpy(), ring["rewards"])
    np.testing.assert_array_equal(eng.total_reward.cpu().numpy(), co.total_reward)
    assert buf.states.sum() > 0
