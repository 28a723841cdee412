"""GPU: the reference-shaped API (Environment / Gridworld / Agent / ObservationSpec / Buffer) on the
treasurehunt example, against the oracle."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import sb_testutil as util
from oracle import restatement as R

pytestmark = pytest.mark.gpu

CONFIG = {
    "experiment": {"epochs": 1, "max_turns": 30, "record_period": 50},
    "model": {"agent_vision_radius": 2, "epsilon_decay": 0.0005, "observation_spec": "onehot", "num_agents": 3,
              "memory_size": 64, "n_frames": 3},
    "world": {"height": 11, "width": 13, "gem_value": 10, "food_value": 5, "bone_value": -10, "spawn_prob": 0.03},
}


def _make(num_envs, seed, model_factory=None, **over):
    from sorrel_b200.examples.treasurehunt.entities import EmptyEntity
    from sorrel_b200.examples.treasurehunt.env import TreasurehuntEnv
    from sorrel_b200.examples.treasurehunt.world import TreasurehuntWorld

    cfg = {k: dict(v) for k, v in CONFIG.items()}
    for k, v in over.items():
        cfg["model"][k] = v
    world = TreasurehuntWorld(cfg, EmptyEntity(), num_envs=num_envs, device="cuda:0", seed=seed)
    return TreasurehuntEnv(world, cfg, model_factory=model_factory), cfg


def _oracle_spec():
    w, m = CONFIG["world"], CONFIG["model"]
    return R.treasurehunt_spec(w["height"], w["width"], m["num_agents"], m["agent_vision_radius"], w["spawn_prob"])


def test_example_env_fused_matches_oracle(cuda_device):
    """treasurehunt with RandomModel agents: take_turn()/take_turns() == oracle with the same stream."""
    N, seed, T = 24, 77, 20
    env, _ = _make(N, seed)
    env.add_empty()
    env.take_turn()
    env.take_turns(T - 1)
    env.engine.check_errors()
    spec = _oracle_spec()
    grids = env.engine.grid_numpy()
    for n in range(N):
        o = R.OracleEnv(spec, R.PhiloxDraws(seed), env_id=n, capacity=64, n_frames=3)
        o.reset()
        for b in o.buffers:
            b.add_empty()
        for _ in range(T):
            o.take_turn()
        np.testing.assert_array_equal(grids[n], o.grid)
        for a in range(3):
            mem = env.agents[a].model.memory
            np.testing.assert_array_equal(mem.states[n].cpu().numpy(), o.buffers[a].states)
            np.testing.assert_array_equal(mem.rewards[:, n].cpu().numpy(), o.buffers[a].rewards)
            np.testing.assert_array_equal(mem.current_state()[n].cpu().numpy(), o.buffers[a].current_state())
        assert float(env.world.total_reward[n]) == o.total_reward
    # world.observe(env=...) reads a device instance back as entity prototypes
    yx = env.engine.agent_yx()[0, 0]
    assert env.world.observe((int(yx[0]), int(yx[1]), 1), env=0).kind == "TreasurehuntAgent"
    s = env.reward_summary()
    assert s["count"] == N and abs(s["mean"] - float(env.world.total_reward.mean())) < 1e-9


def test_split_phase_host_policy_matches_oracle(cuda_device):
    """A policy that runs on the host side of the launch boundary (here: a deterministic function of
    the observation) goes through Agent.transition = pov -> get_action -> act per agent slot, and
    reproduces the oracle stepping with the same observation-dependent actions."""
    from sorrel_b200.models import BaseModel

    class ObsPolicy(BaseModel):
        """action = (number of non-zero entries in the stacked frames) mod 4"""

        def take_action(self, state):
            return (state != 0).sum(dim=1).to(torch.int32) % 4

    N, seed, T = 10, 5, 12
    env, _ = _make(N, seed, model_factory=lambda inp, na: ObsPolicy(inp, na, memory_size=64, n_frames=3))
    assert not env._fusable
    env.add_empty()
    for _ in range(T):
        env.take_turn()
    env.engine.check_errors()
    spec = _oracle_spec()
    grids = env.engine.grid_numpy()

    class Draws(R.PhiloxDraws):
        def action(self, n_actions, env_id, tick, agent):
            buf = self.env.buffers[agent]
            obs = R.observe(self.env.spec, self.env.grid, *self.env.pos[agent]).reshape(-1).astype(np.float32)
            stacked = np.concatenate((buf.current_state().reshape(-1), obs))
            return int((stacked != 0).sum() % 4)

    for n in range(N):
        d = Draws(seed)
        o = R.OracleEnv(spec, d, env_id=n, capacity=64, n_frames=3)
        d.env = o
        o.reset()
        for b in o.buffers:
            b.add_empty()
        for _ in range(T):
            o.take_turn()
        np.testing.assert_array_equal(grids[n], o.grid, err_msg=f"instance {n}")
        for a in range(3):
            mem = env.agents[a].model.memory
            np.testing.assert_array_equal(mem.states[n].cpu().numpy(), o.buffers[a].states)
            np.testing.assert_array_equal(mem.actions[:, n].cpu().numpy(), o.buffers[a].actions)
            np.testing.assert_array_equal(mem.rewards[:, n].cpu().numpy(), o.buffers[a].rewards)
            np.testing.assert_array_equal(env.buffer.extra_data["positions"][:, n * 3 + a].cpu().numpy(),
                                          o.buffers[a].positions)
        assert float(env.world.total_reward[n]) == o.total_reward


def test_run_experiment_and_observe_from(cuda_device):
    env, cfg = _make(16, 3)
    hist = env.run_experiment(logging=False)
    assert len(hist) == cfg["experiment"]["epochs"] + 1 and hist[0]["count"] == 16
    # ObservationSpec.observe(world, location): the crop from an arbitrary cell in every instance
    spec = env.agents[0].observation_spec
    obs = spec.observe(env.world, (1, 1, 1)).cpu().numpy()
    full = env.engine.grid_numpy()
    for n in range(16):
        ref = R.observe(_oracle_spec(), full[n], 1, 1).astype(np.float32)
        np.testing.assert_array_equal(obs[n], ref)
    allobs = spec.observe(env.world)
    assert tuple(allobs.shape) == (16, 3, 6, 5, 5)
    # sample() through an agent's memory view: shapes/dtypes of Buffer.sample (buffers.py:98-124)
    out = env.agents[1].model.memory.sample(8)
    assert [tuple(o.shape) for o in out] == [(8, 450), (8, 1), (8, 1), (8, 450), (8, 1), (8, 1)]
    assert [o.dtype for o in out] == [torch.float32, torch.int64, torch.float32, torch.float32, torch.float32,
                                      torch.float64]


# ------------------------------------------------------------------------------------------------
# examples/cleanup through the reference-shaped API
# ------------------------------------------------------------------------------------------------
CLEANUP_CONFIG = {
    "experiment": {"name": "cleanup", "epochs": 1, "max_turns": 24, "epsilon_decay": 0.0001},
    "env": {"height": 13, "width": 17, "layers": 3, "default_object": "EmptyEntity", "full_mdp": False,
            "pollution_threshold": 0.5, "initial_apples": 6, "apple_spawn_chance": 0.03,
            "pollution_spawn_chance": 0.06, "mode": "DEFAULT"},
    "model": {"iqn": {"parameters": {"memory_size": 32}}, "epsilon_decay": 0.0001},
    "agent": {"agent": {"num": 4, "beam_radius": 2, "obs": {"channels": 8, "vision": 3, "n_frames": 2,
                                                             "embeddings": 3}}},
}


def _make_cleanup(num_envs, seed, model_factory=None):
    from sorrel_b200.examples.cleanup.entities import EmptyEntity
    from sorrel_b200.examples.cleanup.env import CleanupEnv
    from sorrel_b200.examples.cleanup.world import CleanupWorld

    world = CleanupWorld(CLEANUP_CONFIG, EmptyEntity(), num_envs=num_envs, device="cuda:0", seed=seed)
    return CleanupEnv(world, CLEANUP_CONFIG, model_factory=model_factory)


def _cleanup_oracle_spec():
    e, a = CLEANUP_CONFIG["env"], CLEANUP_CONFIG["agent"]["agent"]
    return R.cleanup_spec(e["height"], e["width"], a["num"], a["obs"]["vision"], e["initial_apples"],
                          e["apple_spawn_chance"], e["pollution_spawn_chance"], a["beam_radius"])


def _check_cleanup_against(env, oracles, A, check_actions=False):
    grids = env.engine.grid_numpy()
    for n, o in oracles.items():
        np.testing.assert_array_equal(grids[n], o.grid, err_msg=f"instance {n}")
        np.testing.assert_array_equal(env.engine.agent_facing.cpu().numpy()[n], o.facing)
        for a in range(A):
            mem = env.agents[a].model.memory
            np.testing.assert_array_equal(mem.states[n].cpu().numpy(), o.buffers[a].states)
            np.testing.assert_array_equal(mem.rewards[:, n].cpu().numpy(), o.buffers[a].rewards)
            if check_actions:
                np.testing.assert_array_equal(mem.actions[:, n].cpu().numpy(), o.buffers[a].actions)
        assert float(env.world.total_reward[n]) == o.total_reward


def test_cleanup_example_env_fused_matches_oracle(cuda_device):
    """CleanupEnv with RandomModel agents (in-kernel policy): the rule table compiled from the example's
    entity classes, stepped fused on the device, equals the oracle on the same counter-based stream."""
    N, seed, T, A = 12, 31, 24, 4
    env = _make_cleanup(N, seed)
    assert env._fusable
    env.add_empty()
    env.take_turn()
    env.take_turns(T - 1)
    env.engine.check_errors()
    spec = _cleanup_oracle_spec()
    oracles = {}
    for n in range(0, N, 5):
        o = R.OracleEnv(spec, R.PhiloxDraws(seed), env_id=n, capacity=32, n_frames=2)
        o.reset()
        for b in o.buffers:
            b.add_empty()
        for _ in range(T):
            o.take_turn()
        oracles[n] = o
    _check_cleanup_against(env, oracles, A, check_actions=True)
    assert 0.0 <= env.world.measure_pollution(0) <= 1.0
    obs = env.agents[0].observation_spec.observe(env.world)
    assert tuple(obs.shape) == (N, A, spec.obs_size)


def test_cleanup_split_phase_host_policy_matches_oracle(cuda_device):
    """Host-side policy through Agent.transition (pov -> get_action -> act) on the cleanup rule set: beams,
    facings and all-layer rewards in split-phase launches equal the oracle with the same policy."""
    from sorrel_b200.models import BaseModel

    class ObsPolicy(BaseModel):
        def take_action(self, state):
            return (state != 0).sum(dim=1).to(torch.int32) % 6

    N, seed, T, A = 6, 8, 14, 4
    env = _make_cleanup(N, seed, model_factory=lambda inp, na: ObsPolicy(inp, na, memory_size=32, n_frames=2))
    assert not env._fusable
    env.add_empty()
    for _ in range(T):
        env.take_turn()
    env.engine.check_errors()
    spec = _cleanup_oracle_spec()

    class Draws(R.PhiloxDraws):
        def action(self, n_actions, env_id, tick, agent):
            o, buf = self.env, self.env.buffers[agent]
            y, x = o.pos[agent]
            obs = np.concatenate((R.observe(o.spec, o.grid, y, x).reshape(-1), o.spec.obs_extra[y, x]))
            stacked = np.concatenate((buf.current_state().reshape(-1), obs.astype(np.float32)))
            return int((stacked != 0).sum() % 6)

    oracles = {}
    for n in range(N):
        d = Draws(seed)
        o = R.OracleEnv(spec, d, env_id=n, capacity=32, n_frames=2)
        d.env = o
        o.reset()
        for b in o.buffers:
            b.add_empty()
        for _ in range(T):
            o.take_turn()
        oracles[n] = o
    _check_cleanup_against(env, oracles, A, check_actions=True)


def test_generate_memories_saved_games_format(cuda_device, tmp_path):
    """Environment.generate_memories (environment.py:213-300): agent{i}.npz in the SavedGames layout, games back to
    back; game g of agent i is instance g's episode (checked against the oracle on the same stream)."""
    N, seed, games = 8, 123, 11  # 11 games from 8 instances: 2 episodes, the second one partly kept
    env, cfg = _make(N, seed)
    T = cfg["experiment"]["max_turns"]
    paths = env.generate_memories(num_games=games, output_dir=tmp_path)
    assert [p.name for p in paths] == ["agent0.npz", "agent1.npz", "agent2.npz"]
    spec = _oracle_spec()
    oracles = []
    for n in (0, 5):
        o = R.OracleEnv(spec, R.PhiloxDraws(seed), env_id=n, capacity=T)
        # the constructor of Environment resets once (environment.py:52-54), generate_memories resets again
        o.reset()
        o.reset()
        for _ in range(T):
            o.take_turn()
        oracles.append((n, o))
    for a, path in enumerate(paths):
        with np.load(path) as d:
            assert d["states"].shape == (games * T, 150) and d["states"].dtype == np.float32
            assert d["actions"].shape == (games * T,) and d["actions"].dtype == np.int64
            assert d["rewards"].dtype == np.float32 and d["dones"].dtype == np.float32
            assert d["positions"].shape == (games * T, 2) and int(d["n_frames"]) == 3
            for n, o in oracles:
                sl = slice(n * T, (n + 1) * T)
                np.testing.assert_array_equal(d["states"][sl], o.buffers[a].states)
                np.testing.assert_array_equal(d["actions"][sl], o.buffers[a].actions)
                np.testing.assert_array_equal(d["rewards"][sl], o.buffers[a].rewards)
                np.testing.assert_array_equal(d["positions"][sl], o.buffers[a].positions)


def test_iowa_example_env_fused_matches_oracle(cuda_device):
    """GamblingEnv (examples/iowa) through the reference-shaped API: value states compiled from `ValueRedraw`,
    float rewards in the ring, float64 totals -- equal to the oracle on the same counter-based stream."""
    from sorrel_b200.examples.iowa.entities import EmptyEntity
    from sorrel_b200.examples.iowa.env import GamblingEnv
    from sorrel_b200.examples.iowa.world import GamblingWorld

    cfg = {"experiment": {"epochs": 1, "max_turns": 30}, "world": {"height": 12, "width": 14, "spawn_prob": 0.05},
           "model": {"agent_vision_radius": 2, "parameters": {"memory_size": 32, "n_frames": 1}}}
    N, seed, T = 20, 9, 30
    world = GamblingWorld(cfg, EmptyEntity(), num_envs=N, device="cuda:0", seed=seed)
    env = GamblingEnv(world, cfg)
    env.take_turns(T)
    env.engine.check_errors()
    spec = R.iowa_spec(12, 14, 2, 2, 0.05)
    grids = env.engine.grid_numpy()
    seen = set()
    for n in range(N):
        o = R.OracleEnv(spec, R.PhiloxDraws(seed), env_id=n, capacity=32)
        o.reset()
        for _ in range(T):
            o.take_turn()
        np.testing.assert_array_equal(grids[n], o.grid)
        assert float(env.world.total_reward[n]) == o.total_reward
        for a in range(2):
            mem = env.agents[a].model.memory
            np.testing.assert_array_equal(mem.states[n].cpu().numpy(), o.buffers[a].states)
            np.testing.assert_array_equal(mem.rewards[:, n].cpu().numpy(), o.buffers[a].rewards)
            seen.update(np.unique(o.buffers[a].rewards).tolist())
    assert len(seen) >= 5  # several deck value states were collected


def test_tag_example_env_matches_oracle(cuda_device):
    """TagEnv (examples/tag) on the base Gridworld through the reference-shaped API, fused and split-phase."""
    from sorrel_b200.entities import EmptyEntity
    from sorrel_b200.examples.tag.env import TagEnv
    from sorrel_b200.models import BaseModel
    from sorrel_b200.worlds import Gridworld

    cfg = {"experiment": {"epochs": 1, "max_turns": 20}, "model": {"memory_size": 32, "n_frames": 1},
           "agent": {"num_agents": 5, "vision_radius": 3, "it_agent": 1}, "world": {"height": 9, "width": 10, "layers": 1}}
    N, seed, T = 16, 44, 20
    spec = R.tag_spec(9, 10, 5, 3, 10, it_agent=1)

    def oracle(n, draws):
        o = R.OracleEnv(spec, draws, env_id=n, capacity=32)
        if hasattr(draws, "env"):
            draws.env = o
        o.reset()
        for _ in range(T):
            o.take_turn()
        return o

    # fused: in-kernel uniform policy
    env = TagEnv(Gridworld(**cfg["world"], default_entity=EmptyEntity(), num_envs=N, device="cuda:0", seed=seed), cfg)
    env.take_turns(T)
    env.engine.check_errors()
    grids = env.engine.grid_numpy()
    for n in range(0, N, 3):
        o = oracle(n, R.PhiloxDraws(seed))
        np.testing.assert_array_equal(grids[n], o.grid)
        np.testing.assert_array_equal(env.engine.agent_flag.cpu().numpy()[n], o.flags)
        assert float(env.world.total_reward[n]) == o.total_reward
        for a in range(5):
            np.testing.assert_array_equal(env.agents[a].model.memory.states[n].cpu().numpy(), o.buffers[a].states)

    # split-phase: a host policy that looks at the observation (its last entry is the agent's own `it` flag)
    class FlagPolicy(BaseModel):
        def take_action(self, state):
            return ((state != 0).sum(dim=1) + 2 * state[:, -1].to(torch.int64)).to(torch.int32) % 4

    class Draws(R.PhiloxDraws):
        env = None

        def action(self, n_actions, env_id, tick, agent):
            o = self.env
            y, x = o.pos[agent]
            obs = np.concatenate((R.observe(o.spec, o.grid, y, x).reshape(-1), [float(o.flags[agent])]))
            return int(((obs != 0).sum() + 2 * int(o.flags[agent])) % 4)

    env = TagEnv(Gridworld(**cfg["world"], default_entity=EmptyEntity(), num_envs=N, device="cuda:0", seed=seed), cfg,
                 model_factory=lambda inp, na: FlagPolicy(inp, na, memory_size=32, n_frames=1))
    assert not env._fusable
    for _ in range(T):
        env.take_turn()
    env.engine.check_errors()
    grids = env.engine.grid_numpy()
    for n in range(0, N, 5):
        o = oracle(n, Draws(seed))
        np.testing.assert_array_equal(grids[n], o.grid, err_msg=f"instance {n}")
        assert float(env.world.total_reward[n]) == o.total_reward
        for a in range(5):
            mem = env.agents[a].model.memory
            np.testing.assert_array_equal(mem.actions[:, n].cpu().numpy(), o.buffers[a].actions)
            np.testing.assert_array_equal(mem.rewards[:, n].cpu().numpy(), o.buffers[a].rewards)


def test_run_experiment_animate_writes_gif(cuda_device, tmp_path):
    """ImageRenderer readback hook (utils/visualization.py:227-273): one frame per turn of a recorded epoch, painted
    from a single device instance."""
    from PIL import Image

    from sorrel_b200.visualization import render_instance

    env, cfg = _make(8, 3)
    cfg["experiment"]["record_period"] = 1
    env.config.experiment.record_period = 1
    env.run_experiment(animate=True, logging=False, output_dir=tmp_path)
    gifs = sorted((tmp_path / "gifs").glob("*.gif"))
    assert len(gifs) == cfg["experiment"]["epochs"] + 1
    with Image.open(gifs[0]) as im:
        assert im.n_frames == cfg["experiment"]["max_turns"] and im.size == (13 * 16, 11 * 16)
    frame = render_instance(env.world, env=2)
    y, x = env.engine.agent_yx()[2, 0]
    spec = env.engine.spec
    assert frame.shape == (11 * 16, 13 * 16, 3) and frame[y * 16 + 8, x * 16 + 8].any()  # the agent's tile is painted
    assert not frame[5 * 16 + 8, 0 * 16 + 8].tolist() == [0, 0, 0]  # the wall border is visible
    assert spec.kind_names[int(env.engine.grid_instance(2)[y, x, 1])] == "TreasurehuntAgent"


def test_buffer_add_from_buffer(cuda_device):
    """Buffer.add_from_buffer (buffers.py:71-96) against the reference semantics restated on numpy: appended at the
    head, truncated at the end of the ring, head not wrapped, extra_data copied."""
    from sorrel_b200.buffers import Buffer

    rows, obs = 3, 7
    src = Buffer(6, (obs,), rows=rows, device=cuda_device, positions=(2,))
    dst = Buffer(10, (obs,), rows=rows, device=cuda_device, positions=(2,))
    rng = np.random.RandomState(0)
    for t in range(4):
        src.add(torch.from_numpy(rng.rand(rows, obs).astype(np.float32)), torch.from_numpy(rng.randint(0, 4, rows)),
                torch.from_numpy(rng.rand(rows).astype(np.float32)), torch.zeros(rows),
                positions=torch.from_numpy(rng.randint(0, 9, (rows, 2))))
    dst.idx = 8  # only 2 of the 4 experiences fit
    dst.add_from_buffer(src)
    assert dst.idx == 10
    np.testing.assert_array_equal(dst.states[:, 8:10].cpu().numpy(), src.states[:, :2].cpu().numpy())
    np.testing.assert_array_equal(dst.actions[8:10].cpu().numpy(), src.actions[:2].cpu().numpy())
    np.testing.assert_array_equal(dst.extra_data["positions"][8:10].cpu().numpy(),
                                  src.extra_data["positions"][:2].cpu().numpy())
    assert not dst.states[:, :8].any()
    with pytest.raises(AssertionError):
        dst.add_from_buffer(Buffer(4, (obs + 1,), rows=rows, device=cuda_device))


def test_ring_layouts_agree_across_the_buffer_api(cuda_device):
    """`Buffer(layout="slot")` (a turn's rows contiguous in HBM: what Environment picks for multi-agent worlds) and the
    default agent-major layout are the same buffer to every caller: fused step writes, Buffer.add, current_state (with
    ring wrap), sample / sample_transformer gathers on the same indices, add_from_buffer, save / load."""
    import tempfile

    from sorrel_b200.buffers import Buffer
    from sorrel_b200.engine import Engine

    spec = R.treasurehunt_spec(21, 21, 3, 2, 0.02)
    N, cap, F, seed = 37, 12, 4, 99
    bufs = {}
    for layout in ("agent", "slot"):
        eng = Engine(util.engine_spec(spec), N, device=cuda_device, seed=seed)
        buf = Buffer(cap, (spec.obs_size,), n_frames=F, rows=N * 3, device=cuda_device, layout=layout, positions=(2,))
        assert tuple(buf.states.shape) == (N * 3, cap, spec.obs_size)
        eng.reset()
        buf.add_empty()
        eng.step(7, buffer=buf)
        eng.step(9, buffer=buf)  # wraps the 12-deep ring
        obs = eng.observe().reshape(N * 3, -1)
        buf.add(obs, torch.arange(N * 3, device=cuda_device), torch.ones(N * 3, device=cuda_device), 0.0,
                positions=torch.zeros((N * 3, 2), dtype=torch.int64, device=cuda_device))
        eng.check_errors()
        bufs[layout] = buf
    a, s = bufs["agent"], bufs["slot"]
    assert s.states.stride(1) == N * 3 * spec.obs_size and a.states.stride(1) == spec.obs_size
    assert (a.idx, a.size) == (s.idx, s.size)
    assert torch.equal(a.states, s.states) and torch.equal(a.actions, s.actions) and torch.equal(a.rewards, s.rewards)
    # and both equal the oracle's ring (reset, add_empty, 16 turns; the extra add() above only touched slot 7)
    o = R.OracleEnv(spec, R.PhiloxDraws(seed), env_id=5, capacity=cap, n_frames=F)
    o.reset()
    for b in o.buffers:
        b.add_empty()
    for _ in range(16):
        o.take_turn()
    keep = [i for i in range(cap) if i != 7]
    for k in range(3):
        np.testing.assert_array_equal(s.row(5 * 3 + k)["states"][keep], o.buffers[k].states[keep])
    for idx in (2, 7):  # current_state with and without wrap
        a.idx = s.idx = idx
        assert torch.equal(a.current_state(), s.current_state())
    rows = torch.randint(0, N * 3, (64,), device=cuda_device)
    starts = torch.randint(0, cap - F - 1, (64,), device=cuda_device)
    for x, y in zip(a.gather(rows, starts), s.gather(rows, starts)):
        assert torch.equal(x, y)
    for x, y in zip(a.gather_transformer(rows, starts), s.gather_transformer(rows, starts)):
        assert torch.equal(x, y)
    a.check_errors()
    s.check_errors()
    big = Buffer(3 * cap, (spec.obs_size,), n_frames=F, rows=N * 3, device=cuda_device, layout="slot", positions=(2,))
    big.add_from_buffer(a)
    assert torch.equal(big.states[:, : a.size], a.states[:, : a.size])
    with tempfile.TemporaryDirectory() as d:
        s.save(f"{d}/b.npz", row=4)
        back = Buffer.load(f"{d}/b.npz", device=cuda_device)
        assert torch.equal(back.states[0], a.states[4])
    with pytest.raises(ValueError):
        Buffer(4, (6,), rows=2, device=cuda_device, layout="diagonal")


@pytest.mark.parametrize("A,H,lpe", [(1, 21, 4), (3, 21, None), (20, 40, 128)])
def test_uint8_action_tapes_equal_int32(A, H, lpe, cuda_device):
    """Action tapes may be uint8 (ActionSpec sizes <= 8): same step as the int32 tape, a quarter of the bytes; eager
    launches, the host pipeline and its CUDA-graph form."""
    from sorrel_b200.buffers import Buffer
    from sorrel_b200.engine import Engine
    from sorrel_b200.hostio import HostStepPipeline

    spec = R.treasurehunt_spec(H, H, A, 2, 0.02)
    N, T = 33, 6
    acts = torch.randint(0, 4, (3, T, N, A), dtype=torch.int32)
    outs = {}
    for mode in ("i32", "u8", "pipe", "pipe_graph"):
        eng = Engine(util.engine_spec(spec), N, device=cuda_device, seed=5, tuning={"lpe": lpe} if lpe else None)
        buf = Buffer(3 * T, (spec.obs_size,), rows=N * A, device=cuda_device, positions=(2,))
        eng.reset()
        rew = []
        if mode in ("i32", "u8"):
            for k in range(3):
                tape = acts[k].to(cuda_device) if mode == "i32" else acts[k].to(torch.uint8).to(cuda_device)
                r = torch.zeros((T, N, A), dtype=torch.float32, device=cuda_device)
                eng.step(T, buffer=buf, actions=tape, reward_out=r)
                rew.append(r.cpu())
        else:
            eng.step(T, buffer=buf, actions=acts[0].to(cuda_device))  # (kernels loaded before any capture)
            rew.append(None)
            pipe = HostStepPipeline(eng, buf, T, action_dtype=torch.uint8, graphs=mode == "pipe_graph")
            for k in (1, 2):
                h = pipe.submit(acts[k].to(torch.uint8).pin_memory())
                pipe.drain()
                rew.append(h.clone())
        eng.check_errors()
        outs[mode] = (eng.grid_numpy(), buf.states.cpu(), buf.actions.cpu(), rew)
    for mode in ("u8", "pipe", "pipe_graph"):
        np.testing.assert_array_equal(outs[mode][0], outs["i32"][0])
        assert torch.equal(outs[mode][1], outs["i32"][1]) and torch.equal(outs[mode][2], outs["i32"][2])
        for k in (1, 2):
            assert torch.equal(outs[mode][3][k], outs["i32"][3][k])


@pytest.mark.parametrize("per_slot", [False, True])
def test_turn_graph_equals_eager_split_phase(per_slot, cuda_device):
    """`Environment.enable_turn_graph()`: one whole split-phase turn (per agent slot: observe, frame stacking, the
    policy, act) captured as a CUDA graph with device-resident tick / ring slot -- one graph, or one per ring slot
    reading the observation straight from the ring.  Replays must equal the eager split-phase path turn for turn,
    across episode boundaries (reset, add_empty) and a ring wrap."""
    from sorrel_b200.models import BaseModel

    class ObsPolicy(BaseModel):
        """capturable: action = (number of non-zero entries in the stacked frames) mod 4"""

        def take_action(self, state):
            return (state != 0).sum(dim=1).to(torch.int32) % 4

    N, seed = 24, 17
    envs = []
    for graph in (False, True):
        env, _ = _make(N, seed, model_factory=lambda inp, na: ObsPolicy(inp, na, memory_size=16, n_frames=3))
        env.add_empty()
        for _ in range(3):
            env.take_turn()  # eager turns first: the graph starts from a non-trivial clock
        if graph:
            env.enable_turn_graph(per_slot=per_slot)  # (runs one eager turn itself, then captures)
        else:
            env.take_turn()
        for _ in range(9):
            env.take_turn()
        env.reset()
        env.add_empty()
        for _ in range(11):  # 3 + 1 + 9 + 2 (add_empty) + 11 > 16: the ring wraps
            env.take_turn()
        env.engine.check_errors()
        envs.append(env)
    e, g = envs
    assert g._turn_graph is not None
    assert (e.turn, e.engine.tick, e.buffer.idx, e.buffer.size) == (g.turn, g.engine.tick, g.buffer.idx, g.buffer.size)
    assert torch.equal(e.engine.grid, g.engine.grid) and torch.equal(e.engine.agent_pos, g.engine.agent_pos)
    assert torch.equal(e.engine.total_reward, g.engine.total_reward)
    assert torch.equal(e.buffer.states, g.buffer.states) and torch.equal(e.buffer.actions, g.buffer.actions)
    assert torch.equal(e.buffer.rewards, g.buffer.rewards)
    assert torch.equal(e.buffer.extra_data["positions"], g.buffer.extra_data["positions"])
    assert int((e.buffer.actions != 0).sum()) > 0  # the policy did act


def test_stacked_view_equals_concatenated_frames(cuda_device):
    """Agent.get_action's zero-copy frame stack (a row-strided view of the agent-major ring) holds exactly what the
    reference builds with np.vstack((prev_states, state)) (treasurehunt/agents.py:34-41); where the frames are not
    adjacent (ring wrap) the concatenating path is taken."""
    from sorrel_b200.models import BaseModel

    seen = []

    class Recorder(BaseModel):
        def take_action(self, state):
            seen.append(state)
            return torch.zeros(state.shape[0], dtype=torch.int32, device=state.device)

    N = 9
    env, _ = _make(N, 5, model_factory=lambda inp, na: Recorder(inp, na, memory_size=8, n_frames=3), num_agents=1)
    assert env.buffer.layout == "agent" and len(env.agents) == 1
    env.add_empty()
    views = 0
    for t in range(13):  # crosses the ring wrap (capacity 8)
        idx = env.buffer.idx
        seen.clear()
        env.take_turn()
        (state,) = seen
        prev = [env.buffer.states[:, (idx - 2 + f) % 8] for f in range(2)]
        want = torch.cat(prev + [env.buffer.states[:, idx]], dim=1)
        assert torch.equal(state, want), f"turn {t}, head {idx}"
        views += int(not state.is_contiguous())
    assert views >= 6  # most turns took the zero-copy path
