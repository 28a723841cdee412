"""Time the UNMODIFIED reference's own CPU path (build container only: needs /root/reference).

TEST / MEASUREMENT INFRASTRUCTURE (see oracle/__init__.py).  BASELINE.md §5: the reference publishes no numbers, so
its CPU path is timed here with the IQN swapped for the reference's own `RandomModel` (env + observation + buffer
only -- the same work the CUDA kernels do):

  1. one process, treasurehunt example shape (21x21x2, 1 agent, 5x5 one-hot view)       -> agent-env-steps/s
  2. P = os.cpu_count() independent processes, one environment each                      -> aggregate + P
  3. 128x128 map, 64 agents, 9x9 view (config 4), a few turns                            -> agent-env-steps/s
  4. Buffer.add and Buffer.sample(64) micro-timings (sorrel/buffers.py:46-63, 98-124)    -> calls/s

    python -m oracle.time_reference            # writes profiles/reference_cpu_timings.json

A Python reference cannot travel to the GPU box, so bench.py reports these recorded figures beside the C-port
baseline it times live on the box's host cores (and times the reference live whenever /root/reference exists).
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _build(H, W, A, v, p, T=100):
    from oracle import shim

    shim.install()
    import numpy as np
    from sorrel.examples.treasurehunt.entities import EmptyEntity
    from sorrel.examples.treasurehunt.env import TreasurehuntEnv
    from sorrel.examples.treasurehunt.world import TreasurehuntWorld
    from sorrel.models.base_model import RandomModel
    from sorrel.utils.helpers import set_seed

    config = {"experiment": {"epochs": 1, "max_turns": T, "record_period": 50},
              "model": {"agent_vision_radius": v, "epsilon_decay": 0.0005, "observation_spec": "onehot",
                        "num_agents": A},
              "world": {"height": H, "width": W, "gem_value": 10, "food_value": 5, "bone_value": -10,
                        "spawn_prob": p}}
    set_seed(0)
    world = TreasurehuntWorld(config=config, default_entity=EmptyEntity())
    env = TreasurehuntEnv(world, config)
    for agent in env.agents:  # SURVEY 8c harness: the reference's own RandomModel instead of the IQN
        size = int(np.prod(agent.model.input_size)) if hasattr(agent.model, "input_size") else 0
        agent.model = RandomModel(size, 4, 1024)
        agent.model.memory.extra_data["positions"] = np.zeros((1024, 2), dtype=np.int64)
        agent.get_action = (lambda state, ag=agent: ag.model.take_action(state))
    env.reset()
    return env


def single_process_rate(H=21, W=21, A=1, v=2, p=0.005, seconds=5.0, max_turns=100) -> float:
    env = _build(H, W, A, v, p, max_turns)
    done, turns, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        if turns >= max_turns:
            env.reset()
            turns = 0
        env.take_turn()
        turns += 1
        done += A
    return done / (time.perf_counter() - t0)


def _worker(args):
    return single_process_rate(*args)


def buffer_micro(seconds=2.0):
    from oracle import shim

    shim.install()
    import numpy as np
    from sorrel.buffers import Buffer

    buf = Buffer(capacity=1024, obs_shape=(150,), n_frames=5)
    obs = np.zeros(150)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        buf.add(obs, 1, 1.0, False)
        n += 1
    adds = n / (time.perf_counter() - t0)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        buf.sample(64)
        n += 1
    return adds, n / (time.perf_counter() - t0)


def main():
    P = os.cpu_count() or 1
    out = {"where": "build container (CPU only), unmodified /root/reference, RandomModel policy",
           "cpu_count": P, "unit": "agent-env-steps/s"}
    out["treasurehunt_21x21_1agent_1process"] = single_process_rate(seconds=8.0)
    with mp.get_context("spawn").Pool(P) as pool:
        rates = pool.map(_worker, [(21, 21, 1, 2, 0.005, 8.0)] * P)
    out["treasurehunt_21x21_1agent_P_processes"] = {"P": P, "aggregate": sum(rates), "per_process_min": min(rates)}
    out["treasurehunt_21x21_8agents_1process"] = single_process_rate(21, 21, 8, 2, 0.005, 6.0)
    t0 = time.perf_counter()
    env = _build(128, 128, 64, 4, 0.005)
    out["large_128x128_64agents_build_s"] = time.perf_counter() - t0
    turns, t0 = 0, time.perf_counter()
    while turns < 3:
        env.take_turn()
        turns += 1
    out["large_128x128_64agents_1process"] = 64 * turns / (time.perf_counter() - t0)
    adds, samples = buffer_micro()
    out["buffer_add_calls_per_s"], out["buffer_sample64_calls_per_s"] = adds, samples
    path = os.path.join(ROOT, "profiles", "reference_cpu_timings.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
