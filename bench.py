#!/usr/bin/env python
"""bench.py -- agent-env-steps/s (including observations) of the batched gridworld engine.

    python bench.py --gpus N --steps K --warmup W             # this engine (one rank per GPU), BASELINE configs[4]
    python bench.py --config 2|3|4|4obs|4obs19|5|cleanup ...   # the other BASELINE shapes, same line schema
    python bench.py --impl reference --gpus N --steps K ...    # CPU arm: the oracle port on the host cores

Default workload (BASELINE.json configs[4], the one the metric "at 1/2/4/8 B200" is quoted on): 2^20 treasurehunt
worlds in total (21x21x2, 1 agent, 5x5 one-hot view, C=6, spawn_prob 0.005, episodes of 100 turns), sharded over the
ranks with no collective on the step path.  The GPU arm is driven through the product API only --
`sorrel_b200.examples.treasurehunt.{TreasurehuntWorld, TreasurehuntEnv}` and `Environment.begin_epoch / take_turns /
sample_device` -- exactly what `Environment.run_experiment` calls.  One *step* = `take_turns(T)` (one launch of T fused
turns for every instance of the rank: entity sweep, observation written straight into the replay ring, move, reward,
Buffer.add) followed by one Buffer.sample gather of T*N*A/64 transitions (5 stacked frames); every 100 turns
`begin_epoch()` reduces the episode's reward statistics (the only collective), resets the worlds and `add_empty`s the
ring.  `value` times that with everything resident in HBM; `e2e` drives the same step from pinned HOST action buffers
and reads the rewards back to the host every step (`sorrel_b200.hostio.HostStepPipeline`); observations stay in the
device ring by design (they are the learner's input), so they are not part of the host traffic.

`oracle/` (the CPU restatement) is imported by the `cpu_baseline` leg and by `--impl reference` only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "agent_env_steps_per_sec_incl_obs"
UNIT = "agent-env-steps/s"

# BASELINE.json configs[1..4] (+ the cleanup example): argument presets
PRESETS = {
    "2": dict(envs_total=4096, turns=100),
    "3": dict(envs_total=16384, agents=8, turns=25),
    "4": dict(envs_total=2048, agents=64, size=128, vision=4, turns=10, capacity=16),
    "4obs": dict(obs_study=True, envs_total=4096, agents=64, size=128, vision=4),
    "4obs19": dict(obs_study=True, envs_total=4096, agents=64, size=128, vision=9),
    "5": dict(),
    "cleanup": dict(world="cleanup", agents=10, vision=5, envs_total=32768, turns=25, capacity=8),
    # live (on-GPU, not in-kernel) policy: one split-phase turn per take_turn(), eager launches vs the turn graph
    "live": dict(live=True, turns=1, n_frames=1),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(PRESETS),
                    help="BASELINE.json shape preset (explicit flags given after it still override)")
    ap.add_argument("--envs-total", type=int, default=1 << 20, help="world instances over all ranks")
    ap.add_argument("--turns", type=int, default=25, help="fused turns per launch (one step)")
    ap.add_argument("--agents", type=int, default=1)
    ap.add_argument("--size", type=int, default=21)
    ap.add_argument("--vision", type=int, default=2)
    ap.add_argument("--spawn-prob", type=float, default=0.005)
    ap.add_argument("--capacity", type=int, default=32, help="replay ring rows per agent")
    ap.add_argument("--n-frames", type=int, default=5)
    ap.add_argument("--max-turns", type=int, default=100)
    ap.add_argument("--sample-div", type=int, default=64, help="sample 1 transition per this many agent-steps; 0 = off")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: do not sample clocks")
    ap.add_argument("--rng", default="philox", choices=["philox", "sparse"],
                    help="counter-based spawn sampling: philox = one random byte per cell per turn (dense, like the "
                         "reference's draw-per-cell loop); sparse = Binomial count + distinct cells (same law)")
    ap.add_argument("--no-alt", action="store_true", help="skip the extra sparse-sampler timing")
    ap.add_argument("--no-graphs", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--world", default="treasurehunt", choices=["treasurehunt", "cleanup"],
                    help="rule set: treasurehunt (BASELINE configs) or the reference's cleanup example (SURVEY 8f rank 3: "
                         "21x31x3, 10 agents, 11x11 one-hot view + positional code, beams, column rules)")
    ap.add_argument("--obs-study", action="store_true",
                    help="config 4: time the standalone observation kernel (sb_observe) against the HBM roofline")
    ap.add_argument("--live", action="store_true",
                    help="live-policy cadence: Environment.take_turn() through the split-phase launches with a device "
                         "policy between observe and act, eager and as one CUDA graph per turn")
    ap.add_argument("--actions-i32", action="store_true", help="e2e: ship int32 actions instead of uint8")
    ap.add_argument("--no-numa", action="store_true", help="diagnostic: do not bind ranks to their GPU's NUMA node")
    ap.add_argument("--lpe", type=int, default=0, help="diagnostic: pin the lanes per instance (SbStepArgs.tune_lpe)")
    # presets first, then the command line on top of them
    pre = ap.parse_known_args()[0]
    if pre.config:
        ap.set_defaults(**PRESETS[pre.config])
    return ap.parse_args()


def workload_name(a):
    if a.world == "cleanup":
        return (f"cleanup 21x31x3, {a.agents} agent(s)/env, v={a.vision} one-hot C=9 + 12 positional floats, "
                f"{a.envs_total} envs total, {a.turns} turns/launch, ring cap {a.capacity}, F={a.n_frames}, "
                f"sample 1/{a.sample_div}")
    return (f"treasurehunt {a.size}x{a.size}x2, {a.agents} agent(s)/env, v={a.vision} one-hot C=6, "
            f"{a.envs_total} envs total, {a.turns} turns/launch, ring cap {a.capacity}, F={a.n_frames}, "
            f"sample 1/{a.sample_div}")


def shape_key(a, n_per_gpu: int, rng: str | None = None) -> str:
    """Key of profiles/traffic.json: one ncu-measured DRAM byte count per (shape, sampler)."""
    return f"{a.world}:{a.size}:{a.agents}:v{a.vision}:n{n_per_gpu}:t{a.turns}:c{a.capacity}:{rng or a.rng}"


def alg_bytes_per_env_step(a) -> int:
    """2*H*W*L_dyn + A*(C*S^2*4 + 36) + 8."""
    S = 2 * a.vision + 1
    if a.world == "cleanup":  # 3 dynamic planes of 21x31, rows of 9*S*S + 12 floats, + 8 B facing r/w per agent
        return 2 * 21 * 31 * 3 + a.agents * ((9 * S * S + 12) * 4 + 36 + 8) + 8
    return 2 * a.size * a.size * 1 + a.agents * (6 * S * S * 4 + 36) + 8


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(key: str):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get("step_kernel_dram_bytes_per_launch", {}).get(key)


# ----------------------------------------------------------------------------------------------------
# product-API construction of the world family (no oracle / tests import on this path)
# ----------------------------------------------------------------------------------------------------
def make_env(a, num_envs: int, device, env_id_base: int = 0, tuning=None):
    """The reference's example wiring with the world / env classes swapped."""
    if a.world == "cleanup":
        from sorrel_b200.examples.cleanup.entities import EmptyEntity
        from sorrel_b200.examples.cleanup.env import CleanupEnv
        from sorrel_b200.examples.cleanup.main import CONFIG
        from sorrel_b200.examples.cleanup.world import CleanupWorld
        import copy

        cfg = copy.deepcopy(CONFIG)
        cfg["experiment"]["max_turns"] = a.max_turns
        cfg["model"]["iqn"]["parameters"]["memory_size"] = a.capacity
        cfg["agent"]["agent"]["num"] = a.agents
        cfg["agent"]["agent"]["obs"].update(vision=a.vision, n_frames=a.n_frames)
        world = CleanupWorld(cfg, EmptyEntity(), num_envs=num_envs, device=device, seed=a.seed,
                             env_id_base=env_id_base)
        world.tuning = tuning
        return CleanupEnv(world, cfg)
    from sorrel_b200.examples.treasurehunt.entities import EmptyEntity
    from sorrel_b200.examples.treasurehunt.env import TreasurehuntEnv
    from sorrel_b200.examples.treasurehunt.world import TreasurehuntWorld

    cfg = {"world": dict(height=a.size, width=a.size, spawn_prob=a.spawn_prob, gem_value=10, food_value=5,
                         bone_value=-10),
           "model": dict(agent_vision_radius=a.vision, observation_spec="onehot", num_agents=a.agents,
                         memory_size=a.capacity, n_frames=a.n_frames),
           "experiment": dict(epochs=1, max_turns=a.max_turns),
           "engine": dict(turns_per_launch=a.turns)}
    world = TreasurehuntWorld(cfg, EmptyEntity(), num_envs=num_envs, device=device, seed=a.seed,
                              env_id_base=env_id_base)
    world.tuning = tuning
    return TreasurehuntEnv(world, cfg)


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C port (the reference is pure Python and cannot travel to the GPU box)
# ----------------------------------------------------------------------------------------------------
def oracle_spec(a):
    from oracle import restatement as R

    if a.world == "cleanup":
        return R.cleanup_spec(21, 31, a.agents, a.vision)
    return R.treasurehunt_spec(a.size, a.size, a.agents, a.vision, a.spawn_prob)


def cpu_port_throughput(a, seconds: float, threads: int | None = None):
    from oracle import coracle

    threads = threads or coracle.max_threads()
    spec = oracle_spec(a)
    n = max(threads * 64, 1024) if a.size <= 32 else max(threads * 2, 32)
    co = coracle.COracle(spec, n, capacity=a.capacity, n_frames=a.n_frames, rng=coracle.RNG_NUMPY, seed=a.seed,
                         threads=threads)
    co.reset()
    co.add_empty()
    co.step(a.turns)  # warm-up
    done, t0 = 0, time.perf_counter()
    turns_in_episode = a.turns
    while True:
        if turns_in_episode >= a.max_turns:
            co.reset()
            co.add_empty()
            turns_in_episode = 0
        co.step(a.turns)
        turns_in_episode += a.turns
        done += n * a.agents * a.turns
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    return {"value": done / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} envs x {done // (n * a.agents)} turns of the same workload (C oracle port, "
                      f"numpy-legacy MT19937 draws, {threads} pthreads), {dt:.1f} s"}


def python_port_throughput(a, seconds: float = 3.0):
    """The Python restatement (same loop structure as the reference's Python) on one core."""
    from oracle import restatement as R

    env = R.OracleEnv(oracle_spec(a), R.NumpyDraws(a.seed), capacity=a.capacity, n_frames=a.n_frames)
    env.reset()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        env.take_turn()
        n += a.agents
    return n / (time.perf_counter() - t0)


def reference_python_rate(a, seconds: float = 3.0):
    """The UNMODIFIED reference's own take_turn() loop on one core, when /root/reference is present (build container
    only: a Python reference cannot travel to the GPU box).  Elsewhere: the figures recorded by
    oracle/time_reference.py in the build container (profiles/reference_cpu_timings.json)."""
    recorded = None
    rec_path = os.path.join(ROOT, "profiles", "reference_cpu_timings.json")
    if os.path.exists(rec_path):
        recorded = json.load(open(rec_path))
    if a.world != "treasurehunt" or not os.path.isdir("/root/reference/sorrel"):
        return {"live": None, "note": "reference not on box (pure-Python package, build container only)",
                "recorded_in_build_container": recorded}
    try:
        from oracle import time_reference as TR

        rate = TR.single_process_rate(a.size, a.size, a.agents, a.vision, a.spawn_prob, seconds)
        return {"live": rate, "unit": UNIT, "cores": 1, "note": "sorrel TreasurehuntEnv.take_turn, RandomModel policy",
                "recorded_in_build_container": recorded}
    except Exception as exc:  # noqa: BLE001
        return {"live": None, "note": f"reference import failed: {exc!r}", "recorded_in_build_container": recorded}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    per_step = max(1.0, a.cpu_seconds / max(1, a.steps + a.warmup))
    vals = []
    for i in range(a.warmup + a.steps):
        r = cpu_port_throughput(a, per_step)
        if i >= a.warmup:
            vals.append(r["value"])
    value = sum(vals) / len(vals)
    r["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "CPU arm: oracle C port on host cores, bounded sample"},
        "cpu_baseline": r,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / throttle reasons DURING the timed region through NVML in a thread
    (in-process and cheap, so that sampling does not perturb the launches it observes)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int, period: float = 0.05):
        self.rows, self.ok, self.period, self._stop = [], False, period, threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as exc:  # noqa: BLE001
            self.err = repr(exc)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((time.perf_counter(), sm, reasons, power))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def stop(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {getattr(self, 'err', '')}"]}
        self._stop.set()
        self.thread.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        sm = sorted(r[1] for r in rows)
        bits = 0
        for r in rows:
            bits |= r[2]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm,
                "reasons": [n for b, n in self.REASONS.items() if bits & b], "samples": len(rows),
                "power_w_max": max((r[3] for r in rows), default=None)}


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1 and not a.no_numa:  # host staging buffers and the thread that fills them next to the rank's GPU
        from sorrel_b200.hostio import bind_to_gpu_numa_node

        numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert a.envs_total % world == 0
    N = a.envs_total // world
    T, A = a.turns, a.agents
    env = make_env(a, N, dev, env_id_base=rank * N, tuning={"lpe": a.lpe} if a.lpe else None)
    eng, buf = env.engine, env.buffer
    lib = eng.lib
    B = (T * N * A) // a.sample_div if a.sample_div else 0
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)

    episode_stats = []
    seg = {"boundary": [], "step": [], "sample": []}  # (event, event) pairs per segment, timed steps only

    def mark(name, record):
        if not record:
            return None
        e = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        seg[name].append(e)
        e[0].record()
        return e

    def one_step(record=False, rng=a.rng):
        """Environment.run_experiment's inner loop: epoch boundary every max_turns turns, T fused turns, one
        learner batch."""
        if env.turn >= a.max_turns:  # environment.py:128-133,191-199
            e = mark("boundary", record)
            episode_stats.append(env.begin_epoch())  # stats all_gather (40 B per rank), reset, add_empty
            if e:
                e[1].record()
        e = mark("step", record)
        env.take_turns(T, rng=rng)
        if e:
            e[1].record()
        out = None
        if B:
            e = mark("sample", record)
            out = env.sample_device(B, rng=rng, generator=gen)
            if e:
                e[1].record()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(rng):
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(a.steps):
            one_step(record=True, rng=rng)
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        return float(ts.item())

    # ---------------- device-resident throughput ("value") ----------------
    env.turn = a.max_turns  # the constructor already reset once: start the first step with a full boundary
    # warm-up: at least W steps AND one full episode boundary (stats + reset kernels get loaded)
    for _ in range(max(a.warmup, a.max_turns // T + 1)):
        one_step()
    if not a.no_graphs:  # the ring is full and every kernel has run once: capture, then warm the replays
        env.enable_graphs(T, sample_batch=B, rng=a.rng)
        for _ in range(3):
            one_step()
    plan = eng.step_plan(T, rng=a.rng)
    # NVML is initialised BEFORE the barrier: doing it after would make rank 0 enter the timed region a few
    # milliseconds late and every other rank wait for it at the first all-gather (measured: +4.6 ms)
    clocks = ClockSampler(local) if (rank == 0 and not a.no_clocks) else None
    l0 = env.kernel_launches()  # direct launches + library kernels inside replayed graphs
    t_wall0 = time.perf_counter()
    ms_total = timed(a.rng)
    t_wall1 = time.perf_counter()
    launches = env.kernel_launches() - l0
    eng.check_errors()
    kernel_ms = [e0.elapsed_time(e1) for e0, e1 in seg["step"]]
    breakdown = {k: sum(e0.elapsed_time(e1) for e0, e1 in v) / a.steps for k, v in seg.items()}
    clock_info = clocks.stop(t_wall0, t_wall1) if clocks else None
    steps_total = a.envs_total * A * T * a.steps
    value = steps_total / (ms_total * 1e-3)

    # ---------------- the same workload with the sparse spawn sampler (reported beside `value`) ----------------
    alt = None
    if a.rng == "philox" and eng.sparse_available and not a.no_alt:
        seg = {"boundary": [], "step": [], "sample": []}
        for _ in range(3):
            one_step(rng="sparse")
        if not a.no_graphs:
            env.enable_graphs(T, sample_batch=B, rng="sparse")
            for _ in range(3):
                one_step(rng="sparse")
        ms_alt = timed("sparse")
        sk = [e0.elapsed_time(e1) for e0, e1 in seg["step"]]
        k_alt = sum(sk) / len(sk)
        peak, _ = hbm_peak()
        tr = measured_traffic(shape_key(a, N, "sparse"))
        alt = {"rng": "sparse (Binomial count + distinct cells; same law as one draw per cell)",
               "value": steps_total / (ms_alt * 1e-3), "unit": UNIT, "ms_per_step": ms_alt / a.steps,
               "kernel_ms": k_alt, "frac": alg_bytes_per_env_step(a) * N * T / (k_alt * 1e-3) / 1e9 / peak,
               "traffic": tr, "frac_dram": None if tr is None else tr / (k_alt * 1e-3) / 1e9 / peak}
        eng.check_errors()

    # ---------------- end-to-end with host buffers ----------------
    e2e = None
    if not a.no_e2e:
        from sorrel_b200.hostio import HostStepPipeline

        act_dtype = torch.int32 if a.actions_i32 else torch.uint8  # action indices fit a byte (<= 8 actions)
        pipe = HostStepPipeline(eng, buf, T, rng=a.rng, timing=True, action_dtype=act_dtype)
        n_act = env.agents[0].action_spec.n_actions
        h_actions = [torch.randint(0, n_act, (T, N, A), dtype=act_dtype).pin_memory() for _ in range(2)]

        def e2e_step(i):
            """The same step as `one_step` (boundary statistics + reset, fused turns, Buffer.sample), with the
            step's actions coming from pinned host memory and its rewards read back to the host."""
            if env.turn >= a.max_turns:
                episode_stats.append(env.begin_epoch())
            pipe.submit(h_actions[i % 2])
            env.turn += T

        for i in range(max(3, a.warmup)):  # eager: loads the tape-policy path
            e2e_step(i)
        pipe.drain()
        if not a.no_graphs:  # per-slot graphs of [step, clock advance, index draw, sample] over fixed buffers
            pipe = HostStepPipeline(eng, buf, T, rng=a.rng, timing=True, graphs=True, sample_batch=B,
                                    action_dtype=act_dtype)
        else:
            pipe.B = B
        for i in range(3):
            e2e_step(i)
        pipe.drain()
        barrier()
        t0 = time.perf_counter()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(a.steps):
            e2e_step(i)
        results = pipe.drain()
        c1.record()
        barrier()
        e2e_ms = max(c0.elapsed_time(c1), 1e3 * (time.perf_counter() - t0))
        te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ms = float(te.item())
        e2e = {"value": steps_total / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": pipe.h2d_bytes * world, "d2h_bytes_per_step": pipe.d2h_bytes * world,
               "ms_per_step": e2e_ms / a.steps, "reward_checksum": float(results),
               "step_kernel_ms": sum(e0.elapsed_time(e1) for e0, e1 in pipe.kernel_events[-a.steps:]) / a.steps,
               "api": "sorrel_b200.hostio.HostStepPipeline.submit (pinned host " + ("int32" if a.actions_i32 else "uint8") +
                      " actions in, float32 host rewards out, "
                      "3 streams, 2 buffers in flight" + ("" if a.no_graphs else ", compute stage = CUDA graph") + ")",
               "note": "observations are written straight into the device replay ring (the learner's input) and "
                       "stay in HBM by design; only actions and rewards cross PCIe",
               "numa": numa}
        eng.check_errors()
        buf.check_errors()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = hbm_peak()
    bytes_per_launch = alg_bytes_per_env_step(a) * N * T
    k_ms = sum(kernel_ms) / len(kernel_ms)
    achieved = bytes_per_launch / (k_ms * 1e-3) / 1e9
    traffic = measured_traffic(shape_key(a, N))  # ncu dram bytes of ONE launch of exactly this shape, else null
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "frac_dram": None if traffic is None else traffic / (k_ms * 1e-3) / 1e9 / peak,
                "kernel": plan["kernel"], "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": peak_src,
                "step_kernel_share_of_step": k_ms * a.steps / ms_total, "breakdown_ms_per_step": breakdown,
                "note": "frac = algorithmic bytes (SURVEY 8d: grid r/w every turn) / time / peak; frac_dram = the "
                        "DRAM bytes ncu measured for one launch of this shape / time / peak (turn fusion keeps the "
                        "grid in shared memory, so it moves fewer bytes than the algorithmic count)"}

    cpu = None
    if not a.no_cpu:
        cpu = cpu_port_throughput(a, a.cpu_seconds)
        if a.size <= 32:
            cpu["python_port_1core"] = python_port_throughput(a, 2.0)
        cpu["reference_python"] = reference_python_rate(a)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(a), "envs_per_gpu": N, "parallelism": f"env-shard x{world}",
                   "l2": "inputs larger than L2 (grid + ring slots touched per step >> 126 MB)"
                         if N * A * T * eng.spec.obs_size * 4 > 256e6 else
                         "working set of a step fits L2 at this batch size (small-batch config as named)",
                   "policy": "in-kernel uniform (RandomModel)",
                   "api": "TreasurehuntWorld/TreasurehuntEnv -> Environment.begin_epoch / take_turns / sample_device"
                          if a.world == "treasurehunt" else
                          "CleanupWorld/CleanupEnv -> Environment.begin_epoch / take_turns / sample_device",
                   "launch": "eager" if a.no_graphs else "CUDA graph replay (device-resident tick/slot)",
                   "world": a.world, "variant": plan,
                   "rng": "philox4x32-10, " + ("one byte per cell per turn" if a.rng == "philox" else
                                               "sparse (Binomial count + distinct cells)")},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clock_info,
        "sparse_rng": alt,
        "episodes_reduced": len(episode_stats),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_obs_study(a):
    """BASELINE configs[3]: large map, many agents, one-hot crop written at HBM speed (sb_observe alone)."""
    import torch

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    N, A = a.envs_total, a.agents
    env = make_env(a, N, dev)
    eng, spec = env.engine, env.engine.spec
    # ~35 % of the interior pre-seeded with items (steady-state density of the reference)
    g = eng.grid[:, 0, : a.size * a.size]
    items = torch.randint(2, 5, g.shape, device=dev, dtype=torch.uint8)
    mask = (g == 0) & (torch.rand(g.shape, device=dev) < 0.35)
    g[mask] = items[mask]
    S = 2 * a.vision + 1
    obs = spec.channels * S * S
    # rotate over several output tensors so that consecutive launches do not hit lines still in L2
    n_out = max(2, int(200e6 // (N * A * obs * 4)) + 1)
    outs = [torch.empty((N, A, spec.channels, S, S), dtype=torch.float32, device=dev) for _ in range(n_out)]
    for i in range(max(3, a.warmup)):
        eng.observe(outs[i % n_out])
    torch.cuda.synchronize()
    clocks = None if a.no_clocks else ClockSampler(dev.index or 0, period=0.002)
    reps = max(a.steps, 200)  # long enough for the NVML thread to see the region (launches are 0.1-0.4 ms)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    t0 = time.perf_counter()
    for i in range(reps):
        evs[i][0].record()
        eng.observe(outs[i % n_out])
        evs[i][1].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    clock_info = clocks.stop(t0, t1) if clocks else None
    ms = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    k_ms = sum(ms) / len(ms)
    alg = N * A * obs * 4 + N * a.size * a.size + N * A * 4
    peak, peak_src = hbm_peak()
    ach = alg / (k_ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "observations_per_sec", "value": N * A / (k_ms * 1e-3), "unit": "obs/s", "n_gpus": 1,
        "steps": reps, "warmup": max(3, a.warmup), "ms_per_step": k_ms, "higher_is_better": True, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": f"obs-kernel study: {N} envs x {A} agents, {a.size}x{a.size} map, {S}x{S} one-hot "
                               f"window, C={spec.channels}, {obs * 4} B per observation, {n_out} rotating outputs",
                   "api": "TreasurehuntEnv -> Engine.observe (sb_observe)",
                   "l2": f"{n_out} rotating output tensors, {N * A * obs * 4 / 1e6:.0f} MB each"},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": None, "kernel": "sb::observe_kernel", "kernel_ms": k_ms, "kernel_ms_min": ms[0],
                     "algorithmic_bytes_per_launch": alg, "peak_source": peak_src},
        "gpu_launches": reps, "clocks": clock_info,
    }), flush=True)


def run_live(a):
    """Item 9 of the round-1 review: a policy that is NOT evaluated in-kernel (any torch module on the GPU) forces one
    turn per take_turn().  Times Environment.take_turn() with a trivial device policy, eagerly and through
    Environment.enable_turn_graph()."""
    import torch

    from sorrel_b200.models import BaseModel

    class DevicePolicy(BaseModel):
        """uniform random actions drawn by torch on the device (stands in for a network)"""

        def take_action(self, state):
            return torch.randint(0, self.action_space, (state.shape[0],), dtype=torch.int32, device=state.device)

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    N, A = a.envs_total, a.agents
    out = {}
    for mode in ("eager", "graph"):
        env = make_env(a, N, dev)
        for ag in env.agents:  # swap the in-kernel RandomModel for the device policy
            m = DevicePolicy(ag.model.input_size, ag.model.action_space, ag.model.memory_size, n_frames=a.n_frames)
            m.memory = ag.model.memory
            ag.model = m
        env._fusable = False
        env.add_empty()
        for _ in range(3):
            env.take_turn()
        if mode == "graph":
            # large batches: one graph per ring slot (the policy reads the ring row itself; no second copy of the rows)
            env.enable_turn_graph(per_slot=N * A * env.engine.spec.obs_size * 4 > 64e6)
        clocks = None if (a.no_clocks or mode == "eager") else ClockSampler(dev.index or 0, period=0.01)
        turns = max(a.steps, 100)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        done = 0
        while done < turns:
            if env.turn >= a.max_turns:
                env.begin_epoch()
            env.take_turn()
            done += 1
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms = max(e0.elapsed_time(e1), 1e3 * (t1 - t0)) / turns
        out[mode] = {"ms_per_turn": ms, "value": N * A / (ms * 1e-3), "clocks": clocks.stop(t0, t1) if clocks else None}
        env.engine.check_errors()
        del env
        torch.cuda.empty_cache()
    peak, peak_src = hbm_peak()
    alg = alg_bytes_per_env_step(a) * N
    g = out["graph"]
    best = "graph" if g["ms_per_turn"] <= out["eager"]["ms_per_turn"] else "eager"
    print(json.dumps({
        "metric": METRIC, "value": out[best]["value"], "unit": UNIT, "n_gpus": 1, "steps": max(a.steps, 100),
        "warmup": 3, "ms_per_step": g["ms_per_turn"], "higher_is_better": True, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(a) + ", live device policy (torch.randint) between observe and act",
                   "api": "Environment.take_turn() (split-phase: Agent.pov / get_action / act); value = the faster of "
                          "eager launches and one CUDA graph per turn (Environment.enable_turn_graph): " + best
                          + "; ms_per_step = the graph form, eager.ms_per_turn = the eager form",
                   "l2": "inputs larger than L2" if N * A * 600 > 256e6 else "working set fits L2 at this batch size"},
        "roofline": {"bound": "hbm", "achieved": alg / (g["ms_per_turn"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (g["ms_per_turn"] * 1e-3) / 1e9 / peak, "traffic": None,
                     "kernel": "split-phase sb::step_kernel launches + policy", "peak_source": peak_src,
                     "note": "whole turn incl. the policy's torch kernels and the second copy of every row the policy "
                             "reads"},
        "eager": out["eager"], "gpu_launches": None, "clocks": g["clocks"],
    }), flush=True)


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.live:
        run_live(a)
    elif a.obs_study:
        run_obs_study(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
