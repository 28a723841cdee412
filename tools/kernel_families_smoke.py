"""Small launches of every kernel family and variant (group widths, staged / unstaged planes, agent-parallel path,
general observation path, cleanup / iowa / tag rule sets) with the device error word checked after each -- a quick
functional smoke on a GPU box, also suitable as the workload of `compute-sanitizer` where that tool is available
(it is closed on the build pool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import sb_testutil as util  # noqa: E402
from oracle import restatement as R  # noqa: E402
from sorrel_b200.buffers import Buffer  # noqa: E402
from sorrel_b200.engine import Engine  # noqa: E402

dev = torch.device("cuda:0")


def run(name, spec, N, T, rng="philox", tuning=None):
    eng = Engine(util.engine_spec(spec), N, device=dev, seed=3, tuning=tuning)
    buf = Buffer(8, (spec.obs_size,), n_frames=2, rows=N * spec.num_agents, device=dev, positions=(2,))
    eng.reset()
    eng.step(T, buffer=buf, rng=rng)
    eng.step(1, buffer=buf, rng=rng)
    eng.observe()
    rows = torch.randint(0, N * spec.num_agents, (16,), device=dev)
    starts = torch.randint(0, 3, (16,), device=dev)
    buf.gather(rows, starts)
    buf.current_state()
    eng.check_errors()
    torch.cuda.synchronize()
    print("ok", name, eng.step_plan(T, rng=rng)["kernel"], flush=True)


run("treasurehunt lpe4", R.treasurehunt_spec(), 37, 5)
run("treasurehunt sparse", R.treasurehunt_spec(), 37, 5, rng="sparse")
run("treasurehunt 8 agents lpe8", R.treasurehunt_spec(21, 21, 8, 2, 0.01), 21, 4)
run("treasurehunt rgb (general path)", R.treasurehunt_spec(11, 11, 2, 2, 0.05, obs="rgb"), 9, 4)
run("64x64 lpe32 par", R.treasurehunt_spec(64, 64, 16, 3, 0.01), 5, 4, tuning={"lpe": 32})
run("64x64 unstaged par", R.treasurehunt_spec(64, 64, 16, 3, 0.01), 5, 4, tuning={"lpe": 32, "staged": -1})
run("64x64 block per instance", R.treasurehunt_spec(64, 64, 16, 3, 0.01), 5, 4, tuning={"lpe": 128})
run("128x128 config 4", R.treasurehunt_spec(128, 128, 64, 4, 0.005), 3, 2)
run("torus", R.torus_spec(), 9, 4)
run("cleanup", R.cleanup_spec(), 7, 4)
run("iowa", R.iowa_spec(), 9, 4)
run("tag", R.tag_spec(), 11, 4)
print("kernel families smoke done")
