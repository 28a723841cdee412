import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
import sb_testutil as util
from oracle import restatement as R
from sorrel_b200.engine import Engine
from sorrel_b200.buffers import Buffer
dev = torch.device("cuda:0")
N, T = 1 << 20, 25
spec = util.engine_spec(R.treasurehunt_spec())
eng = Engine(spec, N, device=dev); buf = Buffer(32, (150,), n_frames=5, rows=N, device=dev, positions=(2,))
eng.reset(); buf.add_empty()
for _ in range(4): eng.step(T, buffer=buf)
torch.cuda.synchronize()
def run(K, reset_every, use_events):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    host = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(K):
        if reset_every and i % reset_every == 0:
            eng.reset(); buf.add_empty()
        h0 = time.perf_counter()
        if use_events: evs[i][0].record()
        eng.step(T, buffer=buf)
        if use_events: evs[i][1].record()
        host.append(time.perf_counter() - h0)
    torch.cuda.synchronize(); wall = time.perf_counter() - t0
    k = [a.elapsed_time(b) for a, b in evs] if use_events else [0]
    print(f"K={K} reset_every={reset_every} events={use_events}: wall/step={1e3*wall/K:.3f} ms, kernel avg={sum(k)/len(k):.3f} ms, kernels={['%.2f'%x for x in k[:8]]}, host/step max={1e3*max(host):.3f} ms")
run(8, 0, True); run(8, 0, False); run(8, 4, True); run(8, 1, True); run(16, 4, True)
# slot pattern: does the time depend on the ring position?
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); idx = buf.idx; eng.step(T, buffer=buf); e1.record(); torch.cuda.synchronize()
    print("idx", idx, "ms", e0.elapsed_time(e1))
