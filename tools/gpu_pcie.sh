cd $GRAFT_REPO_ROOT
python - <<'EOF'
import torch, time
dev=torch.device('cuda:0')
n=104857600
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device=dev)
h2=torch.empty(n,dtype=torch.uint8).pin_memory(); d2=torch.empty(n,dtype=torch.uint8,device=dev)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def t(f,reps=10):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps*1e3
print('h2d ms %.2f'%t(lambda: d.copy_(h,non_blocking=True)))
print('d2h ms %.2f'%t(lambda: h2.copy_(d2,non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
print('both ms %.2f'%t(both))
EOF
nvidia-smi topo -m 2>/dev/null | head -5; lscpu | grep -i "numa\|model name\|^CPU(s)" | head -6
for i in 1 2; do python bench.py --no-cpu --no-alt --steps 16 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.3e e2e %.3e e2e_ms %.2f'%(d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"; done
