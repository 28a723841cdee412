cd $GRAFT_REPO_ROOT
ARGS="--world cleanup --agents 10 --vision 5 --envs-total 32768 --turns 25 --capacity 8 --n-frames 5 --steps 6 --warmup 4 --no-e2e --no-cpu --sample-div 0"
for L in 8 16 32; do
SB_LPE=$L python bench.py $ARGS 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('LPE $L', '%.3e'%d['value'], 'kernel_ms %.3f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'])"
done
