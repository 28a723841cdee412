cd $GRAFT_REPO_ROOT
run() {
for cfg in "--size 128 --agents 64 --vision 4 --envs-total 4096" "--size 128 --agents 64 --vision 9 --envs-total 1024" "--size 21 --agents 8 --vision 2 --envs-total 262144" "--size 21 --agents 1 --vision 2 --envs-total 1048576"; do
python bench.py --obs-study $cfg --steps 20 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$1 $cfg', 'ms %.3f min %.3f'%(r['kernel_ms'], r['kernel_ms_min']), 'GB/s %.0f frac %.3f'%(r['achieved'], r['frac']))"
done
}
run minb6


timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "observe or obs" 2>&1 | tail -2
