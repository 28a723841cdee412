cd $GRAFT_REPO_ROOT
TAG=${1:-o}
for cfg in "--size 128 --agents 64 --vision 4 --envs-total 4096" "--size 128 --agents 64 --vision 9 --envs-total 1024" "--size 21 --agents 8 --vision 2 --envs-total 262144" "--size 21 --agents 1 --vision 2 --envs-total 1048576"; do
python bench.py --obs-study $cfg --steps 20 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$cfg', 'ms %.3f min %.3f'%(r['kernel_ms'], r['kernel_ms_min']), 'GB/s %.0f frac %.3f'%(r['achieved'], r['frac']))"
done
if [ -n "$2" ]; then
CMD="python bench.py --obs-study --size 128 --agents 64 --vision 4 --envs-total 4096 --steps 3 --warmup 2"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:observe_kernel -s 3 -c 1 -o gpurun_out/prof_obs_${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_${TAG}.log
fi
