cd $GRAFT_REPO_ROOT
for extra in "" "--no-graphs" "--sample-div 0" "--turns 100"; do
SB_BENCH_DEBUG=1 python bench.py --envs-total 131072 --no-e2e --no-cpu --no-alt --steps 32 $extra 2> gpurun_out/small.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['roofline']['breakdown_ms_per_step']; print('$extra', 'ms/step %.3f'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'sum %.3f'%sum(b.values()))"
grep "rank 0" gpurun_out/small.err | tail -1
done
