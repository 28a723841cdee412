cd $GRAFT_REPO_ROOT
TAG=${1:-clb}
ARGS="--world cleanup --agents 10 --vision 5 --envs-total 32768 --turns 25 --capacity 8 --n-frames 5 --steps 8 --warmup 4 --no-e2e --no-cpu"
python bench.py $ARGS > gpurun_out/${TAG}_cleanup.json 2> gpurun_out/${TAG}.err; tail -3 gpurun_out/${TAG}.err; cat gpurun_out/${TAG}_cleanup.json | cut -c1-400
CMD="python bench.py $ARGS --steps 2 --warmup 3 --no-graphs"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o gpurun_out/prof_cleanup_${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log
