cd $GRAFT_REPO_ROOT
TAG=${1:-q}
CL="--world cleanup --agents 10 --vision 5 --envs-total 32768 --turns 25 --capacity 8 --n-frames 5 --steps 6 --warmup 4 --no-e2e --no-cpu --sample-div 0"
python bench.py $CL > gpurun_out/${TAG}_cleanup.json 2> gpurun_out/${TAG}.err
python bench.py --no-cpu --no-e2e --steps 8 --no-alt > gpurun_out/${TAG}_default.json 2>> gpurun_out/${TAG}.err
for f in cleanup default; do python tools/show_bench.py gpurun_out/${TAG}_$f.json; done
timeout 600 python -m pytest tests/test_gpu_cleanup.py -m gpu -x -q 2>&1 | tail -2
if [ -n "$2" ]; then
CMD="python bench.py $CL --steps 2 --warmup 3 --no-graphs"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o gpurun_out/prof_cleanup_${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_${TAG}.log
fi
