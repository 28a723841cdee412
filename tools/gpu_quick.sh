cd $GRAFT_REPO_ROOT
TAG=$1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu --no-e2e --steps 12 > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}.err; tail -3 gpurun_out/${TAG}.err
python bench.py --turns 1 --no-e2e --no-cpu --steps 8 > gpurun_out/${TAG}_T1.json 2>> gpurun_out/${TAG}.err
python bench.py --no-e2e --no-cpu --steps 8 --agents 8 --envs-total 262144 --turns 25 --capacity 8 > gpurun_out/${TAG}_c3big.json 2>> gpurun_out/${TAG}.err
python bench.py --no-e2e --no-cpu --steps 4 --agents 64 --size 128 --vision 4 --envs-total 2048 --turns 10 --capacity 16 > gpurun_out/${TAG}_c4.json 2>> gpurun_out/${TAG}.err
CMD="python bench.py --envs-total 262144 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 1 -o gpurun_out/prof_step_${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_${TAG}.log
