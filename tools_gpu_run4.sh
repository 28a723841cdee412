set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --no-cpu > gpurun_out/r4_default.json 2> gpurun_out/r4_default.err; tail -3 gpurun_out/r4_default.err
for T in 1 100; do python bench.py --turns $T --no-e2e --no-cpu --steps 8 > gpurun_out/r4_T$T.json 2>> gpurun_out/r4_var.err; done
python bench.py --no-e2e --no-cpu --steps 8 --agents 8 --envs-total 262144 --turns 25 --capacity 8 > gpurun_out/r4_c3big.json 2>> gpurun_out/r4_var.err
python bench.py --no-e2e --no-cpu --steps 4 --agents 64 --size 128 --vision 4 --envs-total 2048 --turns 10 --capacity 16 > gpurun_out/r4_c4.json 2>> gpurun_out/r4_var.err
CMD="python bench.py --envs-total 262144 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 1 -o gpurun_out/prof_step_r4 $CMD > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/ncu4.log
tail -20 gpurun_out/r4_var.err
