set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --no-cpu > gpurun_out/r3_default.json 2> gpurun_out/r3_default.err; tail -3 gpurun_out/r3_default.err
python bench.py --no-cpu --no-e2e --no-clocks --steps 8 > gpurun_out/r3_noclocks.json 2>> gpurun_out/r3_var.err
python bench.py --no-cpu --no-e2e --steps 8 --sample-div 0 > gpurun_out/r3_nosample.json 2>> gpurun_out/r3_var.err
for T in 1 100; do python bench.py --turns $T --no-e2e --no-cpu --steps 8 > gpurun_out/r3_T$T.json 2>> gpurun_out/r3_var.err; done
SB_LPE=8 python bench.py --no-e2e --no-cpu --steps 8 > gpurun_out/r3_lpe8.json 2>> gpurun_out/r3_var.err
python bench.py --no-e2e --no-cpu --steps 8 --agents 8 --envs-total 262144 --turns 25 --capacity 8 > gpurun_out/r3_c3big.json 2>> gpurun_out/r3_var.err
python bench.py --no-e2e --no-cpu --steps 4 --agents 64 --size 128 --vision 4 --envs-total 2048 --turns 10 --capacity 16 > gpurun_out/r3_c4.json 2>> gpurun_out/r3_var.err
CMD="python bench.py --envs-total 262144 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/prof_step_r3 $CMD > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
tail -20 gpurun_out/r3_var.err
