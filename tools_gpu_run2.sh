set -x
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
for T in 1 5 100; do python bench.py --turns $T --no-e2e --no-cpu --steps 8 > gpurun_out/bench_T$T.json 2>> gpurun_out/bench_var.err; done
for L in 8 16 32; do SB_LPE=$L python bench.py --no-e2e --no-cpu --steps 8 > gpurun_out/bench_lpe$L.json 2>> gpurun_out/bench_var.err; done
python bench.py --no-e2e --no-cpu --steps 8 --sample-div 0 > gpurun_out/bench_nosample.json 2>> gpurun_out/bench_var.err
python bench.py --no-e2e --no-cpu --steps 8 --agents 8 --envs-total 16384 --turns 25 > gpurun_out/bench_c3.json 2>> gpurun_out/bench_var.err
python bench.py --no-e2e --no-cpu --steps 8 --agents 8 --envs-total 262144 --turns 25 --capacity 8 > gpurun_out/bench_c3big.json 2>> gpurun_out/bench_var.err
python bench.py --no-e2e --no-cpu --steps 4 --agents 64 --size 128 --vision 4 --envs-total 2048 --turns 10 --capacity 16 > gpurun_out/bench_c4.json 2>> gpurun_out/bench_var.err
python bench.py --no-e2e --no-cpu --steps 8 --envs-total 4096 --turns 100 > gpurun_out/bench_c2.json 2>> gpurun_out/bench_var.err
CMD="python bench.py --envs-total 262144 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/prof_step_r1 $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
cat gpurun_out/bench_var.err | tail -20
