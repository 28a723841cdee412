# Run on the B200 box: bash tools/gpu_profiles.sh rNN   (ncu evidence of the default bench command)
cd $GRAFT_REPO_ROOT
TAG=${1:-r02}
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --no-graphs"
# 1. plain run first (must exit 0), then the launch list of the SAME command
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
# 2. full captures of the dominant kernels (steady-state launches)
bash tools/gpu_ncu.sh ${TAG}_step_kernel
bash tools/gpu_ncu.sh ${TAG}_sparse_kernel --rng sparse
SKIP=3 bash tools/gpu_ncu.sh ${TAG}_c2_kernel --config 2   # (few launches: skip fewer)
bash tools/gpu_ncu.sh ${TAG}_c3_kernel --config 3
bash tools/gpu_ncu.sh ${TAG}_c4_kernel --config 4
CMD2="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-alt --no-graphs --no-clocks"
ncu --set full --clock-control none -k regex:buffer_sample -s 3 -c 1 -f -o gpurun_out/${TAG}_sample_kernel $CMD2 > gpurun_out/${TAG}_ncu_sample.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_sample.log
