# Run on the B200 box: bash tools/gpu_profiles.sh rNN   (one ncu family per gpurun call)
cd $GRAFT_REPO_ROOT
TAG=${1:-r01}
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu"
# 1. plain run first (must exit 0), then the launch list of the SAME command
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
# 2. one full capture of the dominant kernel (steady-state launch)
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 1 -o gpurun_out/${TAG}_step_kernel $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:buffer_sample -s 3 -c 1 -o gpurun_out/${TAG}_sample_kernel $CMD > gpurun_out/${TAG}_ncu_full2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
