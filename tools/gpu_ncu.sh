# one ncu --set full capture of a steady-state step_kernel launch:  bash tools/gpu_ncu.sh NAME [bench flags]
cd $GRAFT_REPO_ROOT
NAME=$1; shift
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-alt --no-graphs --no-clocks $*"
$CMD > gpurun_out/${NAME}_plain.json 2> gpurun_out/${NAME}_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s ${SKIP:-6} -c 1 -f -o gpurun_out/${NAME} $CMD > gpurun_out/${NAME}_ncu.log 2>&1
tail -2 gpurun_out/${NAME}_ncu.log
