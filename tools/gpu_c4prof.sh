cd $GRAFT_REPO_ROOT
TAG=${1:-c4p}
C4="--agents 64 --size 128 --vision 4 --envs-total 8192 --turns 10 --capacity 8 --no-e2e --no-cpu --no-alt --sample-div 0"
python bench.py $C4 --steps 4 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}.err; python tools/show_bench.py gpurun_out/${TAG}_c4.json
CMD="python bench.py $C4 --steps 2 --warmup 3 --no-graphs"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 6 -c 1 -o gpurun_out/prof_c4_${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_${TAG}.log
