cd $GRAFT_REPO_ROOT
C4="--agents 64 --size 128 --vision 4 --turns 10 --capacity 8 --no-e2e --no-cpu --no-alt --sample-div 0 --envs-total 8192 --steps 6"
for i in 1 2 3; do python bench.py $C4 2>/dev/null > gpurun_out/c4v_$i.json; python tools/show_bench.py gpurun_out/c4v_$i.json; done
SB_NO_PAR=1 python bench.py $C4 2>/dev/null > gpurun_out/c4v_nopar.json; python tools/show_bench.py gpurun_out/c4v_nopar.json
nvidia-smi --query-gpu=clocks.sm,clocks.mem,temperature.gpu,power.draw --format=csv
