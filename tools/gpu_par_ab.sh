cd $GRAFT_REPO_ROOT
run() {
C4="--agents 64 --size 128 --vision 4 --turns 10 --capacity 8 --no-e2e --no-cpu --no-alt --sample-div 0"
python bench.py $C4 --envs-total 8192 --steps 4 2>/dev/null > gpurun_out/par_$1_8192.json; python tools/show_bench.py gpurun_out/par_$1_8192.json
python bench.py $C4 --envs-total 2048 --steps 4 2>/dev/null > gpurun_out/par_$1_2048.json; python tools/show_bench.py gpurun_out/par_$1_2048.json
python bench.py --agents 64 --size 128 --vision 9 --turns 5 --capacity 8 --no-e2e --no-cpu --no-alt --sample-div 0 --envs-total 1024 --steps 4 2>/dev/null > gpurun_out/par_$1_v9.json; python tools/show_bench.py gpurun_out/par_$1_v9.json
}
for P in 8 2; do
touch sorrel_b200/csrc/sb_step.cu; make -C sorrel_b200/csrc EXTRA=-DSB_PAR=$P > /dev/null 2>&1; run par$P
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k agent_parallel 2>&1 | tail -1
done
touch sorrel_b200/csrc/sb_step.cu; make -C sorrel_b200/csrc > /dev/null 2>&1
SB_NO_PAR=1 run nopar
