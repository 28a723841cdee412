cd $GRAFT_REPO_ROOT
TAG=${1:-c4q}
SB_STAGED=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
C4="--agents 64 --size 128 --vision 4 --turns 10 --capacity 8 --no-e2e --no-cpu --no-alt --sample-div 0"
python bench.py $C4 --envs-total 8192 --steps 4 > gpurun_out/${TAG}_c4_8192.json 2> gpurun_out/${TAG}.err; python tools/show_bench.py gpurun_out/${TAG}_c4_8192.json
python bench.py $C4 --envs-total 2048 --steps 4 > gpurun_out/${TAG}_c4_2048.json 2>> gpurun_out/${TAG}.err; python tools/show_bench.py gpurun_out/${TAG}_c4_2048.json
python bench.py --agents 64 --size 128 --vision 9 --turns 5 --capacity 8 --no-e2e --no-cpu --no-alt --sample-div 0 --envs-total 1024 --steps 4 > gpurun_out/${TAG}_c4_v9.json 2>> gpurun_out/${TAG}.err; python tools/show_bench.py gpurun_out/${TAG}_c4_v9.json
python bench.py --no-cpu --no-e2e --steps 8 --no-alt > gpurun_out/${TAG}_default.json 2>> gpurun_out/${TAG}.err; python tools/show_bench.py gpurun_out/${TAG}_default.json
