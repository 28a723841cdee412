cd $GRAFT_REPO_ROOT
TAG=${1:-p}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
SB_STAGED=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or philox" 2>&1 | tail -1
B="--no-e2e --no-cpu --no-alt"
python bench.py $B --steps 8 > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}.err
python bench.py $B --steps 6 --agents 8 --envs-total 262144 --turns 25 --capacity 8 --sample-div 0 > gpurun_out/${TAG}_c3.json 2>> gpurun_out/${TAG}.err
python bench.py $B --steps 4 --agents 64 --size 128 --vision 4 --turns 10 --capacity 8 --sample-div 0 --envs-total 8192 > gpurun_out/${TAG}_c4.json 2>> gpurun_out/${TAG}.err
python bench.py $B --steps 8 --turns 1 > gpurun_out/${TAG}_T1.json 2>> gpurun_out/${TAG}.err
python bench.py $B --steps 6 --world cleanup --agents 10 --vision 5 --envs-total 32768 --turns 25 --capacity 8 --n-frames 5 --sample-div 0 > gpurun_out/${TAG}_cleanup.json 2>> gpurun_out/${TAG}.err
python bench.py $B --steps 8 --rng sparse > gpurun_out/${TAG}_sparse.json 2>> gpurun_out/${TAG}.err
for f in default c3 c4 T1 cleanup sparse; do python tools/show_bench.py gpurun_out/${TAG}_$f.json; done
