# cleanup / iowa / tag GPU tests + the cleanup bench line
cd $GRAFT_REPO_ROOT
TAG=${1:-cl}
timeout 600 python -m pytest tests/test_gpu_cleanup.py -m gpu -x -q 2>&1 | tail -2
python bench.py --world cleanup --agents 10 --vision 5 --envs-total 32768 --turns 25 --capacity 8 --n-frames 5 --steps 6 --warmup 4 --no-e2e --no-cpu --no-alt --sample-div 0 > gpurun_out/${TAG}_cleanup.json 2> gpurun_out/${TAG}.err
python tools/show_bench.py gpurun_out/${TAG}_cleanup.json
