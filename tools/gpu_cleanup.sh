cd $GRAFT_REPO_ROOT
TAG=${1:-cl}
timeout 600 python -m pytest tests/test_gpu_cleanup.py -m gpu -x -q > gpurun_out/${TAG}_cleanup.log 2>&1; tail -30 gpurun_out/${TAG}_cleanup.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_all.log 2>&1; tail -5 gpurun_out/${TAG}_all.log
python bench.py --no-cpu --no-e2e --steps 12 > gpurun_out/${TAG}_default.json 2> gpurun_out/${TAG}.err; tail -3 gpurun_out/${TAG}.err
