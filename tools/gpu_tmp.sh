cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "block_per or config4" 2>&1 | tail -3
python bench.py --config 4 --no-e2e --no-cpu --no-alt --steps 12 --warmup 4 > gpurun_out/ab26_c4.json 2> gpurun_out/ab26_c4.err
python bench.py --config 4 --envs-total 8192 --capacity 8 --no-e2e --no-cpu --no-alt --steps 8 --warmup 4 > gpurun_out/ab26_c4big.json 2> gpurun_out/ab26_c4big.err
python tools/show_bench.py "gpurun_out/ab26_*.json"
