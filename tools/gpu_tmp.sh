cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_named_sizes.py -m gpu -x -q -k "block_per or config4 or parallel_path or philox_mode" 2>&1 | tail -5
for C in 4 3; do
  python bench.py --config $C --no-e2e --no-cpu --no-alt --steps 12 --warmup 4 > gpurun_out/ab7_main_c$C.json 2> gpurun_out/ab7_main_c$C.err || tail -3 gpurun_out/ab7_main_c$C.err
done
python bench.py --config 4 --lpe 32 --no-e2e --no-cpu --no-alt --steps 12 --warmup 4 > gpurun_out/ab7_lpe32_c4.json 2> gpurun_out/ab7_lpe32_c4.err || tail -3 gpurun_out/ab7_lpe32_c4.err
python bench.py --config 4 --envs-total 8192 --capacity 8 --no-e2e --no-cpu --no-alt --steps 8 --warmup 4 > gpurun_out/ab7_main_c4big.json 2> gpurun_out/ab7_main_c4big.err || tail -3 gpurun_out/ab7_main_c4big.err
python bench.py --config 4 --vision 9 --envs-total 1024 --turns 5 --capacity 8 --no-e2e --no-cpu --no-alt --steps 8 --warmup 4 > gpurun_out/ab7_main_c4v9.json 2> gpurun_out/ab7_main_c4v9.err || tail -3 gpurun_out/ab7_main_c4v9.err
python tools/show_bench.py "gpurun_out/ab7_*.json"
