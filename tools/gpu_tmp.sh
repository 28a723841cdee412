cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "philox_mode or block_per or config3 or config4 or config5" 2>&1 | tail -3
for C in 5 3 4; do
python bench.py --config $C --no-e2e --no-cpu --steps 12 --warmup 4 > gpurun_out/ab17_main_c$C.json 2> gpurun_out/ab17_main_c$C.err
done
python tools/show_bench.py "gpurun_out/ab17_*.json"
