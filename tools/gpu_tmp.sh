cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for C in 5 3 2 4; do
python bench.py --config $C --no-e2e --no-cpu --steps 12 --warmup 4 > gpurun_out/ab24_main_c$C.json 2> gpurun_out/ab24_main_c$C.err
done
python tools/show_bench.py "gpurun_out/ab24_*.json"
