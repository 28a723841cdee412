# full round-end style validation: GPU tests, smoke, both bench arms, then one line per BASELINE shape
cd $GRAFT_REPO_ROOT
TAG=${1:-val}
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -6 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -4 gpurun_out/${TAG}_bench.err
( time python bench.py --impl reference --steps 4 --warmup 1 ) > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; tail -4 gpurun_out/${TAG}_ref.err
for C in 2 3 4 cleanup 4obs; do
  python bench.py --config $C --no-e2e --no-cpu --steps 8 --warmup 4 > gpurun_out/${TAG}_c$C.json 2> gpurun_out/${TAG}_c$C.err || tail -5 gpurun_out/${TAG}_c$C.err
done
python tools/show_bench.py "gpurun_out/${TAG}_*.json"
