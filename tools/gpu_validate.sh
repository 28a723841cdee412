# full round-end style validation: GPU tests, smoke, both bench arms
cd $GRAFT_REPO_ROOT
TAG=${1:-val}
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -4 gpurun_out/${TAG}_bench.err
( time python bench.py --impl reference --steps 4 --warmup 1 ) > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; tail -4 gpurun_out/${TAG}_ref.err
