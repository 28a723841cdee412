cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for C in 5 3 4 2 cleanup; do
  python bench.py --config $C --no-e2e --no-cpu --steps 12 --warmup 4 > gpurun_out/ab5_main_c$C.json 2> gpurun_out/ab5_main_c$C.err || tail -3 gpurun_out/ab5_main_c$C.err
done
python tools/show_bench.py "gpurun_out/ab5_*.json"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab5_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['roofline']['breakdown_ms_per_step'])
PY
