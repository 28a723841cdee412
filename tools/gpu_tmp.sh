cd $GRAFT_REPO_ROOT
for V in main mb7 mb6 copy4 copy16; do
  if [ "$V" = main ]; then unset SB_LIB_PATH; else export SB_LIB_PATH=$GRAFT_REPO_ROOT/sorrel_b200/lib/variants/libsb_$V.so; fi
  python bench.py --config 5 --no-e2e --no-cpu --no-alt --steps 12 --warmup 4 > gpurun_out/ab25_${V}.json 2> gpurun_out/ab25_${V}.err
done
unset SB_LIB_PATH
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab25_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.3e'%d['value'], {k:round(v,3) for k,v in d['roofline']['breakdown_ms_per_step'].items()})
PY
