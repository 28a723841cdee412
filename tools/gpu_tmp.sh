cd $GRAFT_REPO_ROOT
python bench.py --config live > gpurun_out/ab22_live_1m.json 2> gpurun_out/ab22_live_1m.err || tail -5 gpurun_out/ab22_live_1m.err
python bench.py --config live --envs-total 4096 > gpurun_out/ab22_live_4k.json 2> gpurun_out/ab22_live_4k.err || tail -5 gpurun_out/ab22_live_4k.err
python bench.py --config live --envs-total 16384 --agents 8 > gpurun_out/ab22_live_c3.json 2> gpurun_out/ab22_live_c3.err || tail -5 gpurun_out/ab22_live_c3.err
python - <<'PY'
import json
for f in ['ab22_live_1m','ab22_live_4k','ab22_live_c3']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'graph %.3e (%.3f ms/turn, frac %.3f)  eager %.3e (%.3f ms/turn)'%(d['value'],d['ms_per_step'],d['roofline']['frac'],d['eager']['value'],d['eager']['ms_per_turn']))
    except Exception as e: print(f,'ERR',e)
PY
