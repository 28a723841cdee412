cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "uint8 or graph or api" 2>&1 | tail -3
python bench.py --no-cpu --no-alt --steps 16 > gpurun_out/ab21_u8.json 2> gpurun_out/ab21_u8.err
python bench.py --no-cpu --no-alt --steps 16 --actions-i32 > gpurun_out/ab21_i32.json 2> gpurun_out/ab21_i32.err
python - <<'PY'
import json
for f in ['ab21_u8','ab21_i32']:
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); e=d['e2e']
    print(f, 'value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']), 'e2e %.3e ms/step %.3f kernel %.3f h2d %d'%(e['value'],e['ms_per_step'],e['step_kernel_ms'],e['h2d_bytes_per_step']), 'launches', d['gpu_launches'])
PY
