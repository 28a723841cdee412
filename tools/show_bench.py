import json,sys,glob
for f in sorted(glob.glob(sys.argv[1])):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        s=d.get('sparse_rng') or {}
        print(f"{f.split('/')[-1]:34s} value={d['value']:.3e} ms/step={d['ms_per_step']:.3f} kernel_ms={r['kernel_ms']:.3f} frac={r['frac']:.3f} "
              f"share={r.get('step_kernel_share_of_step',0):.2f} sparse_kernel_ms={s.get('kernel_ms',0):.3f} clk={(d.get('clocks') or {}).get('sm_mhz')}")
        if d.get('e2e'): print('    e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
