cd $GRAFT_REPO_ROOT
for extra in "" "--sample-div 0" "--no-graphs --sample-div 0"; do python bench.py --no-cpu --no-alt --steps 16 $extra 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$extra value %.3e kernel_ms %.2f e2e %.3e e2e_ms %.2f e2e_kernel_ms %.2f'%(d['value'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['step_kernel_ms']))"; done
