cd $GRAFT_REPO_ROOT
N=${1:-2}; TOT=${2:-1048576}; TAG=${3:-m}
$PRE python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --envs-total $TOT --no-cpu --no-alt --steps 32 $EXTRA > gpurun_out/${TAG}_g$N.json 2> gpurun_out/${TAG}_g$N.err
python - <<EOF
import json
d=json.loads(open('gpurun_out/${TAG}_g$N.json').read().strip().splitlines()[-1]); b=d['roofline']['breakdown_ms_per_step']
print('N=$N total=$TOT value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], {k:round(v,3) for k,v in b.items()}, 'sum %.3f'%sum(b.values()), 'e2e %.3e'%(d['e2e']['value'] if d['e2e'] else 0))
EOF
