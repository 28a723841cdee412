cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
SB_STAGED=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for S in 1 0; do
  SB_STAGED=$S python bench.py --no-e2e --no-cpu --no-clocks --no-alt --steps 4 --agents 64 --size 128 --vision 4 --envs-total 2048 --turns 10 --capacity 16 > gpurun_out/c4_staged${S}_2048.json 2>> gpurun_out/c4.err
  SB_STAGED=$S python bench.py --no-e2e --no-cpu --no-clocks --no-alt --steps 4 --agents 64 --size 128 --vision 4 --envs-total 8192 --turns 10 --capacity 8 > gpurun_out/c4_staged${S}_8192.json 2>> gpurun_out/c4.err
  SB_STAGED=$S python bench.py --no-e2e --no-cpu --no-clocks --no-alt --steps 4 --agents 64 --size 128 --vision 4 --envs-total 8192 --turns 10 --capacity 8 --rng sparse > gpurun_out/c4_staged${S}_8192_sparse.json 2>> gpurun_out/c4.err
done
SB_STAGED=0 python bench.py --no-e2e --no-cpu --no-clocks --no-alt --steps 8 > gpurun_out/c4_staged0_default.json 2>> gpurun_out/c4.err
tail -3 gpurun_out/c4.err
