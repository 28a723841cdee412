cd $GRAFT_REPO_ROOT
for L in 4 8 16 32; do
  SB_LPE=$L python bench.py --no-e2e --no-cpu --no-clocks --steps 8 --agents 8 --envs-total 262144 --turns 25 --capacity 8 > gpurun_out/lpe${L}_c3.json 2>> gpurun_out/lpe.err
  SB_LPE=$L python bench.py --no-e2e --no-cpu --no-clocks --steps 8 --agents 8 --envs-total 16384 --turns 25 --capacity 32 > gpurun_out/lpe${L}_c3small.json 2>> gpurun_out/lpe.err
done
for L in 16 32; do
  SB_LPE=$L python bench.py --no-e2e --no-cpu --no-clocks --steps 4 --agents 64 --size 128 --vision 4 --envs-total 2048 --turns 10 --capacity 16 > gpurun_out/lpe${L}_c4.json 2>> gpurun_out/lpe.err
  SB_LPE=$L python bench.py --no-e2e --no-cpu --no-clocks --steps 4 --agents 64 --size 128 --vision 4 --envs-total 8192 --turns 10 --capacity 8 > gpurun_out/lpe${L}_c4big.json 2>> gpurun_out/lpe.err
done
tail -3 gpurun_out/lpe.err
