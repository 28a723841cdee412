cd $GRAFT_REPO_ROOT
for V in "$@"; do
  export SB_LIB_PATH=$GRAFT_REPO_ROOT/sorrel_b200/lib/ab/$V.so
  python bench.py --no-e2e --no-cpu --no-clocks --steps 8 > gpurun_out/ab_${V}_default.json 2>> gpurun_out/ab.err
  python bench.py --no-e2e --no-cpu --no-clocks --steps 8 --agents 8 --envs-total 262144 --turns 25 --capacity 8 > gpurun_out/ab_${V}_c3.json 2>> gpurun_out/ab.err
  python bench.py --no-e2e --no-cpu --no-clocks --steps 4 --agents 64 --size 128 --vision 4 --envs-total 2048 --turns 10 --capacity 16 > gpurun_out/ab_${V}_c4.json 2>> gpurun_out/ab.err
  python bench.py --no-e2e --no-cpu --no-clocks --steps 8 --turns 1 > gpurun_out/ab_${V}_T1.json 2>> gpurun_out/ab.err
done
tail -3 gpurun_out/ab.err
