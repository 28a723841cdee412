# All BASELINE.json shapes on one GPU -> gpurun_out/${TAG}_shapes.jsonl (one bench.py JSON line per shape)
cd $GRAFT_REPO_ROOT
TAG=${1:-r01}
OUT=gpurun_out/${TAG}_shapes.jsonl
: > $OUT
B="python bench.py --no-e2e --no-cpu --steps 8 --warmup 4"
$B --envs-total 4096 --turns 100 >> $OUT 2>> gpurun_out/${TAG}_shapes.err                                   # config 2
$B --envs-total 4096 --turns 1 >> $OUT 2>> gpurun_out/${TAG}_shapes.err                                     # config 2, one turn per launch
$B --envs-total 16384 --agents 8 --turns 25 >> $OUT 2>> gpurun_out/${TAG}_shapes.err                        # config 3
$B --envs-total 262144 --agents 8 --turns 25 --capacity 8 >> $OUT 2>> gpurun_out/${TAG}_shapes.err          # config 3 shape, bandwidth-sized batch
$B --envs-total 2048 --agents 64 --size 128 --vision 4 --turns 10 --capacity 16 >> $OUT 2>> gpurun_out/${TAG}_shapes.err   # config 4 (S=9)
$B --envs-total 8192 --agents 64 --size 128 --vision 4 --turns 10 --capacity 8 >> $OUT 2>> gpurun_out/${TAG}_shapes.err
$B --envs-total 1024 --agents 64 --size 128 --vision 9 --turns 5 --capacity 8 >> $OUT 2>> gpurun_out/${TAG}_shapes.err     # config 4, S=19 reading
$B --envs-total 1048576 --turns 25 >> $OUT 2>> gpurun_out/${TAG}_shapes.err                                 # config 5 (the default)
$B --envs-total 1048576 --turns 1 >> $OUT 2>> gpurun_out/${TAG}_shapes.err                                  # config 5, one turn per launch (live-policy cadence)
$B --world cleanup --agents 10 --vision 5 --envs-total 32768 --turns 25 --capacity 8 --n-frames 5 >> $OUT 2>> gpurun_out/${TAG}_shapes.err   # cleanup example
python bench.py --obs-study --envs-total 4096 --agents 64 --size 128 --vision 4 --steps 20 --warmup 5 >> $OUT 2>> gpurun_out/${TAG}_shapes.err
python bench.py --obs-study --envs-total 4096 --agents 64 --size 128 --vision 9 --steps 20 --warmup 5 >> $OUT 2>> gpurun_out/${TAG}_shapes.err
python bench.py --obs-study --envs-total 1048576 --agents 1 --steps 20 --warmup 5 >> $OUT 2>> gpurun_out/${TAG}_shapes.err
python bench.py --obs-study --envs-total 262144 --agents 8 --steps 20 --warmup 5 >> $OUT 2>> gpurun_out/${TAG}_shapes.err
tail -3 gpurun_out/${TAG}_shapes.err
wc -l $OUT
