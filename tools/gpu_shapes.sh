# Every BASELINE.json shape on one GPU through bench.py's presets -> gpurun_out/${TAG}_shapes.jsonl (one JSON line each)
cd $GRAFT_REPO_ROOT
TAG=${1:-r02}
OUT=gpurun_out/${TAG}_shapes.jsonl
: > $OUT
B="python bench.py --no-e2e --no-cpu --steps 12 --warmup 4"
$B --config 2 >> $OUT 2>> gpurun_out/${TAG}_shapes.err                       # 4096 x 1, 100 turns per launch
$B --config 2 --turns 1 >> $OUT 2>> gpurun_out/${TAG}_shapes.err             # ... one turn per launch
$B --config 3 >> $OUT 2>> gpurun_out/${TAG}_shapes.err                       # 16384 x 8 agents
$B --config 3 --envs-total 262144 --capacity 8 >> $OUT 2>> gpurun_out/${TAG}_shapes.err   # same shape, bandwidth-sized batch
$B --config 4 >> $OUT 2>> gpurun_out/${TAG}_shapes.err                       # 128x128, 64 agents, 9x9 windows, 2048 instances
$B --config 4 --envs-total 8192 --capacity 8 >> $OUT 2>> gpurun_out/${TAG}_shapes.err
$B --config 4 --vision 9 --envs-total 1024 --turns 5 --capacity 8 >> $OUT 2>> gpurun_out/${TAG}_shapes.err   # 19x19 reading of "9x9 vision radius"
$B --config 5 >> $OUT 2>> gpurun_out/${TAG}_shapes.err                       # 2^20 x 1 (the default)
$B --config 5 --turns 1 >> $OUT 2>> gpurun_out/${TAG}_shapes.err             # live-policy cadence
$B --config cleanup >> $OUT 2>> gpurun_out/${TAG}_shapes.err                 # the reference's cleanup example
python bench.py --config 4obs >> $OUT 2>> gpurun_out/${TAG}_shapes.err      # observation kernel alone (config 4 study)
python bench.py --config 4obs19 >> $OUT 2>> gpurun_out/${TAG}_shapes.err
python bench.py --obs-study --envs-total 1048576 --agents 1 >> $OUT 2>> gpurun_out/${TAG}_shapes.err
python bench.py --obs-study --envs-total 262144 --agents 8 >> $OUT 2>> gpurun_out/${TAG}_shapes.err
tail -3 gpurun_out/${TAG}_shapes.err
wc -l $OUT
