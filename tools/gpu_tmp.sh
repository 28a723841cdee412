cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_named_sizes.py -x -q -m gpu 2>&1 | tail -3
bash tools/gpu_ab.sh t1 "main noasync main noasync" "5 2 3" --turns 1 2>&1 | tail -14
bash tools/gpu_ab.sh t25 "main noasync main noasync" "5 3" 2>&1 | tail -10
