cd $GRAFT_REPO_ROOT
B="--no-e2e --no-cpu --no-alt"
for lib in "" "sorrel_b200/lib/alt_oldcopy.so"; do
export SB_LIB_PATH=$lib; [ -z "$lib" ] && unset SB_LIB_PATH
python bench.py $B --steps 4 --agents 64 --size 128 --vision 4 --turns 10 --capacity 8 --sample-div 0 --envs-total 8192 2>/dev/null > gpurun_out/alt_c4.json; echo "lib=$lib"; python tools/show_bench.py gpurun_out/alt_c4.json
python bench.py $B --steps 6 --agents 8 --envs-total 262144 --turns 25 --capacity 8 --sample-div 0 2>/dev/null > gpurun_out/alt_c3.json; python tools/show_bench.py gpurun_out/alt_c3.json
python bench.py $B --steps 8 2>/dev/null > gpurun_out/alt_def.json; python tools/show_bench.py gpurun_out/alt_def.json
done
