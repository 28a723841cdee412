# A/B of library variants on one box:  bash tools/gpu_ab.sh TAG "VARIANTS" "CONFIGS" [extra bench flags]
#   VARIANTS: names under sorrel_b200/lib/variants (libsb_NAME.so); "main" = the product library
#   CONFIGS:  bench.py --config presets
cd $GRAFT_REPO_ROOT
TAG=$1; VARS=$2; CFGS=$3; shift 3
for C in $CFGS; do
  for V in $VARS; do
    if [ "$V" = main ]; then unset SB_LIB_PATH; else export SB_LIB_PATH=$GRAFT_REPO_ROOT/sorrel_b200/lib/variants/libsb_$V.so; fi
    python bench.py --config $C --no-e2e --no-cpu --steps 12 --warmup 4 "$@" > gpurun_out/${TAG}_${V}_c$C.json 2> gpurun_out/${TAG}_${V}_c$C.err || tail -3 gpurun_out/${TAG}_${V}_c$C.err
  done
done
unset SB_LIB_PATH
python tools/show_bench.py "gpurun_out/${TAG}_*.json"
