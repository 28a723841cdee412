# Build an A/B variant of the library next to the product one:  bash tools/build_variant.sh NAME "-DSB_X=.. -DSB_Y=.."
# -> sorrel_b200/lib/variants/libsb_NAME.so  (select it with SB_LIB_PATH, read by the PYTHON loader only)
NAME=$1; shift
mkdir -p "$(dirname "$0")/../sorrel_b200/lib/variants"
make -C "$(dirname "$0")/../sorrel_b200/csrc" OUT=../lib/variants/libsb_$NAME.so OBJ=../lib/obj_$NAME EXTRA="$*" 2>&1 | grep -E "error|Error" ; ls -la "$(dirname "$0")/../sorrel_b200/lib/variants/libsb_$NAME.so"
