#!/bin/bash
# Builds tuning variants of libnimmt_b200.so (one translation unit recompiled with different -D knobs, the rest reused):
#   profiles/tools/build_variants.sh <unit.cu> <name> "<-D flags>" [<name> "<-D flags>" ...]
# Variants land in profiles/tools/variants/lib_<name>.so (git-ignored, shipped to the GPU box) and are selected with
# NIMMT_B200_LIB=<path> (rl-6-nimmt_b200/_native.py).
set -e
cd "$(dirname "$0")/../.."
make -s -j8 -C rl-6-nimmt_b200/csrc
unit=$1; shift
mkdir -p profiles/tools/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $flags \
       -c rl-6-nimmt_b200/csrc/$unit -o profiles/tools/variants/${unit%.cu}_$name.o &
done
wait
for o in profiles/tools/variants/${unit%.cu}_*.o; do
  name=$(basename $o .o); name=${name#${unit%.cu}_}
  others=$(ls rl-6-nimmt_b200/lib/obj/*.o | grep -v "/${unit%.cu}.o")
  nvcc $ARCH -shared -o profiles/tools/variants/lib_$name.so $o $others -lcudart
  rm $o
done
ls -la profiles/tools/variants/
