#!/bin/bash
# Builds profiles/tools/variants/lib_checked.so: the whole library with -DNIMMT_BOUNDS_CHECK (device asserts on every
# data-dependent shared-memory index).  Run the small cases and the GPU tests against it:
#   NIMMT_B200_LIB=$PWD/profiles/tools/variants/lib_checked.so python profiles/tools/sanitize_small.py
set -e
cd "$(dirname "$0")/../.."
mkdir -p profiles/tools/variants/obj_checked
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in rl-6-nimmt_b200/csrc/*.cu; do
  nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -DNIMMT_BOUNDS_CHECK \
       -c $f -o profiles/tools/variants/obj_checked/$(basename ${f%.cu}).o &
done
wait
nvcc $ARCH -shared -o profiles/tools/variants/lib_checked.so profiles/tools/variants/obj_checked/*.o -lcudart
rm -rf profiles/tools/variants/obj_checked
ls -la profiles/tools/variants/lib_checked.so
