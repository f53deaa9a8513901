#!/bin/sh
# Builds lib/dbg/libnimmt_b200.so = the product library with policy_rollouts.cu and policy_kernels.cu compiled -DNIMMT_PHASE_CLOCKS
# (bring-up instrument read by profiles/tools/policy_phase_clocks.py).  Run from the repo root after `make`.
set -e
cd "$(dirname "$0")/../../rl-6-nimmt_b200/csrc"
mkdir -p ../lib/dbg
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    --expt-relaxed-constexpr -DNIMMT_PHASE_CLOCKS -c policy_rollouts.cu -o ../lib/dbg/policy_rollouts.o
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    --expt-relaxed-constexpr -DNIMMT_PHASE_CLOCKS -c policy_kernels.cu -o ../lib/dbg/policy_kernels.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/dbg/libnimmt_b200.so ../lib/dbg/policy_rollouts.o ../lib/dbg/policy_kernels.o \
    $(ls ../lib/obj/*.o | grep -v "policy_rollouts\|policy_kernels") -lcudart
