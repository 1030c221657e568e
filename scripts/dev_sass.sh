#!/bin/bash
# Offline check of a kernel variant (no GPU): compile only the headline kernel (fused emissivity, RK4) with the given
# extra flags, print ptxas's register / spill line, and run the SASS issue model on its main loop.
#   scripts/dev_sass.sh <name> [exact|fast] [MINB] [extra nvcc flags...]
set -e
cd "$(dirname "$0")/../raytrace_cpu_b200/csrc"
export PATH=/usr/local/cuda/bin:$PATH
name=$1; mode=${2:-exact}; minb=${3:-4}; shift; shift || true; shift || true
mkdir -p /tmp/dev
if [ "$mode" = fast ]; then M="-fmad=true -DKT_FAST=1"; NS=4fast; else M="-fmad=false -DKT_FAST=0"; NS=5exact; fi
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off \
     -Xptxas -v $M -DKT_DEV_MIN=$minb "$@" -c kt_kernels.cu -o /tmp/dev/$name.o 2> /tmp/dev/$name.log
K=$(grep -o "_ZN2kt${NS}16trace_persistentILi1ENS0_7EmisIOTILb0EEE[A-Za-z0-9_]*" /tmp/dev/$name.log | head -1)
grep -A2 "Compiling entry function '$K'" /tmp/dev/$name.log | tail -2
cuobjdump -sass -fun "$K" /tmp/dev/$name.o > /tmp/dev/$name.sass
python ../../scripts/sass_sim.py /tmp/dev/$name.sass auto
