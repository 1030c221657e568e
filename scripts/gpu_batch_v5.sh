#!/bin/bash
# GPU session 4: occupancy sweep of the non-pipelined guard-free RK4 step, unrolled stages, 6-operation quotients, RK45 occupancy.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/variants3.log
: > $L
run() { echo "### lib=${1:-default} minb=$2 $3 $4" >> $L; KT_LIB=$1 KT_MIN_BLOCKS=$2 timeout 300 python scripts/gpu_variant.py $3 $4 2>&1 | grep -v "h=5.0" >> $L; }
for mb in 3 4 5 6; do run "" $mb exact; done
for mb in 3 4 5; do run u $mb exact; done
run nd6 3 exact; run nd6 4 exact; run ud6 4 exact
for mb in 3 4 5 6; do run "" $mb fast; done
run u 4 fast; run u 5 fast
for mb in 3 4 5 6; do echo "### rk45 default minb=$mb" >> $L; KT_MIN_BLOCKS=$mb timeout 300 python scripts/gpu_variant.py exact rk45 2>&1 | grep RK45 >> $L; done
echo "### rk45 fast v1 / default" >> $L
KT_LIB=v1 timeout 300 python scripts/gpu_variant.py fast rk45 2>&1 | grep RK45 >> $L
timeout 300 python scripts/gpu_variant.py fast rk45 2>&1 | grep RK45 >> $L
KT_MIN_BLOCKS=4 timeout 300 python scripts/gpu_variant.py fast rk45 2>&1 | grep RK45 >> $L
grep -v "h=10" $L
