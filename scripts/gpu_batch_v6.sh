#!/bin/bash
# GPU session 5: merged-acceptance-branch RK4 step, rolled / unrolled, occupancy per mode; RK45 occupancy.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/variants4.log
: > $L
run() { echo "### lib=${1:-default} minb=$2 $3 $4" >> $L; KT_LIB=$1 KT_MIN_BLOCKS=$2 timeout 300 python scripts/gpu_variant.py $3 $4 2>&1 | grep -v "h=5.0" >> $L; }
run v1 6 exact
run "" 4 exact; run "" 5 exact; run "" 3 exact
run u 4 exact; run u 5 exact; run u 3 exact
run ud6 4 exact
run "" 4 fast; run "" 5 fast; run "" 6 fast
run u 4 fast; run u 5 fast; run u 6 fast
echo "### rk45 exact minb=3 / fast minb 3 4 5" >> $L
KT_MIN_BLOCKS=3 timeout 300 python scripts/gpu_variant.py exact rk45 2>&1 | grep RK45 >> $L
for mb in 3 4 5; do KT_MIN_BLOCKS=$mb timeout 300 python scripts/gpu_variant.py fast rk45 2>&1 | grep RK45 >> $L; done
grep -v "h=10" $L; grep "h=10" $L | awk '{print $NF}' | sort | uniq -c
