#!/bin/bash
# GPU session 3: A/B of the zero-tolerant guard-free RK4 step (pipelined / not, 8- / 6-operation quotients) against the v1 build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/variants2.log
: > $L
run() { echo "### lib=${1:-default} minb=$2 $3 $4" >> $L; KT_LIB=$1 KT_MIN_BLOCKS=$2 timeout 300 python scripts/gpu_variant.py $3 $4 >> $L 2>&1; }
run v1 6 exact
run "" 6 exact
run "" 5 exact
run "" 4 exact
run np 6 exact
run np 5 exact
run np 4 exact
run d6 6 exact
run d6 5 exact
run nd6 6 exact
run nd6 5 exact
run "" 5 fast
run np 5 fast
KT_LIB=v1 timeout 300 python scripts/gpu_variant.py exact rk45 2>&1 | grep RK45 >> $L
timeout 300 python scripts/gpu_variant.py exact rk45 2>&1 | grep RK45 >> $L
KT_LIB=np timeout 300 python scripts/gpu_variant.py exact rk45 2>&1 | grep RK45 >> $L
cat $L
