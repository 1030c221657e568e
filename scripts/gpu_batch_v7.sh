#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/variants5.txt
: > $L
run() { echo "### lib=${1:-default} minb=$2 $3 $4" >> $L; KT_LIB=$1 KT_MIN_BLOCKS=$2 timeout 600 python scripts/gpu_variant.py $3 $4 >> $L 2>&1; }
run "" 0 exact rk45
run cold 0 exact rk45
run v1 0 exact rk45
timeout 900 python scripts/gpu_configs.py > gpurun_out/configs.txt 2>&1
cat $L gpurun_out/configs.txt
