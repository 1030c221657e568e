#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/variants6.txt
: > $L
run() { echo "### lib=${1:-default} minb=$2 $3 $4" >> $L; KT_LIB=$1 KT_MIN_BLOCKS=$2 timeout 300 python scripts/gpu_variant.py $3 $4 2>&1 | grep -v "h=5.0" >> $L; }
for v in "" e1 e3 e6 e7 ""; do run "$v" 0 exact; done
for v in "" e1 e6 e7; do run "$v" 0 fast; done
grep -v "h=10" $L | cut -c1-60; grep "h=10" $L | awk '{print $NF}' | sort | uniq -c
