#!/bin/bash
# GPU session 2: FP64 pipe micro-benchmarks + ncu captures (raw + per-instruction source pages exported as CSV on the box).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./build/fp64_probe > gpurun_out/fp64_probe.txt 2>&1
cap() { # name lib minb math
  KT_LIB=$2 timeout 600 ncu --set full --clock-control none -k regex:trace_persistent -s 1 -c 1 -f -o /tmp/$1 \
     python scripts/gpu_profile.py --rays 1.5e7 --math $4 --min-blocks $3 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  gzip -f gpurun_out/$1_source.csv
}
cap r1c_v1_exact6 v1 6 exact
cap r1c_v2_exact6 "" 6 exact
cap r1c_v2_exact4 "" 4 exact
cap r1c_v1_exact4 v1 4 exact
du -sh gpurun_out; cat gpurun_out/fp64_probe.txt | head -100
