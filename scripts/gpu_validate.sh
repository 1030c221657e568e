#!/bin/bash
# One GPU session that produces everything profiles/ holds: the GPU test suite, bench.py, the ncu launch list and the full captures
# (raw + per-instruction source pages exported as CSV on the box; feed the .ncu-rep files to scripts/summarise_ncu.py).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/gpu_selftest.py > gpurun_out/selftest.log 2>&1
fi
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_2e6.csv \
    python bench.py --rays-per-gpu 2e6 --e2e-rays 2e6 --steps 2 --warmup 3 --no-cpu-baseline --no-other-mode > gpurun_out/ncu_launches.log 2>&1
cap() { # name minb math
  timeout 600 ncu --set full --clock-control none -k regex:trace_persistent -s 1 -c 1 -f -o /tmp/$1 \
     python scripts/gpu_profile.py --rays 1.5e7 --math $3 --min-blocks $2 > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  gzip -f gpurun_out/$1_source.csv
  cp /tmp/$1.ncu-rep gpurun_out/ 2>/dev/null
}
cap r2_exact 0 exact
cap r2_fast 0 fast
du -sh gpurun_out; tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/selftest.log; cat gpurun_out/bench.json
# the ray-set kernel (HBM-resident records: the write-back traffic)
timeout 600 ncu --set full --clock-control none -k regex:trace_persistent -c 1 -f -o /tmp/r2_rayset python scripts/gpu_profile.py --rays 1.6e7 --rayset > gpurun_out/ncu_r2_rayset.log 2>&1
cp /tmp/r2_rayset.ncu-rep gpurun_out/ 2>/dev/null
