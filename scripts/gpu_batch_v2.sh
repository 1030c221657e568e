#!/bin/bash
# One GPU session: A/B of the RK4 step variants (parity statistics + 1.5e7-ray timing), the GPU test suite, ncu captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/variants.log
: > $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L 2>&1
run() { echo "### lib=${1:-default} minb=$2 $3" >> $L; KT_LIB=$1 KT_MIN_BLOCKS=$2 timeout 300 python scripts/gpu_variant.py $3 >> $L 2>&1; }
run v1 6 exact
run "" 6 exact
run "" 5 exact
run "" 4 exact
run "" 7 exact
run d6 6 exact
run d6 5 exact
run d6 4 exact
run v1 5 fast
run "" 5 fast
run "" 4 fast
run "" 6 fast
echo "### selftest default" >> $L; timeout 300 python scripts/gpu_selftest.py >> $L 2>&1
echo "### selftest d6" >> $L; KT_LIB=d6 timeout 300 python scripts/gpu_selftest.py >> $L 2>&1
# the GPU suite on the default library
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $L
# ncu: full capture of the new kernel at two occupancies (second launch of each run)
for mb in 6 5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_persistent -s 1 -c 1 -f -o gpurun_out/prof_r1c_exact_mb$mb \
     python scripts/gpu_profile.py --rays 1.5e7 --math exact --min-blocks $mb > gpurun_out/ncu_exact_mb$mb.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_persistent -s 1 -c 1 -f -o gpurun_out/prof_r1c_fast_mb5 \
     python scripts/gpu_profile.py --rays 1.5e7 --math fast --min-blocks 5 > gpurun_out/ncu_fast_mb5.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
cat $L | grep -v "^\[.*h=" | tail -60
