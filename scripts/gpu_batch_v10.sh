#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/variants7.txt
: > $L
for mb in 0 4; do echo "### rk45 exact minb=$mb" >> $L; KT_MIN_BLOCKS=$mb timeout 300 python scripts/gpu_variant.py exact rk45 2>&1 | grep -v "h=5.0" >> $L; done
echo "### rk45 cold" >> $L; KT_LIB=cold timeout 300 python scripts/gpu_variant.py exact rk45 2>&1 | grep RK45 >> $L
for mb in 0 4; do echo "### rk45 fast minb=$mb" >> $L; KT_MIN_BLOCKS=$mb timeout 300 python scripts/gpu_variant.py fast rk45 2>&1 | grep "RK45\|1.5e7" >> $L; done
CFG_ONLY=C3 CFG_TOLS=1e-8 timeout 600 python scripts/gpu_configs.py >> $L 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "rk45 or cold or golden" > gpurun_out/pytest_rk45.txt 2>&1; tail -3 gpurun_out/pytest_rk45.txt >> $L
cut -c1-260 $L
