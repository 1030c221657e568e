#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | cut -c1-400; cat gpurun_out/bench_ref.json | cut -c1-300
