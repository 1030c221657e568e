#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cold_paths or nan_centre or deterministic" > gpurun_out/pytest_cold.txt 2>&1; tail -5 gpurun_out/pytest_cold.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --rays-per-gpu 2e7 --e2e-rays 1e7 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; python -c "
import json; d=json.load(open('gpurun_out/bench_small.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['other_math_mode']['value'])"
