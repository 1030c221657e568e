#!/bin/bash
# A/B of library variants on one GPU (gpurun -- 'bash scripts/gpu_ab.sh "<variant>:<CTAs/SM>:<mode>[:rk45]" ...').
# A variant is a build `make -C raytrace_cpu_b200/csrc VARIANT=x EXTRA="-DKT_...=1"` -> libkerrtrace_x.so ("" = the default library);
# every run prints parity statistics against the oracle, a digest of all per-ray results (identical digests = bit-identical
# builds) and the time of a 1.5e7-ray fused launch (scripts/gpu_variant.py).  ~7 s of GPU time per run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/ab.txt
: > $L
for spec in "$@"; do
  IFS=: read -r lib mb mode extra <<< "$spec"
  echo "### lib=${lib:-default} CTAs/SM=${mb:-0} ${mode:-exact} $extra" >> $L
  KT_LIB=$lib KT_MIN_BLOCKS=${mb:-0} timeout 300 python scripts/gpu_variant.py ${mode:-exact} $extra >> $L 2>&1
done
cat $L
