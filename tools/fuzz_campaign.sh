#!/bin/bash
# Longer randomised parity campaign (GPU vs the float64 oracle), default and opt-in moving-event kernels:
#   tools/fuzz_campaign.sh 100-299 > profiles/r02_fuzz_campaign.txt
range=${1:-100-199}
for mode in 0 2 1; do
  echo "== ALR_FUSED=$mode seeds $range"
  ALR_FUSED=$mode ALR_FUZZ_SEEDS=$range python -m pytest tests/test_gpu_fuzz.py -m gpu -q -k random_batch 2>&1 | tail -3
done
