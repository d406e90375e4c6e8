#!/bin/bash
# A/B kernel variants: runs the device-resident benchmark once per library and prints the per-kernel times.
#   tools/ab_variants.sh audiblelight_b200/v_*.so
for lib in "$@"; do
  ALR_LIBRARY=$PWD/$lib python bench.py --no-cpu-baseline --no-e2e --steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['roofline']['kernel_ms'].items()})"
done
