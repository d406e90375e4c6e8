timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "fused" 2>&1 | tail -3
for v in "" _nolead; do
ALR_LIBRARY=$PWD/audiblelight_b200/libalrender$v.so ALR_FUSED=2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/tmp.json 2> gpurun_out/tmp.err
python - "$v" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/tmp.json").read().strip().splitlines()[-1]); print("variant", sys.argv[1] or "leader", round(d["ms_per_step"],3), "fused", round(d["roofline"]["kernel_ms"]["fused"],3))
except Exception as e: print("ERR", e, open("gpurun_out/tmp.err").read()[-1500:])
PY
done
