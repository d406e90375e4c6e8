timeout 900 python -m pytest tests/test_augment.py -m gpu -q 2>&1 | tail -6
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/tmp.json 2> gpurun_out/tmp.err
python - <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/tmp.json").read().strip().splitlines()[-1]); print(round(d["ms_per_step"],3), d["roofline"]["kernel_ms"])
except Exception as e: print("ERR", e, open("gpurun_out/tmp.err").read()[-1500:])
PY
python tools/small_rir_bench.py | tee gpurun_out/r02_small_rir.txt
