timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/b.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_1gpu.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "cpu", round(d["cpu_baseline"]["value"],1), "pipe", round(d["roofline"]["pipeline"]["frac"],3), "fp32", round(d["roofline"]["fp32"]["frac"],3), d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "traffic", d["roofline"]["traffic"], "alg/launch", d["roofline"]["algorithmic_bytes_per_launch"], d["clocks"], d["gpu_launches"])
PY
