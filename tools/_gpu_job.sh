timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/b.err
timeout 600 python bench.py --workload c1 --scenes-per-gpu 1024 --steps 20 --warmup 3 > gpurun_out/r02_bench_c1.json 2>/dev/null
timeout 600 python bench.py --workload c2 --scenes-per-gpu 128 --steps 20 --warmup 3 > gpurun_out/r02_bench_c2.json 2>/dev/null
timeout 900 python bench.py --workload c4 --scenes-per-gpu 16 --e2e-scenes 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_c4.json 2>/dev/null
python - <<'PY'
import json
for w in ("1gpu","c1","c2","c4"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{w}.json").read().strip().splitlines()[-1])
        print(w, "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "cpu", round(d["cpu_baseline"]["value"],1), d["cpu_baseline"]["kind"], "pipe", round(d["roofline"]["pipeline"]["frac"],3), "fp32", round(d["roofline"]["fp32"]["frac"],3), d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "traffic", d["roofline"]["traffic"], d["clocks"])
    except Exception as e: print(w,"ERR",e)
PY
