timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/tmp.json 2> gpurun_out/tmp.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/tmp.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), round(d["value"]), {k:round(v,2) for k,v in d["roofline"]["kernel_ms"].items()}); e=d["e2e"]
print("e2e", round(e["value"]), {k:(round(v["value"]) if "value" in v else v) for k,v in e.items() if isinstance(v,dict)})
print(d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic"], d["config"]["partition"])
PY
for w in c1 c2 c4; do
  n=128; [ $w = c1 ] && n=1024; [ $w = c4 ] && n=16
  timeout 600 python bench.py --workload $w --scenes-per-gpu $n --e2e-scenes 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/tmp_$w.json 2>/dev/null
  python - $w <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/tmp_{sys.argv[1]}.json").read().strip().splitlines()[-1]); print(sys.argv[1], round(d["ms_per_step"],3), round(d["value"]))
PY
done
