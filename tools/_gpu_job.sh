timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/tmp.json 2> gpurun_out/tmp.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/tmp.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3)); 
import pprint; pprint.pprint(d["e2e"])
PY
grep -v INFO gpurun_out/tmp.err | tail -5
