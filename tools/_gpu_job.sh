TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 1 2 4 8; do
  $TR --nproc-per-node $n --master-port $((29500+n)) tools/pcie_ranks_probe.py 2>/dev/null | grep "^ranks" >> gpurun_out/r02_pcie_ranks.txt
done
cat gpurun_out/r02_pcie_ranks.txt
$TR --nproc-per-node 8 --master-port 29600 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/b8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_8gpu.json").read().strip().splitlines()[-1])
print("8gpu value", round(d["value"]), "ms", round(d["ms_per_step"],3)); e=d["e2e"]
print("e2e", round(e["value"]), "all", round(e["all_outputs"]["value"]), "pcm", round(e["dataset_mode"]["value"]), "2ctx", e["two_contexts"].get("value"), "obj", e["objects_mode"].get("value"))
PY
grep -v INFO gpurun_out/b8.err | tail -5
nproc; free -g | head -2
