# usage: bash tools/scale_sweep.sh "1 2 4 8"     (driver-style invocations; N > 1 = frame-window sharding + replicas)
for n in $1; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-ref-kernel > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print("N=$n", d["config"]["parallelism"][:14], "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["scaling"], "replicas", d.get("replicas", {}).get("value"))
except Exception as e:
    print("N=$n failed", e)
PY
done
