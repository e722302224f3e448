#!/bin/bash
# N-GPU run of bench.py on one box (weak-scaling headline + single-chain strong-scaling legs): bash tools/gpu_scale.sh 8
N=${1:-8}
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/r2_scale$N.json 2> gpurun_out/r2_scale$N.err; tail -c 600 gpurun_out/r2_scale$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_scale$N.json").read().strip().splitlines()[-1])
print("value", d["value"], "n_gpus", d["n_gpus"], "e2e", d["e2e"])
for s in d["strong_scaling"]:
    print({k:v for k,v in s.items() if k not in ("sharding","collective")})
PY
