#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --size-mb 1024 --steps 5 --mb-frames 60 --cabac-frames 0 --cfg4-streams 0 --frame-mb 0 > gpurun_out/r2w_bench2.json 2> gpurun_out/r2w_bench2.err
echo "rc=$?"; tail -8 gpurun_out/r2w_bench2.err | cut -c1-300; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2w_bench2.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","n_gpus","merged_table","strong_scaling","scan_only","e2e","roofline"):
        print(k, json.dumps(d.get(k))[:600])
    print(d["config"]["parallelism"], d["config"]["numa"])
except Exception as e: print("parse fail",e)
PY
