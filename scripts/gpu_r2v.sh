#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --size-mb 1024 --steps 5 --mb-frames 120 --cabac-frames 60 --cfg4-streams 2 --frame-mb 256 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
echo "rc=$?"; tail -5 gpurun_out/r2v_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2v_bench.json").read().strip().splitlines()[-1])
    print(json.dumps({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if not isinstance(vv,(dict,str))}) for k,v in d.items() if k not in("config",)},indent=0)[:3500])
except Exception as e: print("parse fail",e)
PY
