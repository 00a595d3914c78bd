#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k3e2e}
timeout 900 python -m pytest tests/test_gpu_annexb.py -m gpu -x -q -k "frame or writer" 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
timeout 900 python bench.py --size-mb 256 --steps 3 --warmup 3 --e2e-steps 1 --mb-frames 2 --cabac-frames 0 --cfg4-streams 0 --frame-mb 1024 --no-reader > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print(json.dumps(d["extra"]["writer_frame"], indent=1))
PY
