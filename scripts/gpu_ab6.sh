#!/bin/bash
# A/B of the in-place scan kernel generations (6 vs 5) + parity tests.  Output in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-ab6}
echo "== pytest gpu (in-place scan)"; H264GPU_INPLACE_GEN=${PYGEN:-6} timeout 600 python -m pytest tests/test_scan_inplace.py tests/test_gpu_annexb.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
Q="--steps 10 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --frame-mb 0"
for v in ${VARIANTS:-6:5 5:5}; do
  set -- ${v//:/ }
  echo "== bench gen $1 ctas $2 flags ${3:-0}"
  H264GPU_SCAN6_CTAS=$2 H264GPU_SCAN6_FLAGS=${3:-0} timeout 600 python bench.py $Q --scan-gen $1 > gpurun_out/${TAG}_g$1_c$2_f${3:-0}.json 2> gpurun_out/${TAG}_g$1_c$2_f${3:-0}.err
  tail -2 gpurun_out/${TAG}_g$1_c$2_f${3:-0}.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_g$1_c$2_f${3:-0}.json').read())
print('value', d['value'], 'frac', d['roofline']['frac'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
