#!/bin/bash
# A/B of the writer kernel's resident-CTA budget + the scan kernel without trace marks.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-fab}
timeout 300 python -m pytest tests/test_gpu_annexb.py tests/test_scan_inplace.py -m gpu -x -q 2>&1 | tail -3
Q="--size-mb 256 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --frame-mb 1024"
for c in 4 5 6; do
  H264GPU_FRAME_CTAS=$c timeout 300 python bench.py $Q > gpurun_out/${TAG}_f$c.json 2> gpurun_out/${TAG}_f$c.err
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_f$c.json').read())
print('frame ctas $c', d['extra']['writer_frame']['gb_per_s'], d['extra']['writer_frame']['roofline_frac'])"
done
Q2="--steps 10 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --frame-mb 0"
timeout 300 python bench.py $Q2 > gpurun_out/${TAG}_scan.json 2> gpurun_out/${TAG}_scan.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_scan.json').read())
print('scan', d['value'], d['roofline']['frac'], d['roofline']['traffic'])"
