#!/bin/bash
# Same-box A/B of two builds of libh264gpu.so (libh264_b200/_ab/{base,new}.so): scan and writer legs.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-ab}
Q="--steps 10 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --frame-mb 1024"
for v in base new base new; do
  cp libh264_b200/_ab/$v.so libh264_b200/libh264gpu.so
  timeout 300 python bench.py $Q > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_$v.json').read())
print('$v scan', round(d['value'],1), 'frame', round(d['extra']['writer_frame']['gb_per_s'],1))"
done
cp libh264_b200/_ab/new.so libh264_b200/libh264gpu.so
echo "== pytest (new)"; timeout 900 python -m pytest tests/test_gpu_annexb.py tests/test_scan_inplace.py -m gpu -x -q 2>&1 | tail -3
