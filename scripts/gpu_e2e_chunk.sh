#!/bin/bash
# End-to-end leg (h264gpu_split_strip_host) vs the pipeline chunk size (H264GPU_CHUNK_MB).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-e2e}
Q="--steps 3 --warmup 3 --e2e-steps 3 --no-cpu --mb-frames 2 --cabac-frames 0 --frame-mb 0"
for mb in 128 32 64 256 16; do
  H264GPU_CHUNK_MB=$mb timeout 300 python bench.py $Q > gpurun_out/${TAG}_$mb.json 2> gpurun_out/${TAG}_$mb.err
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_$mb.json').read())
print('chunk $mb MB: e2e', round(d['e2e']['value'],2), 'GB/s', round(d['e2e']['ms_per_step'],1), 'ms')"
done
