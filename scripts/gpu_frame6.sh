#!/bin/bash
# K3 gen 6 session: writer parity tests, A/B of frame_kernel (gen 1) vs frame6_kernel at the
# config-5 size, launch list + one full ncu capture of frame6_kernel.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-f6}
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest (writer + scan)"; timeout 600 python -m pytest tests/test_gpu_annexb.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
Q="--size-mb 256 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --frame-mb 1024"
for gen in 1 6 64; do
  H264GPU_FRAME_CTAS=$((gen == 64 ? 4 : 5)) H264GPU_FRAME_GEN=$((gen == 1 ? 1 : 6)) timeout 300 python bench.py $Q > gpurun_out/${TAG}_g$gen.json 2> gpurun_out/${TAG}_g$gen.err
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_g$gen.json').read())
print('frame gen $gen', d['extra']['writer_frame'])"
done
echo "== ncu frame6"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:frame6_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_frame6 python bench.py $Q > gpurun_out/${TAG}_ncu_frame6.log 2>&1
ls -la gpurun_out/${TAG}_*
