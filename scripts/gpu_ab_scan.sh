#!/bin/bash
# same-box A/B of the in-place scan kernel: current tree vs the tree exported to _ab_old/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-ab}
for i in 1 2; do
  echo "== new $i"; timeout 300 python scripts/scan_ab.py --size-mb 4096 --steps 10 2>&1 | grep gen7_ | tee -a gpurun_out/${T}_new.log
  echo "== old $i"; (cd _ab_old && timeout 300 python scripts/scan_ab.py --size-mb 4096 --steps 10 2>&1 | grep gen7_ | tee -a ../gpurun_out/${T}_old.log)
done
echo "== bench scan only legs"; timeout 600 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --cfg4-streams 0 --frame-mb 0 --no-reader 2>/dev/null | cut -c1-200 | tee gpurun_out/${T}_bench.log
echo "== gpu tests of the new K4"; timeout 900 python -m pytest tests/test_cavlc.py tests/test_host_lib_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${T}_pytest.log
echo "== K4 sweep"
for pl in 1 2; do
H264GPU_CAVLC_PER_LANE=$pl timeout 600 python scripts/bench_parse.py --cabac 0 --frames 1000 --lanes auto,2,3,4,5 2>&1 | sed "s/^/per_lane=$pl /" | tee -a gpurun_out/${T}_parse_cavlc.log
done
H264GPU_CAVLC_GEN=1 timeout 600 python scripts/bench_parse.py --cabac 0 --frames 1000 --lanes auto 2>&1 | sed "s/^/gen1 /" | tee -a gpurun_out/${T}_parse_cavlc.log
timeout 600 python scripts/bench_parse.py --cabac 0 --frames 300 --slices-per-frame 1 --lanes auto,0 2>&1 | tee -a gpurun_out/${T}_parse_cavlc.log
H264GPU_CAVLC_GEN=1 timeout 600 python scripts/bench_parse.py --cabac 0 --frames 300 --slices-per-frame 1 --lanes auto 2>&1 | sed "s/^/gen1 /" | tee -a gpurun_out/${T}_parse_cavlc.log
