#!/bin/bash
# Writer kernel gen 7 vs gen 6: parity tests, then the A/B at 1 GiB (config 5), then one ncu capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k3}
echo "== gpu tests (writer)"; timeout 900 python -m pytest tests/test_gpu_annexb.py tests/test_gpu_vs_reference.py -m gpu -x -q -k "frame or writer or insert" 2>&1 | tail -6 | tee gpurun_out/${T}_pytest.log
L=gpurun_out/${T}_frame_ab.log
: > $L
H264GPU_FRAME_GEN=6 timeout 300 python scripts/frame_ab.py 2>&1 | tail -1 | tee -a $L
for nw in 1 2 3; do
H264GPU_FRAME7_NW=$nw timeout 300 python scripts/frame_ab.py 2>&1 | tail -1 | tee -a $L
done
H264GPU_FRAME7_ROWS=6 timeout 300 python scripts/frame_ab.py 2>&1 | tail -1 | tee -a $L
H264GPU_FRAME7_ROWS=4 timeout 300 python scripts/frame_ab.py 2>&1 | tail -1 | tee -a $L
H264GPU_FRAME7_ROWS=6 H264GPU_FRAME7_NW=3 timeout 300 python scripts/frame_ab.py 2>&1 | tail -1 | tee -a $L
echo "== ncu frame7"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:frame7_kernel -s 3 -c 1 -f -o gpurun_out/${T}_prof_frame7 python scripts/frame_ab.py --size-mb 256 --steps 2 --hash 0 > gpurun_out/${T}_ncu_frame7.log 2>&1
tail -2 gpurun_out/${T}_ncu_frame7.log
echo "== K4 residual elements per vote"
for b in 1 4; do
H264GPU_CAVLC_REPEAT=$b timeout 600 python scripts/bench_parse.py --cabac 0 --frames 1000 --lanes auto,3 2>&1 | sed "s/^/repeat=$b /" | tee -a gpurun_out/${T}_parse_cavlc.log
done
