#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k3b}
L=gpurun_out/${T}_frame_ab.log
: > $L
H264GPU_FRAME_GEN=6 timeout 300 python scripts/frame_ab.py 2>&1 | tail -1 | tee -a $L
H264GPU_FRAME7_NW=1 timeout 300 python scripts/frame_ab.py 2>&1 | tail -1 | tee -a $L
H264GPU_FRAME7_ROWS=6 timeout 300 python scripts/frame_ab.py 2>&1 | tail -1 | tee -a $L
H264GPU_FRAME7_NW=3 timeout 300 python scripts/frame_ab.py 2>&1 | tail -1 | tee -a $L
echo "== ncu frame7"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:frame7_kernel -s 3 -c 1 -f -o gpurun_out/${T}_prof_frame7 python scripts/frame_ab.py --size-mb 256 --steps 2 --hash 0 > gpurun_out/${T}_ncu_frame7.log 2>&1
tail -2 gpurun_out/${T}_ncu_frame7.log
bash scripts/gpu_ws7.sh ${T}
