#!/bin/bash
# Writer kernel: gen 6 (default) vs gen 7 (warp-autonomous spans) at 1 GiB (config 5), phase trace
# and one ncu capture of gen 7.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k3}
L=gpurun_out/${T}_frame_ab.log
: > $L
W="--warmup 10 --steps 20"
timeout 300 python scripts/frame_ab.py $W 2>&1 | tail -1 | tee -a $L
H264GPU_FRAME_GEN=8 timeout 300 python scripts/frame_ab.py $W 2>&1 | tail -1 | tee -a $L
H264GPU_FRAME_GEN=7 timeout 300 python scripts/frame_ab.py $W 2>&1 | tail -1 | tee -a $L
H264GPU_FRAME_GEN=7 H264GPU_FRAME7_NBUF=1 timeout 300 python scripts/frame_ab.py $W 2>&1 | tail -1 | tee -a $L
H264GPU_FRAME_GEN=7 H264GPU_FRAME7_ROWS=6 timeout 300 python scripts/frame_ab.py $W 2>&1 | tail -1 | tee -a $L
H264GPU_FRAME_GEN=7 python scripts/frame7_trace.py --size-mb 1024 2>&1 | tail -30 | tee gpurun_out/${T}_trace.txt
rm -f gpurun_out/frame7_trace.bin
echo "== ncu frame7"
H264GPU_FRAME_GEN=7 timeout 600 ncu --set full --clock-control none --import-source on -k regex:frame7_kernel -s 3 -c 1 -f -o gpurun_out/${T}_prof_frame7 python scripts/frame_ab.py --size-mb 1024 --steps 2 --hash 0 > gpurun_out/${T}_ncu_frame7.log 2>&1
tail -2 gpurun_out/${T}_ncu_frame7.log
