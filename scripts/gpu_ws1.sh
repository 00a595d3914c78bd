#!/bin/bash
# slow workspace placement reproduced on one GPU (--stage 1): do the launch knobs or a spacer move it?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export H264GPU_DEBUG_WS=1
O=gpurun_out/ws1.log
: > $O
A="--order strip --steps 10"
run() { echo "== $*" >> $O; env "$@" timeout 100 python scripts/scan_ab.py $A --stage 1 2>&1 | grep -E "ws7|gen7_strip \{" | cut -c1-90 >> $O; }
echo "== stage 6 (fast order)" >> $O; timeout 100 python scripts/scan_ab.py $A --stage 6 2>&1 | grep -E "ws7|gen7_strip \{" | cut -c1-90 >> $O
run X=1
run H264GPU_SCAN7_REGIONS=370
run H264GPU_SCAN7_REGIONS=5920
run H264GPU_SCAN7_NAP=32
run H264GPU_SCAN7_NAP=4096
run H264GPU_SCAN7_PF=0
run H264GPU_SCAN7_PF=4
run H264GPU_WS7_PAD_MB=2
run H264GPU_WS7_PAD_MB=64
run H264GPU_WS7_PAD_MB=1024
run H264GPU_WS7_PAD_MB=4100
cat $O
