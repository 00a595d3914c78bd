#!/bin/bash
# Scan workspace placement: cudaMalloc vs a virtual range of its own (H264GPU_WS7_VMM=1), in the
# allocation order that is slow with cudaMalloc (--stage 1) and in the fast one (--stage 6).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-ws7}
L=gpurun_out/${T}_placement.log
: > $L
for stage in 1 6; do for vmm in 0 1; do
echo "== --stage $stage H264GPU_WS7_VMM=$vmm" | tee -a $L
H264GPU_DEBUG_WS=1 H264GPU_WS7_VMM=$vmm timeout 600 python scripts/scan_ab.py --stage $stage --order strip 2>&1 | grep -v "^$" | cut -c1-200 | tail -6 | tee -a $L
done; done
