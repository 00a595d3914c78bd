#!/bin/bash
# Does the scan workspace placement depend on the device index or on other ranks running? (4-GPU box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export H264GPU_DEBUG_WS=1
O=gpurun_out/ws4.log
: > $O
A="--stage 6 --order strip --steps 10 --regrow 2 --threads 16"
echo "== alone, dev 0" >> $O; timeout 120 python scripts/scan_ab.py $A --dev 0 >> $O 2>&1
echo "== alone, dev 3" >> $O; timeout 120 python scripts/scan_ab.py $A --dev 3 >> $O 2>&1
echo "== four at once" >> $O
for d in 0 1 2 3; do timeout 150 python scripts/scan_ab.py $A --dev $d > gpurun_out/ws4_dev$d.log 2>&1 & done
wait
cat gpurun_out/ws4_dev[0-3].log >> $O
grep -v Warning $O | tail -60
