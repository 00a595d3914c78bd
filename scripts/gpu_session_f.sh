#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest cabac+cavlc gpu"; timeout 900 python -m pytest tests/test_cabac.py tests/test_cavlc.py -m gpu -q 2>&1 | tail -4
echo "== parse bench"; timeout 900 python scripts/bench_cabac.py 2>&1 | tail -12 | tee gpurun_out/f_cabac3.log
