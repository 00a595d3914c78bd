#!/bin/bash
# last GPU call of round 2: short bench line of the final build + a few launch knobs with 8 ticket counters
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --cabac-frames 0 --cfg4-streams 0 --mb-frames 60 --frame-mb 0 --no-reader --e2e-steps 2 > gpurun_out/r02g_bench_short.json 2> gpurun_out/r02g_bench_short.err
cut -c1-250 gpurun_out/r02g_bench_short.json
O=gpurun_out/r02g_knobs.log
: > $O
run() { echo "== $*" >> $O; env "$@" timeout 60 python scripts/scan_ab.py --order strip --steps 10 --stage 6 2>&1 | grep -E "gen7_strip \{" | cut -c1-100 >> $O; }
run H264GPU_SCAN7_REGIONS=740
run H264GPU_SCAN7_REGIONS=2044
run H264GPU_SCAN7_REGIONS=370
run H264GPU_SCAN7_PF=2
cat $O
