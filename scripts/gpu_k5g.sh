#!/bin/bash
# K5: source-level counters of the current build at 4000 slices (cheap capture)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-k5g}
timeout 600 ncu --section SourceCounters --section WarpStateStats --section LaunchStats --section Occupancy --section SchedulerStats --section InstructionStats \
  --import-source on --clock-control none -k regex:cabac_parse_kernel -s 1 -c 1 -o gpurun_out/${T}_prof_cabac -f \
  python scripts/bench_parse.py --cabac 1 --frames 250 --lanes 0 --reps 1 > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
