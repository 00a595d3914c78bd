#!/bin/bash
# Round-2 closing session: parity tests, smoke, bench (both arms), K5 knobs, launch list, K5 capture
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02f}
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== K5 knobs"
timeout 300 python scripts/bench_parse.py --cabac 1 --frames 250 --copies 4 --lanes auto,0,1,2,3 2>&1 | tee gpurun_out/${TAG}_parse_cabac.log
timeout 300 python scripts/bench_parse.py --cabac 1 --frames 250 --lanes auto,1 2>&1 | tee -a gpurun_out/${TAG}_parse_cabac.log
echo "== bench"; timeout 1200 python bench.py 2> gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/${TAG}_bench_ref.err > gpurun_out/${TAG}_bench_ref.json; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
SHORT="--size-mb 1024 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 120 --cabac-frames 60 --cfg4-streams 0 --frame-mb 256"
echo "== ncu launch list"
timeout 600 python bench.py $SHORT > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $SHORT > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "== ncu: cabac"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cabac_parse -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_cabac python bench.py $SHORT > gpurun_out/${TAG}_ncu_cabac.log 2>&1
if [ -f gpurun_out/${TAG}_prof_cabac.ncu-rep ]; then
  python tools/ncu_summary.py gpurun_out/${TAG}_prof_cabac.ncu-rep gpurun_out/${TAG}_cabac_kernel_raw.csv > /dev/null 2>&1
  python tools/ncu_lines.py gpurun_out/${TAG}_prof_cabac.ncu-rep 40 gpurun_out/${TAG}_cabac_kernel_source_hotspots.csv > /dev/null 2>&1
  rm -f gpurun_out/${TAG}_prof_cabac.ncu-rep
fi
ls -la gpurun_out/${TAG}_*
