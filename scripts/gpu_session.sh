#!/bin/bash
# One GPU session (round 2): parity tests, smoke, full bench (both arms), ncu launch list, full
# captures of the scan / writer / parse kernels, DRAM traffic of the scan kernel at the bench
# size, phase trace.  Everything lands in gpurun_out/ under the given tag.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== bench"; timeout 1500 python bench.py 2> gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/${TAG}_bench_ref.err > gpurun_out/${TAG}_bench_ref.json; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
echo "== phase trace"; timeout 300 python scripts/scan7_trace.py --size-mb 4096 2>&1 | tail -8 | tee gpurun_out/${TAG}_scan7_phase_trace.txt
rm -f gpurun_out/scan7_trace.npy gpurun_out/scan7_trace.bin
SHORT="--size-mb 1024 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 120 --cabac-frames 60 --cfg4-streams 0 --frame-mb 256"
echo "== ncu launch list"
timeout 600 python bench.py $SHORT > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $SHORT > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "== ncu full: scan, scan-only, frame, cavlc, cabac"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan7_kernel -s 6 -c 2 -f -o gpurun_out/${TAG}_prof_scan python bench.py $SHORT > gpurun_out/${TAG}_ncu_scan.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan7_only -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_scanonly python bench.py $SHORT > gpurun_out/${TAG}_ncu_scanonly.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frame6_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_frame python bench.py $SHORT > gpurun_out/${TAG}_ncu_frame.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cavlc_steps -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_cavlc python bench.py $SHORT > gpurun_out/${TAG}_ncu_cavlc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cabac_parse -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_cabac python bench.py $SHORT > gpurun_out/${TAG}_ncu_cabac.log 2>&1
# summaries made here (gpurun brings back at most 64 MiB: the reports themselves stay on the box)
for k in scan scanonly frame cavlc cabac; do
  if [ -f gpurun_out/${TAG}_prof_$k.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/${TAG}_prof_$k.ncu-rep gpurun_out/${TAG}_${k}_kernel_raw.csv > /dev/null 2>&1
    python tools/ncu_lines.py gpurun_out/${TAG}_prof_$k.ncu-rep 40 gpurun_out/${TAG}_${k}_kernel_source_hotspots.csv > /dev/null 2>&1
    rm -f gpurun_out/${TAG}_prof_$k.ncu-rep
  fi
done
ls -la gpurun_out/${TAG}_*
echo "== dram traffic of the scan kernels at the bench size"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:scan7_kernel -s 6 -c 2 --csv --log-file gpurun_out/${TAG}_traffic.csv python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --cfg4-streams 0 --frame-mb 0 --no-reader > gpurun_out/${TAG}_traffic_bench.json 2> gpurun_out/${TAG}_traffic.err
python - <<PY
import csv, json
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_traffic.csv")) if len(r) > 10 and r[0].isdigit()]
d = json.loads(open("gpurun_out/${TAG}_traffic_bench.json").read().strip().splitlines()[-1])
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = {}
for r in rows:  # main pass + second pass of one step
    tot[r[-3]] = tot.get(r[-3], 0) + float(r[-1].replace(",", "")) * mul[r[-2]]
out = {"kernel": "annexb7::scan7_kernel<8> (main pass + second pass of one step)", "bytes_in": d["config"]["bytes_per_gpu"],
       "dram_bytes_read": tot["dram__bytes_read.sum"], "dram_bytes_write": tot["dram__bytes_write.sum"],
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, both scan7_kernel launches of the 4th step of bench.py's default workload"}
json.dump(out, open("gpurun_out/${TAG}_traffic.json", "w"), indent=1)
print(out)
PY
echo "== CAVLC / CABAC parse at 16000 slices vs slices per warp"
timeout 600 python scripts/bench_parse.py --cabac 0 --frames 1000 --lanes auto,2,3,5 2>&1 | tee gpurun_out/${TAG}_parse_cavlc.log
timeout 600 python scripts/bench_parse.py --cabac 1 --frames 250 --lanes auto,0,2 2>&1 | tee gpurun_out/${TAG}_parse_cabac.log
echo "== writer: gen 6 vs gen 7, phase trace"
bash scripts/gpu_k3.sh ${TAG}
if [ -f gpurun_out/${TAG}_prof_frame7.ncu-rep ]; then
  python tools/ncu_summary.py gpurun_out/${TAG}_prof_frame7.ncu-rep gpurun_out/${TAG}_frame7_kernel_raw.csv > /dev/null 2>&1
  python tools/ncu_lines.py gpurun_out/${TAG}_prof_frame7.ncu-rep 40 gpurun_out/${TAG}_frame7_kernel_source_hotspots.csv > /dev/null 2>&1
  rm -f gpurun_out/${TAG}_prof_frame7.ncu-rep
fi
echo "== reader phases"
timeout 600 python scripts/reader_phases.py 2>&1 | tail -12 | tee gpurun_out/${TAG}_reader_phases.txt
