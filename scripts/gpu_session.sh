#!/bin/bash
# One GPU session: parity tests, smoke, full bench (both arms), ncu launch list and
# full captures of the three kernels.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-s}
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/${TAG}_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
echo "== bench"; timeout 900 python bench.py 2> gpurun_out/${TAG}_bench.err > gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/${TAG}_bench_ref.err > gpurun_out/${TAG}_bench_ref.json; cut -c1-400 gpurun_out/${TAG}_bench_ref.json
SHORT="--size-mb 1024 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 120 --cabac-frames 60 --frame-mb 256"
echo "== ncu launch list"
timeout 600 python bench.py $SHORT > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $SHORT > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "== ncu full: scan, cavlc, cabac, frame"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan6_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_scan python bench.py $SHORT > gpurun_out/${TAG}_ncu_scan.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cavlc_parse -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_cavlc python bench.py $SHORT > gpurun_out/${TAG}_ncu_cavlc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cabac_parse -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_cabac python bench.py $SHORT > gpurun_out/${TAG}_ncu_cabac.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:frame6_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_frame python bench.py $SHORT > gpurun_out/${TAG}_ncu_frame.log 2>&1
ls -la gpurun_out/${TAG}_*
echo "== dram traffic of the scan kernel at the bench size"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:scan6_kernel -s 3 -c 1 --csv --log-file gpurun_out/${TAG}_traffic.csv python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu --mb-frames 2 --cabac-frames 0 --frame-mb 0 > gpurun_out/${TAG}_traffic_bench.json 2> gpurun_out/${TAG}_traffic.err
python - <<PY
import csv, json
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_traffic.csv")) if len(r) > 10 and r[0].isdigit()]
d = json.loads(open("gpurun_out/${TAG}_traffic_bench.json").read())
val = {r[-3]: float(r[-1].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[-2]] for r in rows}
out = {"kernel": "annexb6::scan6_kernel<8,strip>", "bytes_in": d["config"]["bytes_per_gpu"],
       "dram_bytes_read": val["dram__bytes_read.sum"], "dram_bytes_write": val["dram__bytes_write.sum"],
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, 4th launch of bench.py default workload"}
json.dump(out, open("gpurun_out/${TAG}_traffic.json", "w"), indent=1)
print(out)
PY
