#!/bin/bash
# round 2, GPU session 18: scan CTAs meet at a barrier every second tile; launch lists of C4 and C2
mkdir -p gpurun_out
s=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/s18_pytest_all.txt; echo "pytest wall $(( $(date +%s) - s )) s" | tee -a gpurun_out/s18_pytest_all.txt
s=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_C3_r02.json 2> gpurun_out/s18_bench_default.err; echo "default bench rc=$? wall $(( $(date +%s) - s )) s" | tee gpurun_out/s18_bench_wall.txt
tail -1 gpurun_out/bench_C3_r02.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value %.2f G rows/s, %.2f ms/step, kernel frac %.3f (%.2f ms), step frac %.3f, parity %s" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["step_frac_of_peak"], d["parity"] and d["parity"]["ok"])); print("e2e %.1f ms" % d["e2e"]["ms_per_step"]); print("extra", {k: (round(v["value"]/1e9,2), round(v["roofline"]["frac"],3)) if "value" in v else v for k,v in (d.get("extra_workloads") or {}).items()})' 2>&1 | tee gpurun_out/s18_bench_summary.txt
timeout 600 python tools/sweep.py --rows 100000000 --steps 5 --warmup 3 --workloads C2,C3 --out gpurun_out/sweep_c23.json > gpurun_out/s18_sweep.log 2>&1
python - <<'PY' | tee gpurun_out/s18_sweep_summary.txt
import json
for r in json.load(open('gpurun_out/sweep_c23.json')):
    if r['hot_block_permille'] == 1000: print(r['workload'], 'hit', r['hit_row_permille'], '%.1f G rows/s' % (r['rows_per_s'] / 1e9), 'kernel frac', round(r['scan_kernel_frac_of_peak'], 3))
PY
for wl in C4 C2; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${wl}_r02.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_substr_scan -s 3 -c 1 -o gpurun_out/prof_scan_C3_1B_r02 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/s18_ncu_scan.log 2>&1
ls gpurun_out | grep -E "s18|launches_C"
