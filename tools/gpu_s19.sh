#!/bin/bash
# round 2, GPU session 19: per-tile barrier back (session 18 measured the every-second-tile variant slower: 0.799 vs 0.880), k_row_match capped at 64 registers
mkdir -p gpurun_out
s=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/s19_pytest_all.txt; echo "pytest wall $(( $(date +%s) - s )) s" | tee -a gpurun_out/s19_pytest_all.txt
s=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_C3_r02.json 2> gpurun_out/s19_bench_default.err; echo "default bench rc=$? wall $(( $(date +%s) - s )) s" | tee gpurun_out/s19_bench_wall.txt
tail -1 gpurun_out/bench_C3_r02.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value %.2f G rows/s, %.2f ms/step, kernel frac %.3f (%.2f ms), step frac %.3f, parity %s" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["step_frac_of_peak"], d["parity"] and d["parity"]["ok"])); print("e2e %.1f ms" % d["e2e"]["ms_per_step"]); print("extra", {k: (round(v["value"]/1e9,2), round(v["ms_per_step"],2), round(v["roofline"]["frac"],3)) if "value" in v else v for k,v in (d.get("extra_workloads") or {}).items()})' 2>&1 | tee gpurun_out/s19_bench_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_C4_r02.csv python bench.py --workload C4 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_substr_scan -s 3 -c 1 -o gpurun_out/prof_scan_C3_1B_r02 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/s19_ncu_scan.log 2>&1
ls gpurun_out | grep -E "s19|launches_C"
