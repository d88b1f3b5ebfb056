#!/bin/bash
# round 2, GPU session 14: converged verification (candidates of a vector collected first, one verification loop), one queue atomic per round; sweep; part reader timing
mkdir -p gpurun_out
s=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/s14_pytest_all.txt; echo "pytest wall $(( $(date +%s) - s )) s" | tee -a gpurun_out/s14_pytest_all.txt
timeout 900 python tools/sweep.py --rows 100000000 --steps 5 --warmup 3 --out gpurun_out/sweep_r02.json > gpurun_out/s14_sweep.log 2>&1
python - <<'PY' | tee gpurun_out/s14_sweep_summary.txt
import json
for r in json.load(open('gpurun_out/sweep_r02.json')):
    print(r['workload'], 'hot', r['hot_block_permille'], 'hit', r['hit_row_permille'], 'sel %.4f' % r['selectivity'], '%.1f G rows/s' % (r['rows_per_s'] / 1e9), 'kernel frac', None if r['scan_kernel_frac_of_peak'] is None else round(r['scan_kernel_frac_of_peak'], 3))
PY
s=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_C3_r02.json 2> gpurun_out/s14_bench_default.err; echo "default bench rc=$? wall $(( $(date +%s) - s )) s" | tee gpurun_out/s14_bench_wall.txt
tail -1 gpurun_out/bench_C3_r02.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value %.2f G rows/s, %.2f ms/step, kernel frac %.3f (%.2f ms), step frac %.3f, parity %s" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["step_frac_of_peak"], d["parity"] and d["parity"]["ok"])); print("e2e", json.dumps(d["e2e"])[:300]); print("bloom-first", json.dumps(d.get("e2e_bloom_first_staging"))[:700]); print("extra", {k: (round(v["value"]/1e9,2), round(v["roofline"]["frac"],3)) if "value" in v else v for k,v in (d.get("extra_workloads") or {}).items()})' 2>&1 | tee gpurun_out/s14_bench_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_substr_scan -s 3 -c 1 -o gpurun_out/prof_scan_C3_1B_r02 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/s14_ncu_scan.log 2>&1; tail -1 gpurun_out/s14_ncu_scan.log | cut -c1-120
timeout 900 python tools/part_bench.py --rows 60000000 --out gpurun_out/part_bench_r02.json > gpurun_out/s14_part.log 2> gpurun_out/s14_part.err; tail -6 gpurun_out/s14_part.err | cut -c1-400
ls gpurun_out | grep s14
