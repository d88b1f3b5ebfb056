#!/bin/bash
# round 2, GPU session 10: vector-granular candidate queue, tapered launch groups, parallel packing of pageable pieces; part reader timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/s10_pytest_all.txt
e2e_line='import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print("   %s e2e: %.1f ms/step, %.0f M rows/s h2d %.2f GB ok=%s/%s" % (sys.argv[1], e["ms_per_step"], e["value"]/1e6, e["h2d_bytes_per_step"]/1e9, e.get("matched_equals_resident"), e.get("digest_equals_resident")))'
{
for cfg in "0 0" "1 0" "1 4" "1 2"; do
  set -- $cfg
  echo "overlap=$1 group_scale=$2 (0 = tapered default):"
  VLSCAN_ZSTD_OVERLAP=$1 VLSCAN_ZSTD_GROUP_SCALE=$2 timeout 400 python bench.py --workload C2 --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "$e2e_line" C2
  VLSCAN_ZSTD_OVERLAP=$1 VLSCAN_ZSTD_GROUP_SCALE=$2 timeout 400 python bench.py --workload C3 --rows 100000000 --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "$e2e_line" C3
done
} 2>&1 | tee gpurun_out/s10_decoder.txt
timeout 900 python tools/sweep.py --rows 100000000 --steps 5 --warmup 3 --out gpurun_out/sweep_r02b.json > gpurun_out/s10_sweep.log 2>&1
python - <<'PY' | tee gpurun_out/s10_sweep_summary.txt
import json
for r in json.load(open('gpurun_out/sweep_r02b.json')):
    if r['hot_block_permille'] == 1000:
        print(r['workload'], 'hit', r['hit_row_permille'], 'sel %.4f' % r['selectivity'], '%.1f G rows/s' % (r['rows_per_s'] / 1e9), 'kernel frac', None if r['scan_kernel_frac_of_peak'] is None else round(r['scan_kernel_frac_of_peak'], 3))
PY
s=$(date +%s); timeout 900 python bench.py > gpurun_out/s10_bench_default.json 2> gpurun_out/s10_bench_default.err; echo "default bench rc=$? wall $(( $(date +%s) - s )) s" | tee gpurun_out/s10_bench_wall.txt
tail -1 gpurun_out/s10_bench_default.json | cut -c1-300
timeout 1200 python tools/part_bench.py --rows 220000000 --out gpurun_out/part_bench_r02.json > gpurun_out/s10_part.log 2>&1; tail -5 gpurun_out/s10_part.log | cut -c1-400
