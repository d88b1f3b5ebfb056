#!/bin/bash
# round 2, GPU session 4: rewritten k_execute + two-stream decode pipeline + smaller launch groups + bloom skipping: parity, phase times, e2e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zstd.py tests/test_gpu_zz_part.py tests/test_gpu_zzz_workers.py -x -q 2>&1 | tail -8 | tee gpurun_out/s4_pytest_zstd.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/s4_pytest_all.txt
for wl in C2 C3; do
  VLSCAN_DEBUG_TIMING=1 timeout 400 python bench.py --workload $wl --rows 100000000 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extra > gpurun_out/s4_timing_$wl.json 2> gpurun_out/s4_timing_$wl.err
  grep "vlscan upload\|vlscan zstd" gpurun_out/s4_timing_$wl.err | tail -4
  timeout 400 python bench.py --workload $wl --rows 100000000 --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; print('$wl e2e: %.1f ms/step, %.0f M rows/s, h2d %.2f GB, matched_ok %s digest_ok %s' % (e['ms_per_step'], e['value']/1e6, e['h2d_bytes_per_step']/1e9, e.get('matched_equals_resident'), e.get('digest_equals_resident')))"
done 2>&1 | tee gpurun_out/s4_e2e.txt
# ncu: DRAM traffic of the dominant kernel at the bench's launch size (1 B rows), and the launch list of 2 steps
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_substr_scan -s 3 -c 1 -o gpurun_out/prof_scan_r02 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/s4_ncu_scan.log 2>&1; tail -2 gpurun_out/s4_ncu_scan.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/s4_ncu_launches.log 2>&1; tail -1 gpurun_out/s4_ncu_launches.log | cut -c1-120
