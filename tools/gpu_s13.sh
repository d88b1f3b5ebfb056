#!/bin/bash
# round 2, GPU session 13: source-level ncu captures of the scan at dense candidates (C3 regexp, C2 phrase; every second row a candidate); part reader timing
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_substr_scan -s 2 -c 1 -o gpurun_out/prof_scan_dense_C3_r02 -f python bench.py --workload C3 --rows 30000000 --hit-row-permille 500 --vocab-focus 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/s13_ncu_dense3.log 2>&1; tail -1 gpurun_out/s13_ncu_dense3.log | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_substr_scan -s 2 -c 1 -o gpurun_out/prof_scan_dense_C2_r02 -f python bench.py --workload C2 --rows 30000000 --hit-row-permille 500 --vocab-focus 2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/s13_ncu_dense2.log 2>&1; tail -1 gpurun_out/s13_ncu_dense2.log | cut -c1-160
timeout 900 python tools/part_bench.py --rows 60000000 --out gpurun_out/part_bench_r02.json > gpurun_out/s13_part.log 2> gpurun_out/s13_part.err; tail -6 gpurun_out/s13_part.err | cut -c1-300
ls -la gpurun_out | grep -v s11 | head -30
