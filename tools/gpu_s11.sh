#!/bin/bash
# round 2, GPU session 11 (after the container was re-created): full GPU suite on the current build, the default bench line, e2e phases, launch list + ncu capture at 1 B rows
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | tee gpurun_out/s11_gpu.txt
lscpu | grep -E "Model name|^CPU\(s\)|NUMA node\(s\)" | tee -a gpurun_out/s11_gpu.txt
s=$(date +%s); timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/s11_pytest_all.txt; echo "pytest wall $(( $(date +%s) - s )) s" | tee -a gpurun_out/s11_pytest_all.txt
s=$(date +%s); timeout 900 python bench.py > gpurun_out/bench_C3_r02.json 2> gpurun_out/s11_bench_default.err; echo "default bench rc=$? wall $(( $(date +%s) - s )) s" | tee gpurun_out/s11_bench_wall.txt
tail -1 gpurun_out/bench_C3_r02.json | cut -c1-1500
e2e_line='import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print("   %s e2e: %.1f ms/step, %.0f M rows/s h2d %.2f GB ok=%s/%s" % (sys.argv[1], e["ms_per_step"], e["value"]/1e6, e["h2d_bytes_per_step"]/1e9, e.get("matched_equals_resident"), e.get("digest_equals_resident")))'
{
VLSCAN_DEBUG_TIMING=1 VLSCAN_ZSTD_OVERLAP=0 timeout 400 python bench.py --workload C3 --rows 100000000 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/s11_t3.err
echo "phases C3 (serial):"; grep "vlscan zstd\]" gpurun_out/s11_t3.err | tail -3
VLSCAN_DEBUG_TIMING=1 VLSCAN_ZSTD_OVERLAP=0 timeout 400 python bench.py --workload C2 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/s11_t2.err
echo "phases C2 (serial):"; grep "vlscan zstd\]" gpurun_out/s11_t2.err | tail -3
timeout 400 python bench.py --workload C2 --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | tee gpurun_out/bench_C2_r02.json | python -c "$e2e_line" C2
} 2>&1 | tee gpurun_out/s11_decoder.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/s11_ncu_launches.log 2>&1; tail -2 gpurun_out/s11_ncu_launches.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_substr_scan -s 3 -c 1 -o gpurun_out/prof_scan_C3_1B_r02 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/s11_ncu_scan.log 2>&1; tail -2 gpurun_out/s11_ncu_scan.log | cut -c1-200
ls -la gpurun_out | head -40
