#!/bin/bash
# One GPU session that re-validates everything that changed on the host side since the last measured commit (multi-threaded header walk,
# part reader) and refreshes the numbers.  Run from the repo root under gpurun, e.g.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_check.sh'
# Every step has its own timeout; outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
# parity first: the whole GPU suite (the new ones sort last: test_gpu_zstd::..._host_threads, test_gpu_zz_part)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
# where the end-to-end time goes: header walk / work lists / table staging / decode phases, per upload
VLSCAN_DEBUG_TIMING=1 timeout 400 python bench.py --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline > gpurun_out/bench_timing.json 2> gpurun_out/bench_timing.err
grep "vlscan upload\|vlscan zstd" gpurun_out/bench_timing.err | tail -8
# the walk on 0 / 1 / 16 / 32 threads (same tables, different end-to-end time)
for t in 0 1 16 32; do
    VLSCAN_HOST_THREADS=$t timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('host threads $t: e2e %.1f ms, %.0f M rows/s, matches resident: %s' % (d['e2e']['ms_per_step'], d['e2e']['value'] / 1e6, d['e2e'].get('matched_equals_resident')))"
done 2>&1 | tee gpurun_out/host_threads.txt
# the numbers of record
timeout 500 python bench.py > gpurun_out/BENCH_final.json 2> gpurun_out/BENCH_final.err; tail -1 gpurun_out/BENCH_final.json | cut -c1-150
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/BENCH_ref_final.json 2>/dev/null; tail -1 gpurun_out/BENCH_ref_final.json | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_e2e.csv python bench.py --rows 30000000 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_e2e.log 2>&1; tail -1 gpurun_out/ncu_e2e.log | cut -c1-100
