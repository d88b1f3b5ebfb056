#!/bin/bash
# round 2, GPU session 2: queue-based verification (parity + sweeps), the new default bench line, decoder timing + ncu captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gen.py -x -q 2>&1 | tail -4 | tee gpurun_out/s2_pytest.txt
run() { # name, args...
  local name=$1; shift
  timeout 300 python bench.py "$@" --no-e2e --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$name: %.2f ms/step kernel %.3f ms frac %.3f step_frac %.3f launches/step %d matched %d' % (d['ms_per_step'], r['kernel_ms_per_launch'], r['frac'], d['step_frac_of_peak'], d['gpu_launches']/d['steps'], d['rows_matched_per_gpu']))"
}
{
for h in 1 60 100 500 1000; do run "C2 100M hit $h" --workload C2 --steps 5 --hit-row-permille $h; done
for h in 1 60 500; do run "C1 100M hit $h" --workload C1 --rows 100000000 --steps 5 --hit-row-permille $h; done
for h in 60 500 1000; do run "C3 100M hit $h" --workload C3 --rows 100000000 --steps 5 --hit-row-permille $h; done
run "C4 125M hit 60" --workload C4 --steps 5
} 2>&1 | tee gpurun_out/s2_summary.txt
# the default line, timed as the driver would run it
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/s2_bench_default.json 2> gpurun_out/s2_bench_default.err; tail -1 gpurun_out/s2_bench_default.json | cut -c1-400; grep -E "Elapsed|Maximum resident" gpurun_out/s2_bench_default.err
# decoder: per-phase device times of the e2e leg (C2, 100M rows) and full ncu captures of its three big kernels at 30M rows
VLSCAN_DEBUG_TIMING=1 timeout 400 python bench.py --workload C2 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extra > gpurun_out/s2_timing.json 2> gpurun_out/s2_timing.err
grep "vlscan upload\|vlscan zstd" gpurun_out/s2_timing.err | tail -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_seq_decode|k_huf_decode|k_execute" -c 6 -o gpurun_out/prof_zstd_r02 python bench.py --workload C2 --rows 30000000 --steps 1 --warmup 1 --e2e-rows 30000000 --e2e-steps 1 --no-cpu-baseline --no-extra > gpurun_out/s2_ncu_zstd.log 2>&1; tail -2 gpurun_out/s2_ncu_zstd.log | cut -c1-200
