#!/bin/bash
# round 2, GPU session 6: rewritten decoder kernels (quad sequence decoder, container-reload Huffman, batched executor): parity, phase times, e2e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zstd.py tests/test_gpu_zzzz_time.py -x -q 2>&1 | tail -12 | tee gpurun_out/s6_pytest_zstd.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/s6_pytest_all.txt
{
for cfg in "0 4" "1 4" "0 1" "1 1"; do
  set -- $cfg
  VLSCAN_ZSTD_OVERLAP=$1 VLSCAN_ZSTD_GROUP_SCALE=$2 VLSCAN_DEBUG_TIMING=1 timeout 400 python bench.py --workload C2 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/s6_t.err
  echo "overlap=$1 group_scale=$2:"; grep "vlscan zstd\] [0-9]" gpurun_out/s6_t.err | tail -1
  VLSCAN_ZSTD_OVERLAP=$1 VLSCAN_ZSTD_GROUP_SCALE=$2 timeout 400 python bench.py --workload C2 --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; print('   C2 e2e: %.1f ms/step, %.0f M rows/s h2d %.2f GB ok=%s/%s' % (e['ms_per_step'], e['value']/1e6, e['h2d_bytes_per_step']/1e9, e.get('matched_equals_resident'), e.get('digest_equals_resident')))"
  VLSCAN_ZSTD_OVERLAP=$1 VLSCAN_ZSTD_GROUP_SCALE=$2 VLSCAN_DEBUG_TIMING=1 timeout 400 python bench.py --workload C3 --rows 100000000 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/s6_t.err
  grep "vlscan zstd\] [0-9]" gpurun_out/s6_t.err | tail -1
  VLSCAN_ZSTD_OVERLAP=$1 VLSCAN_ZSTD_GROUP_SCALE=$2 timeout 400 python bench.py --workload C3 --rows 100000000 --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; print('   C3 e2e: %.1f ms/step, %.0f M rows/s h2d %.2f GB ok=%s/%s' % (e['ms_per_step'], e['value']/1e6, e['h2d_bytes_per_step']/1e9, e.get('matched_equals_resident'), e.get('digest_equals_resident')))"
done
} 2>&1 | tee gpurun_out/s6_decoder.txt
timeout 600 python bench.py --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('C3 1B: %.2f ms/step %.1f G rows/s kernel frac %.3f step frac %.3f launches/step %d' % (d['ms_per_step'], d['value']/1e9, r['frac'], d['step_frac_of_peak'], d['gpu_launches']/d['steps'])); [print(k, '%.2f ms/step step_hbm %.0f GB/s kernel frac %.3f' % (v['ms_per_step'], v['step_hbm_gbs'], v['roofline']['frac'])) for k,v in (d.get('extra_workloads') or {}).items()]" | tee gpurun_out/s6_resident.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_seq_decode|k_huf_decode|k_execute" -c 6 -o gpurun_out/prof_zstd_r02b python bench.py --workload C2 --rows 30000000 --steps 1 --warmup 1 --e2e-rows 30000000 --e2e-steps 1 --no-cpu-baseline --no-extra > gpurun_out/s6_ncu_zstd.log 2>&1; tail -1 gpurun_out/s6_ncu_zstd.log | cut -c1-200
