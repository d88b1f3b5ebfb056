#!/bin/bash
# round 2, GPU session 8: executor with far matches batched beside the literals; focus sweep; decoder ncu for the docs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zstd.py tests/test_gpu_zzzz_time.py tests/test_gpu_gen.py -x -q 2>&1 | tail -12 | tee gpurun_out/s8_pytest_zstd.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/s8_pytest_all.txt
e2e_line='import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print("   %s e2e: %.1f ms/step, %.0f M rows/s h2d %.2f GB ok=%s/%s" % (sys.argv[1], e["ms_per_step"], e["value"]/1e6, e["h2d_bytes_per_step"]/1e9, e.get("matched_equals_resident"), e.get("digest_equals_resident")))'
{
VLSCAN_ZSTD_OVERLAP=0 VLSCAN_ZSTD_GROUP_SCALE=4 VLSCAN_DEBUG_TIMING=1 timeout 400 python bench.py --workload C2 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/s8_t.err
echo "phases C2 (serial, scale 4):"; grep "vlscan zstd\] [0-9]" gpurun_out/s8_t.err | tail -1
VLSCAN_ZSTD_OVERLAP=0 VLSCAN_ZSTD_GROUP_SCALE=4 VLSCAN_DEBUG_TIMING=1 timeout 400 python bench.py --workload C3 --rows 100000000 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/s8_t.err
echo "phases C3 (serial, scale 4):"; grep "vlscan zstd\] [0-9]" gpurun_out/s8_t.err | tail -1
for cfg in "0 4" "1 4" "1 2" "1 1"; do
  set -- $cfg
  echo "overlap=$1 group_scale=$2:"
  VLSCAN_ZSTD_OVERLAP=$1 VLSCAN_ZSTD_GROUP_SCALE=$2 timeout 400 python bench.py --workload C2 --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "$e2e_line" C2
  VLSCAN_ZSTD_OVERLAP=$1 VLSCAN_ZSTD_GROUP_SCALE=$2 timeout 400 python bench.py --workload C3 --rows 100000000 --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "$e2e_line" C3
done
} 2>&1 | tee gpurun_out/s8_decoder.txt
timeout 900 python tools/sweep.py --rows 100000000 --steps 5 --warmup 3 --out gpurun_out/sweep_r02.json > gpurun_out/s8_sweep.log 2>&1; tail -2 gpurun_out/s8_sweep.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_seq_decode|k_huf_decode|k_execute" -c 6 -o gpurun_out/prof_zstd_r02c python bench.py --workload C3 --rows 30000000 --steps 1 --warmup 1 --e2e-rows 30000000 --e2e-steps 1 --no-cpu-baseline --no-extra > gpurun_out/s8_ncu_zstd.log 2>&1; tail -1 gpurun_out/s8_ncu_zstd.log | cut -c1-200
