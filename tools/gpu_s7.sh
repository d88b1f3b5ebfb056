#!/bin/bash
# round 2, GPU session 7: leaner single-lane sequence decoder, decoder enqueued before any blocking copy (true DMA/decode overlap), new filter kinds + gather
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zstd.py tests/test_gpu_zzzz_time.py -x -q 2>&1 | tail -12 | tee gpurun_out/s7_pytest_zstd.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/s7_pytest_all.txt
e2e_line='import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print("   %s e2e: %.1f ms/step, %.0f M rows/s h2d %.2f GB ok=%s/%s" % (sys.argv[1], e["ms_per_step"], e["value"]/1e6, e["h2d_bytes_per_step"]/1e9, e.get("matched_equals_resident"), e.get("digest_equals_resident")))'
{
VLSCAN_ZSTD_OVERLAP=0 VLSCAN_ZSTD_GROUP_SCALE=4 VLSCAN_DEBUG_TIMING=1 timeout 400 python bench.py --workload C2 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/s7_t.err
echo "phases C2 (serial, scale 4):"; grep "vlscan zstd\] [0-9]" gpurun_out/s7_t.err | tail -1; grep "vlscan upload\] blocks" gpurun_out/s7_t.err | tail -1
VLSCAN_ZSTD_OVERLAP=0 VLSCAN_ZSTD_GROUP_SCALE=4 VLSCAN_DEBUG_TIMING=1 timeout 400 python bench.py --workload C3 --rows 100000000 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/s7_t.err
echo "phases C3 (serial, scale 4):"; grep "vlscan zstd\] [0-9]" gpurun_out/s7_t.err | tail -1; grep "vlscan upload\] blocks" gpurun_out/s7_t.err | tail -1
for cfg in "0 4" "1 4" "0 2" "0 1" "1 1"; do
  set -- $cfg
  echo "overlap=$1 group_scale=$2:"
  VLSCAN_ZSTD_OVERLAP=$1 VLSCAN_ZSTD_GROUP_SCALE=$2 timeout 400 python bench.py --workload C2 --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "$e2e_line" C2
  VLSCAN_ZSTD_OVERLAP=$1 VLSCAN_ZSTD_GROUP_SCALE=$2 timeout 400 python bench.py --workload C3 --rows 100000000 --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "$e2e_line" C3
done
} 2>&1 | tee gpurun_out/s7_decoder.txt
s=$(date +%s); timeout 900 python bench.py > gpurun_out/s7_bench_default.json 2> gpurun_out/s7_bench_default.err; echo "default bench rc=$? wall $(( $(date +%s) - s )) s" | tee gpurun_out/s7_bench_wall.txt
tail -1 gpurun_out/s7_bench_default.json | cut -c1-400
