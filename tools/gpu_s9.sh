#!/bin/bash
# round 2, GPU session 9: executor on 4-byte chunks assembled in the ring, Huffman register window; source-level profile of the scan at dense candidates
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zstd.py -x -q 2>&1 | tail -12 | tee gpurun_out/s9_pytest_zstd.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/s9_pytest_all.txt
e2e_line='import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print("   %s e2e: %.1f ms/step, %.0f M rows/s h2d %.2f GB ok=%s/%s" % (sys.argv[1], e["ms_per_step"], e["value"]/1e6, e["h2d_bytes_per_step"]/1e9, e.get("matched_equals_resident"), e.get("digest_equals_resident")))'
{
VLSCAN_ZSTD_OVERLAP=0 VLSCAN_ZSTD_GROUP_SCALE=4 VLSCAN_DEBUG_TIMING=1 timeout 400 python bench.py --workload C2 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/s9_t.err
echo "phases C2 (serial, scale 4):"; grep "vlscan zstd\] [0-9]" gpurun_out/s9_t.err | tail -1
VLSCAN_ZSTD_OVERLAP=0 VLSCAN_ZSTD_GROUP_SCALE=4 VLSCAN_DEBUG_TIMING=1 timeout 400 python bench.py --workload C3 --rows 100000000 --steps 3 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extra > /dev/null 2> gpurun_out/s9_t.err
echo "phases C3 (serial, scale 4):"; grep "vlscan zstd\] [0-9]" gpurun_out/s9_t.err | tail -1
for cfg in "0 4" "1 4" "1 2" "1 1"; do
  set -- $cfg
  echo "overlap=$1 group_scale=$2:"
  VLSCAN_ZSTD_OVERLAP=$1 VLSCAN_ZSTD_GROUP_SCALE=$2 timeout 400 python bench.py --workload C2 --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "$e2e_line" C2
  VLSCAN_ZSTD_OVERLAP=$1 VLSCAN_ZSTD_GROUP_SCALE=$2 timeout 400 python bench.py --workload C3 --rows 100000000 --steps 3 --warmup 3 --e2e-steps 5 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "$e2e_line" C3
done
} 2>&1 | tee gpurun_out/s9_decoder.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_substr_scan" -s 2 -c 1 -o gpurun_out/prof_scan_dense_r02 python bench.py --workload C3 --rows 30000000 --hit-row-permille 500 --vocab-focus 4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/s9_ncu_dense.log 2>&1; tail -1 gpurun_out/s9_ncu_dense.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_execute|k_huf_decode" -c 4 -o gpurun_out/prof_zstd_r02d python bench.py --workload C3 --rows 30000000 --steps 1 --warmup 1 --e2e-rows 30000000 --e2e-steps 1 --no-cpu-baseline --no-extra > gpurun_out/s9_ncu_zstd.log 2>&1; tail -1 gpurun_out/s9_ncu_zstd.log | cut -c1-200
