#!/bin/bash
# round 2, GPU session 15: persistent packing threads + lazy (group by group) staging of pageable compressed bytes: parity, part reader timing with the upload's timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_part.py tests/test_gpu_zstd.py tests/test_gpu_parity.py tests/test_gpu_gen.py tests/test_gpu_zzz_workers.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/s15_pytest.txt
for th in 16 48; do
  echo "VLSCAN_HOST_THREADS=$th" | tee -a gpurun_out/s15_part_summary.txt
  VLSCAN_HOST_THREADS=$th VLSCAN_DEBUG_TIMING=1 timeout 600 python tools/part_bench.py --rows 60000000 --passes 2 --out gpurun_out/part_bench_r02_t$th.json > gpurun_out/s15_part_$th.log 2> gpurun_out/s15_part_$th.err
  grep -E "^\{\"pass\"|vlscan upload\] blocks" gpurun_out/s15_part_$th.err | cut -c1-420 | tail -8 | tee -a gpurun_out/s15_part_summary.txt
done
