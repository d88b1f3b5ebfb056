#!/bin/bash
# round 2, GPU session 16: pointer-query cache in the upload (part descriptors come as tens of thousands of small pageable pieces); part reader timing without the debug timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_part.py tests/test_gpu_zzz_workers.py tests/test_gpu_zzzz_time.py tests/test_gpu_gen.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/s16_pytest.txt
timeout 600 python tools/part_bench.py --rows 60000000 --passes 3 --out gpurun_out/part_bench_r02.json > gpurun_out/s16_part.log 2> gpurun_out/s16_part.err; grep -E "^\{\"pass\"|wrote" gpurun_out/s16_part.err | cut -c1-300 | tee gpurun_out/s16_part_summary.txt
VLSCAN_DEBUG_TIMING=1 timeout 600 python tools/part_bench.py --rows 60000000 --passes 2 > /dev/null 2> gpurun_out/s16_part_dbg.err; grep -E "vlscan upload\] blocks" gpurun_out/s16_part_dbg.err | cut -c1-400 | tail -3 | tee -a gpurun_out/s16_part_summary.txt
