#!/bin/bash
# round 2, GPU session 3: the newly wired filter kinds (f3), whole GPU suite, the default bench line as the driver runs it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_next_filters.py -x -q 2>&1 | tail -12 | tee gpurun_out/s3_pytest_next.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/s3_pytest_all.txt
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/s3_bench_default.json 2> gpurun_out/s3_bench_default.err
echo "default bench wall seconds: $(( $(date +%s) - t0 ))" | tee gpurun_out/s3_bench_wall.txt
tail -1 gpurun_out/s3_bench_default.json | cut -c1-300; tail -5 gpurun_out/s3_bench_default.err
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/s3_bench_reference.json 2> gpurun_out/s3_bench_reference.err
echo "reference arm wall seconds: $(( $(date +%s) - t0 ))" | tee -a gpurun_out/s3_bench_wall.txt
tail -1 gpurun_out/s3_bench_reference.json | cut -c1-600
