mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_zstd.py -m gpu -x -q > gpurun_out/pytest_zstd.log 2>&1; rc=$?; echo "zstd pytest rc=$rc"; tail -5 gpurun_out/pytest_zstd.log
[ $rc -ne 0 ] && exit 1
timeout 400 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_zstd.py > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/pytest_gpu.log
[ $rc -ne 0 ] && exit 1
VLSCAN_DEBUG_TIMING=1 timeout 400 python bench.py --no-cpu-baseline --steps 3 --e2e-steps 1 > gpurun_out/bench_C2_zd.log 2> gpurun_out/bench_C2_zd.err; grep "vlscan upload\|vlscan zstd" gpurun_out/bench_C2_zd.err | tail -2
timeout 400 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_C2_z.log 2> gpurun_out/bench_C2_z.err; tail -1 gpurun_out/bench_C2_z.log | python -c "import sys,json; print(json.loads(sys.stdin.read())['e2e'])"
