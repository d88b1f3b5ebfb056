set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_C2_final.log 2> gpurun_out/bench_C2_final.err; tail -1 gpurun_out/bench_C2_final.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 51 -c 51 --csv --log-file gpurun_out/launches_final.csv python bench.py --rows 30000000 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_substr_scan -s 4 -c 1 -f -o gpurun_out/prof_scan_final python bench.py --rows 30000000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | tail -5
