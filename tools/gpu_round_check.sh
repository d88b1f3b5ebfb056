mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 500 python bench.py > gpurun_out/BENCH_final.json 2> gpurun_out/BENCH_final.err; tail -1 gpurun_out/BENCH_final.json | cut -c1-150
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/BENCH_ref_final.json 2>/dev/null; tail -1 gpurun_out/BENCH_ref_final.json | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_e2e.csv python bench.py --rows 30000000 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_e2e.log 2>&1; tail -1 gpurun_out/ncu_e2e.log | cut -c1-100
