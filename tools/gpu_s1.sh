#!/bin/bash
# round 2, GPU session 1: parity of the new scan kernel, then resident numbers at the bench knobs and in the dense regime
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | tee gpurun_out/s1_gpu.txt
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" | tee -a gpurun_out/s1_gpu.txt; free -g | head -2 | tee -a gpurun_out/s1_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/s1_pytest.txt
for wl in C2 C1 C4; do
  timeout 300 python bench.py --workload $wl --rows 100000000 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/s1_$wl.err | tail -1 > gpurun_out/s1_$wl.json
  python - <<PY
import json
d=json.load(open("gpurun_out/s1_$wl.json")); r=d["roofline"]
print("$wl: %.2f ms/step, %.1f G rows/s, kernel %.3f ms frac %.3f share %.2f launches/step %d matched %d" % (d["ms_per_step"], d["value"]/1e9, r["kernel_ms_per_launch"], r["frac"], r["kernel_share_of_step"] or 0, d["gpu_launches"]/d["steps"], d["rows_matched_per_gpu"]))
PY
done 2>&1 | tee gpurun_out/s1_summary.txt
for h in 1 10 100 500 1000; do
  timeout 300 python bench.py --workload C2 --rows 100000000 --steps 5 --warmup 3 --hit-row-permille $h --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('C2 hit_row_permille $h: %.2f ms/step kernel %.3f ms frac %.3f matched %d' % (d['ms_per_step'], r['kernel_ms_per_launch'], r['frac'], d['rows_matched_per_gpu']))"
done 2>&1 | tee -a gpurun_out/s1_summary.txt
for h in 60 500 1000; do
  timeout 300 python bench.py --workload C3 --rows 100000000 --steps 5 --warmup 3 --hit-row-permille $h --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('C3 100M hit_row_permille $h: %.2f ms/step kernel %.3f ms frac %.3f matched %d' % (d['ms_per_step'], r['kernel_ms_per_launch'], r['frac'], d['rows_matched_per_gpu']))"
done 2>&1 | tee -a gpurun_out/s1_summary.txt
timeout 600 python bench.py --workload C3 --rows 1000000000 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/s1_C3_1B.err | tail -1 > gpurun_out/s1_C3_1B.json
python -c "
import json
d=json.load(open('gpurun_out/s1_C3_1B.json')); r=d['roofline']
print('C3 1B: %.2f ms/step, %.1f G rows/s, kernel %.3f ms frac %.3f share %.2f gen %.1fs %s' % (d['ms_per_step'], d['value']/1e9, r['kernel_ms_per_launch'], r['frac'], r['kernel_share_of_step'] or 0, d['config']['gen_seconds'], d['config']['l2']))" 2>&1 | tee -a gpurun_out/s1_summary.txt
tail -3 gpurun_out/s1_C3_1B.err
