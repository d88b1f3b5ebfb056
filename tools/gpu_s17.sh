#!/bin/bash
# round 2, GPU session 17 (2 GPUs): the driver's multi-rank launch of bench.py, both arms
mkdir -p gpurun_out
s=$(date +%s); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_C3_2gpu_r02.json 2> gpurun_out/s17_bench2.err; echo "2-GPU bench rc=$? wall $(( $(date +%s) - s )) s" | tee gpurun_out/s17_wall.txt
tail -1 gpurun_out/bench_C3_2gpu_r02.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value %.2f G rows/s n_gpus %d, %.2f ms/step, kernel frac %.3f" % (d["value"]/1e9, d["n_gpus"], d["ms_per_step"], d["roofline"]["frac"])); print("e2e", json.dumps(d["e2e"])[:300]); print("config", json.dumps(d["config"])[:600]); print("extra", json.dumps(d.get("extra_workloads"))[:500]); print("totals", d.get("allreduced_totals_over_timed_steps"))' 2>&1 | tee gpurun_out/s17_summary.txt
tail -5 gpurun_out/s17_bench2.err | cut -c1-300
