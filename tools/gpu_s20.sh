#!/bin/bash
# round 2, GPU session 20 (8 GPUs): the driver's 8-rank launch of bench.py
mkdir -p gpurun_out
free -g | head -2 | tee gpurun_out/s20_mem.txt
s=$(date +%s); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_C3_8gpu_r02.json 2> gpurun_out/s20_bench8.err; echo "8-GPU bench rc=$? wall $(( $(date +%s) - s )) s" | tee gpurun_out/s20_wall.txt
tail -1 gpurun_out/bench_C3_8gpu_r02.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value %.2f G rows/s n_gpus %d, %.2f ms/step, kernel frac %.3f" % (d["value"]/1e9, d["n_gpus"], d["ms_per_step"], d["roofline"]["frac"])); print("e2e", json.dumps(d["e2e"])[:260]); print("affinity", d["config"].get("host_affinity")); print("extra", json.dumps(d.get("extra_workloads"))[:330]); print("totals", d.get("allreduced_totals_over_timed_steps"))' 2>&1 | tee gpurun_out/s20_summary.txt
tail -4 gpurun_out/s20_bench8.err | cut -c1-300
