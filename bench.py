#!/usr/bin/env python3
"""bench.py -- rows scanned/s of the LogsQL block-scan hot path on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # GPU arm (libvlscan.so)
    python bench.py --impl reference [--gpus N] [--steps K] ...    # reference arm: the CPU algorithm on the host cores

A "step" is one pass of the hot path over one batch of synthetic blocks:  bm.init/setBits + filter.applyToBlockSearch for every
block (lib/logstorage/block_search.go:207-215).  Workload at N=1: BASELINE.json configs[1], `_msg:"timeout" AND level:error`
over 100 M vlogsgenerator-shaped rows (16 fields => 3000 rows/block by the 2 MB rule), resident in HBM when the timed region
starts (`value`), and again through the C-ABI call on pinned HOST buffers with the copies inside the timed region (`e2e`).
N>1: every rank scans its own 100 M-row shard (blocks are independent: weak scaling, no data-path collective) and the ranks
all-reduce {rows, rows_matched, blocks_matched, values_bytes} over NCCL once per step, inside the timed region.

The reference (Go) cannot run here (no Go toolchain); the reference arm / cpu_baseline time the oracle's restatement of the same
per-block algorithm (kind "port") on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20250718
WORKLOADS = {
    # name: (rows per rank, fields, rows/block, columns_mask, filter builder, logsql)
    "C2": dict(rows=100_000_000, fields=16, rows_per_block=3000, mask=0b0011, logsql='_msg:"timeout" AND level:error',
               tree=lambda F: F.and_([F.phrase("_msg", "timeout"), F.phrase("level", "error")])),
    "C1": dict(rows=1_000_000, fields=8, rows_per_block=6400, mask=0b0001, logsql='_msg:"error"', tree=lambda F: F.phrase("_msg", "error")),
    "C3": dict(rows=1_000_000_000, fields=32, rows_per_block=2000, mask=0b0001, logsql='_msg:~"conn.*refused"', tree=lambda F: F.regexp("_msg", "conn.*refused")),
    "C4": dict(rows=125_000_000, fields=32, rows_per_block=2000, mask=0b1101, logsql='_msg:"GET" AND path:api* AND status:in(500,502,503)',
               tree=lambda F: F.and_([F.phrase("_msg", "GET"), F.prefix("path", "api"), F.in_("status", ["500", "502", "503"])])),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vlscan", choices=["vlscan", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="rows per rank (default: the workload's)")
    ap.add_argument("--hot-block-permille", type=int, default=1000, help="block clustering knob: fraction of blocks holding vocabulary rows")
    ap.add_argument("--hit-row-permille", type=int, default=60, help="selectivity knob: vocabulary rows inside hot blocks")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-stage", default="ondisk", choices=["ondisk", "decoded"], help="form of the host blocks handed to vlscan_scan_batch")
    ap.add_argument("--cpu-sample-rows", type=int, default=12_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": reasons}


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def cpu_port(wl, gen_kw, sample_rows, threads, target_secs=10.0):
    """The reference's per-block algorithm restated on the CPU (oracle/), all host threads, bounded sample. -> rows/s, info"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vloracle
    cfg = vloracle.GenConfig(**gen_kw)
    nb = max(1, min(sample_rows // wl["rows_per_block"], (gen_kw["total_rows"] + wl["rows_per_block"] - 1) // wl["rows_per_block"]))
    flt = wl["tree"](vloracle.Filter)
    # calibrate the number of passes so that the timed region holds ~target_secs of CPU work (bounded sample, repeated)
    r = vloracle.scan_generated(cfg, flt, 0, nb, threads)
    passes = int(max(1, min(400, target_secs / max(r["secs"], 1e-4))))
    if passes > 1:
        r = vloracle.scan_generated(cfg, flt, 0, nb, threads, passes=passes)
    rows = int(r["stats"][1])
    rate = rows * passes / r["secs"]
    return rate, dict(rows=rows, blocks=nb, secs=r["secs"], passes=passes, matches=r["matches"], values_bytes=int(r["stats"][3]))


def run_reference(args, wl, gen_kw, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t0 = time.time()
    rates = []
    sample = args.cpu_sample_rows
    for i in range(args.warmup + args.steps):
        rate, info = cpu_port(wl, gen_kw, sample, threads, target_secs=min(10.0, 200.0 / (args.warmup + args.steps)))
        if i >= args.warmup:
            rates.append(rate)
        if time.time() - t0 > 240:
            break
    value = statistics.median(rates) if rates else rate
    out = {
        "impl": "reference", "metric": "log rows scanned/sec (LogsQL phrase+regex)", "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": len(rates),
        "warmup": args.warmup, "ms_per_step": 1000.0 * info["rows"] * info["passes"] / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic (deterministic vlogsgenerator-shaped rows, seed %d)" % SEED,
        "config": {"workload": "%s: %s" % (args.workload, wl["logsql"]), "rows_per_step": info["rows"] * info["passes"], "rows_per_block": wl["rows_per_block"], "fields": wl["fields"],
                   "hot_block_permille": gen_kw["hot_block_permille"], "hit_row_permille": gen_kw["hit_row_permille"],
                   "note": "Go toolchain absent: the reference's per-block algorithm restated in C++ (oracle/), ZSTD-compressed values blocks included, all host threads"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port", "sample": "%d rows (%d blocks) of the %s workload x %d passes per step" % (info["rows"], info["blocks"], args.workload, info["passes"])},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    rows = args.rows or wl["rows"]
    rows -= rows % wl["rows_per_block"] if rows % wl["rows_per_block"] and rows % wl["rows_per_block"] < 64 else 0
    nb = (rows + wl["rows_per_block"] - 1) // wl["rows_per_block"]
    gen_kw = dict(seed=SEED, total_rows=rows * world, rows_per_block=wl["rows_per_block"], hot_block_permille=args.hot_block_permille,
                  hit_row_permille=args.hit_row_permille, columns_mask=wl["mask"])
    if rows % wl["rows_per_block"]:
        gen_kw["total_rows"] = nb * wl["rows_per_block"] * (world - 1) + rows if world > 1 else rows

    if args.impl == "reference":
        run_reference(args, wl, gen_kw, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from victorialogs_b200 import scan as vs, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libvlscan has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = vs.Ctx(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    gcfg = vs.GenConfig(**gen_kw)
    block_lo = rank * nb
    t_gen = time.time()
    batch = ctx.generate(gcfg, block_lo, block_lo + nb)
    t_gen = time.time() - t_gen
    prog = vs.Program(wl["tree"](vs.Filter))

    class _Arr:   # zero-copy torch view of the library's 4 x u64 totals vector
        def __init__(self, ptr):
            self.__cuda_array_interface__ = {"shape": (4,), "typestr": "<i8", "data": (ptr, False), "version": 2}

    def step():
        ctx.scan_resident(prog, batch, want_stats=False)
        if world > 1:
            _, _, totals = ctx.result_device_ptrs()
            with torch.cuda.stream(stream):
                t = torch.as_tensor(_Arr(totals), device="cuda")
                shard.reduce_counters(t)     # the only collective of the path: final NCCL reduce of the match counters

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    sync_all()
    # the timed region is a few tens of milliseconds, shorter than one nvidia-smi sampling period: keep the same load running (untimed)
    # until the sampler has seen ~1 s of it, so that `clocks` really is the SM clock / throttle state under this workload
    t_hold = time.perf_counter()
    while time.perf_counter() - t_hold < 1.0:
        ctx.scan_resident(prog, batch, want_stats=False)
        ctx.sync()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    st = ctx.last_scan_stats()
    totals_host = None
    if world > 1:
        _, _, totals = ctx.result_device_ptrs()
        totals_host = torch.as_tensor(_Arr(totals), device="cuda").cpu().tolist()

    # roofline of the dominant kernel (k_substr_scan): algorithmic bytes / its CUDA-event duration, averaged over fresh scans
    kms, kbytes, gms = [], 0, []
    for _ in range(min(args.steps, 10)):
        ctx.scan_resident(prog, batch, want_stats=False)
        s = ctx.last_scan_stats()
        kms.append(s.scan_kernel_ms)
        gms.append(s.gpu_ms)
        kbytes = s.scan_kernel_bytes
    peak, peak_src = hbm_peak()
    k_avg = statistics.mean(kms) if kms else 0.0
    achieved = (kbytes / 1e9) / (k_avg / 1e3) if k_avg > 0 else 0.0
    step_bytes = st.values_bytes + st.bloom_probe_bytes + st.bitmap_bytes

    # end to end through the C ABI on pinned host buffers: H2D of every block + scan + D2H of bitmaps and counts, every step
    e2e = None
    if not args.no_e2e:
        try:
            host = ctx.download(batch)
            t_comp = 0.0
            if args.e2e_stage == "ondisk":
                # the reference's writer re-encodes every values block into its on-disk form (ZSTD frames); the scan call then ships the
                # compressed bytes and regenerates them on the device.  Not timed: it is the ingestion side.
                t1 = time.perf_counter()
                disk = host.compress(threads=max(1, (os.cpu_count() or 1) // world))   # ranks share the host cores
                t_comp = time.perf_counter() - t1
                del host
                host = disk
            # host threads the library may use inside an upload (ZSTD header walk); the ranks of a box share its cores
            os.environ.setdefault("VLSCAN_HOST_THREADS", str(max(1, min(32, (os.cpu_count() or 1) // world))))
            nwords = sum((r + 63) // 64 for r in host.rows)
            words = np.zeros(max(nwords, 1), dtype=np.uint64)
            counts = np.zeros(max(host.nblocks, 1), dtype=np.uint32)
            ctx.scan_batch(prog, host, words, counts)   # warm-up (allocations)
            sync_all()
            t0 = time.perf_counter()
            est = None
            for _ in range(args.e2e_steps):
                _, _, est = ctx.scan_batch(prog, host, words, counts)
            sync_all()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
            e2e = {"value": rows * world * args.e2e_steps / dt, "unit": "rows/s", "h2d_bytes_per_step": int(est.h2d_bytes) * world, "d2h_bytes_per_step": int(est.d2h_bytes) * world,
                   "ms_per_step": 1000 * dt / args.e2e_steps, "steps": args.e2e_steps, "matched": int(counts.sum()),
                   "input_stage": "on-disk values blocks (ZSTD frames, decoded on the device)" if args.e2e_stage == "ondisk" else "decoded values blocks",
                   "host_bytes": int(host.bytes), "writer_compress_seconds": round(t_comp, 2), "host_threads": int(os.environ["VLSCAN_HOST_THREADS"]),
                   "matched_equals_resident": int(counts.sum()) == int(st.rows_matched)}
            del host
        except Exception as e:   # pinned host memory for the full data set may not be available
            e2e = {"value": None, "unit": "rows/s", "error": str(e)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, info = cpu_port(wl, gen_kw, args.cpu_sample_rows, threads)
        cpu = {"value": rate, "unit": "rows/s", "cores": threads, "kind": "port",
               "sample": "first %d rows (%d blocks) of the %s workload x %d passes, ZSTD-compressed values blocks, %d threads, %.2f s" % (info["rows"], info["blocks"], args.workload, info["passes"], threads, info["secs"])}

    if rank == 0:
        out = {
            "metric": "log rows scanned/sec (LogsQL phrase+regex)", "value": rows * world * args.steps / (ms / 1e3), "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic (deterministic vlogsgenerator-shaped rows generated on the device, seed %d)" % SEED,
            "config": {"workload": "%s: %s over %d rows/GPU, %d fields" % (args.workload, wl["logsql"], rows, wl["fields"]), "rows_per_gpu": rows, "rows_per_block": wl["rows_per_block"],
                       "blocks_per_gpu": nb, "hot_block_permille": args.hot_block_permille, "hit_row_permille": args.hit_row_permille,
                       "l2": "inputs (%.1f GB/GPU) are far larger than the 126 MB L2; no flush between iterations" % (batch.device_bytes / 1e9),
                       "parallelism": "blocks sharded over %d GPU(s), one NCCL all-reduce of 4 counters per step" % world if world > 1 else "1 GPU",
                       "gen_seconds": round(t_gen, 2)},
            "rows_matched_per_gpu": int(st.rows_matched), "blocks_matched_per_gpu": int(st.blocks_matched),
            "algorithmic_bytes_per_step_per_gpu": int(step_bytes),
            "step_hbm_gbs_per_gpu": (step_bytes / 1e9) / (ms / args.steps / 1e3),
            "gpu_launches": int(st.gpu_launches) * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_substr_scan", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": None,   # no ncu capture exists for this exact launch; the one that does (30 M rows of the same workload):
                         "traffic_capture": {"rows": 30000000, "dram_bytes_per_launch": 4403560752, "algorithmic_bytes_per_launch": 3806642565, "ratio": 1.157,
                                             "source": "profiles/ncu_k_substr_scan_r01.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"},
                         "peak_source": peak_src, "kernel_ms_per_launch": k_avg, "algorithmic_bytes_per_launch": int(kbytes),
                         "kernel_share_of_step": (k_avg / statistics.mean(gms)) if gms and statistics.mean(gms) > 0 else None},
            "e2e": e2e, "cpu_baseline": cpu,
        }
        if totals_host is not None:
            out["allreduced_totals"] = {"rows": totals_host[0], "rows_matched": totals_host[1], "blocks_matched": totals_host[2]}
        print(json.dumps(out), flush=True)
    batch.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
