"""Multi-GPU sharding of the block scan: blocks are independent (one stream, own headers / bloom / values), so the path
partitions with no data-path collective: rank r owns a contiguous range of the block list and the only
exchange is one all-reduce of the match counters at the end of a scan (`| stats count()` needs nothing else,
lib/logstorage/stats_count.go:39-46).  Used by bench.py (NCCL) and by the CPU gloo tests."""
import torch
import torch.distributed as dist


def shard_range(nblocks, world, rank):
    """contiguous, balanced block range [lo, hi) of `rank` (sizes differ by at most one)"""
    base, extra = divmod(nblocks, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_ranges_by_bytes(block_bytes, world):
    """contiguous block ranges [lo, hi) for every rank, balanced by the decoded bytes of the blocks rather than by their number:
    rank r ends at the first block where the running sum reaches (r + 1) / world of the total."""
    total = float(sum(block_bytes))
    out, lo, acc = [], 0, 0.0
    n = len(block_bytes)
    for r in range(world):
        hi = lo
        target = total * (r + 1) / world
        while hi < n and (r == world - 1 or acc + block_bytes[hi] / 2.0 < target):
            acc += block_bytes[hi]
            hi += 1
        out.append((lo, hi))
        lo = hi
    return out


def reduce_counters(counters, group=None):
    """sum {rows, rows_matched, blocks_matched, values_bytes} over the ranks; `counters` is a 4-element int64 tensor (in place)"""
    assert counters.dtype == torch.int64 and counters.numel() == 4
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters


def gather_hit_prefix(local_hits, group=None):
    """global exclusive prefix of per-rank hit counts (for consumers that want global hit-row offsets)"""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return 0, int(local_hits)
    t = torch.tensor([int(local_hits)], dtype=torch.int64)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    counts = [int(x.item()) for x in out]
    rank = dist.get_rank(group)
    return sum(counts[:rank]), sum(counts)
