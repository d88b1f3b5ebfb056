"""Glue between the oracle's encoded blocks and the product's vlscan_block descriptors (tests only)."""
import numpy as np

import vloracle
from victorialogs_b200 import scan as vs


def oracle_block_to_desc(blk, stage="ondisk"):
    """vloracle.Block -> dict accepted by victorialogs_b200.scan.HostBlocks"""
    cols = []
    for name, value in blk.consts:
        cols.append(dict(field=name, kind="const", value=value))
    for c in blk.columns:
        d = dict(field=c.name, kind="values", value_type=c.value_type, min_value=c.min_value, max_value=c.max_value, dict=c.dict, bloom=c.bloom)
        if stage == "ondisk":
            d["values_block"] = c.values_block
        else:
            d["lens_items"], d["data"] = vloracle.decode_values_block(c.values_block)
        cols.append(d)
    d = dict(rows=blk.rows, columns=cols)
    try:
        d["timestamps"] = blk.timestamps_block()   # (encoded bytes, marshalType, minTimestamp, maxTimestamp)
    except ValueError:
        pass
    return d


def field_names_of(blocks):
    names = []
    for b in blocks:
        for n, _ in b.consts:
            if n not in names:
                names.append(n)
        for c in b.columns:
            if c.name not in names:
                names.append(c.name)
    return names


def host_blocks_from_oracle(blocks, stage="ondisk"):
    return vs.HostBlocks(field_names_of(blocks) or [b"_msg"], [oracle_block_to_desc(b, stage) for b in blocks])


def gpu_rows(ctx, flt, blocks, stage="ondisk"):
    """run the product end to end through vlscan_scan_batch -> list of matching row lists, counts, stats"""
    hb = host_blocks_from_oracle(blocks, stage)
    prog = vs.Program(flt)
    words, counts, st = ctx.scan_batch(prog, hb)
    per = vs.split_bitmaps(words, [b.rows for b in blocks])
    return [vloracle.bitmap_rows(np.ascontiguousarray(w), b.rows) for w, b in zip(per, blocks)], counts, st
