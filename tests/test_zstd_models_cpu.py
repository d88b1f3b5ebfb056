"""Executable models of two algorithms of the device ZSTD decoder (victorialogs_b200/csrc/vl_zstd.cuh), checked on the CPU against
straightforward references.  The kernels themselves need a GPU (tests/test_gpu_zstd.py); these models pin the parts whose correctness
is an argument rather than a format rule, in the exact form the kernels use:

* k_execute: per group of 32 sequences, all literal runs first; `dep` = last match of the group whose destination overlaps the source
  (5 binary-search probes over ascending destinations); runs of matches with dep < first-of-run copied "simultaneously" (every read of a
  run happens against the state before the run - which is also why the kernel may issue the loads of several copy steps before the first
  store, as it does since round 2); the per-warp ring in shared memory with its validity rule (`gend - q < RING`, group span
  < RING, `ring_lo` after an oversized group).
  Since round 2 the unit of a copy is a chunk of up to 4 bytes, a group is assembled in the ring and flushed to HBM at its end (so HBM may
  only be read in front of the group), and groups that span the ring keep the byte-per-lane path.
* SeqLane (the sequence decoder's bit window): a 96-bit window funnel-shifted out of four aligned words of a 256-byte ring that is topped
  up by one 16-byte chunk per sequence; three fields per sequence (offset bits | match + literal length bits | the three state updates)."""
import random

RING = 4096
M64 = (1 << 64) - 1


def lz_reference(lits, seqs, prefix):
    out = bytearray(prefix)
    lp = 0
    for ll, ml, off in seqs:
        out += lits[lp:lp + ll]
        lp += ll
        for _ in range(ml):
            out.append(out[len(out) - off])
    out += lits[lp:]
    return bytes(out)


def execute_model(lits, seqs, prefix, ring_size=RING):
    """k_execute, round 2.  `dst` is HBM, `ring` the per-warp ring.  Fast path (group output < ring): literal runs and matches are cut into
    chunks of up to 4 bytes, assembled in the ring only, and [out_run, gend) is flushed to HBM at the end of the group - so a source byte
    read from HBM must lie in front of the group.  Slow path: one byte per lane per step straight to HBM, mirrored in the ring."""
    RS = ring_size
    total = len(prefix) + len(lits) + sum(m for _, m, _ in seqs)
    dst = bytearray(prefix) + bytearray(b"\xee" * (total - len(prefix)))      # 0xEE: never-written HBM
    ring = bytearray(RS)
    for i, b in enumerate(prefix):
        ring[i & (RS - 1)] = b
    lit_run, out_run, ring_lo = 0, len(prefix), 0
    for g in range(0, len(seqs), 32):
        grp = seqs[g:g + 32]
        cnt = len(grp)
        il, io, a, b = [], [], 0, 0
        for ll, ml, _ in grp:
            a += ll
            b += ll + ml
            il.append(a)
            io.append(b)
        T, O = il[-1], io[-1]
        gend = out_run + O
        o_start = [io[j] - grp[j][0] - grp[j][1] for j in range(cnt)]
        amd = [out_run + o_start[j] + grp[j][0] for j in range(cnt)]
        ml = [q[1] for q in grp]
        off = [q[2] for q in grp]
        dep = []
        for j in range(cnt):
            s = amd[j] - off[j]
            e = s + min(ml[j], off[j])
            lo = 0
            for st in (16, 8, 4, 2, 1):                        # number of matches whose destination starts below e
                idx = lo + st - 1
                if idx < cnt and amd[idx] < e:
                    lo += st
            dep.append(lo - 1 if lo > 0 and amd[lo - 1] + ml[lo - 1] > s else -1)

        def in_ring(q):
            return q >= ring_lo and gend - q <= RS

        if O < RS:
            for j in range(cnt):                               # literal chunks (no dependencies)
                for k in range(grp[j][0]):
                    ring[(out_run + o_start[j] + k) & (RS - 1)] = lits[lit_run + il[j] - grp[j][0] + k]
            cur = 0
            while cur < cnt:
                n = 0
                while cur + n < cnt and dep[cur + n] < cur:
                    n += 1
                n = max(n, 1)
                chunks = [(j, b0) for j in range(cur, cur + n) for b0 in range(0, ml[j], 4)]
                for c0 in range(0, len(chunks), 128):          # four steps of 32 chunks: loads, then stores
                    snap, rsnap, stores = bytes(dst), bytes(ring), []
                    for j, b0 in chunks[c0:c0 + 128]:
                        nb = min(4, ml[j] - b0)
                        sa = amd[j] - off[j] + b0
                        if off[j] >= ml[j] and in_ring(sa):
                            src = [rsnap[(sa + t) & (RS - 1)] for t in range(nb)]
                        elif off[j] >= ml[j] and sa + nb <= out_run:
                            src = [snap[sa + t] for t in range(nb)]
                        else:
                            src = []
                            for t in range(nb):
                                sb = amd[j] - off[j] + (b0 + t if off[j] >= ml[j] else (b0 + t) % off[j])
                                assert in_ring(sb) or sb < out_run
                                src.append(rsnap[sb & (RS - 1)] if in_ring(sb) else snap[sb])
                        stores.append((amd[j] + b0, src))
                    for at, src in stores:
                        for t, v in enumerate(src):
                            ring[(at + t) & (RS - 1)] = v
                cur += n
            for q in range(out_run, gend):                     # flush
                dst[q] = ring[q & (RS - 1)]
        else:
            for j in range(cnt):
                for k in range(grp[j][0]):
                    v = lits[lit_run + il[j] - grp[j][0] + k]
                    at = out_run + o_start[j] + k
                    dst[at] = v
                    ring[at & (RS - 1)] = v
            cur = 0
            while cur < cnt:
                n = 0
                while cur + n < cnt and dep[cur + n] < cur:
                    n += 1
                n = max(n, 1)
                run = [(j, kk) for j in range(cur, cur + n) for kk in range(ml[j])]
                for c0 in range(0, len(run), 32):
                    snap = bytes(dst)
                    for j, kk in run[c0:c0 + 32]:
                        sa = amd[j] - off[j] + (kk if off[j] >= ml[j] else kk % off[j])
                        v = snap[sa]
                        dst[amd[j] + kk] = v
                        ring[(amd[j] + kk) & (RS - 1)] = v
                cur += n
            ring_lo = gend
        lit_run += T
        out_run += O
    for k in range(len(lits) - lit_run):
        dst[out_run + k] = lits[lit_run + k]
        ring[(out_run + k) & (RS - 1)] = lits[lit_run + k]
    return bytes(dst)


def test_executor_model_equals_sequential_lz():
    rng = random.Random(1)
    for trial in range(400):
        mode = trial % 4
        prefix = bytes(rng.getrandbits(8) for _ in range(rng.choice([0, 5, 100, 5000])))
        seqs, pos, nl = [], len(prefix), 0
        for _ in range(rng.randint(1, 150)):
            ll = rng.choice([0, 0, 1, 2, 3, 5, 8, 20]) if mode != 2 else rng.choice([0, 1, 200, 3000])
            pos += ll
            nl += ll
            if pos == 0:
                ll, pos, nl = ll + 1, pos + 1, nl + 1
            ml = rng.choice([3, 3, 4, 5, 8, 12, 30, 100]) if mode != 3 else rng.choice([3, 4, 500, 5000])
            off = rng.randint(1, min(pos, rng.choice([1, 2, 3, 4, 8, 16, 64, 300, 5000, 100000])))
            seqs.append((ll, ml, off))
            pos += ml
        lits = bytes(rng.getrandbits(8) for _ in range(nl + rng.randint(0, 10)))
        want = lz_reference(lits, seqs, prefix)
        assert execute_model(lits, seqs, prefix) == want, trial
        assert execute_model(lits, seqs, prefix, ring_size=256) == want, trial      # a small ring: sources straddle what the ring holds all the time


M32 = 0xFFFFFFFF


def funnelshift_l(lo, hi, sh):       # high 32 bits of (hi:lo) << (sh & 31)
    return (((hi << 32 | lo) << (sh & 31)) >> 32) & M32


def top_bits(x, n):                  # __funnelshift_rc(x, 0, 32 - n): the n = 0..32 highest bits of x
    return x >> (32 - n) if n else 0


class SeqWindowModel:
    """SeqLane (k_seq_decode): a 32-bit bit index `p` relative to a 256-byte aligned origin below the stream; per sequence four aligned
    ring words -> 96-bit left-aligned window c2:c1:c0; offset bits, (match + literal length) bits and the three state updates come out
    with one funnel shift each.  The ring is modelled as the memory itself addressed mod 256 through a dict of fetched chunks, with the
    kernel's top-up rule (one 16-byte chunk per step while the reader is closer than LEAD bytes) - reading a chunk that was never
    requested, or one whose ring slot has been overwritten, fails the test."""
    LEAD = 160

    def __init__(self, mem, st, ln):
        self.mem = mem
        last = mem[st + ln - 1]
        assert last
        self.org = (st & ~255) - 256
        self.s0 = (st - self.org) * 8
        self.p = self.s0 + ln * 8 - (8 - (last.bit_length() - 1))
        self.fc = ((self.p - 1) >> 7) + 1
        self.slot = {}                                         # ring slot (0..15) -> chunk index it holds
        for _ in range(12):
            self.fetch()

    def fetch(self):
        self.fc -= 1
        self.slot[self.fc & 15] = self.fc

    def word(self, j):                                         # aligned 32-bit word j of the origin space, through the ring
        assert self.slot.get((j >> 2) & 15) == j >> 2, "ring does not hold the chunk the reader needs"
        a = self.org + 4 * j
        return int.from_bytes(self.mem[a:a + 4], "little")

    def window(self):
        k, s = (self.p - 1) >> 5, (-self.p) & 31
        w3, w2, w1, w0 = self.word(k), self.word(k - 1), self.word(k - 2), self.word(k - 3)
        return funnelshift_l(w2, w3, s), funnelshift_l(w1, w2, s), funnelshift_l(w0, w1, s)

    def step(self, oc, nM, nL, bL, bM, bO):
        if self.p < self.fc * 128 + self.LEAD * 8 and self.fc > 0:
            self.fetch()
        c2, c1, c0 = self.window()
        n2 = nM + nL
        x_of = top_bits(c2, oc)
        v2 = top_bits(funnelshift_l(c1, c2, oc), n2)
        x_ml, x_ll = v2 >> nL, v2 & ((1 << nL) - 1)
        o3, n3 = oc + n2, bL + bM + bO
        a, b = (c2, c1) if o3 < 32 else (c1, c0)
        v3 = top_bits(funnelshift_l(b, a, o3), n3)
        self.p -= o3 + n3
        return [x_of, x_ml, x_ll, v3 >> (bM + bO), (v3 >> bO) & ((1 << bM) - 1), v3 & ((1 << bO) - 1)]


def test_seq_window_model_equals_plain_bit_reader():
    rng = random.Random(3)
    for trial in range(300):
        ln = rng.choice([1, 2, 7, 40, 300, 3000])
        st = 512 + rng.randint(0, 600)
        mem = bytearray(rng.getrandbits(8) for _ in range(st + ln + 64))
        if mem[st + ln - 1] == 0:
            mem[st + ln - 1] = 1
        mem = bytes(mem)
        w = SeqWindowModel(mem, st, ln)
        value, pos = int.from_bytes(mem[st:st + ln], "little"), w.p - w.s0

        def ref(n):
            nonlocal pos
            pos -= n
            return (value >> pos) & ((1 << n) - 1) if n else 0

        heavy = trial % 3 == 0                                 # worst case: 89 bits per sequence, step after step
        while pos > 0:
            ws = [31, 16, 16, 9, 9, 8] if heavy else [rng.randint(0, 31), rng.randint(0, 16), rng.randint(0, 16), rng.randint(0, 9), rng.randint(0, 9), rng.randint(0, 8)]
            if sum(ws) > pos:
                ws = [min(pos, 5), 0, 0, 0, 0, 0]
            got = w.step(*ws)
            assert got == [ref(x) for x in ws], (trial, ws)
        assert w.p == w.s0
