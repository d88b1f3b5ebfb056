"""Executable models of two algorithms of the device ZSTD decoder (victorialogs_b200/csrc/vl_zstd.cuh), checked on the CPU against
straightforward references.  The kernels themselves need a GPU (tests/test_gpu_zstd.py); these models pin the parts whose correctness
is an argument rather than a format rule, in the exact form the kernels use:

* k_execute: per group of 32 sequences, all literal runs first; `dep` = last match of the group whose destination overlaps the source
  (5 binary-search probes over ascending destinations); runs of matches with dep < first-of-run copied "simultaneously" (every read of a
  run happens against the state before the run - which is also why the kernel may issue the loads of several copy steps before the first
  store, as it does since round 2); the per-warp ring in shared memory with its validity rule (`gend - q < RING`, group span
  < RING, `ring_lo` after an oversized group).
* QuadBitReader (the sequence decoder's bit window; the model class below keeps its round-1 name): a 192-bit window over aligned 8-byte words;
  `field(t, n)` cuts n bits that start t bits below the top; `consume` slides whole words.  The six fields of one sequence are cut at
  precomputed offsets from the same window (since round 2 by three lanes, two fields each)."""
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


def execute_model(lits, seqs, prefix):
    total = len(prefix) + len(lits) + sum(m for _, m, _ in seqs)
    dst = bytearray(prefix) + bytearray(total - len(prefix))
    ring = bytearray(RING)
    for i, b in enumerate(prefix):
        ring[i & (RING - 1)] = b
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
        ring_ok = O < RING
        o_start = [io[j] - grp[j][0] - grp[j][1] for j in range(cnt)]
        for j in range(cnt):                                   # literal runs of the whole group
            for k in range(grp[j][0]):
                v = lits[lit_run + il[j] - grp[j][0] + k]
                at = out_run + o_start[j] + k
                dst[at] = v
                ring[at & (RING - 1)] = v
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
        cur = 0
        while cur < cnt:
            n = 0
            while cur + n < cnt and dep[cur + n] < cur:
                n += 1
            n = max(n, 1)
            snap, rsnap = bytes(dst), bytes(ring)              # a run reads only what existed before the run
            for j in range(cur, cur + n):
                for kk in range(ml[j]):
                    sa = amd[j] - off[j] + (kk if off[j] >= ml[j] else kk % off[j])
                    v = rsnap[sa & (RING - 1)] if (ring_ok and sa >= ring_lo and gend - sa < RING) else snap[sa]
                    dst[amd[j] + kk] = v
                    ring[(amd[j] + kk) & (RING - 1)] = v
            cur += n
        lit_run += T
        out_run += O
        if not ring_ok:
            ring_lo = gend
    for k in range(len(lits) - lit_run):
        dst[out_run + k] = lits[lit_run + k]
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
        assert execute_model(lits, seqs, prefix) == lz_reference(lits, seqs, prefix), trial


class LineReaderModel:
    def __init__(self, mem, p_off, ln):
        self.mem = mem
        last = mem[p_off + ln - 1]
        assert last
        self.pos = ln * 8 - (8 - (last.bit_length() - 1))
        abits = p_off * 8 + self.pos
        wtop = (abits + 63) >> 6
        whi = wtop - 1
        self.wi = whi - 2
        self.w0, self.w1, self.w2 = self.word(whi), self.word(whi - 1), self.word(self.wi)
        self.off = (wtop << 6) - abits

    def word(self, w):
        return int.from_bytes(self.mem[w * 8:w * 8 + 8], "little")

    def field(self, t, n):
        s = t + n or 1
        low = s > 128
        a, b = (self.w1, self.w2) if low else (self.w0, self.w1)
        e = s - 64 if low else s
        if e <= 64:
            v = a >> (64 - e)
        else:
            sh = e - 64
            v = (((a << sh) & M64) if sh < 64 else 0) | (b >> (64 - sh))
        return v & ((1 << n) - 1) & 0xFFFFFFFF

    def consume(self, n):
        self.off += n
        self.pos -= n
        while self.off >= 64:
            self.w0, self.w1 = self.w1, self.w2
            self.off -= 64
            self.wi -= 1
            self.w2 = self.word(self.wi)


def test_line_reader_model_equals_plain_bit_reader():
    rng = random.Random(3)
    for trial in range(300):
        ln = rng.randint(1, 300)
        p_off = 600 + rng.randint(0, 40)
        mem = bytearray(rng.getrandbits(8) for _ in range(p_off + ln + 64))
        if mem[p_off + ln - 1] == 0:
            mem[p_off + ln - 1] = 1
        mem = bytes(mem)
        lr = LineReaderModel(mem, p_off, ln)
        value, pos = int.from_bytes(mem[p_off:p_off + ln], "little"), lr.pos

        def ref(n):
            nonlocal pos
            pos -= n
            return (value >> pos) & ((1 << n) - 1) if n else 0

        while pos > 0:
            if trial % 2 == 0:                                 # one field at a time
                n = rng.randint(0, min(32, pos))
                got = lr.field(lr.off, n)
                lr.consume(n)
                assert got == ref(n), (trial, n)
            else:                                              # the six fields of a sequence from one window (<= 89 bits)
                ws = [rng.randint(0, 31), rng.randint(0, 16), rng.randint(0, 16), rng.randint(0, 9), rng.randint(0, 9), rng.randint(0, 8)]
                if sum(ws) > pos:
                    ws = [min(pos, 5), 0, 0, 0, 0, 0]
                t, got = lr.off, []
                for w in ws:
                    got.append(lr.field(t, w))
                    t += w
                lr.consume(t - lr.off)
                assert got == [ref(w) for w in ws], (trial, ws)
