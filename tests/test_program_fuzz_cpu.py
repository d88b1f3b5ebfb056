"""vlscan_program_create must reject malformed filter trees with an error, never crash: mutated serialisations of real trees (byte flips,
truncation, insertions, appended garbage) go through the program compiler in a child process, whose exit status is the test."""
import os
import subprocess
import sys
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))

CHILD = textwrap.dedent('''
    import sys, random
    sys.path.insert(0, %r); sys.path.insert(0, %r)
    from victorialogs_b200 import scan as vs
    from golden_util import load_filter_cases, build_filter
    rng = random.Random(20250923)
    seeds = [build_filter(vs.Filter, c["filter"]).blob for c in load_filter_cases()[::7]]
    nxt = ("exact_prefix", "len_range", "string_range", "ipv4_range", "value_type")
    seeds += [build_filter(vs.Filter, c["filter"]).blob for c in load_filter_cases("filter_cases_next.json") if c["filter"]["kind"] in nxt][::5]
    seeds.append(vs.Filter.and_([vs.Filter.phrase("a", "b c"), vs.Filter.or_([vs.Filter.regexp("x", "a.*b|c+"), vs.Filter.not_(vs.Filter.in_("y", ["1", "2"]))])]).blob)
    ok = bad = 0
    for i in range(6000):
        b = bytearray(rng.choice(seeds))
        for _ in range(rng.randrange(1, 4)):
            k = rng.randrange(4)
            if k == 0 and b: b[rng.randrange(len(b))] = rng.getrandbits(8)
            elif k == 1 and b: del b[rng.randrange(len(b)):]
            elif k == 2: b += bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 6)))
            else: b.insert(rng.randrange(len(b) + 1), rng.getrandbits(8))
        try:
            p = vs.Program(vs.Filter(bytes(b), "fuzz")); p.fields(); p.leaf_tokens(0) if True else None; ok += 1
        except (vs.VlscanError, IndexError):
            bad += 1
    print("compiled", ok, "rejected", bad)
    assert ok > 100 and bad > 1000
''')


def test_mutated_filter_trees_never_crash_the_compiler():
    r = subprocess.run([sys.executable, "-c", CHILD % (HERE, os.path.dirname(HERE))], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
    assert "compiled" in r.stdout
