"""Several devices behind one handle (tb_multi / MultiContext): the range split, the per-device result slices, the text broadcast with
one index per device and the sharded anchoring must give what one device gives, pair by pair. Device 0 is named two and three times
so that the sharding logic runs on a one-GPU box too; every visible device is added when there is more than one."""
import numpy as np
import pytest

import tracy_b200
from tracy_b200 import AlignConfig, DnaScore, synth

pytestmark = pytest.mark.gpu


def _layouts():
    import torch
    n = torch.cuda.device_count()
    return [[0, 0], [0, 0, 0]] + ([list(range(n))] if n > 1 else [])


def test_multi_gotoh_equals_single(ctx):
    rng = np.random.default_rng(8)
    n = 157
    refs = [synth.random_seq(rng, int(rng.integers(150, 2500))) for _ in range(n)]
    profs = [synth.profile_from_seq(rng, synth.mutate_seq(rng, r[20: 20 + int(rng.integers(60, 900))], 0.02, 0.01), 0.3) for r in refs]
    sc, ac = DnaScore(3, -5, -10, -4), AlignConfig(True, False)
    want = ctx.gotoh("ps", profs, refs, sc, ac, rows=True)
    want_s = ctx.gotoh("ps", profs, refs, sc, ac, traceback=False)[0]
    for devs in _layouts():
        with tracy_b200.MultiContext(devs) as m:
            assert m.size == len(devs)
            got = m.gotoh("ps", profs, refs, sc, ac, rows=True)
            assert np.array_equal(want[0], got[0]) and np.array_equal(want[2], got[2]), devs              # scores, ops_len
            for k in (1, 3, 4):                                                                               # ops, row0, row1: the valid prefixes
                assert all(np.array_equal(want[k][i, : want[2][i]], got[k][i, : want[2][i]]) for i in range(n)), (devs, k)
            r = m.last_ranges
            assert r[0] == 0 and r[-1] == n and all(x <= y for x, y in zip(r, r[1:])) and all(y > x for x, y in zip(r, r[1:])), r
            assert np.array_equal(m.gotoh("ps", profs, refs, sc, ac, traceback=False)[0], want_s)
            st = m.device_stats()
            assert all(s["kernel_launches"] > 0 for s in st), st               # every device took part
            # profile x profile and string x string go the same way
            a = [synth.random_profile(rng, int(rng.integers(40, 300)), "msa") for _ in range(23)]
            b = [synth.random_profile(rng, int(rng.integers(40, 300)), "trace") for _ in range(23)]
            for kind, x, y, cfg in (("pp", a, b, AlignConfig(True, True)),
                                    ("ss", [synth.random_seq(rng, 80 + i) for i in range(9)], [synth.random_seq(rng, 120 - i) for i in range(9)], AlignConfig(False, False))):
                w, g = ctx.gotoh(kind, x, y, sc, cfg), m.gotoh(kind, x, y, sc, cfg)
                assert np.array_equal(w[0], g[0]) and np.array_equal(w[2], g[2]), (devs, kind)
                assert all(np.array_equal(w[1][i, : w[2][i]], g[1][i, : w[2][i]]) for i in range(len(x))), (devs, kind)


def test_multi_broadcast_index_and_anchor(ctx):
    rng = np.random.default_rng(9)
    genome = synth.random_seq(rng, 300000)
    text = genome + b"\n"
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads = []
    for i in range(75):
        p, L = int(rng.integers(0, 298000)), int(rng.integers(300, 1200))
        s = synth.mutate_seq(rng, genome[p: p + L], 0.01, 0.003)
        reads.append(s.translate(comp)[::-1] if i % 2 else s)
    reads.append(synth.random_seq(rng, 500))
    idx = ctx.build_index(text)
    want = ctx.anchor(idx, reads, 50, 50, 15, 3)
    idx.close()
    for devs in _layouts():
        with tracy_b200.MultiContext(devs) as m:
            mi = m.build_index(text)
            got = m.anchor(mi, reads, 50, 50, 15, 3)
            for k in want:
                assert np.array_equal(want[k], got[k]), (devs, k)
            mi.close()
    assert want["anchored"].sum() >= 70 and not want["anchored"][-1]
