"""GPU parity tests for the decompose sweeps (reference src/decompose.h:210-313) through the C ABI."""
import numpy as np
import pytest

from conftest import load_decompose_golden
from tracy_b200 import decompose, synth

pytestmark = pytest.mark.gpu
DGOLD = load_decompose_golden()
DGOLD1000 = load_decompose_golden("decompose1000_golden.npz")


def _gpu_sweep(ctx):
    def sweep(refrow, pri, sec, vi_end, ai, vi, ndel, nins, grid):
        fref, fins, g = ctx.decompose_sweep([refrow], [pri], [sec], [vi_end], [ai], [vi], [ndel], [nins], grid=grid)
        return fref[0, :ndel], fins[0, :nins], (g[0, :nins, :ndel] if grid else None)
    return sweep


@pytest.mark.parametrize("idx", range(len(DGOLD)))
def test_decompose_alleles_golden(ctx, idx):
    """Host glue + CUDA sweeps == the reference's decomposeAlleles outputs (primary, secondary, .decomp table)."""
    c = DGOLD[idx]
    pri, sec, dcp, info = decompose.decompose_alleles(c["row0"], c["row1"], c["pri"], c["sec"], c["trimL"], c["trimR"], c["maxindel"],
                                                      c["madc"], c["bp"], c["nref"], _gpu_sweep(ctx))
    assert pri == c["pri_out"] and sec == c["sec_out"]
    assert np.array_equal(dcp, c["dcp"])


def test_sweep_batch_vs_oracle(ctx, oracle_port):
    rng = np.random.default_rng(17)
    alphabet = b"ACGTNRYSWKM"
    refs, pris, secs, vend, ai, vi, nd, ni = [], [], [], [], [], [], [], []
    for t in range(200):
        L = int(rng.integers(1, 900))
        nbc = int(rng.integers(1, 900))
        ref = bytearray(synth.random_seq(rng, L, b"ACGT-"))
        pri = bytearray(synth.random_seq(rng, nbc, b"ACGTN"))
        sec = bytearray(synth.random_seq(rng, nbc, alphabet))
        for k in range(min(L, nbc)):   # make most positions phaseable so counts are not trivially large
            if rng.random() < 0.7:
                pri[k] = ref[k] if ref[k] != 45 else pri[k]
        refs.append(bytes(ref)); pris.append(bytes(pri)); secs.append(bytes(sec))
        vend.append(int(rng.integers(0, nbc + 1))); ai.append(int(rng.integers(0, L))); vi.append(int(rng.integers(0, nbc)))
        nd.append(int(rng.integers(0, 40))); ni.append(int(rng.integers(0, 40)))
    fref, fins, grid = ctx.decompose_sweep(refs, pris, secs, vend, ai, vi, nd, ni, grid=True)
    for t in range(200):
        wr, wi, wg = oracle_port.decompose_sweep(refs[t], pris[t], secs[t], vend[t], ai[t], vi[t], nd[t], ni[t], grid=True)
        assert np.array_equal(fref[t, : nd[t]], wr), t
        if ni[t] > 0 and nd[t] > 0:
            assert np.array_equal(fins[t, : ni[t]], wi), t
        elif ni[t] > 0:
            assert np.array_equal(fins[t, 1: ni[t]], wi[1:]), t
        assert np.array_equal(grid[t, : ni[t], : nd[t]], wg), t


def test_config3_scale(ctx, oracle_port):
    """Config 3 shape: 10k traces, +-30 bp sweep. Sample-checked against the oracle, plus the identity fins[0] == fref[0]."""
    rng = np.random.default_rng(3)
    N = 10000
    base = synth.random_seq(rng, 5000)
    refs, pris, secs = [], [], []
    for t in range(N):
        o = int(rng.integers(0, 4000))
        L = int(rng.integers(600, 900))
        r = base[o:o + L]
        refs.append(r)
        p = bytearray(r[:L - 40])
        s = bytearray(p)
        for k in rng.integers(0, len(p), 30):
            s[k] = b"ACGTRYSWKMN"[int(rng.integers(0, 11))]
        pris.append(bytes(p)); secs.append(bytes(s))
    vend = [len(p) - 20 for p in pris]
    ai = [200] * N; vi = [205] * N; nd = [30] * N; ni = [30] * N
    fref, fins, _ = ctx.decompose_sweep(refs, pris, secs, vend, ai, vi, nd, ni)
    assert np.array_equal(fref[:, 0], fins[:, 0])
    for t in range(0, N, 997):
        wr, wi, _ = oracle_port.decompose_sweep(refs[t], pris[t], secs[t], vend[t], ai[t], vi[t], 30, 30)
        assert np.array_equal(fref[t], wr) and np.array_equal(fins[t], wi)


@pytest.mark.parametrize("idx", range(len(DGOLD1000)))
def test_decompose_alleles_maxindel_1000_golden(ctx, idx):
    """The CLI default maxindel = 1000 on ~1 kb traces against 4 kb windows (tests/golden/make_golden_decompose1000.py): long
    deletions / insertions and the ins x del fallback grid (hundreds of thousands of shifts for ONE trace, split over many
    blocks) -- primary, secondary and the .decomp table equal the reference's decomposeAlleles."""
    c = DGOLD1000[idx]
    pri, sec, dcp, info = decompose.decompose_alleles(c["row0"], c["row1"], c["pri"], c["sec"], c["trimL"], c["trimR"], c["maxindel"],
                                                      c["madc"], c["bp"], c["nref"], _gpu_sweep(ctx))
    assert pri == c["pri_out"] and sec == c["sec_out"]
    assert np.array_equal(dcp, c["dcp"])


def test_sweep_thousand_shifts_vs_oracle(ctx, oracle_port):
    """ndel = 1000, nins = 400 with the full grid for two traces (4 * 10^5 shifts each) and a 64-trace batch without the grid."""
    rng = np.random.default_rng(1001)
    refs, pris, secs = [], [], []
    for t in range(64):
        r = synth.random_seq(rng, 3900, b"ACGT-" if t % 2 else b"ACGT")
        p = bytearray(r[100:1000])
        s = bytearray(p)
        for k in rng.integers(0, len(p), 120):
            s[k] = b"ACGTRYSWKMN"[int(rng.integers(0, 11))]
            p[k] = b"ACGTN"[int(rng.integers(0, 5))]
        refs.append(r); pris.append(bytes(p)); secs.append(bytes(s))
    vend = [880] * 64; ai = [150] * 64; vi = [60] * 64
    fref, fins, _ = ctx.decompose_sweep(refs, pris, secs, vend, ai, vi, [1000] * 64, [400] * 64)
    for t in range(0, 64, 9):
        wr, wi, _ = oracle_port.decompose_sweep(refs[t], pris[t], secs[t], vend[t], ai[t], vi[t], 1000, 400)
        assert np.array_equal(fref[t, :1000], wr) and np.array_equal(fins[t, :400], wi), t
    f2, i2, g2 = ctx.decompose_sweep(refs[:2], pris[:2], secs[:2], vend[:2], ai[:2], vi[:2], [1000, 1000], [400, 400], grid=True)
    print("grid sweep of 2 traces x 400 000 shifts:", ctx.last_kernel_ms()["sweep_ms"], "ms")
    for t in range(2):
        wr, wi, wg = oracle_port.decompose_sweep(refs[t], pris[t], secs[t], vend[t], ai[t], vi[t], 1000, 400, grid=True)
        assert np.array_equal(f2[t, :1000], wr) and np.array_equal(i2[t, :400], wi)
        assert np.array_equal(g2[t, :400, :1000], wg), t
