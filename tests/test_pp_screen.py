"""The screening rule of the profile x profile kernel (tracy_b200/csrc/gotoh_pp.cu, header item 3), attacked on the CPU: the
short form sum_k p1[k] * ((match - mismatch) * p2[k] + mismatch * sum(p2)) and the reference's literal 16/25-term float
sequence (src/align.h:112-116) are both evaluated in float32 on millions of cells drawn from the value families that put
the sum ON or next to an integer (one-hot, dyadic, k/d fractions of MSA columns, createProfile-like blends, tiny and large
magnitudes). Wherever the kernel's test calls a cell safe, the two truncations must agree -- and the test must not call
everything unsafe either."""
import numpy as np
import pytest

F = np.float32
U = F(5.9604645e-8)
MAGIC = F(12582912.0)


def literal(a, b, match, mismatch):
    """float s = 0; for k1 for k2: s += a[k1] * b[k2] * (k1 == k2 ? match : mismatch); every product and sum rounded to float."""
    nch = a.shape[0]
    s = np.zeros(a.shape[1], F)
    for k1 in range(nch):
        for k2 in range(nch):
            w = F(match if k1 == k2 else mismatch)
            s = (s + ((a[k1] * b[k2]).astype(F) * w).astype(F)).astype(F)
    return s


def fma32(x, y, z):
    return (x.astype(np.float64) * y.astype(np.float64) + z.astype(np.float64)).astype(F)


def screened(a, b, match, mismatch):
    """The kernel's short form, its distance to the nearest integer and the bound it is compared with."""
    nch = a.shape[0]
    wx, wd = F(mismatch), F(match - mismatch)
    psum, pabs = b[0].copy(), np.abs(b[0])
    asum = np.abs(a[0])
    for k in range(1, nch):
        psum = (psum + b[k]).astype(F)
        pabs = (pabs + np.abs(b[k])).astype(F)
        asum = (asum + np.abs(a[k])).astype(F)
    z = (wx * psum).astype(F)
    c = [fma32(np.full_like(b[k], wd), b[k], z) for k in range(nch)]
    ap = (a[0] * c[0]).astype(F)
    for k in range(1, nch):
        ap = fma32(a[k], c[k], ap)
    rn = ((ap + MAGIC).astype(F) - MAGIC).astype(F)
    rem = (ap - rn).astype(F)
    ecoef = F(2.0) * U * (F(nch * nch + 1) * F(max(abs(match), abs(mismatch))) + F(11.0) * abs(wx) + F(6.0) * abs(wd))
    e = fma32((asum * ecoef).astype(F), pabs, np.full_like(pabs, F(1e-30)))
    with np.errstate(invalid="ignore"):
        safe = np.abs(rem) > e
    return ap, safe


def family(rng, kind, nch, n):
    p = np.zeros((nch, n), F)
    if kind == "onehot":
        p[rng.integers(0, 4, n), np.arange(n)] = 1
    elif kind == "dyadic":
        p[:4] = rng.integers(0, 9, (4, n)).astype(F) / F(8)
    elif kind == "fractions":                      # counts / coverage of an MSA column (src/align.h:138-180)
        d = rng.integers(1, 13, n)
        cnt = np.stack([rng.multinomial(int(x), [0.7, 0.1, 0.1, 0.1]) for x in d], 1)
        rot = rng.integers(0, 4, n)
        for k in range(4):
            p[(k + rot) % 4, np.arange(n)] = cnt[k].astype(F) / d.astype(F)
    elif kind == "blend":                          # createProfile: normfac * share + (1 - normfac) * 0.25 (src/profile.h:46-49)
        share = rng.dirichlet([6, 0.3, 0.3, 0.3], n).T.astype(F)
        nf = rng.uniform(0.5, 1.0, n).astype(F)
        rot = rng.integers(0, 4, n)
        for k in range(4):
            p[(k + rot) % 4, np.arange(n)] = nf * share[k] + (F(1) - nf) * F(0.25)
    elif kind == "uniform":
        p[:4] = rng.uniform(0, 1, (4, n)).astype(F)
    elif kind == "tiny":
        p[:4] = (rng.uniform(0, 1, (4, n)) * 10.0 ** rng.integers(-30, -3, n)).astype(F)
    elif kind == "large":
        p[:4] = (rng.uniform(-1, 1, (4, n)) * 10.0 ** rng.integers(0, 5, n)).astype(F)
    if nch == 5 and kind not in ("onehot",):
        nmass = (rng.integers(0, 4, n) == 0).astype(F) * F(0.25)
        p[:4] *= (F(1) - nmass)
        p[4] = nmass
    return p


KINDS = ["onehot", "dyadic", "fractions", "blend", "uniform", "tiny", "large"]


@pytest.mark.parametrize("sc", [(3, -5), (5, -4), (1, -1), (2, -7), (1000, -2000), (0, -3), (4, 4)])
@pytest.mark.parametrize("nch", [4, 5])
def test_safe_cells_truncate_alike(sc, nch):
    rng = np.random.default_rng(1000 * nch + abs(sc[0]) + 7 * abs(sc[1]))
    n = 60000
    flagged = {}
    for ka in KINDS:
        for kb in KINDS:
            a, b = family(rng, ka, nch, n), family(rng, kb, nch, n)
            want = np.trunc(literal(a, b, *sc))
            ap, safe = screened(a, b, *sc)
            got = np.trunc(ap)
            bad = safe & (got != want)
            assert not bad.any(), (ka, kb, sc, nch, a[:, bad][:, :1].ravel(), b[:, bad][:, :1].ravel())
            flagged[(ka, kb)] = 1.0 - safe.mean()
    # generic values pass the screen nearly always (that is the point of it) ...
    if sc[0] != sc[1] and abs(sc[0]) < 100:
        assert flagged[("blend", "blend")] < 2e-3 and flagged[("uniform", "uniform")] < 2e-3, flagged
    # ... and exact-integer families never do
    assert flagged[("onehot", "onehot")] == 1.0


def test_non_finite_operands_are_never_safe():
    rng = np.random.default_rng(5)
    a, b = family(rng, "blend", 4, 1000), family(rng, "blend", 4, 1000)
    a[1, ::7] = np.inf
    b[2, ::11] = np.nan
    with np.errstate(invalid="ignore", over="ignore"):
        _, safe = screened(a, b, 3, -5)
    hit = np.zeros(1000, bool)
    hit[::7] = True
    hit[::11] = True
    assert not safe[hit].any()
