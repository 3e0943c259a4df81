"""Seeded synthetic inputs for the benchmark configs of BASELINE.json.

Shapes follow the reference's data: a trace profile is float32[6][m] whose rows A,C,G,T sum to 1 and whose rows
N,- are exactly 0 (what createProfile guarantees, reference src/profile.h:37-49); a reference window is m..n
characters over ACGT. Generation is vectorised numpy (PCG64), so 10^5 pairs take seconds.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", np.uint8)
_COMP = np.array([3, 2, 1, 0], np.int64)


def align_batch(npairs, m=1000, n=4000, seed=44, sub_rate=0.01, indel_rate=0.003, het_rate=0.05, rc_frac=0.5):
    """Config 2/5 (`batched Gotoh`): returns (profiles float32[N][6][m], windows uint8[N][n] of ACGT chars)."""
    rng = np.random.default_rng(seed)
    win = rng.integers(0, 4, size=(npairs, n), dtype=np.int64)
    off = rng.integers(0, max(n - m - 16, 1), size=npairs)
    # source index walk with 1-3 bp deletions; insertions draw a random base instead of advancing
    u = rng.random((npairs, m))
    is_del = u < indel_rate / 2
    is_ins = (u >= indel_rate / 2) & (u < indel_rate)
    step = np.ones((npairs, m), np.int64)
    step[is_del] += rng.integers(1, 4, size=int(is_del.sum()))
    step[is_ins] = 0
    src = off[:, None] + np.cumsum(step, axis=1) - step
    src = np.minimum(src, n - 1)
    base = np.take_along_axis(win, src, axis=1)
    base = np.where(is_ins, rng.integers(0, 4, size=(npairs, m)), base)
    mut = rng.random((npairs, m)) < sub_rate
    base = np.where(mut, (base + rng.integers(1, 4, size=(npairs, m))) % 4, base)
    rc = rng.random(npairs) < rc_frac
    base[rc] = _COMP[base[rc][:, ::-1]]
    # signal model: primary weight w, occasionally a real secondary peak, the rest spread as background
    w = rng.uniform(0.55, 1.0, size=(npairs, m)).astype(np.float32)
    het = rng.random((npairs, m)) < het_rate
    sec = (base + rng.integers(1, 4, size=(npairs, m))) % 4
    prof = np.zeros((npairs, 6, m), np.float32)
    rest = (np.float32(1.0) - w)
    bg = np.where(het, np.float32(0), rest / np.float32(3.0)).astype(np.float32)
    for k in range(4):
        prof[:, k, :] = bg
    ii = np.arange(npairs)[:, None]
    jj = np.arange(m)[None, :]
    prof[ii, base, jj] = w
    hp, hq = np.nonzero(het)
    prof[hp, sec[hp, hq], hq] = rest[hp, hq]
    windows = ACGT[win]
    return prof, np.ascontiguousarray(windows)


def random_profile(rng, m, style="trace"):
    """One float32[6][m] profile. style: 'trace' (rows 0-3 sum to 1), 'ties' (coarse values -> many equal scores),
    'msa' (column frequencies with N and gap mass, as _createProfile(char MSA) produces, reference src/align.h:138-180)."""
    p = np.zeros((6, m), np.float32)
    if style == "trace":
        x = rng.random((4, m)).astype(np.float32) ** 3
        x[rng.integers(0, 4, m), np.arange(m)] += np.float32(2.0)
        p[:4] = x / x.sum(0, keepdims=True)
    elif style == "ties":
        idx = rng.integers(0, 4, m)
        p[idx, np.arange(m)] = 1.0
        flat = rng.random(m) < 0.2
        p[:4, flat] = 0.25
        half = rng.random(m) < 0.2
        p[:4, half] = 0.0
        p[rng.integers(0, 4, m)[half], np.nonzero(half)[0]] += 0.5
        p[rng.integers(0, 4, m)[half], np.nonzero(half)[0]] += 0.5
    else:
        depth = rng.integers(1, 9, m)
        for j in range(m):
            cnt = rng.multinomial(depth[j], [0.3, 0.2, 0.2, 0.2, 0.03, 0.07])
            p[:, j] = cnt.astype(np.float32) / np.float32(depth[j])
    return p


def random_seq(rng, n, alphabet=b"ACGT"):
    a = np.frombuffer(alphabet, np.uint8)
    return bytes(a[rng.integers(0, len(a), n)])


def mutate_seq(rng, s, sub=0.03, indel=0.02):
    out = bytearray()
    for ch in s:
        u = rng.random()
        if u < indel / 2:
            continue
        if u < indel:
            out.append(b"ACGT"[rng.integers(0, 4)])
        if rng.random() < sub:
            out.append(b"ACGT"[rng.integers(0, 4)])
        else:
            out.append(ch)
    return bytes(out)


def profile_from_seq(rng, s, noise=0.3):
    """A trace-like profile whose consensus is `s` (ACGT bytes)."""
    m = len(s)
    code = np.array([b"ACGT".index(bytes([c])) for c in s], np.int64)
    w = rng.uniform(1.0 - noise, 1.0, m).astype(np.float32)
    p = np.zeros((6, m), np.float32)
    for k in range(4):
        p[k] = (np.float32(1.0) - w) / np.float32(3.0)
    p[code, np.arange(m)] = w
    return p


# ---- synthetic trace FILES (tests / benches of the ingest path) ---------------------------------------------------------
def abif_bytes(channels, order, ploc, basecalls, qual, basecalls2=None, extra=()):
    """A minimal valid ABIF file (big-endian directory of 28-byte entries, reference src/abif.h:300-378): DATA.9-12 =
    `channels` in file order, FWO_.1 = `order` (e.g. b"GATC"), PLOC.2, PBAS.2, PCON.2, optionally P2BA.1.
    extra: further (name, number, etype, esize, payload bytes) records."""
    import struct
    recs = []
    for i, ch in enumerate(channels):
        recs.append((b"DATA", 9 + i, 4, 2, np.asarray(ch, ">i2").tobytes(), len(ch)))
    recs.append((b"FWO_", 1, 2, 1, bytes(order), len(order)))
    recs.append((b"PLOC", 2, 4, 2, np.asarray(ploc, ">i2").tobytes(), len(ploc)))
    recs.append((b"PBAS", 2, 2, 1, bytes(basecalls), len(basecalls)))
    recs.append((b"PCON", 2, 2, 1, np.asarray(qual, np.uint8).tobytes(), len(qual)))      # stored as char; readab forces etype 1
    if basecalls2 is not None:
        recs.append((b"P2BA", 1, 2, 1, bytes(basecalls2), len(basecalls2)))
    for name, number, etype, esize, payload in extra:
        recs.append((name, number, etype, esize, bytes(payload), len(payload) // max(esize, 1)))
    body = bytearray()
    dirent = []
    data_start = 128
    for name, number, etype, esize, payload, ne in recs:
        dsize = len(payload)
        if dsize > 4:
            doff = data_start + len(body)
            body += payload
            field = struct.pack(">i", doff)
        else:
            field = payload.ljust(4, b"\0")
        dirent.append(name + struct.pack(">ihhii", number, etype, esize, ne, dsize) + field + struct.pack(">i", 0))
    dir_off = data_start + len(body)
    head = b"ABIF" + struct.pack(">h", 101) + b"tdir" + struct.pack(">ihhiiii", 1, 1023, 28, len(dirent), 28 * len(dirent), dir_off, 0)
    return head.ljust(data_start, b"\0") + bytes(body) + b"".join(dirent) + b"\0" * 2


def scf_bytes(channels, ploc, version=b"3.00"):
    """A minimal SCF file (reference src/scf.h:56-93): 128-byte header, then for 3.x the four channels one after the other
    as twice-differenced big-endian int16, else interleaved plain samples; then int32 basecall positions."""
    import struct
    ch = [np.asarray(c, np.int64) for c in channels]
    ns = len(ch[0])
    if float(version) > 2.9:
        enc = []
        for c in ch:
            d = c.copy()
            for _ in range(2):
                d[1:] = d[1:] - d[:-1]
            enc.append(((d + 32768) % 65536 - 32768).astype(">i2").tobytes())
        samples = b"".join(enc)
    else:
        samples = np.stack(ch, 1).astype(">i2").tobytes()
    off = 128
    bases_off = off + len(samples)
    head = b".scf" + struct.pack(">iii", ns, off, len(ploc)) + struct.pack(">ii", 0, 0) + struct.pack(">i", bases_off) + struct.pack(">ii", 0, 0) + version
    return head.ljust(128, b"\0") + samples + np.asarray(ploc, ">i4").tobytes()
