"""Host glue of the assemble path around the batched DP kernels.

Mirrors reference src/msa.h (distanceMatrix :33-42, closestPair/updateDistanceMatrix/upgma :44-87, palign :89-160,
consensus :162-239, revSeqBasedOnDist :243-328, msa :330-368), `_createProfile(char MSA)` (src/align.h:138-180),
`_profileConsChar` (src/align.h:254-270) and the exclusion loop of assemble() (src/assemble.h:428-448).

Every gotohScore()/gotoh() call of those functions goes to the GPU through a `Context` (profile x profile,
AlignConfig<true,true>), batched wherever the reference's loop order leaves the calls independent:
  * distanceMatrix / the initial matrix of revSeqBasedOnDist: all N(N-1)/2 pairs in one call;
  * revSeqBasedOnDist: the 4 N (N-1) scores of all ordered pairs in all orientation combinations in one call (the
    orientation table); the sequential accept rule is replayed on the host over those exact numbers;
  * palign: all guide-tree nodes of equal height in one call;
  * the exclusion loop: the best-scoring partner of every trace first (existence of a hit is all that decides), then the
    remaining candidates in a few geometric rounds.
Tree building, row merging and consensus are integer/byte logic kept literal (tie-breaks, C++ integer division).
"""
import numpy as np

from .api import PP, AlignConfig, Arena, DnaScore

_END_FREE = AlignConfig(True, True)


class _Pool:
    """The profiles of one stage packed ONCE into a flat float32 array (reference layout per item) with spare slots; batches
    of pairs are then arenas of offsets into it -- a pair list of N(N-1)/2 entries costs two index arrays, not a copy of
    every profile per pair. The C ABI copies only the extent an arena touches."""

    def __init__(self, profiles, spare=0):
        ps = [np.ascontiguousarray(p, np.float32) for p in profiles]
        self.cap = 6 * max([p.shape[1] for p in ps] + [1])
        n = len(ps)
        self.base = np.zeros((n + spare) * self.cap, np.float32)
        self.off = np.arange(n + spare, dtype=np.int64) * self.cap
        self.lens = np.zeros(n + spare, np.int32)
        for i, p in enumerate(ps):
            self.put(i, p)

    def put(self, slot, p):
        p = np.ascontiguousarray(p, np.float32)
        self.lens[slot] = p.shape[1]
        self.base[self.off[slot]: self.off[slot] + p.size] = p.reshape(-1)

    def arena(self, idx):
        idx = np.asarray(idx, np.int64)
        return Arena(self.base, np.ascontiguousarray(self.off[idx]), np.ascontiguousarray(self.lens[idx]))


def _revcomp(p):
    """reverseComplementProfile (src/profile.h:74-90) on the host: columns reversed, rows A<->T and C<->G swapped, N and '-'
    kept -- a pure permutation, no arithmetic (Context.revcomp_profile is the batched device version)."""
    return np.ascontiguousarray(np.asarray(p, np.float32)[[3, 2, 1, 0, 4, 5], ::-1])


# ---- small literal helpers -------------------------------------------------------------------------------------------
def profile_cons_chars(p):
    """_profileConsChar for every column (src/align.h:254-270): first strict maximum over the six rows; index >= 4 -> 'N'."""
    p = np.asarray(p, np.float32)
    idx = np.argmax(p.astype(np.float64), axis=0)          # argmax returns the first maximum, like the strict '>' scan
    return np.frombuffer(b"ACGTNN", np.uint8)[idx].tobytes()


def profile_from_alignment(rows):
    """_createProfile(boost::multi_array<char,2>, p), src/align.h:138-180. rows: uint8[nrow][ncol] -> float32[6][ncol].
    A row takes part between its first and last non-gap character; the counts are accumulated over those spans only (a trace
    covers ~1 000 of the tens of thousands of columns of a large assembly)."""
    a = np.asarray(rows, np.uint8)
    nrow, ncol = a.shape
    notgap = a != 0x2D
    has = notgap.any(axis=1)
    first = notgap.argmax(axis=1)
    last = ncol - 1 - notgap[:, ::-1].argmax(axis=1)
    cnt = np.zeros((6, ncol), np.int64)
    total = np.zeros(ncol, np.int64)
    nfull = int((~has).sum())           # a row without any nucleotide keeps first = -1, last = ncol: covered everywhere (src/align.h:147-158)
    if nfull:
        cnt[5] += nfull
        total += nfull
    for i in np.nonzero(has)[0]:
        f, l = int(first[i]), int(last[i]) + 1
        seg = a[i, f:l]
        up = seg & 0xDF                                     # upper-case letters; '-' (0x2d) becomes 0x0d, never a letter
        known = seg == 0x2D
        cnt[5, f:l] += known
        for k, ch in enumerate(b"ACGTN"):
            hit = up == ch
            cnt[k, f:l] += hit
            known = known | hit
        total[f:l] += known                                 # `else --sum`: unknown characters do not count
    p = cnt.astype(np.float32)
    nz = total > 0
    p[:, nz] = p[:, nz] / total[nz].astype(np.float32)
    return p


def onehot_profile(seq):
    """_createProfile(std::string, p) (reference src/align.h:119-136): one column per character, 1 in the row of A, C, G, T, N (either
    case) or '-', all zero for anything else. (The '-' row never enters _score, but it decides the column's consensus character.)"""
    raw = np.frombuffer(bytes(seq), np.uint8)
    up = raw & 0xDF
    p = np.zeros((6, len(raw)), np.float32)
    letter = ((raw | 0x20) >= 0x61) & ((raw | 0x20) <= 0x7A)            # a byte that is a letter: folding is only meaningful there
    for k, ch in enumerate(b"ACGTN"):
        p[k, letter & (up == ch)] = 1
    p[5, raw == 0x2D] = 1
    return p


def _trunc_div2(x):
    return x // 2 if x >= 0 else -((-x) // 2)              # C++ integer division truncates toward zero


# ---- distance matrix, guide tree --------------------------------------------------------------------------------------
def distance_matrix(ctx, profiles, sc):
    """distanceMatrix (src/msa.h:33-42): d[i][j] = gotohScore(sps[i], sps[j], <true,true>) for i < j, one GPU batch."""
    n = len(profiles)
    ii, jj = np.triu_indices(n, 1)
    d = np.zeros((n, n), np.int64)
    if len(ii):
        pool = _Pool(profiles)
        s, _, _ = ctx.gotoh(PP, pool.arena(ii), pool.arena(jj), sc, _END_FREE, traceback=False)
        d[ii, jj] = s
    return d


def upgma(dist, num):
    """upgma (src/msa.h:72-87) on the (2*num+1)^2 matrix the reference uses: returns (phylogeny int[2num+1][3], root).
    closestPair starts at dMax = -1 with a strict '>', so pairs with a negative score are never joined (src/msa.h:47-50).
    The reference rescans the whole matrix for every join (O(n^3)); here every open row keeps its maximum and the first column
    that holds it, which a join changes in O(n): the global scan is then a scan over row maxima, and taking the lowest row among
    equal maxima (and the lowest column inside it) is the reference's first-maximum-in-(i, j)-order tie-break."""
    size = 2 * num + 1
    d = np.full((size, size), -1, np.int64)
    d[:num, :num] = np.where(np.triu(np.ones((num, num), bool), 1), dist[:num, :num], -1)
    p = np.full((size, 3), -1, np.int64)
    rowmax = np.full(size, -1, np.int64)
    rowarg = np.zeros(size, np.int64)

    def rescan(i, hi):                                       # row i over columns i+1 .. hi-1
        seg = d[i, i + 1: hi]
        if len(seg):
            k = int(np.argmax(seg))
            rowmax[i], rowarg[i] = seg[k], i + 1 + k
        else:
            rowmax[i] = -1
    for i in range(num):
        rescan(i, num)
    nn = num
    live = np.arange(num)                                    # nodes without a parent, ascending
    while nn < 2 * num + 1 and len(live) >= 2:
        a = int(np.argmax(rowmax[live]))
        di = int(live[a])
        if rowmax[di] <= -1:
            break
        dj = int(rowarg[di])
        p[di, 0] = nn; p[dj, 0] = nn; p[nn, 1] = di; p[nn, 2] = dj
        io = live[(live != di) & (live != dj)]              # updateDistanceMatrix, src/msa.h:60-70: the open nodes
        if len(io):
            x = np.where(di < io, d[di, io], d[io, di])
            y = np.where(dj < io, d[dj, io], d[io, dj])
            t = x + y
            d[io, nn] = np.where(t >= 0, t // 2, -((-t) // 2))   # C++ integer division truncates toward zero
        d[:di, di] = -1; d[di, di + 1: nn + 1] = -1
        d[:dj, dj] = -1; d[dj, dj + 1: nn + 1] = -1
        rowmax[di] = rowmax[dj] = -1
        if len(io):
            stale = (rowarg[io] == di) | (rowarg[io] == dj)
            for i in io[stale]:
                rescan(int(i), nn + 1)
            fresh = io[~stale]
            better = d[fresh, nn] > rowmax[fresh]           # the new column is the last one: it wins only when strictly larger
            rowmax[fresh[better]] = d[fresh[better], nn]
            rowarg[fresh[better]] = nn
        live = np.append(io, nn)
        nn += 1
    return p, (nn - 1 if nn > 0 else 0)


# ---- progressive alignment --------------------------------------------------------------------------------------------
def _merge_rows(rows1, rows2, new0, new1):
    """src/msa.h:126-146: lay the two row blocks out along the profile alignment's gap pattern."""
    ncol = len(new0)
    g0 = np.frombuffer(new0, np.uint8) != 0x2D
    g1 = np.frombuffer(new1, np.uint8) != 0x2D
    out = np.full((rows1.shape[0] + rows2.shape[0], ncol), 0x2D, np.uint8)
    out[: rows1.shape[0], g0] = rows1[:, : int(g0.sum())]
    out[rows1.shape[0]:, g1] = rows2[:, : int(g1.sum())]
    return out


def palign(ctx, sps, phylo, root, sc):
    """palign (src/msa.h:89-160), iteratively: returns (align uint8[nseq][ncol], profile float32[6][ncol], seqidx list).
    Nodes of equal height are independent, so each height level is one batched gotoh() call."""
    height, order = {}, []

    def visit(node):                                         # post-order without recursion limits
        stack = [(node, False)]
        while stack:
            v, done = stack.pop()
            l, r = int(phylo[v, 1]), int(phylo[v, 2])
            if l == -1 and r == -1:
                height[v] = 0
                continue
            if done:
                height[v] = 1 + max(height[l], height[r])
                order.append(v)
            else:
                stack.append((v, True)); stack.append((r, False)); stack.append((l, False))
    visit(root)
    res = {}
    for v, h in height.items():
        if h == 0:                                           # leaf: consensus characters of the trace profile, src/msa.h:92-96
            prof = np.ascontiguousarray(sps[v], np.float32)
            res[v] = (np.frombuffer(profile_cons_chars(prof), np.uint8).reshape(1, -1).copy(), prof.copy(), [v])
    for h in sorted(set(height[v] for v in order)):
        nodes = [v for v in order if height[v] == h]
        left = [res[int(phylo[v, 1])] for v in nodes]
        right = [res[int(phylo[v, 2])] for v in nodes]
        _, ops, ol = ctx.gotoh(PP, [x[1] for x in left], [x[1] for x in right], sc, _END_FREE, traceback=True)
        for k, v in enumerate(nodes):
            o = ops[k, : ol[k]]
            new0 = np.where(o == ord("h"), 0x2D, 0x58).astype(np.uint8).tobytes()   # only the gap pattern is consumed
            new1 = np.where(o == ord("v"), 0x2D, 0x58).astype(np.uint8).tobytes()
            rows = _merge_rows(left[k][0], right[k][0], new0, new1)
            res[v] = (rows, profile_from_alignment(rows), left[k][2] + right[k][2])
            res.pop(int(phylo[v, 1]), None); res.pop(int(phylo[v, 2]), None)
    return res[root]


def msa(ctx, profiles, sc, dist=None):
    """msa (src/msa.h:330-368): distance matrix -> UPGMA -> progressive alignment. Returns (align rows, seqidx, dist).
    dist: the matrix distanceMatrix() would compute, when the caller already holds it (oriented_distance)."""
    n = len(profiles)
    d = distance_matrix(ctx, profiles, sc) if dist is None else np.asarray(dist, np.int64)
    phylo, root = upgma(d, n)
    rows, _, seqidx = palign(ctx, profiles, phylo, root, sc)
    return rows, seqidx, d


def consensus(rows, fraction_called=0.5, ignore_last=False):
    """consensus (src/msa.h:162-239): returns (gapped, cs, qstr) byte strings. Counts are accumulated over each row's span."""
    a = np.asarray(rows, np.uint8)
    nrow = a.shape[0] - (1 if ignore_last else 0)
    ncol = a.shape[1]
    a = a[:nrow]
    notgap = a != 0x2D
    has = notgap.any(axis=1)                                 # a row of gaps only: start = ncol, end = -1 -> covers nothing
    start = notgap.argmax(axis=1)
    end = ncol - 1 - notgap[:, ::-1].argmax(axis=1)
    cov = np.zeros(ncol, np.int64)
    counts = np.zeros((5, ncol), np.int64)
    for i in np.nonzero(has)[0]:
        f, l = int(start[i]), int(end[i]) + 1
        up = a[i, f:l] & 0xDF
        cov[f:l] += 1
        for k, ch in enumerate(b"ACGT"):
            counts[k, f:l] += up == ch
    counts[4] = cov - counts[:4].sum(axis=0)
    thr = int(np.float32(fraction_called) * np.float32(nrow))   # (int32_t)(float * size_t): float arithmetic
    called = (cov >= 1) & (cov >= thr)
    max_idx = np.where(called, np.argmax(counts, axis=0), 4)  # first maximum, strict '>' scan
    top = counts[np.minimum(max_idx, 4), np.arange(ncol)]
    # qualval is only assigned in called columns and carried over otherwise, but it is only WRITTEN where max_idx < 4, i.e. in a
    # called column with a nucleotide majority, where it was just assigned: no carry-over is ever visible
    qv = (47 + top * 10 // max(nrow, 1)) & 0xFF
    cons = np.where(max_idx < 4, np.frombuffer(b"ACGT-", np.uint8)[np.minimum(max_idx, 4)], 0x2D).astype(np.uint8)
    qual = np.where(max_idx < 4, qv, ord("#")).astype(np.uint8)
    sel = cons != 0x2D
    return cons.tobytes(), cons[sel].tobytes(), qual[sel].tobytes()


# ---- orientation ------------------------------------------------------------------------------------------------------
def orientation_table(ctx, profiles, sc):
    """T[i][k][oi][ok] = gotohScore(P_i in orientation oi, P_k in orientation ok, AlignConfig<true,true>) for every ORDERED pair
    i != k (a1 = i, a2 = k; orientation 1 = reverseComplementProfile of the input) in ONE batched call: 4 N (N-1) fills.
    Everything revSeqBasedOnDist (src/msa.h:243-328) and the distance matrix of msa() (src/msa.h:33-42) ever ask the DP is an
    entry of this table, so the sequential accept rule becomes a host walk over exact numbers. Neither a1/a2 symmetry nor
    reverse-complement symmetry is assumed: the float substitution sum (src/align.h:112-116) changes its order under both."""
    num = len(profiles)
    T = np.zeros((num, num, 2, 2), np.int64)
    if num < 2:
        return T
    pool = _Pool(list(profiles) + [_revcomp(p) for p in profiles])       # slot i: as given, slot num + i: flipped
    ii, kk = np.nonzero(~np.eye(num, dtype=bool))
    a_idx = np.concatenate([ii + oi * num for oi in (0, 1) for _ in (0, 1)])
    b_idx = np.concatenate([kk + ok * num for _ in (0, 1) for ok in (0, 1)])
    s, _, _ = ctx.gotoh(PP, pool.arena(a_idx), pool.arena(b_idx), sc, _END_FREE, traceback=False)
    s = np.asarray(s, np.int64).reshape(2, 2, len(ii))
    for oi in (0, 1):
        for ok in (0, 1):
            T[ii, kk, oi, ok] = s[oi, ok]
    return T


def rev_seq_based_on_dist(ctx, profiles, fwd, sc, table=None, with_state=False):
    """revSeqBasedOnDist (src/msa.h:243-328). profiles: list of float32[6][len] (replaced in place when a flip is kept),
    fwd: list of bool (toggled in place). Returns the final symmetric score matrix (with_state: also the orientation table and
    the per-trace orientation bits, from which msa()'s distance matrix is read without another DP call).
    The reference runs one trial (num - 1 fills against the flipped profile) after the other; here all fills any trial can
    ask for are in the orientation table and the loop below is the reference's accept rule replayed on it."""
    seq = profiles
    num = len(seq)
    T = orientation_table(ctx, seq, sc) if table is None else table
    o = np.zeros(num, np.int64)                                        # orientation relative to the input
    idx = np.arange(num)
    d = np.zeros((num, num), np.int64)
    iu, ju = np.triu_indices(num, 1)
    d[iu, ju] = T[iu, ju, 0, 0]                                        # d[i][j] = gotohScore(seq[i], seq[j]), i < j (src/msa.h:251-260)
    d[ju, iu] = d[iu, ju]
    total = int(d[iu, ju].sum()) if num > 1 else 0
    iterate = num > 0
    while iterate:
        quality = [k for _, k in sorted((int(d[i].sum()), i) for i in range(num))]   # worst row sum first, src/msa.h:270-282
        for k in quality:
            new_d = T[idx, k, o, 1 - o[k]].copy()                     # gotohScore(seq[i], flipped seq[k]) for all i, src/msa.h:293
            new_d[k] = 0
            old = int(d[:, k].sum()) - int(d[k, k])
            if int(new_d.sum()) >= old:                               # scoreSum >= oldScoreSum, src/msa.h:298
                o[k] ^= 1
                fwd[k] = not fwd[k]
                d[:, k] = new_d
                d[k, :] = new_d
        updated = int(d.sum())
        if total < updated:
            total = updated
        else:
            iterate = False
    for k in range(num):
        if o[k]:
            seq[k] = _revcomp(seq[k])
    return (d, T, o) if with_state else d


def oriented_distance(T, o, keep=None):
    """distanceMatrix (src/msa.h:33-42) of the oriented traces read from the orientation table: d[i][j] = gotohScore(sps[i], sps[j])
    for i < j over the traces kept (index order)."""
    idx = np.arange(len(o)) if keep is None else np.asarray(keep, np.int64)
    n = len(idx)
    d = np.zeros((n, n), np.int64)
    iu, ju = np.triu_indices(n, 1)
    d[iu, ju] = T[idx[iu], idx[ju], o[idx[iu]], o[idx[ju]]]
    return d


def exclude_unmatched(ctx, profiles, sc, match_fraction, dist=None):
    """The exclusion loop of assemble() (src/assemble.h:428-448): trace i is kept iff some j != i aligns with > 10 % of i aligned,
    > 25 aligned columns and a score above the match-fraction threshold. The reference tries j = 0, 1, ... and stops at the first
    hit; only EXISTENCE of a hit decides, so the candidates may be tried in any order: with `dist` (the symmetric score matrix
    revSeqBasedOnDist leaves) the best-scoring partner goes first and an overlapping trace is settled by one alignment. Rounds of
    1, 3, 12 and then all remaining candidates per still-unmatched trace: a handful of batched calls instead of one per
    candidate rank. Returns a list of bool (True = keep)."""
    n = len(profiles)
    keep = [False] * n
    if n < 2:
        return keep
    if dist is not None:
        dm = -np.array(dist, np.int64)
        np.fill_diagonal(dm, np.iinfo(np.int64).max)            # the trace itself sorts last and is cut off
        order = np.argsort(dm, axis=1, kind="stable")[:, : n - 1]
    else:
        order = np.array([[j for j in range(n) if j != i] for i in range(n)], np.int64)
    pos = np.zeros(n, np.int64)
    pending = list(range(n))
    pool = _Pool(profiles)
    f32, mf = np.float32, np.float32(match_fraction)
    for block in (1, 3, 12, n):
        if not pending:
            break
        ia, ja = [], []
        for i in pending:
            js = order[i, pos[i]: pos[i] + block]
            pos[i] += len(js)
            ia += [i] * len(js)
            ja += [int(j) for j in js]
        s, ops, ol = ctx.gotoh(PP, pool.arena(ia), pool.arena(ja), sc, _END_FREE, traceback=True)
        ops = np.asarray(ops)
        col = np.arange(ops.shape[1])[None, :]
        aligned = ((ops == ord("s")) & (col < np.asarray(ol)[:, None])).sum(axis=1)
        hit = set()
        for q, i in enumerate(ia):
            num_aligned = int(aligned[q])
            frac = num_aligned / float(np.asarray(profiles[i]).shape[1])
            na = f32(num_aligned)                                      # float arithmetic, as in C++
            thr = float(f32(f32(na * mf) * f32(sc.match)) + f32(f32(na * f32(f32(1) - mf)) * f32(sc.mismatch)))
            if frac > 0.1 and num_aligned > 25 and int(s[q]) > thr:
                hit.add(i)
        for i in hit:
            keep[i] = True
        pending = [i for i in pending if i not in hit and pos[i] < n - 1]
    return keep
