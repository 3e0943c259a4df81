"""Host glue of the assemble path around the batched DP kernels.

Mirrors reference src/msa.h (distanceMatrix :33-42, closestPair/updateDistanceMatrix/upgma :44-87, palign :89-160,
consensus :162-239, revSeqBasedOnDist :243-328, msa :330-368), `_createProfile(char MSA)` (src/align.h:138-180),
`_profileConsChar` (src/align.h:254-270) and the exclusion loop of assemble() (src/assemble.h:428-448).

Every gotohScore()/gotoh() call of those functions goes to the GPU through a `Context` (profile x profile,
AlignConfig<true,true>), batched wherever the reference's loop order leaves the calls independent:
  * distanceMatrix / the initial matrix of revSeqBasedOnDist: all N(N-1)/2 pairs in one call;
  * one orientation trial of revSeqBasedOnDist: the N-1 scores against the flipped profile in one call (the trials
    themselves stay sequential: a flip changes the inputs of the next trial);
  * palign: all guide-tree nodes of equal height in one call;
  * the exclusion loop: round r pairs every still-unmatched trace with its r-th candidate.
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
    return bytes(b"ACGTNN"[int(k)] for k in idx)


def profile_from_alignment(rows):
    """_createProfile(boost::multi_array<char,2>, p), src/align.h:138-180. rows: uint8[nrow][ncol] -> float32[6][ncol]."""
    a = np.asarray(rows, np.uint8)
    nrow, ncol = a.shape
    notgap = a != 0x2D
    first = np.where(notgap.any(axis=1), notgap.argmax(axis=1), -1)
    last = np.where(notgap.any(axis=1), ncol - 1 - notgap[:, ::-1].argmax(axis=1), ncol)
    # a row without any nucleotide keeps first = -1, last = ncol: covered everywhere (src/align.h:147-158)
    cols = np.arange(ncol)[None, :]
    cover = (first[:, None] <= cols) & (cols <= last[:, None])
    up = a & 0xDF                                           # upper-case letters; '-' (0x2d) becomes 0x0d, never a letter
    p = np.zeros((6, ncol), np.float32)
    known = np.zeros((nrow, ncol), bool)
    for k, ch in enumerate(b"ACGTN"):
        hit = cover & (up == ch) & notgap
        p[k] = hit.sum(axis=0)
        known |= hit
    gap = cover & ~notgap
    p[5] = gap.sum(axis=0)
    known |= gap
    total = (cover & known).sum(axis=0)                     # `else --sum`: unknown characters do not count
    nz = total > 0
    p[:, nz] = p[:, nz] / total[nz].astype(np.float32)
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
    closestPair starts at dMax = -1 with a strict '>', so pairs with a negative score are never joined (src/msa.h:47-50)."""
    size = 2 * num + 1
    d = np.full((size, size), -1, np.int64)
    d[:num, :num] = np.where(np.triu(np.ones((num, num), bool), 1), dist[:num, :num], -1)
    p = np.full((size, 3), -1, np.int64)
    nn = num
    while nn < 2 * num + 1:
        sub = np.where(np.triu(np.ones((nn, nn), bool), 1), d[:nn, :nn], -1)
        flat = int(np.argmax(sub))                          # first maximum in (i, j) row-major order = the reference's scan
        di, dj = divmod(flat, nn)
        if nn < 2 or sub[di, dj] <= -1:
            break
        p[di, 0] = nn; p[dj, 0] = nn; p[nn, 1] = di; p[nn, 2] = dj
        for i in range(nn):                                 # updateDistanceMatrix, src/msa.h:60-70
            if p[i, 0] == -1:
                a = d[di, i] if di < i else d[i, di]
                b = d[dj, i] if dj < i else d[i, dj]
                d[i, nn] = _trunc_div2(int(a) + int(b))
        d[:di, di] = -1; d[di, di + 1: nn + 1] = -1
        d[:dj, dj] = -1; d[dj, dj + 1: nn + 1] = -1
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


def msa(ctx, profiles, sc):
    """msa (src/msa.h:330-368): distance matrix -> UPGMA -> progressive alignment. Returns (align rows, seqidx, dist)."""
    n = len(profiles)
    d = distance_matrix(ctx, profiles, sc)
    phylo, root = upgma(d, n)
    rows, _, seqidx = palign(ctx, profiles, phylo, root, sc)
    return rows, seqidx, d


def consensus(rows, fraction_called=0.5, ignore_last=False):
    """consensus (src/msa.h:162-239): returns (gapped, cs, qstr) byte strings."""
    a = np.asarray(rows, np.uint8)
    nrow = a.shape[0] - (1 if ignore_last else 0)
    ncol = a.shape[1]
    a = a[:nrow]
    notgap = a != 0x2D
    has = notgap.any(axis=1)
    start = np.where(has, notgap.argmax(axis=1), ncol)       # a row of gaps only: start = ncol, end = -1 -> covers nothing
    end = np.where(has, ncol - 1 - notgap[:, ::-1].argmax(axis=1), -1)
    cols = np.arange(ncol)[None, :]
    fl = (start[:, None] <= cols) & (cols <= end[:, None])
    cov = fl.sum(axis=0)
    thr = int(np.float32(fraction_called) * np.float32(nrow))   # (int32_t)(float * size_t): float arithmetic
    up = a & 0xDF
    counts = np.stack([(fl & (up == ch) & notgap).sum(axis=0) for ch in b"ACGT"] + [np.zeros(ncol, np.int64)])
    counts[4] = cov - counts[:4].sum(axis=0)
    cons = bytearray(b"-" * ncol)
    qual = bytearray(b"#" * ncol)
    qualval = 33
    for j in range(ncol):
        max_idx = 4
        if cov[j] >= 1 and cov[j] >= thr:
            max_idx = int(np.argmax(counts[:, j]))           # first maximum, strict '>' scan
            qualval = 47 + int(counts[max_idx, j]) * 10 // nrow
        if max_idx < 4:
            cons[j] = b"ACGT"[max_idx]
            qual[j] = qualval & 0xFF
    cs = bytes(c for c in cons if c != 0x2D)
    qs = bytes(q for c, q in zip(cons, qual) if c != 0x2D)
    return bytes(cons), cs, qs


# ---- orientation ------------------------------------------------------------------------------------------------------
class _DeviceTrials:
    """One orientation trial of revSeqBasedOnDist = every profile against the flipped one. With a device context the pool stays
    in HBM for the whole stage: the arenas of a trial (all `num` slots against the spare slot) never change, so a trial is one
    21 KB upload of the flipped profile, one tb_gotoh_pp on device pointers and one read of `num` scores -- no per-trial copy
    of the pool through pinned staging (that was 8 ms per trial at 512 traces, against ~1 ms of kernel)."""

    def __init__(self, ctx, pool, num):
        import torch
        self.torch, self.ctx, self.num, self.cap = torch, ctx, num, pool.cap
        dev = torch.device("cuda", ctx.device)
        self.base = torch.from_numpy(pool.base).to(dev)
        self.off1 = torch.from_numpy(np.ascontiguousarray(pool.off[:num])).to(dev)
        self.len1 = torch.from_numpy(np.ascontiguousarray(pool.lens[:num])).to(dev)
        self.off2 = torch.full((num,), int(pool.off[num]), dtype=torch.int64, device=dev)
        self.len2 = torch.zeros(num, dtype=torch.int32, device=dev)
        self.scores = torch.zeros(num, dtype=torch.int32, device=dev)
        self.stage = torch.zeros(pool.cap, dtype=torch.float32).pin_memory()

    def put(self, slot, p):
        p = np.ascontiguousarray(p, np.float32)
        self.stage[: p.size] = self.torch.from_numpy(p.reshape(-1))
        self.base[slot * self.cap: slot * self.cap + p.size].copy_(self.stage[: p.size], non_blocking=False)
        return p.shape[1]

    def trial(self, s_rc, sc):
        self.len2.fill_(self.put(self.num, s_rc))
        self.torch.cuda.synchronize(self.base.device)               # torch's stream -> the context's stream
        self.ctx.gotoh_device(PP, self.base.data_ptr(), self.off1.data_ptr(), self.len1.data_ptr(), self.base.data_ptr(), self.off2.data_ptr(),
                              self.len2.data_ptr(), self.num, self.scores.data_ptr(), sc=sc, ac=_END_FREE)
        return self.scores.cpu().numpy().astype(np.int64)


def rev_seq_based_on_dist(ctx, profiles, fwd, sc):
    """revSeqBasedOnDist (src/msa.h:243-328). profiles: list of float32[6][len] (replaced in place when a flip is kept),
    fwd: list of bool (toggled in place). Returns the final symmetric score matrix."""
    seq = profiles
    num = len(seq)
    d = np.zeros((num, num), np.int64)
    ii, jj = np.triu_indices(num, 1)
    pool = _Pool(seq, spare=1)                                             # slot `num` holds the flip under trial
    if len(ii):
        s, _, _ = ctx.gotoh(PP, pool.arena(ii), pool.arena(jj), sc, _END_FREE, traceback=False)
        d[ii, jj] = s
        d[jj, ii] = s
    total = int(d[ii, jj].sum()) if len(ii) else 0
    dev = _DeviceTrials(ctx, pool, num) if num > 1 and hasattr(ctx, "gotoh_device") else None   # test doubles serve ctx.gotoh only
    iterate = True
    while iterate:
        quality = sorted((int(d[i].sum()), i) for i in range(num))      # worst row sum first, src/msa.h:270-282
        for _, k in quality:
            s_rc = _revcomp(seq[k])
            others = [i for i in range(num) if i != k]
            new_d = np.zeros(num, np.int64)
            if others and dev is not None:
                new_d = dev.trial(s_rc, sc)
                new_d[k] = 0                                             # the pair (k, flipped k) rides along and is dropped
            elif others:
                pool.put(num, s_rc)
                sc_new, _, _ = ctx.gotoh(PP, pool.arena(others), pool.arena([num] * len(others)), sc, _END_FREE, traceback=False)
                new_d[others] = sc_new
            if int(new_d.sum()) >= int(d[others, k].sum()):              # scoreSum >= oldScoreSum, src/msa.h:298
                seq[k] = s_rc
                pool.put(k, s_rc)
                if dev is not None:
                    dev.put(k, s_rc)
                fwd[k] = not fwd[k]
                d[:, k] = new_d
                d[k, :] = new_d
        updated = int(d.sum())
        if total < updated:
            total = updated
        else:
            iterate = False
    return d


def exclude_unmatched(ctx, profiles, sc, match_fraction):
    """The exclusion loop of assemble() (src/assemble.h:428-448): trace i is kept iff some j != i (first hit in index order
    is enough, so only existence matters) aligns with > 10 % of i aligned, > 25 aligned columns and a score above the
    match-fraction threshold. Returns a list of bool (True = keep)."""
    n = len(profiles)
    keep = [False] * n
    cand = {i: [j for j in range(n) if j != i] for i in range(n)}
    pending = [i for i in range(n) if cand[i]]
    pool = _Pool(profiles)
    while pending:
        js = [cand[i].pop(0) for i in pending]
        s, ops, ol = ctx.gotoh(PP, pool.arena(pending), pool.arena(js), sc, _END_FREE, traceback=True)
        nxt = []
        for k, i in enumerate(pending):
            num_aligned = int((ops[k, : ol[k]] == ord("s")).sum())
            frac = num_aligned / float(np.asarray(profiles[i]).shape[1])
            f32, mf, na = np.float32, np.float32(match_fraction), np.float32(num_aligned)   # float arithmetic, as in C++
            thr = float(f32(f32(na * mf) * f32(sc.match)) + f32(f32(na * f32(f32(1) - mf)) * f32(sc.mismatch)))
            if frac > 0.1 and num_aligned > 25 and int(s[k]) > thr:
                keep[i] = True
            elif cand[i]:
                nxt.append(i)
        pending = nxt
    return keep
