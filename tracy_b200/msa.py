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
    """Orientation trials of revSeqBasedOnDist (every profile against one flipped profile) on a pool that stays in HBM for the
    whole stage, several trials per call. A trial fills only `num` of the GPU's ~1 800 warp slots and lasts as long as one pair
    on one warp, so up to T consecutive trials (in the reference's order) go out together, each against the state at the start
    of the group. The reference runs them one after the other and a kept flip of k_s changes ONE input of a later trial t of
    the group: the pair (k_s, flip of k_t). Those T(T-1)/2 pairs (flip of k_s against flip of k_t) ride along in the same call,
    and the host replays the reference's sequential accept rule on exact numbers -- same pairs, same kernel, same integers.
    Arena order: trial t = `num` pairs (slot i, spare t) followed by its t fix-up pairs (spare s, spare t), s < t, so a shorter
    last group is a prefix."""

    def __init__(self, ctx, pool, num, T):
        import torch
        self.torch, self.ctx, self.num, self.cap, self.T = torch, ctx, num, pool.cap, T
        dev = torch.device("cuda", ctx.device)
        self.base = torch.from_numpy(pool.base).to(dev)
        self.start = [t * num + t * (t - 1) // 2 for t in range(T + 1)]
        npairs = self.start[T]
        a1_off, a2_off = np.zeros(npairs, np.int64), np.zeros(npairs, np.int64)
        self.a1_len, self.a2_len = np.zeros(npairs, np.int32), np.zeros(npairs, np.int32)
        for t in range(T):
            b = self.start[t]
            a1_off[b: b + num] = pool.off[:num]
            self.a1_len[b: b + num] = pool.lens[:num]                     # a flip keeps the length
            a1_off[b + num: b + num + t] = pool.off[num: num + t]
            a2_off[b: b + num + t] = pool.off[num + t]
        self.a1_off, self.a2_off = torch.from_numpy(a1_off).to(dev), torch.from_numpy(a2_off).to(dev)
        self.d_a1_len = torch.zeros(npairs, dtype=torch.int32, device=dev)
        self.d_a2_len = torch.zeros(npairs, dtype=torch.int32, device=dev)
        self.scores = torch.zeros(npairs, dtype=torch.int32, device=dev)
        self.stage = torch.zeros(pool.cap, dtype=torch.float32).pin_memory()

    def put(self, slot, p):
        p = np.ascontiguousarray(p, np.float32)
        self.stage[: p.size] = self.torch.from_numpy(p.reshape(-1))
        self.base[slot * self.cap: slot * self.cap + p.size].copy_(self.stage[: p.size], non_blocking=False)
        return p.shape[1]

    def group(self, flips, sc):
        """flips: the flipped profiles of up to T consecutive trials. Returns (base int64[len(flips)][num], fix) with
        base[t][i] = score(slot i, flip t) and fix[t][s] = score(flip s, flip t) for s < t."""
        num, g = self.num, len(flips)
        lens = [self.put(num + t, p) for t, p in enumerate(flips)]
        for t in range(g):
            b = self.start[t]
            self.a1_len[b + num: b + num + t] = lens[:t]
            self.a2_len[b: b + num + t] = lens[t]
        n = self.start[g]
        self.d_a1_len[:n].copy_(self.torch.from_numpy(self.a1_len[:n]))
        self.d_a2_len[:n].copy_(self.torch.from_numpy(self.a2_len[:n]))
        self.torch.cuda.synchronize(self.base.device)               # torch's stream -> the context's stream
        self.ctx.gotoh_device(PP, self.base.data_ptr(), self.a1_off.data_ptr(), self.d_a1_len.data_ptr(), self.base.data_ptr(),
                              self.a2_off.data_ptr(), self.d_a2_len.data_ptr(), n, self.scores.data_ptr(), sc=sc, ac=_END_FREE)
        sco = self.scores[:n].cpu().numpy().astype(np.int64)
        base = [sco[self.start[t]: self.start[t] + num] for t in range(g)]
        fix = [sco[self.start[t] + num: self.start[t] + num + t] for t in range(g)]
        return base, fix


class _HostTrials:
    """The same grouping with the pool in host memory, for contexts without device entry points (the CPU test double of
    tests/test_glue.py): one ctx.gotoh call per group, arenas in the order _DeviceTrials uses."""

    def __init__(self, ctx, pool, num, T):
        self.ctx, self.pool, self.num, self.T = ctx, pool, num, T

    def put(self, slot, p):                                            # the caller keeps the host pool current itself
        pass

    def group(self, flips, sc):
        num = self.num
        i1, i2, start = [], [], [0]
        for t, p in enumerate(flips):
            self.pool.put(num + t, p)
            i1 += list(range(num)) + [num + s for s in range(t)]
            i2 += [num + t] * (num + t)
            start.append(len(i1))
        sco, _, _ = self.ctx.gotoh(PP, self.pool.arena(i1), self.pool.arena(i2), sc, _END_FREE, traceback=False)
        sco = np.asarray(sco, np.int64)
        return ([sco[start[t]: start[t] + num] for t in range(len(flips))],
                [sco[start[t] + num: start[t + 1]] for t in range(len(flips))])


def rev_seq_based_on_dist(ctx, profiles, fwd, sc):
    """revSeqBasedOnDist (src/msa.h:243-328). profiles: list of float32[6][len] (replaced in place when a flip is kept),
    fwd: list of bool (toggled in place). Returns the final symmetric score matrix."""
    seq = profiles
    num = len(seq)
    d = np.zeros((num, num), np.int64)
    ii, jj = np.triu_indices(num, 1)
    on_device = hasattr(ctx, "gotoh_device")                               # test doubles serve ctx.gotoh only
    T = min(8, max(1, 1700 // max(num, 1)))
    pool = _Pool(seq, spare=T)                                             # slots num.. hold the flips under trial
    if len(ii):
        s, _, _ = ctx.gotoh(PP, pool.arena(ii), pool.arena(jj), sc, _END_FREE, traceback=False)
        d[ii, jj] = s
        d[jj, ii] = s
    total = int(d[ii, jj].sum()) if len(ii) else 0
    trials = (_DeviceTrials if on_device else _HostTrials)(ctx, pool, num, T) if num > 1 else None
    iterate = True
    while iterate:
        quality = [k for _, k in sorted((int(d[i].sum()), i) for i in range(num))]   # worst row sum first, src/msa.h:270-282
        for g0 in range(0, num, T):
            ks = quality[g0: g0 + T]
            flips = [_revcomp(seq[k]) for k in ks]
            if trials is not None:
                base, fix = trials.group(flips, sc)
            kept = []                                                    # trials of this group whose flip was kept
            for t, k in enumerate(ks):
                s_rc = flips[t]
                others = [i for i in range(num) if i != k]
                new_d = np.zeros(num, np.int64)
                if others:
                    new_d = base[t].copy()
                    for s_ in kept:                                      # k_s was flipped after the group went out
                        new_d[ks[s_]] = fix[t][s_]
                    new_d[k] = 0                                         # the pair (k, flipped k) rides along and is dropped
                if int(new_d.sum()) >= int(d[others, k].sum()):          # scoreSum >= oldScoreSum, src/msa.h:298
                    seq[k] = s_rc
                    pool.put(k, s_rc)
                    if trials is not None:
                        trials.put(k, s_rc)
                    kept.append(t)
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
