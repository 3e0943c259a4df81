"""The consensus letters of `tracy consensus` and its output files (reference src/consensus.h). Host logic behind the DP:
the pairwise alignment comes from drivers.consensus_batch (two GPU calls for all trace pairs); what follows per alignment
column is a handful of double operations whose results the reference rounds to integers -- kept on the host in the
reference's own operation order (glibc log10 / pow), so letters and qualities are equal byte for byte.

  gt_letter            gtLetter            src/consensus.h:94-171
  pairwise_consensus   pairwiseConsensus   src/consensus.h:189-238 (consLetter :173-187)
  consensus_fasta / consensus_fastq / plot_clustal_pairwise / consensus_align_fasta    :64-91, :241-327, :561-573
"""
import math

import numpy as np

SMALLEST_GL = -1000.0
_ACGT = "ACGT"
_IUPAC2 = {(0, 2): "R", (1, 3): "Y", (1, 2): "S", (0, 3): "W", (2, 3): "K", (0, 1): "M"}      # iupac(c1, c2), src/abif.h:118-160


def _round(x):
    """boost::math::round for a finite double: half away from zero."""
    f = math.floor(x)
    d = x - f                    # exact
    if x >= 0:
        return f + (1 if d >= 0.5 else 0)
    return f + (1 if d > 0.5 else 0)


def _quality_of(best2nd_pl):
    """The genotype quality as a function of the second-best phred-scaled likelihood (bestPL is always 0 after the rescaling),
    src/consensus.h:136-141."""
    x = 1 - 1 / (math.pow(10.0, -(0.0 / 10.0)) + math.pow(10.0, -(float(best2nd_pl) / 10.0)))
    like = math.log10(x) if x > 0 else -math.inf                    # log10(0) = -inf in C
    like = like if like > SMALLEST_GL else SMALLEST_GL
    q = int(_round(-10 * like))
    return q if q > 0 else 0


_QUAL = {}


def gt_letter(cl, use_iupac=False):
    """gtLetter(c, cl, cons, qual) for one column: cl = the six weights (A, C, G, T, N, '-') as doubles.
    Returns (letter, quality)."""
    cl = [float(x) for x in cl]
    total = 0.0
    for x in cl:
        total += x
    gl = [0.0] * 6
    for k in range(6):
        v = cl[k] / total if total > 0 else 0.0
        if v > 0:
            g = math.log10(v)
            gl[k] = g if g >= SMALLEST_GL else SMALLEST_GL
        else:
            gl[k] = SMALLEST_GL
    best, second = 0, 1
    if gl[best] < gl[second]:
        best, second = 1, 0
    for k in range(2, 6):
        if gl[k] > gl[best]:
            second, best = best, k
        elif gl[k] > gl[second]:
            second = k
    ambiguous = bool(use_iupac) and gl[second] > -1 and best <= 3 and second <= 3
    gbest = gl[best]
    pl2 = int(_round(-10 * (gl[second] - gbest))) & 0xFFFFFFFF
    q = _QUAL.get(pl2)
    if q is None:
        q = _QUAL[pl2] = _quality_of(pl2)
    if ambiguous:
        a, b = (best, second) if best < second else (second, best)
        letter = _IUPAC2.get((a, b), "N")
    else:
        letter = _ACGT[best] if best <= 3 else ("N" if best == 4 else "-")
    return letter, q


def pairwise_consensus(row0, row1, p1, p2, compute_union=True, use_iupac=False):
    """pairwiseConsensus(c, align, trimmedtrace1, trimmedtrace2, cons, qual): row0 / row1 = the gapped rows of the global alignment
    of the two trimmed trace profiles p1 / p2 (float32[6][len]). Returns (consensus bytes, list of uint32 qualities)."""
    row0, row1 = bytes(row0), bytes(row1)
    p1 = np.asarray(p1, np.float32)
    p2 = np.asarray(p2, np.float32)
    s1 = s2 = 0
    cons, qual = [], []

    def emit(cl):
        c, q = gt_letter(cl, use_iupac)
        cons.append(c)
        qual.append(q)

    for a, b in zip(row0, row1):
        if a == 0x2D or b == 0x2D:
            if a != 0x2D:
                if compute_union:
                    emit(p1[:, s1].astype(np.float64))
                s1 += 1
            if b != 0x2D:
                if compute_union:
                    emit(p2[:, s2].astype(np.float64))
                s2 += 1
        else:
            emit((p1[:, s1] + p2[:, s2]).astype(np.float64))          # float + float in float, then widened (consLetter :176)
            s1 += 1
            s2 += 1
    return "".join(cons).encode(), qual


def consensus_fasta(label, cons):
    """consensusFastaOut, src/consensus.h:64-73."""
    return ">" + label + "\n" + bytes(cons).decode() + "\n"


def consensus_fastq(label, cons, qual):
    """consensusFastqOut, src/consensus.h:75-91: qualities + 33, capped at 'z'."""
    return "@" + label + "\n" + bytes(cons).decode() + "\n+\n" + "".join(chr(min(int(q) + 33, 122)) for q in qual) + "\n"


def consensus_align_fasta(stem1, stem2, row0, row1, forward):
    """P.align.fa of consensus(), src/consensus.h:561-573."""
    return ">" + stem1 + "\n" + bytes(row0).decode() + "\n>" + stem2 + (" (forward)" if forward else " (reverse)") + "\n" + bytes(row1).decode() + "\n"


def plot_clustal_pairwise(stem1, stem2, row0, row1, forward, score, linelimit=60):
    """plotClustalPairwise, src/consensus.h:241-327: the two ungapped sequences wrapped at linelimit + 14, the score, then the
    alignment in blocks of `linelimit` columns with a match line, padded to six blocks."""
    r0, r1 = bytes(row0).decode("latin-1"), bytes(row1).decode("latin-1")
    fald = linelimit + 14
    out = []

    def wrapped(row):
        s = row.replace("-", "")
        lines = [s[i: i + fald] for i in range(0, len(s), fald)]
        return "".join(x + "\n" for x in lines)

    out.append(">" + stem1 + "\n" + wrapped(r0))
    out.append(">" + stem2 + (" (forward)\n" if forward else " (reverse)\n") + wrapped(r1))
    bar = "#" + "-" * (fald - 1) + "\n"
    out.append("\nAlignment score: %d\n" % score + bar + "\n")
    f1, f2 = stem1[:8].ljust(8), stem2[:8].ljust(8)
    vi = ri = 1
    blocks = 0
    for s in range(0, len(r0), linelimit):
        a, b = r0[s: s + linelimit], r1[s: s + linelimit]
        out.append(f1 + "%5d " % vi + a + "\n")
        out.append(" " * 14 + "".join("|" if x == y else " " for x, y in zip(a, b)) + "\n")
        out.append(f2 + "%5d " % ri + b + "\n\n")
        vi += len(a) - a.count("-")
        ri += len(b) - b.count("-")
        blocks += 1
    if blocks < 6:
        out.append("\n" * (4 * (6 - blocks)))
    out.append(bar + bar + "\n\n")
    return "".join(out)


def pairwise_consensus_native(row0, row1, p1, p2, compute_union=True, use_iupac=False):
    """pairwise_consensus through tb_pairwise_consensus (csrc/trimq.cu): the same letters and qualities, without the per-column
    interpreter loop. Returns (consensus bytes, uint32 array)."""
    import ctypes as C
    from . import capi
    row0, row1 = bytes(row0), bytes(row1)
    a, b = np.ascontiguousarray(p1, np.float32), np.ascontiguousarray(p2, np.float32)
    cap = 2 * len(row0) + 1
    cons = C.create_string_buffer(cap)
    qual = np.zeros(cap, np.uint32)
    n = C.c_int32(0)
    rc = capi.lib().tb_pairwise_consensus(row0, row1, len(row0), a.ctypes.data, a.shape[1], b.ctypes.data, b.shape[1], int(bool(compute_union)), int(bool(use_iupac)),
                                          cons, qual.ctypes.data, C.byref(n))
    if rc != capi.TB_OK:
        raise ValueError("tb_pairwise_consensus: %d" % rc)
    return cons.raw[: n.value], qual[: n.value].copy()
