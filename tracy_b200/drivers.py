"""Batch drivers: the DP sequences of tracy's subcommand drivers, run for MANY traces at once.

  align_batch      sage()     for single-FASTA references   reference src/sage.h:222-260, :311
  align_genome_batch  sage()  for an indexed genome         reference src/sage.h:216-222, :258-260, :311; src/fmindex.h:236-326
  decompose_batch  indigo()   for single-FASTA references   reference src/indigo.h:190-388
  assemble_denovo  assemble() de novo branch                reference src/assemble.h:418-471
  consensus_batch  consensus() up to the overlap check      reference src/consensus.h:499-556

Everything between the file readers and the file writers of those drivers: orientation pick, semi-global alignment of the
trimmed trace, reference-slice trimming, final alignment (align); orientation optimisation, exclusion of unmatched traces,
MSA and consensus (assemble). Every DP call is one batched GPU call over all traces; the glue in between is the literal
host logic of tracy_b200.api / tracy_b200.msa. File formats, basecalling and the JSON/plot writers stay with the caller.
"""
import numpy as np

from . import decompose, msa
from .api import PP, PS, SS, AlignConfig, DnaScore, find_breakpoint, reference_slice, rows_from_ops, trim_reference_slice

_SEMIGLOBAL = AlignConfig(True, False)          # AlignConfig<true, false>, src/sage.h:165
_COMP = {ord("A"): ord("T"), ord("C"): ord("G"), ord("G"): ord("C"), ord("T"): ord("A"), ord("N"): ord("N")}


_UPPER_TABLE = np.frombuffer(bytes(range(256)).upper(), np.uint8)
_COMP_TABLE = np.zeros(256, np.uint8)
for _k, _v in _COMP.items():
    _COMP_TABLE[_k] = _v


def reverse_complement_seq(seq):
    """reverseComplement(std::string&), reference src/fmindex.h:11-26: upper-cased reverse, A<->T, C<->G, N kept; any other
    character leaves the ORIGINAL character of that position in place (the reference's `default: break`)."""
    seq = bytes(seq)
    a = np.frombuffer(seq, np.uint8)
    mapped = _COMP_TABLE[_UPPER_TABLE[a[::-1]]]
    return np.where(mapped != 0, mapped, a).astype(np.uint8).tobytes()


def _gap_rows(ops):
    """The gap pattern of gotoh()'s two rows from an s/h/v string ('X' where the reference has a nucleotide)."""
    o = np.frombuffer(bytes(ops), np.uint8)
    row0 = np.where(o == ord("h"), 0x2D, 0x58).astype(np.uint8).tobytes()
    row1 = np.where(o == ord("v"), 0x2D, 0x58).astype(np.uint8).tobytes()
    return row0, row1


def align_batch(ctx, trimmed_profiles, full_profiles, references, sc=DnaScore(3, -5, -10, -4), trim_left=50, trim_right=50):
    """`tracy align` against single-FASTA references for a batch of traces (reference src/sage.h:233-260, :311).

    trimmed_profiles / full_profiles: createProfile() of each trace with and without the quality trim (float32[6][len]);
    references: one reference sequence (bytes over ACGTN) per trace.
    Returns one dict per trace: forward, refslice (after trimReferenceSlice), pos, score, row0, row1 (the final alignment)."""
    n = len(references)
    refs = [bytes(r) for r in references]
    rc = [reverse_complement_seq(r) for r in refs]
    tls, trs = np.broadcast_to(trim_left, (n,)), np.broadcast_to(trim_right, (n,))         # one trim pair for all, or one per trace (-t)
    # gsFwd / gsRev (src/sage.h:239-240): one score-only batch of 2N pairs
    s = ctx.gotoh(PS, list(trimmed_profiles) + list(trimmed_profiles), refs + rc, sc, _SEMIGLOBAL, traceback=False)[0]
    forward = [bool(s[i] > s[n + i]) for i in range(n)]                    # strict '>', src/sage.h:247
    pref = [refs[i] if forward[i] else rc[i] for i in range(n)]
    # gotoh(trimmedtrace, prefslice) + trimReferenceSlice (src/sage.h:258-259)
    _, ops, ol = ctx.gotoh(PS, trimmed_profiles, pref, sc, _SEMIGLOBAL)
    slices, pos = [], []
    for i in range(n):
        r0, r1 = _gap_rows(ops[i, : ol[i]])
        sl, p = trim_reference_slice(r0, r1, pref[i], forward[i], 0, int(tls[i]), int(trs[i]))
        slices.append(sl); pos.append(p)
    # gotoh(fulltraceprofile, referenceprofile) (src/sage.h:311): the reported alignment
    score, ops2, ol2 = ctx.gotoh(PS, full_profiles, slices, sc, _SEMIGLOBAL)
    out = []
    for i in range(n):
        row0, row1 = rows_from_ops(PS, full_profiles[i], slices[i], bytes(ops2[i, : ol2[i]]))
        out.append(dict(forward=forward[i], refslice=slices[i], pos=pos[i], score=int(score[i]), row0=row0, row1=row1))
    return out


def align_genome_batch(ctx, index, seqs, consensus, trimmed_profiles, full_profiles, sc=DnaScore(3, -5, -10, -4), trim_left=50,
                       trim_right=50, kmer=15, min_kmer_support=3, maxindel=1000):
    """`tracy align` against an INDEXED GENOME for a batch of traces (reference src/sage.h:216-222, :258-260, :311).

    index: Context.build_index over b"\n".join(seqs) + b"\n" (what `tracy index` dumps); seqs: the genome's sequences;
    consensus: BaseCalls::consensus per trace; trimmed_profiles / full_profiles as in align_batch.
    getReferenceSlice (k-mer anchoring on the GPU, then the slice arithmetic and htslib's inclusive fetch), the semi-global
    alignment of the trimmed trace, trimReferenceSlice and the final alignment. Returns one dict per trace, or None where
    the reference prints "Couldn't anchor the Sanger trace" and gives up."""
    n = len(consensus)
    a = ctx.anchor(index, consensus, trim_left, trim_right, kmer, min_kmer_support)
    seqlen = [len(x) + 1 for x in seqs]                                    # src/fmindex.h:247
    live, pref, meta = [], [], []
    for i in range(n):
        if not a["anchored"][i]:
            continue
        ri, _, s0, s1 = reference_slice(int(a["bestpos"][i]), seqlen, len(consensus[i]), maxindel)
        sl = bytes(seqs[ri][s0: min(s1, len(seqs[ri]) - 1) + 1]).upper()   # faidx_fetch_seq is end-inclusive; to_upper_copy :302
        fw = bool(a["forward"][i])
        live.append(i); pref.append(sl if fw else reverse_complement_seq(sl)); meta.append((ri, s0, fw))
    out = [None] * n
    if not live:
        return out
    _, ops, ol = ctx.gotoh(PS, [trimmed_profiles[i] for i in live], pref, sc, _SEMIGLOBAL)
    slices, pos = [], []
    for j, i in enumerate(live):
        r0, r1 = _gap_rows(ops[j, : ol[j]])
        sl, p = trim_reference_slice(r0, r1, pref[j], meta[j][2], meta[j][1], trim_left, trim_right)
        slices.append(sl); pos.append(p)
    score, ops2, ol2 = ctx.gotoh(PS, [full_profiles[i] for i in live], slices, sc, _SEMIGLOBAL)
    for j, i in enumerate(live):
        row0, row1 = rows_from_ops(PS, full_profiles[i], slices[j], bytes(ops2[j, : ol2[j]]))
        out[i] = dict(forward=meta[j][2], chr=meta[j][0], kmersupport=int(a["kmersupport"][i]), refslice=slices[j], pos=pos[j],
                      score=int(score[j]), row0=row0, row1=row1)
    return out


def consensus_batch(ctx, profiles1, profiles2, sc=DnaScore(3, -5, -10, -4), min_overlap=0, match_fraction=0.0):
    """The DP sequence of `tracy consensus` for a batch of trace PAIRS (reference src/consensus.h:499-556).

    profiles1 / profiles2: createProfile() of the two trimmed traces of every pair (float32[6][len]).
    Orientation of the second trace by two global score fills (strict '>', :517-527), the global alignment (:535), the overlap
    statistics and the minOverlap / matchFraction gate (:538-549). Returns per pair dict(forward, score, row0, row1, num_aligned,
    num_match, ok); pairwiseConsensus and the writers stay with the caller."""
    glob = AlignConfig(True, True)                                         # AlignConfig<true, true>, src/consensus.h:464
    n = len(profiles1)
    rev2 = ctx.revcomp_profile(profiles2)                                  # reverseComplementProfile, :514
    s = ctx.gotoh(PP, list(profiles1) + list(profiles1), list(profiles2) + list(rev2), sc, glob, traceback=False)[0]
    forward = [bool(s[i] > s[n + i]) for i in range(n)]
    second = [profiles2[i] if forward[i] else rev2[i] for i in range(n)]
    score, ops, ol = ctx.gotoh(PP, profiles1, second, sc, glob)
    out = []
    for i in range(n):
        row0, row1 = rows_from_ops(PP, profiles1[i], second[i], bytes(ops[i, : ol[i]]))
        a0, a1 = np.frombuffer(row0, np.uint8), np.frombuffer(row1, np.uint8)
        both = (a0 != 0x2D) & (a1 != 0x2D)
        na, nm = int(both.sum()), int((both & (a0 == a1)).sum())
        frac = nm / na if na else 0.0
        out.append(dict(forward=forward[i], score=int(score[i]), row0=row0, row1=row1, num_aligned=na, num_match=nm,
                        ok=not (na < min_overlap or frac < match_fraction)))
    return out


def trimmed_seq(s, ltrim, rtrim):
    """trimmedSeq, reference src/abif.h:68-75."""
    s = bytes(s)
    if ltrim + rtrim + 1 >= len(s):
        return s
    return s[ltrim: len(s) - rtrim]


def find_homozygous_breakpoint(row0, row1):
    """findHomozygousBreakpoint(align, bp), reference src/decompose.h:59-128: breakpoint of a homozygous indel from the
    mismatch density either side of every alignment column. Returns None where the reference returns false, else
    dict(indelshift, traceleft, breakpoint, bestDiff). bestDiff is a float in the reference (compared as double)."""
    row0, row1 = bytes(row0), bytes(row1)
    L = len(row0)
    gap = 0x2D
    start = end = var_index = 0
    for j in range(L):
        if row0[j] != gap and row1[j] != gap:
            start = j
            break
        if row0[j] != gap:
            var_index += 1
    for j in range(L - 1, -1, -1):
        if row0[j] != gap and row1[j] != gap:
            end = j
            break
    w = 25
    if start >= end or end < start + 2 * w:
        return None
    mism = np.frombuffer(row0, np.uint8) != np.frombuffer(row1, np.uint8)
    csum = np.concatenate([[0], np.cumsum(mism)])
    best = np.float32(0)
    traceleft, bp = True, 0
    for i in range(start, start + w):
        if row0[i] != gap:
            var_index += 1
    for i in range(start + w, end - w):
        if row0[i] != gap:
            var_index += 1
        left = float(csum[i] - csum[i - w]) / float(w)
        right = float(csum[i + w] - csum[i]) / float(w)
        diff = abs(right - left)
        if diff > float(best):
            bp, best, traceleft = var_index, np.float32(diff), left < right
    indelshift = True
    if float(best) < 0.25:
        indelshift, bp, traceleft, best = False, var_index, True, np.float32(0)
    return dict(indelshift=indelshift, traceleft=traceleft, breakpoint=bp, bestDiff=float(best))


def generate_secondary_decomposed(trace, bcpos, primary, secondary):
    """generateSecondaryDecomposed(tr, bc), reference src/decompose.h:378-410: the second allele as plain nucleotides; an
    IUPAC pair is resolved to its higher peak (strict '>' for the first base of the pair)."""
    pair = {ord("R"): (0, 2, b"AG"), ord("Y"): (1, 3, b"CT"), ord("S"): (1, 2, b"CG"), ord("W"): (0, 3, b"AT"), ord("K"): (2, 3, b"GT"),
            ord("M"): (0, 1, b"AC")}
    pri, sec = bytes(primary), bytes(secondary)
    out = bytearray(len(sec))
    for i in range(min(len(pri), len(sec))):
        if pri[i] == sec[i]:
            out[i] = pri[i]
        elif sec[i] in b"ACGT":
            out[i] = sec[i]
        elif sec[i] in pair:
            a, b, ch = pair[sec[i]]
            out[i] = ch[0] if trace[a][bcpos[i]] > trace[b][bcpos[i]] else ch[1]
        else:
            out[i] = ord("N")
    return bytes(out)


def decompose_batch(ctx, traces, bcpos, primaries, secondaries, references, sc=DnaScore(3, -5, -10, -4), trim_left=50, trim_right=50,
                    maxindel=1000, madc=5):
    """`tracy decompose` against single-FASTA references for a batch of basecalled traces: the DP sequence of indigo()
    (reference src/indigo.h:190-388) with every stage batched over the traces.

    traces: int32[4][nsamples] each; bcpos / primaries / secondaries: BaseCalls::bcPos / primary / secondary.
    Returns one dict per trace, or None where indigo() gives up (alignment below its score threshold, no usable homozygous
    breakpoint): forward, refslice, score, row0/row1 (trimmed trace vs reference), breakpoint dict, primary / secondary after
    decomposeAlleles, secDecompose, decomp table, align1 / align2 / align3 = (score, row0, row1, refslice, pos)."""
    n = len(traces)
    refs = [bytes(r) for r in references]
    rc = [reverse_complement_seq(r) for r in refs]
    prof = ctx.create_profile(traces, bcpos, primaries, secondaries, trim_left, trim_right)        # src/indigo.h:190-192
    bps = [find_breakpoint(p) for p in prof]                                                        # src/indigo.h:195-196
    s = ctx.gotoh(PS, prof + prof, refs + rc, sc, _SEMIGLOBAL, traceback=False)[0]                  # src/indigo.h:235-236
    forward = [bool(s[i] > s[n + i]) for i in range(n)]
    rsl = [refs[i] if forward[i] else rc[i] for i in range(n)]
    ali, ops, ol = ctx.gotoh(PS, prof, rsl, sc, _SEMIGLOBAL)                                        # src/indigo.h:302
    out = [None] * n
    items, live = [], []
    for i in range(n):
        seqsize = float(prof[i].shape[1])
        if int(ali[i]) <= seqsize * 0.35 * sc.match + seqsize * (1 - 0.35) * sc.mismatch:           # src/indigo.h:303-309
            continue
        row0, row1 = rows_from_ops(PS, prof[i], rsl[i], bytes(ops[i, : ol[i]]))
        bp = bps[i]
        if not bp["indelshift"]:                                                                    # src/indigo.h:314-317
            bp = find_homozygous_breakpoint(row0, row1)
            if bp is None:
                continue
        out[i] = dict(forward=forward[i], refslice=rsl[i], score=int(ali[i]), row0=row0, row1=row1, breakpoint=bp)
        items.append(dict(row0=row0, row1=row1, primary=bytes(primaries[i]), secondary=bytes(secondaries[i]), trim_left=trim_left,
                          trim_right=trim_right, maxindel=maxindel, madc=madc, breakpoint=bp["breakpoint"], refslice_len=len(rsl[i])))
        live.append(i)
    for i, (pri, sec, dcp, info) in zip(live, decompose.decompose_alleles_batch(ctx, items)):       # src/indigo.h:340
        out[i].update(primary=pri, secondary=sec, decomp=dcp, decompose_mode=info["mode"])
        out[i]["secDecompose"] = generate_secondary_decomposed(traces[i], bcpos[i], pri, sec)       # src/indigo.h:344
    if not live:
        return out
    # allele-specific alignments (src/indigo.h:355-388): string x string, two semi-global rounds and one global
    # allelicFraction (src/indigo.h:350): the 0.01-grid fit of the two alleles' signal shares, one GPU call for all traces;
    # the reference indexes bcPos[i + trimLeft] even when trimmedSeq() left a short read untrimmed (out of bounds there), so
    # such reads keep the start value
    fit = [i for i in live if trim_left + trim_right + 1 < len(out[i]["primary"])]
    fr = ctx.allelic_fraction([traces[i] for i in fit], [bcpos[i] for i in fit], [out[i]["primary"] for i in fit],
                              [out[i]["secDecompose"] for i in fit], trim_left, trim_right) if fit else []
    for i in live:
        out[i]["allele_fractions"] = (0.5, 0.5)
    for i, f in zip(fit, fr):
        out[i]["allele_fractions"] = (float(f[0]), float(f[1]))
    pris = [trimmed_seq(out[i]["primary"], trim_left, trim_right) for i in live]
    secs = [trimmed_seq(out[i]["secDecompose"], trim_left, trim_right) for i in live]
    refl = [out[i]["refslice"] for i in live]
    _, o1, l1 = ctx.gotoh(SS, pris + secs, refl + refl, sc, _SEMIGLOBAL)
    sl = []
    for k in range(2 * len(live)):
        i = live[k % len(live)]
        r0, r1 = _gap_rows(o1[k, : l1[k]])
        sl.append(trim_reference_slice(r0, r1, refl[k % len(live)], out[i]["forward"], 0, trim_left, trim_right))
    s2, o2, l2 = ctx.gotoh(SS, pris + secs, [x[0] for x in sl], sc, _SEMIGLOBAL)
    s3, o3, l3 = ctx.gotoh(SS, pris, secs, sc, AlignConfig(False, False))
    nl = len(live)
    for k, i in enumerate(live):
        for name, kk, q in (("align1", k, pris[k]), ("align2", nl + k, secs[k])):
            rows = rows_from_ops(SS, q, sl[kk][0], bytes(o2[kk, : l2[kk]]))
            out[i][name] = dict(score=int(s2[kk]), row0=rows[0], row1=rows[1], refslice=sl[kk][0], pos=sl[kk][1])
        rows = rows_from_ops(SS, pris[k], secs[k], bytes(o3[k, : l3[k]]))
        out[i]["align3"] = dict(score=int(s3[k]), row0=rows[0], row1=rows[1], refslice=secs[k], pos=0)
    return out


def assemble_denovo(ctx, profiles, sc=DnaScore(3, -5, -10, -4), match_fraction=0.5, fraction_called=0.1, table=None):
    """The de novo branch of `tracy assemble` from the trace profiles on (reference src/assemble.h:418-471):
    revSeqBasedOnDist -> exclusion of traces that match nothing -> msa -> consensus.
    Returns dict(forward, kept (indices into `profiles`), rows, seqidx (into kept), gapped, consensus, quality).
    table: msa.orientation_table(ctx, profiles, sc) when the caller already holds it."""
    profs = [np.ascontiguousarray(p, np.float32).copy() for p in profiles]
    fwd = [True] * len(profs)
    d, table, orient = msa.rev_seq_based_on_dist(ctx, profs, fwd, sc, table=table, with_state=True)   # src/assemble.h:422
    keep = msa.exclude_unmatched(ctx, profs, sc, match_fraction, dist=d)    # src/assemble.h:428-458
    kept = [i for i, k in enumerate(keep) if k]
    if len(kept) < 2:
        return dict(forward=fwd, kept=kept, rows=None, seqidx=[], gapped=b"", consensus=b"", quality=b"")
    # msa()'s distanceMatrix (src/msa.h:33-42) asks for scores the orientation table already holds
    rows, seqidx, _ = msa.msa(ctx, [profs[i] for i in kept], sc, dist=msa.oriented_distance(table, orient, kept))   # src/assemble.h:468
    gapped, cs, qs = msa.consensus(rows, fraction_called, False)            # src/assemble.h:471
    return dict(forward=fwd, kept=kept, rows=rows, seqidx=seqidx, gapped=gapped, consensus=cs, quality=qs)


def assemble_reference(ctx, profiles, reference, sc=DnaScore(3, -5, -10, -4), match_fraction=0.5, fraction_called=0.5, inc_ref=False):
    """The DP sequence of the reference-guided branch of `tracy assemble` (reference src/assemble.h:163-292) from the trace profiles
    on: both orientations of every trace scored against the one-hot reference profile (ONE batched call instead of 2N), traces below
    the match threshold dropped, the rest ranked by score (then input order) and aligned one after the other against the profile of
    the alignment so far (inherently sequential: each alignment changes the next one's input), consensus over the rows.
    Returns dict(rows, idx, forward, excluded, gapped, consensus, quality): rows = the traces in reverse rank order, the reference last."""
    semiglobal = AlignConfig(True, False)
    profs = [np.ascontiguousarray(p, np.float32) for p in profiles]
    n = len(profs)
    pref = msa.onehot_profile(reference)
    revs = [msa._revcomp(p) for p in profs]
    ranked, excluded = [], []
    if n:
        s, _, _ = ctx.gotoh(PP, profs + revs, [pref] * (2 * n), sc, semiglobal, traceback=False)
        f32 = np.float32
        for i in range(n):
            gs_fwd, gs_rev = int(s[i]), int(s[n + i])
            size = float(profs[i].shape[1])                      # double seqsize; matchFraction is a float, (1 - matchFraction) a float too
            thr = size * float(f32(match_fraction)) * sc.match + size * float(f32(1) - f32(match_fraction)) * sc.mismatch
            if gs_fwd > thr or gs_rev > thr:
                ranked.append((max(gs_fwd, gs_rev), i, gs_fwd >= gs_rev))
            else:
                excluded.append(i)
    ranked.sort(key=lambda t: (-t[0], t[1]))                     # TraceScore::operator<, src/assemble.h:42-44
    if not ranked:
        return dict(rows=np.zeros((0, 0), np.uint8), idx=[], forward=[], excluded=excluded, gapped=b"", consensus=b"", quality=b"")
    rows = None
    for _, i, fwd in ranked:
        p = profs[i] if fwd else revs[i]
        target = pref if rows is None else msa.profile_from_alignment(rows)
        _, ops, ol = ctx.gotoh(PP, [p], [target], sc, semiglobal, traceback=True)
        o = ops[0, : ol[0]]
        cons_p = np.frombuffer(msa.profile_cons_chars(p), np.uint8)
        new0 = np.full(len(o), 0x2D, np.uint8)
        new0[o != ord("h")] = cons_p[: int((o != ord("h")).sum())]
        has = o != ord("v")
        old = np.frombuffer(msa.profile_cons_chars(pref), np.uint8).reshape(1, -1) if rows is None else rows
        below = np.full((old.shape[0], len(o)), 0x2D, np.uint8)
        below[:, has] = old[:, : int(has.sum())]
        rows = np.vstack([new0.reshape(1, -1), below])
    gapped, cs, qs = msa.consensus(rows, fraction_called, not inc_ref)
    return dict(rows=rows, idx=[i for _, i, _ in ranked], forward=[f for _, _, f in ranked], excluded=excluded, gapped=gapped, consensus=cs, quality=qs)
