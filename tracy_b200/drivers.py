"""Batch drivers: the DP sequences of tracy's subcommand drivers, run for MANY traces at once.

  align_batch      sage()     for single-FASTA references   reference src/sage.h:222-260, :311
  assemble_denovo  assemble() de novo branch                reference src/assemble.h:418-471

Everything between the file readers and the file writers of those drivers: orientation pick, semi-global alignment of the
trimmed trace, reference-slice trimming, final alignment (align); orientation optimisation, exclusion of unmatched traces,
MSA and consensus (assemble). Every DP call is one batched GPU call over all traces; the glue in between is the literal
host logic of tracy_b200.api / tracy_b200.msa. File formats, basecalling and the JSON/plot writers stay with the caller.
"""
import numpy as np

from . import msa
from .api import PS, AlignConfig, DnaScore, rows_from_ops, trim_reference_slice

_SEMIGLOBAL = AlignConfig(True, False)          # AlignConfig<true, false>, src/sage.h:165
_COMP = {ord("A"): ord("T"), ord("C"): ord("G"), ord("G"): ord("C"), ord("T"): ord("A"), ord("N"): ord("N")}


def reverse_complement_seq(seq):
    """reverseComplement(std::string&), reference src/fmindex.h:11-26: upper-cased reverse, A<->T, C<->G, N kept; any other
    character leaves the ORIGINAL character of that position in place (the reference's `default: break`)."""
    seq = bytes(seq)
    rev = seq[::-1].upper()
    out = bytearray(seq)
    for i, ch in enumerate(rev):
        if ch in _COMP:
            out[i] = _COMP[ch]
    return bytes(out)


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
    # gsFwd / gsRev (src/sage.h:239-240): one score-only batch of 2N pairs
    s = ctx.gotoh(PS, list(trimmed_profiles) + list(trimmed_profiles), refs + rc, sc, _SEMIGLOBAL, traceback=False)[0]
    forward = [bool(s[i] > s[n + i]) for i in range(n)]                    # strict '>', src/sage.h:247
    pref = [refs[i] if forward[i] else rc[i] for i in range(n)]
    # gotoh(trimmedtrace, prefslice) + trimReferenceSlice (src/sage.h:258-259)
    _, ops, ol = ctx.gotoh(PS, trimmed_profiles, pref, sc, _SEMIGLOBAL)
    slices, pos = [], []
    for i in range(n):
        r0, r1 = _gap_rows(ops[i, : ol[i]])
        sl, p = trim_reference_slice(r0, r1, pref[i], forward[i], 0, trim_left, trim_right)
        slices.append(sl); pos.append(p)
    # gotoh(fulltraceprofile, referenceprofile) (src/sage.h:311): the reported alignment
    score, ops2, ol2 = ctx.gotoh(PS, full_profiles, slices, sc, _SEMIGLOBAL)
    out = []
    for i in range(n):
        row0, row1 = rows_from_ops(PS, full_profiles[i], slices[i], bytes(ops2[i, : ol2[i]]))
        out.append(dict(forward=forward[i], refslice=slices[i], pos=pos[i], score=int(score[i]), row0=row0, row1=row1))
    return out


def assemble_denovo(ctx, profiles, sc=DnaScore(3, -5, -10, -4), match_fraction=0.5, fraction_called=0.1):
    """The de novo branch of `tracy assemble` from the trace profiles on (reference src/assemble.h:418-471):
    revSeqBasedOnDist -> exclusion of traces that match nothing -> msa -> consensus.
    Returns dict(forward, kept (indices into `profiles`), rows, seqidx (into kept), gapped, consensus, quality)."""
    profs = [np.ascontiguousarray(p, np.float32).copy() for p in profiles]
    fwd = [True] * len(profs)
    msa.rev_seq_based_on_dist(ctx, profs, fwd, sc)                          # src/assemble.h:422
    keep = msa.exclude_unmatched(ctx, profs, sc, match_fraction)            # src/assemble.h:428-458
    kept = [i for i, k in enumerate(keep) if k]
    if len(kept) < 2:
        return dict(forward=fwd, kept=kept, rows=None, seqidx=[], gapped=b"", consensus=b"", quality=b"")
    rows, seqidx, _ = msa.msa(ctx, [profs[i] for i in kept], sc)            # src/assemble.h:468
    gapped, cs, qs = msa.consensus(rows, fraction_called, False)            # src/assemble.h:471
    return dict(forward=fwd, kept=kept, rows=rows, seqidx=seqidx, gapped=gapped, consensus=cs, quality=qs)
