"""Text outputs of `tracy align` / `tracy decompose`: host-side formatting of what
the batch drivers return, byte-identical to the reference's writers.

  align_fasta          P.align.fa   reference src/sage.h:326-339
  plot_alignment       P.txt, P.align1/.align2/.align3   reference src/fmindex.h:329-427 (plotAlignment)
  write_decomposition  P.decomp     reference src/decompose.h:621-627
  trace_txt            P.abif       reference src/abif.h:512-534 (traceTxtOut: the tab-separated trace table of align / decompose)
  trace_fasta, trace_fastq   basecall subcommand   reference src/fasta.h:98-158
  trace_json           basecall JSON   reference src/json.h:32-117 (traceJsonOut)
  alignment_trace_padding, trace_align_json   P.json of `tracy align`   reference src/json.h:383-479, 120-217, src/sage.h:319-343
  decompose_json       P.json of `tracy decompose`   reference src/json.h:16-30, 249-381 (traceAlleleAlignJsonOut)
  aligned_trace_by_row, assemble_files   P.align.fa / P.json / P.vertical / P.cons.fa|fq of `tracy assemble`   reference src/json.h:220-246, src/assemble.h:473-579
The BCF writer (src/variants.h:141-266, htslib) is not covered.
"""
import os

import numpy as np

EMPTY_TRACE_SIGNAL = -99          # reference src/json.h:12-14

_IUPAC_EXPANDED = {"A": "A", "C": "C", "G": "G", "T": "T", "N": "N", "R": "A|G", "Y": "C|T", "S": "C|G", "W": "A|T", "K": "G|T", "M": "A|C"}


def _g(x):
    """operator<<(double) with the default stream state: %g, six significant digits."""
    return "%g" % x


def align_fasta(trace_name, row0, row1, chr_name, forward):
    """P.align.fa (reference src/sage.h:326-339): the two gapped rows as FASTA records."""
    chr_name = chr_name.decode() if isinstance(chr_name, bytes) else chr_name
    return (">" + trace_name + "\n" + bytes(row0).decode("latin-1") + "\n>" + chr_name + (" (forward)" if forward else " (reverse)") + "\n"
            + bytes(row1).decode("latin-1") + "\n")


def plot_alignment(row0, row1, chr_name, pos, refslice_len, forward, score, key=0, a1a2=(0, 0), linelimit=60):
    """plotAlignment(filename, align, rs, key, score, a1a2, linelimit), reference src/fmindex.h:329-420, as a string.
    key 0: trace vs reference (P.txt); 1 / 2: allele 1 / 2 vs reference (P.align1 / P.align2); 3: allele 1 vs allele 2 (P.align3)."""
    a0, a1 = bytes(row0).decode("latin-1"), bytes(row1).decode("latin-1")
    chr_name = chr_name.decode() if isinstance(chr_name, bytes) else chr_name
    ri, riend, vi = pos + 1, pos + refslice_len, 1
    fald = linelimit + 14
    out = []
    if key == 0:
        out.append(">Alt\n")
    elif key == 2:
        out.append(">Alt2 (Estimated allelic Fraction: " + _g(a1a2[1]) + ")\n")
    else:
        out.append(">Alt1 (Estimated allelic Fraction: " + _g(a1a2[0]) + ")\n")

    def ungapped(row):
        s = row.replace("-", "")
        lines = [s[i:i + fald] for i in range(0, len(s), fald)]
        return "".join(x + "\n" for x in lines)
    out.append(ungapped(a0))
    if key != 3:
        if forward:
            out.append(">Ref %s:%d-%d forward\n" % (chr_name, ri, riend))
        else:
            out.append(">Ref %s:%d-%d reversecomplement\n" % (chr_name, pos + refslice_len - (riend - pos) + 1, pos + refslice_len - (ri - pos) + 1))
    else:
        out.append(">Alt2 (Estimated allelic Fraction: " + _g(a1a2[1]) + ")\n")
    out.append(ungapped(a1))
    out.append("\nAlignment score: %d\n" % score)
    rule = "#" + "-" * (fald - 1) + "\n"
    out.append(rule + "\n")
    blocks, s, e = 0, 0, len(a0)
    while s < e:
        seg0, seg1 = a0[s:s + linelimit], a1[s:s + linelimit]
        out.append(("Alt%10d " % vi if key != 3 else "Alt1%9d " % vi) + seg0 + "\n")
        vi += len(seg0) - seg0.count("-")
        out.append(" " * 14 + "".join("|" if x == y else " " for x, y in zip(seg0, seg1)) + "\n")
        if key != 3:
            out.append("Ref%10d " % (ri if forward else pos + refslice_len - (ri - pos) + 1) + seg1 + "\n")
        else:
            out.append("Alt2%9d " % ri + seg1 + "\n")
        ri += len(seg1) - seg1.count("-")
        out.append("\n")
        s += linelimit
        blocks += 1
    if blocks < 6:
        out.append("\n" * (4 * (6 - blocks)))
    out.append(rule + rule + "\n\n")
    return "".join(out)


def write_decomposition(decomp):
    """writeDecomposition(path, dcp), reference src/decompose.h:621-627: `indel<TAB>decomp` rows."""
    return "indel\tdecomp\n" + "".join("%d\t%d\n" % (int(a), int(b)) for a, b in decomp)


# ---- per-trace outputs ---------------------------------------------------------------------------------------------------
def _called(nsamples, bcpos):
    """The walk all of the reference's per-sample writers share: sample i carries basecall k when it equals the NEXT expected
    basecall position; the expectation only moves forward (a position that is not larger than its predecessor is never met)."""
    if len(bcpos) == 0:
        raise ValueError("the reference reads bcPos[0] unconditionally: at least one basecall is required")
    k, idx = 0, int(bcpos[0])
    for i in range(nsamples):
        if idx == i:
            yield i, k
            if k < len(bcpos) - 1:
                k += 1
                idx = int(bcpos[k])


def _s(x):
    return bytes(x).decode("latin-1") if not isinstance(x, str) else x


def _peaks(acgt):
    return "".join('"peak%s": [%s],\n' % ("ACGT"[k], ", ".join(str(int(v)) for v in acgt[k])) for k in range(4))


def trace_txt(acgt, bcpos, qual, primary, secondary, consensus, trim_left, trim_right):
    """traceTxtOut (reference src/abif.h:512-534): one row per trace sample; samples that carry a basecall show its number,
    the three calls, the quality and whether the call lies in the trimmed ends."""
    acgt = np.asarray(acgt)
    pri, sec, con = _s(primary), _s(secondary), _s(consensus)
    rtr = len(pri) - trim_right if trim_right < len(pri) else 0
    hit = dict(_called(acgt.shape[1], bcpos))
    out = ["pos\tpeakA\tpeakC\tpeakG\tpeakT\tbasenum\tprimary\tsecondary\tconsensus\tqual\ttrim\n"]
    for i in range(acgt.shape[1]):
        row = "%d\t%d\t%d\t%d\t%d\t" % (i + 1, acgt[0][i], acgt[1][i], acgt[2][i], acgt[3][i])
        if i in hit:
            k = hit[i]
            row += "%d\t%s\t%s\t%s\t%d\t%s\n" % (k + 1, pri[k], sec[k], con[k], int(qual[k]), "Y" if (k < trim_left or k >= rtr) else "N")
        else:
            row += "NA\tNA\tNA\tNA\tNA\tNA\n"
        out.append(row)
    return "".join(out)


def trace_json(acgt, bcpos, qual, primary, secondary):
    """traceJsonOut (reference src/json.h:32-117)."""
    acgt = np.asarray(acgt)
    pri, sec = _s(primary), _s(secondary)
    calls = list(_called(acgt.shape[1], bcpos))
    out = ["{\n", '"pos": [%s],\n' % ", ".join(str(i + 1) for i in range(acgt.shape[1])), _peaks(acgt)]
    out.append('"basecallPos": [%s],\n' % ", ".join(str(i + 1) for i, _ in calls))
    out.append('"basecallQual": [%s],\n' % ", ".join(str(int(qual[k])) for _, k in calls))
    items = []
    for i, k in calls:
        v = "%d:%s" % (k + 1, pri[k])
        if pri[k] != sec[k]:
            v += "|" + _IUPAC_EXPANDED.get(sec[k], "N")
        items.append('"%d":"%s"' % (i + 1, v))
    out.append('"basecalls": {%s},\n' % ", ".join(items))
    out.append('"primarySeq": "%s",\n"secondarySeq": "%s"\n\n}\n' % (pri, sec))
    return "".join(out)


def alignment_trace_padding(row, acgt, bcpos, qual, primary, secondary, consensus):
    """alignmentTracePadding (reference src/json.h:383-479): every gap run of the trace's alignment row becomes `step` empty
    samples per gap column, inserted half-way between the two basecalls around it, with a '-' basecall in their middle
    (step = the truncated mean basecall distance, 6 for a single basecall); runs before the first / after the last base only
    count as leading / trailing gaps. Returns dict(acgt, bcpos, qual, primary, secondary, consensus, leading, trailing)."""
    acgt = np.asarray(acgt)
    bcpos = [int(x) for x in bcpos]
    pri, sec, con = _s(primary), _s(secondary), _s(consensus)
    step = 6
    if len(bcpos) > 1:
        step = int(float(sum(bcpos[i] - bcpos[i - 1] for i in range(1, len(bcpos)))) / (len(bcpos) - 1))
    ins_pos, ins_size, pos, gapsize, ingap, leading = [], [], 0, 0, False, 0
    for ch in _s(row):
        if ch == "-":
            gapsize = gapsize + 1 if ingap else 1
            ingap = True
        else:
            if ingap:
                ingap = False
                if pos:
                    ins_pos.append(int((bcpos[pos - 1] + bcpos[pos]) / 2.0))
                    ins_size.append(gapsize)
                else:
                    leading = gapsize
            pos += 1
    trailing = gapsize if ingap else 0
    cols, nbp, nq, npri, nsec, ncon = [], [], [], [], [], []
    k, idx = 0, bcpos[0]
    offset, q, ins_idx = 0, 0, (ins_pos[0] if ins_pos else -1)
    empty = (EMPTY_TRACE_SIGNAL,) * 4
    for t in range(acgt.shape[1]):
        cols.append((int(acgt[0][t]), int(acgt[1][t]), int(acgt[2][t]), int(acgt[3][t])))
        if ins_idx == t:
            for _ in range(ins_size[q]):
                nbp.append(t + offset + int(step / 2.0)); nq.append(0); npri.append("-"); nsec.append("-"); ncon.append("-")
                cols.extend([empty] * step)
                offset += step
            if q < len(ins_pos) - 1:
                q += 1
                ins_idx = ins_pos[q]
        if idx == t:
            nbp.append(idx + offset); nq.append(int(qual[k])); npri.append(pri[k]); nsec.append(sec[k]); ncon.append(con[k])
            if k < len(bcpos) - 1:
                k += 1
                idx = bcpos[k]
    return dict(acgt=np.array(cols, np.int32).T.reshape(4, -1), bcpos=np.array(nbp, np.int32), qual=np.array(nq, np.uint8), primary="".join(npri),
                secondary="".join(nsec), consensus="".join(ncon), leading=leading, trailing=trailing)


def assembly_trace(padded, trace_file_name="trace"):
    """assemblyTrace (reference src/json.h:120-195): the JSON object of one gapped trace."""
    acgt, pri, sec = padded["acgt"], padded["primary"], padded["secondary"]
    calls = list(_called(acgt.shape[1], padded["bcpos"]))
    out = ['{\n"traceFileName": "%s",\n"leadingGaps": %d,\n"trailingGaps": %d,\n' % (trace_file_name, padded["leading"], padded["trailing"]), _peaks(acgt)]
    out.append('"basecallPos": [%s],\n' % ", ".join(str(i + 1) for i, _ in calls))
    out.append('"basecallQual": [%s],\n' % ", ".join(str(int(padded["qual"][k])) for _, k in calls))
    items, gapless = [], 0
    for i, k in calls:
        if pri[k] != "-":
            gapless += 1
            v = "%d:%s" % (gapless, pri[k]) + ("|" + sec[k] if pri[k] != sec[k] else "")
        else:
            v = "-"
        items.append('"%d":"%s"' % (i + 1, v))
    out.append('"basecalls": {%s}\n}\n' % ", ".join(items))
    return "".join(out)


def trace_align_json(acgt, bcpos, qual, primary, secondary, consensus, row0, row1, chr_name, pos, forward):
    """P.json of `tracy align` (reference src/sage.h:319-343): the trace padded along alignment row 0, then traceAlignJsonOut
    (src/json.h:197-217)."""
    padded = alignment_trace_padding(row0, acgt, bcpos, qual, primary, secondary, consensus)
    return ('{\n"gappedTrace":\n' + assembly_trace(padded) + ',\n"refchr": "%s",\n"refpos": %d,\n"altalign": "%s",\n"refalign": "%s",\n"forward": %d\n}\n'
            % (_s(chr_name), pos + 1, _s(row0), _s(row1), 1 if forward else 0))


# ---- P.json of `tracy decompose` ------------------------------------------------------------------------------------------
TRACY_VERSION = "0.9.1"           # reference src/version.h:8 (the writer prints it into the meta block)


def x_window_viewport(bcpos, k):
    """xWindowViewport (reference src/json.h:249-258): the sample range shown around basecall k, 150 samples either side,
    clipped to the first sample and the last basecall."""
    lb = int(bcpos[k]) + 1
    lb = 1 if lb <= 150 else lb - 150
    ub = int(bcpos[k]) + 1
    ub = ub + 150 if ub + 150 < int(bcpos[-1]) else int(bcpos[-1])
    return lb, ub


def sort_variants(var):
    """std::sort over Variant::operator< (reference src/variants.h:20-22): by chromosome, position, base number."""
    return sorted(var, key=lambda v: (v["chr"], v["pos"], v["basenum"]))


def decompose_json(cfg, acgt, bcpos, qual, primary, secondary, var, allele1, allele2, align3, decomp, indelshift, breakpoint, a1a2):
    """traceAlleleAlignJsonOut (reference src/json.h:260-381). cfg: dict(trim_left, trim_right, qual_cut, pratio, input, genome);
    var: variant records (tracy_b200.variants) in output order; allele1 / allele2: (row0, row1, chr, pos, forward, score);
    align3: (allele1 row, allele2 row, score); decomp: (indel, count) pairs; a1a2: the allelic fractions."""
    from .variants import variant_type
    pri = _s(primary)
    tl, tr = int(cfg["trim_left"]), int(cfg["trim_right"])
    out = ['{\n"meta": {"program": "tracy", "version": "%s", "arguments": {"trimLeft": %d, "trimRight": %d, "pratio": %s, "genome": "%s", "input": "%s"}},\n'
           % (TRACY_VERSION, tl, tr, _g(float(np.float32(cfg["pratio"]))), str(cfg["genome"]).rsplit("/", 1)[-1], str(cfg["input"]).rsplit("/", 1)[-1])]
    out.append(trace_json(acgt, bcpos, qual, primary, secondary)[2:-3])       # the body between "{\n" and "\n}\n"
    out.append(",\n")
    out.append('"chartConfig": { "x": { "axis": { "range": [%d, %d] }}},\n' % x_window_viewport(bcpos, tl + int(breakpoint)))
    for n, (r0, r1, chr_name, pos, fwd, score) in ((1, allele1), (2, allele2)):
        out.append('"ref%dchr": "%s",\n"ref%dpos": %d,\n"alt%dalign": "%s",\n"ref%dalign": "%s",\n"ref%dforward": %d,\n"align%dscore": %d,\n'
                   % (n, _s(chr_name), n, pos + 1, n, _s(r0), n, _s(r1), n, 1 if fwd else 0, n, score))
    out.append('"allele1fraction": %s,\n"allele1align": "%s",\n"allele2fraction": %s,\n"allele2align": "%s",\n"align3score": %d,\n'
               % (_g(a1a2[0]), _s(align3[0]), _g(a1a2[1]), _s(align3[1]), align3[2]))
    out.append('"hetindel": %d,\n' % (1 if indelshift else 0))
    out.append('"decomposition": {\n"x": [%s],\n"y": [%s]\n},\n' % (", ".join(str(int(a)) for a, _ in decomp), ", ".join(str(int(b)) for _, b in decomp)))
    out.append('"variants": {\n"columns": ["chr", "pos", "id", "ref", "alt", "qual", "filter", "type", "genotype", "basepos", "signalpos"],\n"rows": [\n')
    fwd1 = bool(allele1[4])
    rows, ranges = [], []
    for v in var:
        k = tl + v["basenum"] - 1 if fwd1 else len(pri) - (tr + v["basenum"])          # the basecall the variant sits on
        q = int(qual[k])
        gt = {0: "hom. REF", 1: "het.", 2: "hom. ALT"}.get(v["gt"], "missing")
        rows.append('["%s", %d, "%s", "%s", "%s", %d, "%s", "%s", "%s", %d, %d]'
                    % (v["chr"], v["pos"], v["id"], v["ref"], v["alt"], q, "LowQual" if q < int(cfg["qual_cut"]) else "PASS", variant_type(v["ref"], v["alt"]), gt,
                       k + 1, int(bcpos[k]) + 1))
        ranges.append("[%d, %d]" % x_window_viewport(bcpos, k))
    out.append(",\n".join(rows) + '],\n"xranges": [\n' + ",\n".join(ranges) + "]\n}\n}\n")
    return "".join(out)


# ---- the files of `tracy assemble` (de novo) ------------------------------------------------------------------------------
def aligned_trace_by_row(rows, row, trace_file_name, forward, ref):
    """alignedTraceByRow (reference src/json.h:220-246): one row of the multiple alignment without its end gaps, as a JSON object."""
    r = bytes(np.asarray(rows, np.uint8)[row]).decode("latin-1")
    lead = len(r) - len(r.lstrip("-"))
    trail = 0
    for ch in r:                                                       # the reference's count: a row of gaps only has lead = trail = len
        trail = 0 if ch != "-" else trail + 1
    return ('{\n"reference": %s,\n"forward": %s,\n"traceFileName": "%s",\n"leadingGaps": "%d",\n"trailingGaps": "%d",\n"align": "%s"\n}\n'
            % ("true" if ref else "false", "true" if forward else "false", trace_file_name, lead, trail, r[lead: max(len(r) - trail, lead)] if lead < len(r) - trail else ""))


def assemble_files(names, forward, rows, gapped, consensus, quality, padded_traces, include_consensus=False, fmt="fasta", reference_last=False):
    """The output section of assemble() (reference src/assemble.h:284-376 reference-guided, :473-579 de novo). names / forward: per TRACE
    in output order (de novo: row i of the alignment, i.e. c.ab[idxMap[seqidx[i]]].stem() and fwd[seqidx[i]]; reference-guided: rank
    order, trace i sitting in row ntraces-1-i with the reference in the last row -- pass reference_last=True); rows: uint8[nrow][ncol];
    gapped / consensus / quality: what msa.consensus returns; padded_traces: per trace the dict alignment_trace_padding returns for its
    row (reverse-complemented first when the trace is not forward). Returns {suffix: text} for .align.fa, .json, .vertical and
    .cons.fa | .cons.fq."""
    a = np.asarray(rows, np.uint8)
    gapped, cs, qs = _s(gapped), _s(consensus), _s(quality)
    nt = len(names)
    row_of = [nt - 1 - i for i in range(nt)] if reference_last else list(range(nt))
    fa = "".join(">%s (%s)\n%s\n" % (names[i], "forward" if forward[i] else "reverse", bytes(a[row_of[i]]).decode("latin-1")) for i in range(nt))
    if reference_last:
        fa += ">Reference\n" + bytes(a[nt]).decode("latin-1") + "\n"
    if include_consensus:
        fa += ">Consensus\n" + gapped + "\n"
    js = ['{\n"gapFreeConsensus": "%s",\n"gappedConsensus": "%s",\n"msa": \n[\n' % (cs, gapped)]
    js.append(",\n".join(aligned_trace_by_row(a, row_of[i], names[i], forward[i], False) for i in range(nt)))
    if reference_last:
        js.append(",\n" + aligned_trace_by_row(a, nt, "", True, True))
    js.append('],\n"gappedTraces": \n[\n')
    js.append(", ".join(assembly_trace(padded_traces[i], names[i]) for i in range(nt)))
    js.append("]\n}\n")
    vertical = "".join(bytes(a[:, j]).decode("latin-1") + "|" + gapped[j] + "\n" for j in range(a.shape[1]))
    out = {".align.fa": fa, ".json": "".join(js), ".vertical": vertical}
    if fmt == "fasta":
        out[".cons.fa"] = ">Consensus\n" + cs + "\n"
    elif fmt == "fastq":
        out[".cons.fq"] = "@Consensus\n" + cs + "\n+\n" + qs + "\n"
    return out


# ---- all files of one trace, as SURVEY appendix B lists them ---------------------------------------------------------------
def align_files(trace_name, acgt, bcpos, qual, primary, secondary, consensus, trim_left, trim_right, row0, row1, chr_name, pos, refslice_len, forward, score,
                linelimit=60):
    """`tracy align -o P`: {suffix: text} for P.abif (reference src/sage.h:188), P.align.fa (:326-339), P.txt (:342), P.json (:319-345).
    The trace arguments are the UNTRIMMED trace and basecalls (the trims only mark rows of P.abif); row0 / row1: the final alignment of
    the full trace against the trimmed reference slice; pos / refslice_len / forward: that slice."""
    return {".abif": trace_txt(acgt, bcpos, qual, primary, secondary, consensus, trim_left, trim_right),
            ".align.fa": align_fasta(trace_name, row0, row1, chr_name, forward),
            ".txt": plot_alignment(row0, row1, chr_name, pos, refslice_len, forward, score, 0, (0, 0), linelimit),
            ".json": trace_align_json(acgt, bcpos, qual, primary, secondary, consensus, row0, row1, chr_name, pos, forward)}


def decompose_files(cfg, acgt, bcpos, qual, primary, secondary, consensus, decomp, var, allele1, allele2, align3, refslice_lens, indelshift, breakpoint, a1a2, linelimit=60):
    """`tracy decompose -o P`: {suffix: text} for P.abif (reference src/indigo.h:187), P.decomp (:341), P.align1 / .align2 / .align3
    (:366, 377, 388) and P.json (:450); P.bcf needs htslib and is not written. allele1 / allele2: (row0, row1, chr, pos, forward, score),
    align3: (allele 1 row, allele 2 row, score), refslice_lens: the lengths of the two reference slices."""
    out = {".abif": trace_txt(acgt, bcpos, qual, primary, secondary, consensus, cfg["trim_left"], cfg["trim_right"]), ".decomp": write_decomposition(decomp)}
    for key, (al, rl) in enumerate(((allele1, refslice_lens[0]), (allele2, refslice_lens[1])), start=1):
        out[".align%d" % key] = plot_alignment(al[0], al[1], al[2], al[3], rl, al[4], al[5], key, a1a2, linelimit)
    # allele 1 against allele 2 (global): the "reference" is the second allele itself, src/indigo.h:380-388
    out[".align3"] = plot_alignment(align3[0], align3[1], b"Alt2", 0, len(bytes(align3[1]).replace(b"-", b"")), True, align3[2], 3, a1a2, linelimit)
    out[".json"] = decompose_json(cfg, acgt, bcpos, qual, primary, secondary, var, allele1, allele2, align3, decomp, indelshift, breakpoint, a1a2)
    return out


# ---- the fasta / fastq formats of the basecall subcommand -------------------------------------------------------------------
def trace_fasta(otype, trim_left, trim_right, primary, secondary, consensus):
    """traceFastaOut (reference src/fasta.h:98-119): the primary, secondary or consensus calls between the trims; any other `otype`
    writes an empty file."""
    seq = {"primary": primary, "secondary": secondary, "consensus": consensus}.get(otype)
    if seq is None:
        return ""
    seq = _s(seq)
    return ">%s\n%s\n" % (otype, seq[trim_left: max(len(seq) - trim_right, trim_left)])


def trace_fastq(otype, trim_left, trim_right, nsamples, bcpos, qual, primary, secondary, consensus):
    """traceFastqOut (reference src/fasta.h:121-158): the record of trace_fasta with '@', then '+' and the estimated qualities
    (Phred + 33) of the basecalls that come up in the writers' walk and lie between the trims (measured on the PRIMARY calls, whichever
    sequence is written)."""
    seq = {"primary": primary, "secondary": secondary, "consensus": consensus}.get(otype)
    head = ""
    if seq is not None:
        seq = _s(seq)
        head = "@%s\n%s\n" % (otype, seq[trim_left: max(len(seq) - trim_right, trim_left)])
    last = len(_s(primary)) - trim_right
    quals = "".join(chr((int(qual[k]) + 33) & 0xFF) for _, k in _called(nsamples, bcpos) if trim_left <= k < last)
    return head + "+\n" + quals + "\n"


# ---- native forms (csrc/writers.cu): the same bytes, formatted and written without the interpreter -------------------------------
def _trace_view(acgt, bcpos, qual, primary, secondary, consensus):
    import ctypes as C
    from . import capi
    acgt = np.ascontiguousarray(acgt, np.int32)
    bcpos = np.ascontiguousarray(bcpos, np.int32)
    qual = np.ascontiguousarray(qual, np.uint8)
    pri, sec, con = (x.encode("latin-1") if isinstance(x, str) else bytes(x) for x in (primary, secondary, consensus))
    keep = (acgt, bcpos, qual, pri, sec, con)
    v = capi.TraceView(acgt.ctypes.data, acgt.shape[1], bcpos.ctypes.data, qual.ctypes.data, C.cast(C.c_char_p(pri), C.c_void_p), C.cast(C.c_char_p(sec), C.c_void_p),
                       C.cast(C.c_char_p(con), C.c_void_p), len(bcpos))
    return v, keep


def write_align_files(prefix, trace_name, acgt, bcpos, qual, primary, secondary, consensus, trim_left, trim_right, row0, row1, chr_name, pos, refslice_len,
                      forward, score, linelimit=60):
    """The four files of `tracy align -o prefix` for one trace (what align_files returns as text), written by tb_write_align_files.
    The call releases the GIL: a pool of writer threads runs them in parallel."""
    import ctypes as C
    from . import capi
    v, keep = _trace_view(acgt, bcpos, qual, primary, secondary, consensus)
    row0, row1 = bytes(row0), bytes(row1)
    chr_b = chr_name.encode("latin-1") if isinstance(chr_name, str) else bytes(chr_name)
    rc = capi.lib().tb_write_align_files(os.fsencode(prefix), trace_name.encode("latin-1"), C.byref(v), int(trim_left), int(trim_right), row0, row1, len(row0),
                                         chr_b, int(pos), int(refslice_len), int(bool(forward)), int(score), int(linelimit))
    if rc != capi.TB_OK:
        raise OSError("tb_write_align_files(%s): %d" % (prefix, rc))


def write_trace_txt(path, acgt, bcpos, qual, primary, secondary, consensus, trim_left, trim_right):
    """P.abif written by tb_write_trace_txt (the bytes of trace_txt)."""
    import ctypes as C
    from . import capi
    v, keep = _trace_view(acgt, bcpos, qual, primary, secondary, consensus)
    rc = capi.lib().tb_write_trace_txt(os.fsencode(path), C.byref(v), int(trim_left), int(trim_right))
    if rc != capi.TB_OK:
        raise OSError("tb_write_trace_txt(%s): %d" % (path, rc))


def write_decompose_files(prefix, cfg, acgt, bcpos, qual, primary, secondary, consensus, decomp, allele1, allele2, align3, refslice_lens, indelshift, breakpoint, a1a2,
                          linelimit=60):
    """P.decomp, P.align1 / .align2 / .align3 and P.json of `tracy decompose -o prefix` (the files decompose_files returns as text, without
    P.abif and with an empty variant table), written by the native writers. Arguments as decompose_files."""
    import ctypes as C
    from . import capi
    L = capi.lib()
    b = lambda x: x.encode("latin-1") if isinstance(x, str) else bytes(x)
    v, keep = _trace_view(acgt, bcpos, qual, primary, secondary, consensus)
    dc = np.ascontiguousarray(np.asarray(decomp, np.int32).reshape(-1, 2))
    pre = os.fsencode(prefix)
    ok = L.tb_write_decomposition(pre + b".decomp", dc.ctypes.data, dc.shape[0]) == 0
    for key, (al, rl) in enumerate(((allele1, refslice_lens[0]), (allele2, refslice_lens[1])), start=1):
        r0, r1 = b(al[0]), b(al[1])
        ok = ok and L.tb_write_plot_alignment(pre + (".align%d" % key).encode(), r0, r1, len(r0), b(al[2]), int(al[3]), int(rl), int(bool(al[4])), int(al[5]), key,
                                              float(a1a2[0]), float(a1a2[1]), int(linelimit)) == 0
    x0, x1 = b(align3[0]), b(align3[1])
    ok = ok and L.tb_write_plot_alignment(pre + b".align3", x0, x1, len(x0), b"Alt2", 0, len(x1.replace(b"-", b"")), 1, int(align3[2]), 3, float(a1a2[0]),
                                          float(a1a2[1]), int(linelimit)) == 0
    a1r0, a1r1, a2r0, a2r1 = b(allele1[0]), b(allele1[1]), b(allele2[0]), b(allele2[1])
    d = capi.DecomposeJson(int(cfg["trim_left"]), int(cfg["trim_right"]), float(cfg["pratio"]), str(cfg["genome"]).rsplit("/", 1)[-1].encode("latin-1"),
                           str(cfg["input"]).rsplit("/", 1)[-1].encode("latin-1"), int(cfg["trim_left"]) + int(breakpoint),
                           b(allele1[2]), int(allele1[3]), a1r0, a1r1, len(a1r0), int(bool(allele1[4])), int(allele1[5]),
                           b(allele2[2]), int(allele2[3]), a2r0, a2r1, len(a2r0), int(bool(allele2[4])), int(allele2[5]),
                           float(a1a2[0]), float(a1a2[1]), x0, x1, len(x0), int(align3[2]), int(bool(indelshift)), dc.ctypes.data, dc.shape[0])
    ok = ok and L.tb_write_decompose_json(pre + b".json", C.byref(v), C.byref(d)) == 0
    if not ok:
        raise OSError("native decompose writers failed for %s" % prefix)


def write_assemble_files(prefix, names, forward, rows, row_of, gapped, consensus, quality, traces, include_consensus=False, fmt="fasta", reference_last=False):
    """The files of `tracy assemble -o prefix` (what assemble_files returns as text) written by tb_write_assemble_files: names / forward /
    row_of per trace in output order, traces: per trace a dict with the UNTRIMMED acgt / bcpos / qual / primary / secondary / consensus and
    the trims tl / tr -- the hard trim, the reverse complement of flipped traces and the padding along the alignment row happen natively,
    one host thread per trace."""
    import ctypes as C
    from . import capi
    a = np.ascontiguousarray(rows, np.uint8)
    nt = len(names)
    arr = (capi.AssembleTrace * max(nt, 1))()
    keep = []
    for i in range(nt):
        t = traces[i]
        v, k = _trace_view(t["acgt"], t["bcpos"], t["qual"], t["primary"], t["secondary"], t["consensus"])
        nm = names[i].encode("latin-1")
        keep.append((k, nm))
        arr[i] = capi.AssembleTrace(nm, int(bool(forward[i])), int(row_of[i]), v, int(t["tl"]), int(t["tr"]))
    b = lambda x: x.encode("latin-1") if isinstance(x, str) else bytes(x)
    rc = capi.lib().tb_write_assemble_files(os.fsencode(prefix), a.ctypes.data, a.shape[0], a.shape[1], arr, nt, b(gapped), b(consensus), b(quality),
                                            int(bool(include_consensus)), 1 if fmt == "fastq" else 0 if fmt == "fasta" else -1, int(bool(reference_last)))
    if rc != capi.TB_OK:
        raise OSError("tb_write_assemble_files(%s): %d" % (prefix, rc))
