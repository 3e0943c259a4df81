"""Text outputs of `tracy align` / `tracy decompose`: host-side formatting of what
the batch drivers return, byte-identical to the reference's writers.

  align_fasta          P.align.fa   reference src/sage.h:326-339
  plot_alignment       P.txt, P.align1/.align2/.align3   reference src/fmindex.h:329-427 (plotAlignment)
  write_decomposition  P.decomp     reference src/decompose.h:621-627
The JSON / BCF writers (src/json.h, src/variants.h) are not covered.
"""


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
