"""Variant calls from the allele alignments of `tracy decompose -v` (reference src/variants.h:34-138, called from
src/indigo.h:405-422): host logic over the gapped rows the DP kernels return. A variant is (pos, basenum, gt, chr, ref, alt);
calling the same variant again (the second allele) raises gt instead of adding a record."""


def _has_n(s):
    return "n" in s or "N" in s


def insert_variant(var, pos, basenum, gt, chr_name, ref, alt):
    """insertVariant (src/variants.h:34-53): an existing (pos, chr, ref, alt) gets gt + 1 (homozygous); new records need pos > 0
    and a reference allele without N."""
    for v in var:
        if v["pos"] == pos and v["chr"] == chr_name and v["ref"] == ref and v["alt"] == alt:
            v["gt"] += 1
            return
    if pos > 0 and not _has_n(ref):
        var.append(dict(pos=pos, basenum=basenum, gt=gt, chr=chr_name, ref=ref, alt=alt, id="."))


def variant_type(ref, alt):
    """variantType (src/variants.h:129-138)."""
    if len(ref) == 1 and len(alt) == 1:
        return "SNV"
    return "Deletion" if len(ref) > len(alt) else "Insertion" if len(ref) < len(alt) else "Complex"


def call_variants(row0, row1, chr_name, pos, var):
    """callVariants(align, rs, var) (src/variants.h:56-126). row0: allele, row1: reference slice, both gapped; pos: rs.pos
    (0-based start of the slice). Columns before the allele's first and after its last base are end gaps and call nothing; a
    mismatch column is an SNV at the 1-based reference position; gap runs become one deletion / insertion anchored on the
    preceding reference base ('N' when the run starts the alignment); an insertion still open at the allele's last base is
    dropped, as there."""
    a0 = bytes(row0).decode("latin-1") if not isinstance(row0, str) else row0
    a1 = bytes(row1).decode("latin-1") if not isinstance(row1, str) else row1
    chr_name = bytes(chr_name).decode("latin-1") if not isinstance(chr_name, str) else chr_name
    ri, start, end = pos, -1, -1
    for j in range(len(a0)):
        if a0[j] != "-":
            if start == -1:
                start = j
            end = j
        if a1[j] != "-" and start == -1:
            ri += 1
    if start < 0:
        return var
    vi, dele, del_start, ins, ins_start, last_ref = 0, "", 0, "", 0, "N"
    for j in range(start, end + 1):
        if dele and a0[j] != "-":
            insert_variant(var, del_start, vi, 1, chr_name, dele, dele[0])
            dele = ""
        if ins and a1[j] != "-":
            insert_variant(var, ins_start, vi, 1, chr_name, ins[0], ins)
            ins = ""
        if a0[j] != "-":
            vi += 1
        if a1[j] != "-":
            ri += 1
        if a0[j] != a1[j]:
            if a0[j] != "-" and a1[j] != "-":
                insert_variant(var, ri, vi, 1, chr_name, a1[j], a0[j])
            elif a0[j] == "-":
                if not dele:
                    dele = last_ref
                    del_start = ri - 1
                dele += a1[j]
            else:
                if not ins:
                    ins = last_ref
                    ins_start = ri
                ins += a0[j]
        if a1[j] != "-":
            last_ref = a1[j]
    return var
