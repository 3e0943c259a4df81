"""Host glue around the decompose sweeps: the parts of decomposeAlleles (reference src/decompose.h:179-376) that
are not the O(maxindel x L) sweeps -- walking the alignment to the breakpoint, median/MAD thresholding, candidate
selection, the .decomp table and applying the chosen shift. The sweeps themselves (src/decompose.h:210-224,
:247-261, :288-313) are delegated to `sweep`, which in the product is Context.decompose_sweep (CUDA).

Kept literal on purpose (uint32 wrap-arounds, nth_element medians, strict/non-strict comparisons): the outputs
must equal the reference's byte for byte.
"""
import numpy as np

_U32 = 0xFFFFFFFF
_PAIRS = {"R": "AG", "Y": "CT", "S": "CG", "W": "AT", "K": "GT", "M": "AC"}
_CODE = {(0, 2): "R", (1, 3): "Y", (1, 2): "S", (0, 3): "W", (2, 3): "K", (0, 1): "M"}


def iupac(one, two):
    """iupac(char, char), reference src/abif.h:141-161: non-ACGT arguments count as 'A'; equal indices give 'N'."""
    a = "ACGT".find(one) if one in "CGT" else 0
    b = "ACGT".find(two) if two in "CGT" else 0
    if b < a:
        a, b = b, a
    return _CODE.get((a, b), "N")


def phase_ref_allele(pri, sec, r):
    """phaseRefAllele, reference src/decompose.h:147-175 (characters as 1-length str)."""
    if r == "-" or sec == "N":
        return "N"
    if sec == r:
        return pri
    pair = _PAIRS.get(sec)
    if pair is None:
        return "N"
    if r == pair[0]:
        return iupac(pri, pair[1])
    if r == pair[1]:
        return iupac(pri, pair[0])
    return "N"


def _median(v):
    """getMedian, reference src/decompose.h:131-136: element n/2 of the sorted range."""
    s = sorted(v)
    return s[len(s) // 2]


def _apply(refrow, pri, sec, j0, vi0, vi_end):
    """The rewrite loops at reference src/decompose.h:319-327, :349-357, :362-370."""
    j, vi = j0, vi0
    L = len(refrow)
    while j < L and vi < vi_end:
        r = chr(refrow[j])
        if r != chr(pri[vi]):
            s = phase_ref_allele(chr(pri[vi]), chr(sec[vi]), r)
            if s != "N":
                pri[vi] = ord(r)
                sec[vi] = ord(s)
        j += 1
        vi += 1


def walk_to_breakpoint(row0, row1, pri, sec, trim_left, breakpoint):
    """reference src/decompose.h:186-208: phase primary/secondary in place up to the breakpoint.
    Returns (alignIndex, varIndex, refPointer)."""
    var_index = ref_pointer = align_index = 0
    vi = trim_left
    bp = (breakpoint + trim_left) & _U32
    for j in range(len(row0)):
        if row0[j] != 0x2D:
            r = chr(row1[j])
            if r != chr(pri[vi]):
                s = phase_ref_allele(chr(pri[vi]), chr(sec[vi]), r)
                if s != "N":
                    pri[vi] = ord(r)
                    sec[vi] = ord(s)
            vi += 1
            if vi == bp:
                align_index, var_index = j, vi
                break
        if row1[j] != 0x2D:
            ref_pointer += 1
    return align_index, var_index, ref_pointer


def sweep_extents(ncons, refslice_len, ref_pointer, trim_right, breakpoint_abs, maxindel):
    """Loop bounds of the sweeps: (ndel, nins, vi_end, maxdel, maxins), reference src/decompose.h:211-213, :249-250."""
    maxdel = 2
    if refslice_len > ref_pointer + trim_right + 2:
        maxdel = refslice_len - (ref_pointer + trim_right)
    ndel = min(maxindel, maxdel // 2)
    maxins = (ncons - (trim_right + breakpoint_abs)) & _U32
    nins = max(1, min(maxindel, maxins // 2))
    vi_end = ncons - trim_right
    return ndel, nins, vi_end, maxdel, maxins


def select_candidates(f, thres):
    """Local-minimum rule, reference src/decompose.h:238-244 / :265-271."""
    out = []
    n = len(f)
    for i in range(n):
        if f[i] < thres:
            if i + 1 < n and 2 * f[i] < f[i + 1]:
                out.append(i)
            elif i > 0 and 2 * f[i] < f[i - 1]:
                out.append(i)
            elif i == 0 and i + 2 < n and 2 * f[i] < f[i + 2]:
                out.append(i)
    return out


def decompose_alleles(row0, row1, primary, secondary, trim_left, trim_right, maxindel, madc, breakpoint, refslice_len, sweep,
                      ncons=None):
    """decomposeAlleles for one trace. row0/row1: the gotoh() alignment (trace row, reference row) as bytes;
    primary/secondary: basecall strings (bytes). `sweep(refrow, pri, sec, vi_end, align_index, var_index, ndel, nins, grid)`
    returns (fref[ndel], fins[nins], grid[nins][ndel] | None).
    Returns (primary', secondary', dcp int32[k][2], info dict)."""
    gen = _decompose_gen(row0, row1, primary, secondary, trim_left, trim_right, maxindel, madc, breakpoint, refslice_len, ncons)
    try:
        req = next(gen)
        while True:
            req = gen.send(sweep(*req))
    except StopIteration as done:
        return done.value


def decompose_alleles_batch(ctx, items):
    """decomposeAlleles for MANY traces with the sweeps of all of them in one GPU call per phase (the indel sweep for every
    trace; then the ins x del grid for the traces that found no candidate, reference src/decompose.h:288-313).
    items: list of dicts with the keyword arguments of decompose_alleles (without `sweep`). Returns one result tuple each."""
    gens = [_decompose_gen(it["row0"], it["row1"], it["primary"], it["secondary"], it["trim_left"], it["trim_right"], it["maxindel"],
                           it["madc"], it["breakpoint"], it["refslice_len"], it.get("ncons")) for it in items]
    results = [None] * len(items)
    pending = {}
    for i, g in enumerate(gens):
        try:
            pending[i] = next(g)
        except StopIteration as done:
            results[i] = done.value
    while pending:
        idx = sorted(pending)
        for want_grid in (False, True):
            sel = [i for i in idx if i in pending and pending[i][8] == want_grid]
            if not sel:
                continue
            req = [pending[i] for i in sel]
            fref, fins, grid = ctx.decompose_sweep([r[0] for r in req], [r[1] for r in req], [r[2] for r in req], [r[3] for r in req],
                                                   [r[4] for r in req], [r[5] for r in req], [r[6] for r in req], [r[7] for r in req], grid=want_grid)
            for k, i in enumerate(sel):
                nd, ni = req[k][6], req[k][7]
                ans = (fref[k, :nd], fins[k, :ni], grid[k, :ni, :nd] if want_grid else None)
                try:
                    pending[i] = gens[i].send(ans)
                except StopIteration as done:
                    results[i] = done.value
                    del pending[i]
    return results


def _decompose_gen(row0, row1, primary, secondary, trim_left, trim_right, maxindel, madc, breakpoint, refslice_len, ncons=None):
    """decomposeAlleles as a generator: yields the arguments of each sweep it needs and is sent the result."""
    pri, sec = bytearray(primary), bytearray(secondary)
    ncons = len(primary) if ncons is None else ncons
    align_index, var_index, ref_pointer = walk_to_breakpoint(row0, row1, pri, sec, trim_left, breakpoint)
    bp_abs = (breakpoint + trim_left) & _U32
    ndel, nins, vi_end, maxdel, maxins = sweep_extents(ncons, refslice_len, ref_pointer, trim_right, bp_abs, maxindel)

    fref, fins, _ = yield (bytes(row1), bytes(pri), bytes(sec), vi_end, align_index, var_index, ndel, nins, False)
    fref = [int(x) for x in fref[:ndel]]
    fins = [int(x) for x in fins[:nins]]
    fins[0] = fref[0]                                         # src/decompose.h:248

    med = _median(fref)                                       # src/decompose.h:226-234
    mad = _median([abs(x - med) for x in fref])
    thres = med - madc * mad if med > madc * mad else 0
    if thres < 10:
        thres = 10
    deldecomp = select_candidates(fref, thres)
    insdecomp = select_candidates(fins, thres)

    nothing = not deldecomp and not insdecomp                 # src/decompose.h:273-285
    defins = 50 if nothing else 15
    for i in insdecomp:
        defins = max(defins, i + 15)
    defins = min(defins, len(fins))
    defdel = 50 if nothing else 15
    for i in deldecomp:
        defdel = max(defdel, i + 15)
    defdel = min(defdel, len(fref))
    dcp = [(-i, fref[i]) for i in range(defdel - 1, -1, -1)] + [(i, fins[i]) for i in range(1, defins)]

    info = dict(align_index=align_index, var_index=var_index, ndel=ndel, nins=nins, thres=thres, deldecomp=deldecomp,
                insdecomp=insdecomp, mode=None, best=None)
    refrow = bytes(row1)
    if nothing:
        # complex mutation: ins x del grid, reference src/decompose.h:288-313
        gi = min(maxindel, maxins // 2)
        gd = ndel
        best_ins = best_del = 0
        best_fr = 1000
        if gi > 0 and gd > 0:
            _, _, grid = yield (refrow, bytes(pri), bytes(sec), vi_end, align_index, var_index, gd, gi, True)
            for ins in range(gi):
                prev = 0
                for d in range(gd):
                    fr = int(grid[ins][d])
                    if 2 * fr < prev and fr < best_fr:
                        best_ins, best_del, best_fr = ins, d, fr
                    prev = fr
        if best_fr != 1000:
            info.update(mode="complex", best=(best_ins, best_del, best_fr))
            _apply(refrow, pri, sec, align_index + best_del + 1, var_index + best_ins, vi_end)
        else:
            info.update(mode="none")
            vi = trim_left                                    # src/decompose.h:331-343: traverse the whole alignment
            for j in range(len(row0)):
                if row0[j] != 0x2D:
                    r = chr(row1[j])
                    if r != chr(pri[vi]):
                        s = phase_ref_allele(chr(pri[vi]), chr(sec[vi]), r)
                        if s != "N":
                            pri[vi] = ord(r)
                            sec[vi] = ord(s)
                    vi += 1
    elif deldecomp:
        info.update(mode="del", best=(0, min(deldecomp)))
        _apply(refrow, pri, sec, align_index + min(deldecomp) + 1, var_index, vi_end)
    else:
        info.update(mode="ins", best=(min(insdecomp), 0))
        _apply(refrow, pri, sec, align_index + 1, var_index + min(insdecomp), vi_end)
    return bytes(pri), bytes(sec), np.array(dcp, np.int32).reshape(-1, 2), info
