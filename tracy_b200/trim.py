"""What every tracy subcommand runs between basecall() and createProfile(): the penalty track over the basecalls, the base
qualities estimated from it and the trimming heuristics (reference src/abif.h:164-253, src/trim.h:35-99). Host logic over a
few thousand basecalls per trace; the arithmetic follows the reference's integer types (uint32 distances, int32 penalties,
double means and thresholds) so that the numbers agree bit for bit."""
import math

import numpy as np

_U32 = 0xFFFFFFFF


def _ambiguous(ch):
    return ch not in "ACGT"


def _s(x):
    return bytes(x).decode("latin-1") if not isinstance(x, str) else x


def _i32(v):
    """Wrap to int32 (two's complement), as the reference's int32_t arithmetic does."""
    v &= _U32
    return v - (1 << 32) if v & 0x80000000 else v


def _cvt_i32(x):
    """(int32_t) of a double as x86 converts it: truncation, INT_MIN for NaN and for anything out of range."""
    return int(x) if (x == x and -2147483649.0 < x < 2147483648.0) else -2147483648


def _tail(n, halfwin):
    """for (uint32_t i = size - halfwin; i < size; ++i): with fewer than halfwin basecalls the start wraps and nothing runs."""
    return range(n - halfwin, n) if n >= halfwin else range(0)


def find_best_trace_section(bcpos, secondary, win=10):
    """findBestTraceSection(bc, penalty, win) (reference src/abif.h:164-219). Returns (penalty list, best index, penalty per base of
    the best 10 % window). Penalty of a basecall = ambiguous secondary calls in the window around it + how far the largest and
    smallest peak distance of the window are from the mean distance."""
    sec = _s(secondary)
    pos = [int(x) for x in bcpos]
    n, halfwin = len(sec), win // 2
    penalty = [0] * n
    amb = sum(1 for ch in sec[:win] if _ambiguous(ch))
    for i in range(min(halfwin, n)):
        penalty[i] = amb
    for i in range(win, n):
        amb += int(_ambiguous(sec[i])) - int(_ambiguous(sec[i - win]))
        penalty[i - halfwin] = amb
    for i in _tail(n, halfwin):
        penalty[i] = amb
    total = float(sum(pos[i] - pos[i - 1] for i in range(1, n)))
    mean = total / (n - 1) if n != 1 else float("nan")
    peak_var = 0
    for i in range(max(n - win, 0)):
        old = (pos[i - 1] if i > 0 else 0) & _U32
        lo, hi = pos[n - 1] & _U32, 0
        for k in range(win):
            dist = (pos[i + k] - old) & _U32
            old = pos[i + k] & _U32
            lo, hi = min(lo, dist), max(hi, dist)
        x = (abs(hi - mean) + abs(lo - mean)) / 2
        # (int32_t) of a double: out of range (peak distances wrap where basecall positions do not increase) gives INT_MIN on x86
        peak_var = _cvt_i32(x) & _U32
        penalty[i + halfwin] = _i32(penalty[i + halfwin] + peak_var)             # int32 += uint32: modulo 2^32
        if i == 0:
            for k in range(halfwin):
                penalty[k] = _i32(penalty[k] + peak_var)
    for i in _tail(n, halfwin):
        penalty[i] = _i32(penalty[i] + peak_var)
    sourcewin = int(0.1 * n)
    best_idx, best_val = 0, 99999999
    for i in range(max(n - sourcewin, 0)):
        v = _i32(sum(penalty[i: i + sourcewin]))
        if v < best_val:
            best_val, best_idx = v, i + sourcewin // 2
    per_base = best_val / sourcewin if sourcewin else (float("nan") if best_val == 0 else math.copysign(float("inf"), best_val))
    return penalty, best_idx, per_base


def estimate_qualities(bcpos, secondary):
    """estimateQualities(bc) (reference src/abif.h:232-253): 60 for the basecall with the smallest penalty down to 0 for the largest.
    A trace whose penalties are all zero gets quality 0 everywhere (60/0 = inf, inf * 0 = NaN, and the conversion of NaN to int
    yields INT_MIN on x86, clamped to 0)."""
    penalty, _, _ = find_best_trace_section(bcpos, secondary)
    max_val = max(penalty + [0])
    if max_val == 0:
        return np.zeros(len(penalty), np.uint8)
    scaling = 60.0 / max_val
    return np.array([min(max(_cvt_i32(60.0 - scaling * p), 0), 60) for p in penalty], np.uint8)


def trim_trace(bcpos, secondary, trim_stringency):
    """trimTrace(c, bc, leftTrim, rightTrim) (reference src/trim.h:35-73): from the best 10 % window walk outwards while the penalty
    of the next `win` basecalls stays below trimStringency * (penalty per base of the best window) * win. Returns (left, right) =
    basecalls to drop at either end."""
    win = 10
    penalty, best_idx, per_base = find_best_trace_section(bcpos, secondary, win)
    n = len(penalty)
    thr = float(np.float32(trim_stringency)) * per_base * win
    right, left = n, 0
    local = float(sum(penalty[best_idx: min(best_idx + win, n)]))
    for i in range(best_idx, max(n - win, best_idx)):
        local += penalty[i + win] - penalty[i]
        if local > thr:
            right = i
            break
    local = float(sum(penalty[best_idx: min(best_idx + win, n)]))
    i = best_idx - 1
    while i >= 0:
        if i + win < n:
            local -= penalty[i + win]
        local += penalty[i]
        if local > thr:
            left = i + win - 1
            break
        i -= 1
    return left, (n - right if right < n else 0)


def trim_basecalls(nsamples, bcpos, qual, primary, secondary, consensus, trim_left, trim_right):
    """trimTrace(tr, bc, trimLeft, trimRight, nbc) (reference src/trim.h:76-99): the basecalls without the trimmed ends; the trace
    samples stay. Follows the writers' walk: only basecall positions that come up in increasing order inside the trace are kept."""
    pri, sec, con = _s(primary), _s(secondary), _s(consensus)
    last = (len(pri) - trim_right) & _U32
    keep = []
    k, idx = 0, int(bcpos[0])
    for t in range(nsamples):
        if idx == t:
            if trim_left <= k < last:
                keep.append(k)
            if k < len(bcpos) - 1:
                k += 1
                idx = int(bcpos[k])
    return dict(bcpos=np.array([int(bcpos[k]) for k in keep], np.int32), qual=np.array([int(qual[k]) for k in keep], np.uint8),
                primary="".join(pri[k] for k in keep), secondary="".join(sec[k] for k in keep), consensus="".join(con[k] for k in keep))


_COMPLEMENT = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N", "H": "D", "V": "B", "M": "K", "Y": "R", "D": "H", "B": "V", "K": "M", "R": "Y", "U": "A",
               "S": "S", "W": "W"}


def reverse_complement_trace(acgt, bcpos, qual, primary, secondary, consensus):
    """reverseComplementTrace (reference src/trim.h:124-151): the samples reversed with channels A<->T, C<->G swapped; the
    basecalls walked from the last one down (positions that do not come up in decreasing order are dropped, like the writers'
    forward walk), complemented over the IUPAC alphabet (:102-123; other characters stay). Returns dict(acgt, bcpos, qual, primary,
    secondary, consensus)."""
    acgt = np.asarray(acgt)
    pri, sec, con = _s(primary), _s(secondary), _s(consensus)
    ns = acgt.shape[1]
    out = np.ascontiguousarray(acgt[::-1, ::-1]).astype(np.int32)
    k = len(bcpos) - 1
    idx = int(bcpos[k])
    nb, nq, np_, ns_, nc = [], [], [], [], []
    for new_pos, t in enumerate(range(ns, 0, -1)):
        if idx == t - 1:
            nb.append(new_pos); nq.append(int(qual[k]))
            np_.append(_COMPLEMENT.get(pri[k], pri[k])); ns_.append(_COMPLEMENT.get(sec[k], sec[k])); nc.append(_COMPLEMENT.get(con[k], con[k]))
            if k > 0:
                k -= 1
                idx = int(bcpos[k])
    return dict(acgt=out, bcpos=np.array(nb, np.int32), qual=np.array(nq, np.uint8), primary="".join(np_), secondary="".join(ns_), consensus="".join(nc))


def nearest_snp(primary, secondary, trim_left, trim_right, rtp):
    """nearestSNP(c, bc, rtp) (reference src/trim.h:11-33): walking outwards from basecall rtp (the reliable trace section), the first
    position inside the trimmed range whose primary and secondary calls differ, as an index into the TRIMMED trace; the right side is
    looked at first at every distance. Without any such position: rtp - trimLeft (or trimLeft when rtp lies in the left trim)."""
    pri, sec = _s(primary), _s(secondary)
    n = min(len(pri), len(sec))
    offset = 0
    while True:
        dead_end = True
        if rtp + offset + trim_right < n:
            if trim_left < rtp + offset and pri[rtp + offset] != sec[rtp + offset]:
                return rtp + offset - trim_left
            dead_end = False
        if offset + trim_left < rtp:
            if pri[rtp - offset] != sec[rtp - offset]:
                return rtp - offset - trim_left
            dead_end = False
        if dead_end:
            break
        offset += 1
    return rtp - trim_left if rtp > trim_left else trim_left


def trace_quality(bcpos, secondary, trim_stringency=None, best=False):
    """estimate_qualities (+ trim_trace when a stringency is given) through the native tb_trace_quality (csrc/trimq.cu): the same
    numbers, without the interpreter in the per-basecall loops. Returns (qual uint8[n], (left, right) or None); with best=True a
    third value, findBestTraceSection(bc) (src/abif.h:221-229)."""
    import ctypes as C
    from . import capi
    pos = np.ascontiguousarray(bcpos, np.int32)
    sec = secondary.encode("latin-1") if isinstance(secondary, str) else bytes(secondary)
    n = min(len(pos), len(sec))
    qual = np.zeros(n, np.uint8)
    left, right = C.c_uint32(0), C.c_uint32(0)
    want = trim_stringency is not None
    bs = C.c_uint32(0)
    rc = capi.lib().tb_trace_quality(pos.ctypes.data, sec, n, float(trim_stringency or 0.0), qual.ctypes.data, C.byref(bs), C.byref(left) if want else None,
                                     C.byref(right) if want else None)
    if rc != capi.TB_OK:
        raise ValueError("tb_trace_quality: %d" % rc)
    lr = (int(left.value), int(right.value)) if want else None
    return (qual, lr, int(bs.value)) if best else (qual, lr)
