"""Host-side Python layer over the C ABI (include/tracy_b200.h).

Batch API (`Context.gotoh`, `Context.gotoh_device`, `Context.decompose_sweep`) plus single-pair mirrors of the
reference's call shapes with the reference's names (`gotohScore`, `gotoh`, `DnaScore`, `AlignConfig`;
reference src/gotoh.h:12-14, :71-73, src/align.h:11-50).  Everything here only marshals buffers; the DP runs in
the CUDA library and nowhere else.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi

PS, PP, SS = "ps", "pp", "ss"
_KIND_ID = {PP: 0, SS: 1, PS: 2}


class TracyError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tracy_b200 error {code}: {msg}")
        self.code = code


@dataclass
class DnaScore:
    """reference src/align.h:11-32 (the CLI always uses the 4-argument constructor, src/sage.h:166)."""
    match: int = 5
    mismatch: int = -4
    go: int = -10
    ge: int = -1

    def c(self):
        return capi.Score(self.match, self.mismatch, self.go, self.ge)


@dataclass
class AlignConfig:
    """AlignConfig<THorizontal, TVertical>, reference src/align.h:37-80."""
    horizontal: bool = False
    vertical: bool = False

    def c(self):
        return capi.AlignConfig(int(self.horizontal), int(self.vertical))


@dataclass
class Arena:
    """One side of a batch in host memory: flat `base` plus per-item element offsets and lengths."""
    base: np.ndarray      # float32 (profiles) or uint8 (sequences), 1-D contiguous
    off: np.ndarray       # int64[N]
    len: np.ndarray       # int32[N]
    trace_profiles: bool = False   # a1 only: every profile has exact zeros in rows 4 (N) and 5 ('-'), as createProfile makes them (TB_A1_TRACE_PROFILES)

    @property
    def n(self):
        return len(self.off)


def pack_profiles(items):
    """list of float[6][len] arrays -> Arena (reference layout per item, row-major)."""
    items = [np.ascontiguousarray(p, dtype=np.float32) for p in items]
    for p in items:
        if p.ndim != 2 or p.shape[0] != 6:
            raise ValueError("a profile is float[6][len] (rows A,C,G,T,N,-)")
    lens = np.array([p.shape[1] for p in items], np.int32)
    sizes = 6 * lens.astype(np.int64)
    off = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64) if len(items) else np.zeros(0, np.int64)
    base = np.concatenate([p.reshape(-1) for p in items]) if len(items) else np.zeros(0, np.float32)
    if base.size == 0:
        base = np.zeros(1, np.float32)
    return Arena(base, off, lens)


def pack_seqs(items):
    items = [bytes(s) for s in items]
    lens = np.array([len(s) for s in items], np.int32)
    off = np.concatenate([[0], np.cumsum(lens.astype(np.int64))[:-1]]).astype(np.int64) if len(items) else np.zeros(0, np.int64)
    base = np.frombuffer(b"".join(items), np.uint8).copy() if sum(map(len, items)) else np.zeros(1, np.uint8)
    return Arena(base, off, lens)


def uniform_profiles(arr, trace_profiles=False):
    """float32[N][6][m] -> Arena without copying. trace_profiles=True: the caller vouches that rows 4 and 5 of every profile are
    exact zeros (createProfile outputs, reference src/profile.h:37); host batches then upload 4 rows of 6."""
    arr = np.ascontiguousarray(arr, np.float32)
    n, six, m = arr.shape
    assert six == 6
    return Arena(arr.reshape(-1), np.arange(n, dtype=np.int64) * (6 * m), np.full(n, m, np.int32), trace_profiles)


def uniform_seqs(arr):
    """uint8[N][n] -> Arena without copying."""
    arr = np.ascontiguousarray(arr, np.uint8)
    n, k = arr.shape
    return Arena(arr.reshape(-1), np.arange(n, dtype=np.int64) * k, np.full(n, k, np.int32))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Context:
    """Owns one tb_ctx (one GPU). Not thread-safe, like the reference's functions it is used from one thread."""

    def __init__(self, device=0):
        self._lib = capi.lib()
        h = C.c_void_p()
        rc = self._lib.tb_ctx_create(C.byref(h), device)
        if rc != capi.TB_OK:
            raise TracyError(rc, self._lib.tb_strerror(rc).decode() + " (tb_ctx_create: is a B200 visible? there is no CPU fallback)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tb_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != capi.TB_OK:
            raise TracyError(rc, self._lib.tb_strerror(rc).decode() + ": " + self._lib.tb_last_error(self._h).decode())

    def pinned_empty(self, shape, dtype=np.uint8):
        """A numpy array in page-locked host memory (tb_host_alloc): batches and result arrays that live there cross PCIe at the
        link's 55 GB/s and overlap with the kernels; pageable memory goes through the driver's staging at ~10 GB/s. The memory is
        returned (tb_host_free) when the array and every view of it are gone; keep the Context alive until then."""
        import weakref
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        p = C.c_void_p()
        self._check(self._lib.tb_host_alloc(self._h, C.byref(p), max(n, 1)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
        lib, owner, addr = self._lib, weakref.ref(self), p.value

        def release():
            c = owner()
            if c is not None and getattr(c, "_h", None):          # a context closed first keeps the block until the process ends
                lib.tb_host_free(c._h, C.c_void_p(addr))
        weakref.finalize(buf, release)
        return arr

    def set_scratch_limit(self, nbytes):
        self._check(self._lib.tb_ctx_set_scratch_limit(self._h, int(nbytes)))

    def stats(self):
        k, a, b = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.tb_ctx_stats(self._h, C.byref(k), C.byref(a), C.byref(b)))
        return {"kernel_launches": k.value, "h2d_bytes": a.value, "d2h_bytes": b.value}

    def last_kernel_ms(self):
        p, g, s = C.c_float(), C.c_float(), C.c_float()
        self._check(self._lib.tb_ctx_last_kernel_ms(self._h, C.byref(p), C.byref(g), C.byref(s)))
        return {"packed_ms": p.value, "general_ms": g.value, "sweep_ms": s.value}

    def last_call_ms(self):
        v = C.c_float()
        self._check(self._lib.tb_ctx_last_call_ms(self._h, C.byref(v)))
        return v.value

    def last_packed_pairs(self):
        v = C.c_uint64()
        self._check(self._lib.tb_ctx_last_packed_pairs(self._h, C.byref(v)))
        return v.value

    def last_big_pairs(self):
        """Pairs of the last "pp" call that were spread over many warps (band-pipelined big pairs)."""
        v = C.c_uint64()
        self._check(self._lib.tb_ctx_last_big_pairs(self._h, C.byref(v)))
        return v.value

    # ---- reference anchoring (reference src/fmindex.h:173-326) -------------------------------------------------
    def build_index(self, text):
        """Index of the reference text for anchor(): `text` is what tracy indexes (upper-cased; a multi-sequence genome joined and
        ended by b"\n", src/index.h:104-121). Returns a KmerIndex (close() it, or let the Context outlive it)."""
        text = np.frombuffer(bytes(text), np.uint8) if not isinstance(text, np.ndarray) else np.ascontiguousarray(text, np.uint8)
        h = C.c_void_p()
        self._check(self._lib.tb_index_build(self._h, _ptr(text), text.size, capi.TB_MEM_HOST, C.byref(h)))
        return KmerIndex(self, h)

    def build_index_device(self, text_ptr, n):
        """build_index for a text that already sits in HBM (e.g. after a broadcast over NCCL): raw device pointer + length."""
        h = C.c_void_p()
        self._check(self._lib.tb_index_build(self._h, C.c_void_p(int(text_ptr)), int(n), capi.TB_MEM_DEVICE, C.byref(h)))
        return KmerIndex(self, h)

    def anchor(self, index, consensus, trim_left=50, trim_right=50, kmer=15, min_kmer_support=3):
        """scanSequence + findMaxFreq + the orientation rule of getReferenceSlice for a batch of consensus strings.
        Returns dict of arrays: anchored (bool), forward (bool), kmersupport (uint32), bestpos (int64), pass_ (uint8)."""
        cons = consensus if isinstance(consensus, Arena) else pack_seqs(consensus)
        n = cons.n
        out = dict(anchored=np.zeros(n, np.uint8), forward=np.ones(n, np.uint8), kmersupport=np.zeros(n, np.uint32),
                   bestpos=np.zeros(n, np.int64), pass_=np.zeros(n, np.uint8))
        a = capi.Arena(_ptr(cons.base), _ptr(cons.off), _ptr(cons.len))
        r = capi.AnchorResult(_ptr(out["anchored"]), _ptr(out["forward"]), _ptr(out["kmersupport"]), _ptr(out["bestpos"]), _ptr(out["pass_"]))
        self._check(self._lib.tb_anchor(self._h, index._h, C.byref(a), n, capi.TB_MEM_HOST,
                                        capi.AnchorConfig(trim_left, trim_right, kmer, min_kmer_support), C.byref(r)))
        out["anchored"] = out["anchored"].astype(bool)
        out["forward"] = out["forward"].astype(bool)
        return out

    def last_anchor_ms(self):
        v = C.c_float()
        self._check(self._lib.tb_ctx_last_anchor_ms(self._h, C.byref(v)))
        return v.value

    # ---- gotoh / gotohScore, batched -------------------------------------------------------------------
    def _fn(self, kind):
        return {PS: self._lib.tb_gotoh_ps, PP: self._lib.tb_gotoh_pp, SS: self._lib.tb_gotoh_ss}[kind]

    def gotoh(self, kind, a1, a2, sc=DnaScore(3, -5, -10, -4), ac=AlignConfig(True, False), traceback=True, out=None, rows=None, packed=False):
        """Batch of pairs in HOST memory. a1/a2: Arena (or lists of profiles / byte strings).
        Returns (scores int32[N], ops uint8[N][stride] or None, ops_len int32[N] or None).
        `out` may carry preallocated (scores, ops, ops_len) arrays (e.g. pinned).
        rows: True, or a preallocated (row0, row1) pair of uint8[N][stride] arrays -- the gapped alignment rows gotoh() leaves in
        `align` (reference src/align.h:196-293), made on the device; they are then returned as a 4th and 5th value.
        packed: ops come back at 2 bits per op (4 per byte; unpack_ops) in an array of stride ceil(max(len1+len2)/4)."""
        if not isinstance(a1, Arena):
            a1 = pack_seqs(a1) if kind == SS else pack_profiles(a1)
        if not isinstance(a2, Arena):
            a2 = pack_profiles(a2) if kind == PP else pack_seqs(a2)
        n = a1.n
        if a2.n != n:
            raise ValueError("a1 and a2 must hold the same number of items")
        maxsum = int((a1.len.astype(np.int64) + a2.len).max()) if n else 1
        if out is not None:
            scores, ops, ops_len = out
        else:
            scores = np.zeros(n, np.int32)
            ops = ops_len = None
            if traceback:
                stride = max((((maxsum + 3) // 4 if packed else maxsum) + 15) // 16 * 16, 16)
                ops = np.zeros((n, stride), np.uint8)
                ops_len = np.zeros(n, np.int32)
        r0 = r1 = None
        if rows is not None and rows is not False:
            if rows is True:
                rs = max((maxsum + 15) // 16 * 16, 16)
                r0, r1 = np.zeros((n, rs), np.uint8), np.zeros((n, rs), np.uint8)
            else:
                r0, r1 = rows
            if ops_len is None:
                ops_len = np.zeros(n, np.int32)
        b = capi.Batch(capi.Arena(_ptr(a1.base), _ptr(a1.off), _ptr(a1.len)),
                       capi.Arena(_ptr(a2.base), _ptr(a2.off), _ptr(a2.len)), n,
                       capi.TB_MEM_HOST | (capi.TB_A1_TRACE_PROFILES if (a1.trace_profiles and kind != SS) else 0))
        want_ops = traceback and ops is not None
        r = capi.Result(_ptr(scores), _ptr(ops) if want_ops else None, ops.shape[1] if want_ops else 0,
                        _ptr(ops_len) if ops_len is not None else None,
                        _ptr(r0) if r0 is not None else None, _ptr(r1) if r1 is not None else None, r0.shape[1] if r0 is not None else 0,
                        1 if (packed and want_ops) else 0)
        self._run_gotoh(kind, b, sc, ac, r)
        if r0 is not None:
            return scores, ops, ops_len, r0, r1
        return scores, ops, ops_len

    def _run_gotoh(self, kind, b, sc, ac, r):
        self._check(self._fn(kind)(self._h, C.byref(b), sc.c(), ac.c(), C.byref(r)))

    def gotoh_device(self, kind, a1_base, a1_off, a1_len, a2_base, a2_off, a2_len, n, scores, ops=None, ops_stride=0, ops_len=None,
                     sc=DnaScore(3, -5, -10, -4), ac=AlignConfig(True, False), row0=None, row1=None, rows_stride=0, packed=False):
        """Everything already resident in HBM: arguments are raw device pointers (ints), e.g. torch tensor .data_ptr()."""
        b = capi.Batch(capi.Arena(a1_base, a1_off, a1_len), capi.Arena(a2_base, a2_off, a2_len), n, capi.TB_MEM_DEVICE)
        r = capi.Result(scores, ops, ops_stride, ops_len, row0, row1, rows_stride, 1 if packed else 0)
        self._check(self._fn(kind)(self._h, C.byref(b), sc.c(), ac.c(), C.byref(r)))

    # ---- decompose sweeps ---------------------------------------------------------------------------------
    def decompose_sweep(self, refrows, primaries, secondaries, vi_end, align_index, var_index, ndel, nins, grid=False):
        """Lists of byte strings (one per trace) plus per-trace ints. Returns (fref[N][S], fins[N][S], grid[N][S][S] | None)
        with S = max(ndel, nins, 1); entries beyond a trace's ndel/nins are 0."""
        ref = pack_seqs(refrows)
        pri = pack_seqs(primaries)
        sec = pack_seqs(secondaries)
        if not np.array_equal(pri.len, sec.len):
            raise ValueError("primary and secondary must have equal lengths")
        n = ref.n
        i32 = lambda x: np.ascontiguousarray(x, np.int32)
        vi_end, align_index, var_index, ndel, nins = map(i32, (vi_end, align_index, var_index, ndel, nins))
        S = int(max(1, ndel.max() if n else 1, nins.max() if n else 1))
        fref = np.zeros((n, S), np.int32)
        fins = np.zeros((n, S), np.int32)
        g = np.zeros((n, S, S), np.int32) if grid else None
        b = capi.SweepBatch(capi.Arena(_ptr(ref.base), _ptr(ref.off), _ptr(ref.len)), capi.Arena(_ptr(pri.base), _ptr(pri.off), _ptr(pri.len)),
                            _ptr(sec.base), _ptr(vi_end), _ptr(align_index), _ptr(var_index), _ptr(ndel), _ptr(nins), n, capi.TB_MEM_HOST)
        r = capi.SweepResult(_ptr(fref), _ptr(fins), S, _ptr(g) if grid else None)
        self._check(self._lib.tb_decompose_sweep(self._h, C.byref(b), C.byref(r)))
        for t in range(n):   # the kernel only writes the evaluated shifts; fins[0] = fref[0] by definition (src/decompose.h:248)
            fref[t, ndel[t]:] = 0
            fins[t, nins[t]:] = 0
        return fref, fins, g

    # ---- profile construction (reference src/profile.h) -------------------------------------------------------
    def create_profile(self, traces, bcpos, primary, secondary, trim_left=None, trim_right=None):
        """createProfile(Trace, BaseCalls, p, trimleft, trimright) for a batch (reference src/profile.h:21-52).
        traces: list of int32[4][nsamples] (Trace::traceACGT); bcpos: list of int32[nbc]; primary/secondary: byte strings.
        Returns a list of float32[6][sz] profiles."""
        n = len(traces)
        tr = [np.ascontiguousarray(t, np.int32) for t in traces]
        for t in tr:
            if t.ndim != 2 or t.shape[0] != 4:
                raise ValueError("a trace is int32[4][nsamples] (channels A,C,G,T)")
        bp = [np.ascontiguousarray(b, np.int32) for b in bcpos]
        tlen = np.array([t.shape[1] for t in tr], np.int32)
        toff = np.concatenate([[0], np.cumsum(4 * tlen.astype(np.int64))[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
        tbase = np.concatenate([t.reshape(-1) for t in tr]) if n else np.zeros(1, np.int32)
        pri, sec = pack_seqs(primary), pack_seqs(secondary)
        blen = np.array([len(b) for b in bp], np.int32)
        if not (np.array_equal(pri.len, blen) and np.array_equal(sec.len, blen)):
            raise ValueError("bcpos, primary and secondary must have one entry per basecall")
        bbase = np.concatenate(bp) if n and blen.sum() else np.zeros(1, np.int32)
        ooff = np.concatenate([[0], np.cumsum(6 * blen.astype(np.int64))[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
        out = np.zeros(max(int(6 * blen.astype(np.int64).sum()), 1), np.float32)
        olen = np.zeros(max(n, 1), np.int32)
        tl = None if trim_left is None else np.ascontiguousarray(np.broadcast_to(trim_left, (n,)), np.int32)
        trr = None if trim_right is None else np.ascontiguousarray(np.broadcast_to(trim_right, (n,)), np.int32)
        b = capi.ProfileBatch(capi.Arena(_ptr(tbase), _ptr(toff), _ptr(tlen)), capi.Arena(_ptr(bbase), _ptr(pri.off), _ptr(blen)),
                              _ptr(pri.base), _ptr(sec.base), _ptr(tl), _ptr(trr), n, capi.TB_MEM_HOST)
        self._check(self._lib.tb_create_profile(self._h, C.byref(b), _ptr(out), _ptr(ooff), _ptr(olen)))
        return [out[ooff[i]: ooff[i] + 6 * olen[i]].reshape(6, olen[i]).copy() for i in range(n)]

    def read_traces(self, files):
        """traceFormat + readab / readscf for a batch of trace files given as bytes (reference src/scf.h:19-102, src/abif.h:286-405);
        the samples are decoded on the GPU. Returns one dict per file: format, ok, status, traceACGT (int32[4][ns]), basecallpos,
        qual, basecalls1, basecalls2 (None for files the library does not unpack)."""
        n = len(files)
        flen = np.array([len(f) for f in files], np.int64)
        foff = np.concatenate([[0], np.cumsum(flen)[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
        blob = np.frombuffer(b"".join(bytes(f) for f in files) or b"\0", np.uint8)
        info = (capi.TraceInfo * max(n, 1))()
        rc = self._lib.tb_trace_scan(_ptr(blob), _ptr(foff), _ptr(flen), n, C.cast(info, C.c_void_p))
        if rc != capi.TB_OK:
            raise TracyError(rc, self._lib.tb_strerror(rc).decode())
        live = [info[i].format >= 0 and info[i].status == 0 for i in range(n)]
        ns = np.array([info[i].nsamples if live[i] else 0 for i in range(n)], np.int64)
        nb = np.array([info[i].nbasecalls if live[i] else 0 for i in range(n)], np.int64)
        soff = np.concatenate([[0], np.cumsum(4 * ns)[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
        boff = np.concatenate([[0], np.cumsum(nb)[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
        samples = np.zeros(max(int(4 * ns.sum()), 1), np.int32)
        tot = max(int(nb.sum()), 1)
        ploc, qual, b1, b2 = np.zeros(tot, np.int32), np.zeros(tot, np.uint8), np.zeros(tot, np.uint8), np.zeros(tot, np.uint8)
        self._check(self._lib.tb_trace_unpack(self._h, _ptr(blob), _ptr(foff), _ptr(flen), n, capi.TB_MEM_HOST, _ptr(samples), _ptr(soff),
                                              _ptr(ploc), _ptr(qual), _ptr(b1), _ptr(b2), _ptr(boff)))
        out = []
        for i in range(n):
            d = dict(format=info[i].format, ok=bool(info[i].ok), status=info[i].status, traceACGT=None, basecallpos=None, qual=None, basecalls1=None, basecalls2=None)
            if info[i].format >= 0 and info[i].status == 0:
                s, b = int(soff[i]), slice(int(boff[i]), int(boff[i] + nb[i]))
                d.update(traceACGT=samples[s: s + 4 * int(ns[i])].reshape(4, int(ns[i])).copy(), basecallpos=ploc[b].copy(), qual=qual[b].copy(),
                         basecalls1=b1[b].tobytes() if info[i].format == 0 else b"",      # readscf leaves both strings empty
                         basecalls2=b2[b].tobytes() if info[i].format == 0 else b"")
            out.append(d)
        return out

    def allelic_fraction(self, traces, bcpos, primary, secdecompose, trim_left=50, trim_right=50):
        """allelicFraction(c, tr, bc) for a batch (reference src/decompose.h:412-617) -> float64[N][2] = (bestI, bestJ)."""
        n = len(traces)
        tr = [np.ascontiguousarray(t, np.int32) for t in traces]
        bp = [np.ascontiguousarray(b, np.int32) for b in bcpos]
        tlen = np.array([t.shape[1] for t in tr], np.int32)
        toff = np.concatenate([[0], np.cumsum(4 * tlen.astype(np.int64))[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
        tbase = np.concatenate([t.reshape(-1) for t in tr]) if n else np.zeros(1, np.int32)
        pri, sec = pack_seqs(primary), pack_seqs(secdecompose)
        blen = np.array([len(b) for b in bp], np.int32)
        if not (np.array_equal(pri.len, blen) and np.array_equal(sec.len, blen)):
            raise ValueError("bcpos, primary and secdecompose must have one entry per basecall")
        bbase = np.concatenate(bp) if n and blen.sum() else np.zeros(1, np.int32)
        out = np.zeros((2, max(n, 1)), np.float64)
        b = capi.FractionBatch(capi.Arena(_ptr(tbase), _ptr(toff), _ptr(tlen)), capi.Arena(_ptr(bbase), _ptr(pri.off), _ptr(blen)),
                               _ptr(pri.base), _ptr(sec.base), trim_left, trim_right, n, capi.TB_MEM_HOST)
        self._check(self._lib.tb_allelic_fraction(self._h, C.byref(b), _ptr(out[0]), _ptr(out[1])))
        return np.ascontiguousarray(out[:, :n].T)

    def basecall(self, traces, ploc, sigratio=0.33):
        """basecall(Trace, BaseCalls, sigratio) for a batch (reference src/abif.h:408-511). traces: list of int32[4][nsamples];
        ploc: the trace files' basecall positions (Trace::basecallpos). Returns a list of dicts bcPos / primary / secondary /
        consensus (estimateQualities is not part of the device path)."""
        n = len(traces)
        tr = [np.ascontiguousarray(t, np.int32) for t in traces]
        pl = [np.ascontiguousarray(p, np.int32) for p in ploc]
        tlen = np.array([t.shape[1] for t in tr], np.int32)
        toff = np.concatenate([[0], np.cumsum(4 * tlen.astype(np.int64))[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
        tbase = np.concatenate([t.reshape(-1) for t in tr]) if n else np.zeros(1, np.int32)
        plen = np.array([len(p) for p in pl], np.int32)
        poff = np.concatenate([[0], np.cumsum(plen.astype(np.int64))[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
        pbase = np.concatenate(pl) if n and plen.sum() else np.zeros(1, np.int32)
        tot = max(int(plen.astype(np.int64).sum()), 1)
        o_pos = np.zeros(tot, np.int32)
        o_pri, o_sec, o_con = (np.zeros(tot, np.uint8) for _ in range(3))
        o_len = np.zeros(max(n, 1), np.int32)
        b = capi.BasecallBatch(capi.Arena(_ptr(tbase), _ptr(toff), _ptr(tlen)), capi.Arena(_ptr(pbase), _ptr(poff), _ptr(plen)), n, capi.TB_MEM_HOST)
        self._check(self._lib.tb_basecall(self._h, C.byref(b), float(sigratio), _ptr(o_pos), _ptr(o_pri), _ptr(o_sec), _ptr(o_con), _ptr(poff), _ptr(o_len)))
        out = []
        for i in range(n):
            sl = slice(int(poff[i]), int(poff[i]) + int(o_len[i]))
            out.append(dict(bcPos=o_pos[sl].copy(), primary=o_pri[sl].tobytes(), secondary=o_sec[sl].tobytes(), consensus=o_con[sl].tobytes()))
        return out

    def revcomp_profile(self, profiles):
        """reverseComplementProfile for a batch of float[6][len] profiles (reference src/profile.h:74-90)."""
        a = profiles if isinstance(profiles, Arena) else pack_profiles(profiles)
        out = np.zeros(max(int(6 * a.len.astype(np.int64).sum()), 1), np.float32)
        ooff = np.concatenate([[0], np.cumsum(6 * a.len.astype(np.int64))[:-1]]).astype(np.int64) if a.n else np.zeros(0, np.int64)
        arena = capi.Arena(_ptr(a.base), _ptr(a.off), _ptr(a.len))
        self._check(self._lib.tb_revcomp_profile(self._h, C.byref(arena), a.n, capi.TB_MEM_HOST, _ptr(out), _ptr(ooff)))
        return [out[ooff[i]: ooff[i] + 6 * a.len[i]].reshape(6, a.len[i]).copy() for i in range(a.n)]

    # ---- helpers ------------------------------------------------------------------------------------------
    def rows_from_ops(self, kind, a1, a2, ops):
        """Gapped rows as gotoh() leaves them in `align` (reference src/align.h:196-293)."""
        return rows_from_ops(kind, a1, a2, ops)


def unpack_ops(packed, ops_len):
    """2-bit packed ops (Context.gotoh(..., packed=True)) of ONE pair -> bytes of 's' / 'h' / 'v'."""
    buf = np.ascontiguousarray(packed, np.uint8)
    out = np.zeros(max(int(ops_len), 1), np.uint8)
    rc = capi.lib().tb_unpack_ops(_ptr(buf), int(ops_len), _ptr(out))
    if rc != capi.TB_OK:
        raise TracyError(rc, "tb_unpack_ops")
    return out[: int(ops_len)].tobytes()


def rows_from_ops(kind, a1, a2, ops):
    lib = capi.lib()
    ops = np.frombuffer(bytes(ops), np.uint8)
    L = len(ops)
    r0, r1 = C.create_string_buffer(L + 1), C.create_string_buffer(L + 1)

    def conv(x, is_prof):
        if is_prof:
            x = np.ascontiguousarray(x, np.float32)
            return x, x.ctypes.data_as(C.c_void_p), x.shape[1]
        x = np.frombuffer(bytes(x), np.uint8)
        return x, x.ctypes.data_as(C.c_void_p), len(x)

    k1, p1, l1 = conv(a1, kind != SS)
    k2, p2, l2 = conv(a2, kind == PP)
    rc = lib.tb_rows_from_ops(_KIND_ID[kind], p1, l1, p2, l2, ops.ctypes.data_as(C.c_void_p), L, C.cast(r0, C.c_void_p), C.cast(r1, C.c_void_p))
    if rc != capi.TB_OK:
        raise TracyError(rc, lib.tb_strerror(rc).decode())
    return r0.raw[:L], r1.raw[:L]


class MultiContext(Context):
    """Several GPUs of one node behind one handle (tb_multi, csrc/multi.cu): gotoh() spreads the pairs of a HOST batch over the
    devices in cost-balanced contiguous ranges (one host thread and one chunk pipeline per device, every device writes its slice of
    the result arrays; no data-path collective), build_index() ships the reference text to every device (PCIe once, then GPU to
    GPU) and indexes it there, anchor() shards the traces. devices: list of device ordinals (None = every visible device); the same
    ordinal may appear more than once. Everything else (profiles, sweeps, ingest) runs on the first device's context."""

    def __init__(self, devices=None):
        self._lib = capi.lib()
        L = self._lib
        L.tb_multi_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]
        L.tb_multi_destroy.argtypes = [C.c_void_p]
        L.tb_multi_destroy.restype = None
        L.tb_multi_size.argtypes = [C.c_void_p]
        L.tb_multi_ctx.argtypes = [C.c_void_p, C.c_int]
        L.tb_multi_ctx.restype = C.c_void_p
        L.tb_multi_last_error.argtypes = [C.c_void_p]
        L.tb_multi_last_error.restype = C.c_char_p
        L.tb_multi_gotoh.argtypes = [C.c_void_p, C.c_int, C.c_void_p, capi.Score, capi.AlignConfig, C.c_void_p, C.POINTER(C.c_size_t)]
        L.tb_multi_index_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
        L.tb_multi_anchor.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t, capi.AnchorConfig, C.c_void_p]
        h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices) if devices else None
        rc = L.tb_multi_create(C.byref(h), devs, len(devices) if devices else 0)
        if rc != capi.TB_OK:
            raise TracyError(rc, L.tb_strerror(rc).decode() + " (tb_multi_create: are the B200s visible? there is no CPU fallback)")
        self._m = h
        self.size = L.tb_multi_size(h)
        self._h = C.c_void_p(L.tb_multi_ctx(h, 0))             # single-device calls of the base class go to the first device
        self.device = devices[0] if devices else 0
        self.last_ranges = None

    def close(self):
        if getattr(self, "_m", None):
            self._lib.tb_multi_destroy(self._m)
            self._m = None
            self._h = None

    def _mcheck(self, rc):
        if rc != capi.TB_OK:
            raise TracyError(rc, self._lib.tb_strerror(rc).decode() + ": " + self._lib.tb_multi_last_error(self._m).decode())

    def device_stats(self):
        out = []
        for i in range(self.size):
            k, a, b = C.c_uint64(), C.c_uint64(), C.c_uint64()
            self._lib.tb_ctx_stats(C.c_void_p(self._lib.tb_multi_ctx(self._m, i)), C.byref(k), C.byref(a), C.byref(b))
            out.append({"kernel_launches": k.value, "h2d_bytes": a.value, "d2h_bytes": b.value})
        return out

    def _run_gotoh(self, kind, b, sc, ac, r):
        first = (C.c_size_t * (self.size + 1))()
        self._mcheck(self._lib.tb_multi_gotoh(self._m, {PP: 0, SS: 1, PS: 2}[kind], C.byref(b), sc.c(), ac.c(), C.byref(r), first))
        self.last_ranges = [int(x) for x in first]

    def build_index(self, text):
        text = np.frombuffer(bytes(text), np.uint8) if not isinstance(text, np.ndarray) else np.ascontiguousarray(text, np.uint8)
        hs = (C.c_void_p * self.size)()
        self._mcheck(self._lib.tb_multi_index_build(self._m, _ptr(text), text.size, hs))
        return MultiKmerIndex(self, [C.c_void_p(h) for h in hs])

    def anchor(self, index, consensus, trim_left=50, trim_right=50, kmer=15, min_kmer_support=3):
        cons = consensus if isinstance(consensus, Arena) else pack_seqs(consensus)
        n = cons.n
        out = dict(anchored=np.zeros(n, np.uint8), forward=np.ones(n, np.uint8), kmersupport=np.zeros(n, np.uint32),
                   bestpos=np.zeros(n, np.int64), pass_=np.zeros(n, np.uint8))
        a = capi.Arena(_ptr(cons.base), _ptr(cons.off), _ptr(cons.len))
        r = capi.AnchorResult(_ptr(out["anchored"]), _ptr(out["forward"]), _ptr(out["kmersupport"]), _ptr(out["bestpos"]), _ptr(out["pass_"]))
        hs = (C.c_void_p * self.size)(*[h.value for h in index._hs])
        self._mcheck(self._lib.tb_multi_anchor(self._m, hs, C.byref(a), n, capi.AnchorConfig(trim_left, trim_right, kmer, min_kmer_support), C.byref(r)))
        out["anchored"] = out["anchored"].astype(bool)
        out["forward"] = out["forward"].astype(bool)
        return out


class MultiKmerIndex:
    """One device-resident index per device of a MultiContext."""

    def __init__(self, mctx, hs):
        self._ctx, self._hs = mctx, hs

    def close(self):
        if self._hs and getattr(self._ctx, "_m", None):
            for i, h in enumerate(self._hs):
                self._ctx._lib.tb_index_destroy(C.c_void_p(self._ctx._lib.tb_multi_ctx(self._ctx._m, i)), h)
        self._hs = None


class KmerIndex:
    """Device-resident sorted k-mer index of one reference text (tb_index)."""

    def __init__(self, ctx, h):
        self._ctx, self._h = ctx, h
        n, b, t = C.c_int64(), C.c_uint64(), C.c_void_p()
        ctx._lib.tb_index_info(h, C.byref(n), C.byref(b), C.byref(t))
        self.text_len, self.device_bytes, self.device_text = n.value, b.value, t.value

    def close(self):
        if self._h and getattr(self._ctx, "_h", None):
            self._ctx._lib.tb_index_destroy(self._ctx._h, self._h)
        self._h = None


def reference_slice(bestpos, seqlen, conslen, maxindel=1000):
    """Slice arithmetic of getReferenceSlice (reference src/fmindex.h:286-299): -> (refindex, chrpos, slicestart, sliceend).
    seqlen: per sequence, length + 1 for a b"\n"-joined genome (src/fmindex.h:247), the plain length for a single sequence.
    For an indexed genome the reference then fetches [slicestart, min(sliceend, len - 1)] INCLUSIVE (htslib faidx_fetch_seq)."""
    lib = capi.lib()
    sl = np.ascontiguousarray(seqlen, np.uint32)
    ri, cp, s0, s1 = C.c_int32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = lib.tb_reference_slice(int(bestpos), _ptr(sl), sl.size, conslen, maxindel, C.byref(ri), C.byref(cp), C.byref(s0), C.byref(s1))
    if rc != capi.TB_OK:
        raise TracyError(rc, lib.tb_strerror(rc).decode())
    return ri.value, cp.value, s0.value, s1.value


def scan_traces(files):
    """traceFormat + the directory walk of readab / readscf (host only, no GPU): one dict per file with format, ok, status,
    nsamples, nbasecalls (reference src/scf.h:19-35, src/abif.h:300-388, src/scf.h:56-93)."""
    lib = capi.lib()
    n = len(files)
    flen = np.array([len(f) for f in files], np.int64)
    foff = np.concatenate([[0], np.cumsum(flen)[:-1]]).astype(np.int64) if n else np.zeros(0, np.int64)
    blob = np.frombuffer(b"".join(bytes(f) for f in files) or b"\0", np.uint8)
    info = (capi.TraceInfo * max(n, 1))()
    rc = lib.tb_trace_scan(_ptr(blob), _ptr(foff), _ptr(flen), n, C.cast(info, C.c_void_p))
    if rc != capi.TB_OK:
        raise TracyError(rc, lib.tb_strerror(rc).decode())
    return [dict(format=info[i].format, ok=bool(info[i].ok), status=info[i].status, nsamples=info[i].nsamples, nbasecalls=info[i].nbasecalls) for i in range(n)]


def trim_reference_slice(row0, row1, refslice, forward=True, pos=0, trim_left=0, trim_right=0):
    """trimReferenceSlice(c, align, rs), reference src/fmindex.h:429-463. Returns (trimmed refslice, new pos)."""
    lib = capi.lib()
    row0, row1, refslice = bytes(row0), bytes(row1), bytes(refslice)
    ri, rsz, npos = C.c_int32(), C.c_int32(), C.c_uint32()
    rc = lib.tb_trim_reference_slice(row0, row1, len(row0), len(refslice), int(bool(forward)), pos, trim_left, trim_right,
                                     C.byref(ri), C.byref(rsz), C.byref(npos))
    if rc != capi.TB_OK:
        raise TracyError(rc, lib.tb_strerror(rc).decode())
    return refslice[ri.value: ri.value + rsz.value], npos.value


def find_breakpoint(profile):
    """findBreakpoint(ptrace, bp), reference src/decompose.h:7-56 -> dict(indelshift, traceleft, breakpoint, bestDiff)."""
    lib = capi.lib()
    p = np.ascontiguousarray(profile, np.float32)
    a, b, c, d = C.c_int32(), C.c_int32(), C.c_uint32(), C.c_float()
    rc = lib.tb_find_breakpoint(p.ctypes.data_as(C.c_void_p), p.shape[1], C.byref(a), C.byref(b), C.byref(c), C.byref(d))
    if rc != capi.TB_OK:
        raise TracyError(rc, lib.tb_strerror(rc).decode())
    return dict(indelshift=bool(a.value), traceleft=bool(b.value), breakpoint=c.value, bestDiff=d.value)


# ---- single-pair mirrors of the reference call shapes --------------------------------------------------------
_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _kind_of(a1, a2):
    s1 = isinstance(a1, (bytes, bytearray, str))
    s2 = isinstance(a2, (bytes, bytearray, str))
    if s1 and s2:
        return SS
    if not s1 and s2:
        return PS
    if not s1 and not s2:
        return PP
    raise TypeError("sequence x profile is not a pairing the reference instantiates")


def _b(x):
    return x.encode() if isinstance(x, str) else x


def gotohScore(a1, a2, ac=AlignConfig(), sc=DnaScore(), ctx=None):
    """int gotohScore(a1, a2, ac, sc) -- reference src/gotoh.h:12-14."""
    ctx = ctx or default_context()
    kind = _kind_of(a1, a2)
    s, _, _ = ctx.gotoh(kind, [_b(a1)], [_b(a2)], sc, ac, traceback=False)
    return int(s[0])


def gotoh(a1, a2, ac=AlignConfig(), sc=DnaScore(), ctx=None):
    """int gotoh(a1, a2, align, ac, sc) -- reference src/gotoh.h:71-73. Returns (score, (row0, row1))."""
    ctx = ctx or default_context()
    kind = _kind_of(a1, a2)
    s, ops, ol = ctx.gotoh(kind, [_b(a1)], [_b(a2)], sc, ac, traceback=True)
    o = bytes(ops[0, : ol[0]])
    return int(s[0]), rows_from_ops(kind, _b(a1), _b(a2), o)
