"""ctypes loaders for the two CPU oracles.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package tracy_b200 never does (tests/test_boundary.py greps for it).

  port()  -> oracle/libgotoh_oracle.so   the plain-C restatement (oracle/gotoh_oracle.c)
  ref()   -> oracle/_ref/libtracy_ref.so the unmodified reference headers (oracle/ref_bridge.cpp); None if absent
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_SC = [C.c_int] * 6  # hfree, vfree, match, mismatch, go, ge


def build(force=False):
    """Compile the C restatement, and the reference bridge when /root/reference exists (make decides)."""
    so = os.path.join(_HERE, "libgotoh_oracle.so")
    src = os.path.join(_HERE, "gotoh_oracle.c")
    need = force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
    ref_so = os.path.join(_HERE, "_ref", "libtracy_ref.so")
    bridge = os.path.join(_HERE, "ref_bridge.cpp")
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(ref_so) or os.path.getmtime(ref_so) < os.path.getmtime(bridge)):
        need = True
    if os.path.isdir("/root/reference/src"):   # the drop-in check binary (tests/cpp/dropin.cpp against the reference headers + include/tracy_b200.hpp)
        exe = os.path.join(_HERE, "_ref", "dropin_test")
        deps = [os.path.join(_HERE, "..", "tests", "cpp", "dropin.cpp"), os.path.join(_HERE, "..", "include", "tracy_b200.hpp"),
                os.path.join(_HERE, "..", "include", "tracy_b200.h")]
        if force or not os.path.exists(exe) or any(os.path.getmtime(exe) < os.path.getmtime(d) for d in deps):
            need = True
    if need:
        subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)


class _Port:
    def __init__(self, path):
        L = self.lib = C.CDLL(path)
        L.orc_gotoh_score_pp.argtypes = [_f32p, C.c_long, _f32p, C.c_long] + _SC
        L.orc_gotoh_score_pp.restype = C.c_int
        L.orc_gotoh_pp.argtypes = [_f32p, C.c_long, _f32p, C.c_long] + _SC + [C.c_char_p, C.POINTER(C.c_long)]
        L.orc_gotoh_pp.restype = C.c_int
        L.orc_gotoh_ps.argtypes = [_f32p, C.c_long, C.c_char_p, C.c_long] + _SC + [C.c_char_p, C.POINTER(C.c_long)]
        L.orc_gotoh_ps.restype = C.c_int
        L.orc_gotoh_ss.argtypes = [C.c_char_p, C.c_long, C.c_char_p, C.c_long] + _SC + [C.c_char_p, C.POINTER(C.c_long)]
        L.orc_gotoh_ss.restype = C.c_int
        L.orc_rows_from_ops.argtypes = [C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_char_p, C.c_long, C.c_char_p, C.c_char_p]
        L.orc_rows_from_ops.restype = None
        L.orc_onehot_profile.argtypes = [C.c_char_p, C.c_long, _f32p]
        L.orc_revcomp_profile.argtypes = [_f32p, C.c_long, _f32p]
        L.orc_phase_ref_allele.argtypes = [C.c_char, C.c_char, C.c_char]
        L.orc_phase_ref_allele.restype = C.c_char
        L.orc_decompose_sweep.argtypes = [C.c_char_p, C.c_long, C.c_char_p, C.c_char_p, C.c_long, C.c_long, C.c_long,
                                          C.c_int, C.c_int, _i32p, _i32p, C.c_void_p]
        L.orc_decompose_sweep.restype = None
        L.orc_bench_gotoh_ps.argtypes = [_f32p, C.c_char_p, C.c_int, C.c_long, C.c_long] + _SC + [C.c_int, _i32p]
        L.orc_bench_gotoh_ps.restype = C.c_longlong

        _i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        L.orc_scan_sequence.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, _i64p, C.c_longlong,
                                        C.POINTER(C.c_longlong), C.POINTER(C.c_uint)]
        L.orc_scan_sequence.restype = C.c_longlong
        L.orc_anchor.argtypes = [C.c_char_p, C.c_longlong, C.c_char_p, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, _i64p, C.c_longlong,
                                 C.POINTER(C.c_int), C.POINTER(C.c_uint), C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
        L.orc_anchor.restype = C.c_int

    def scan_sequence(self, text, cons, trim_left, trim_right, kmer, unique):
        """scanSequence + findMaxFreq of one strand -> (sorted hits int64[], gpos, freq)."""
        cap = max(1, len(cons)) * (1 if unique else 1000)
        hits = np.zeros(cap, np.int64)
        g, f = C.c_longlong(0), C.c_uint(0)
        n = self.lib.orc_scan_sequence(bytes(text), len(text), bytes(cons), len(cons), trim_left, trim_right, kmer, int(unique), hits, cap, C.byref(g), C.byref(f))
        return hits[:n].copy(), int(g.value), int(f.value)

    def anchor(self, text, cons, trim_left, trim_right, kmer, min_support):
        """getReferenceSlice's decision -> (anchored, forward, kmersupport, bestpos, pass)."""
        cap = max(1, len(cons)) * 1000
        scratch = np.zeros(cap, np.int64)
        fw, ks, bp, ps = C.c_int(1), C.c_uint(0), C.c_longlong(0), C.c_int(0)
        ok = self.lib.orc_anchor(bytes(text), len(text), bytes(cons), len(cons), trim_left, trim_right, kmer, min_support, scratch, cap,
                                 C.byref(fw), C.byref(ks), C.byref(bp), C.byref(ps))
        return bool(ok), bool(fw.value), int(ks.value), int(bp.value), int(ps.value)

    @staticmethod
    def _prof(p):
        p = np.ascontiguousarray(p, dtype=np.float32)
        assert p.ndim == 2 and p.shape[0] == 6
        return p

    def gotoh_score_pp(self, p1, p2, hfree, vfree, sc):
        p1, p2 = self._prof(p1), self._prof(p2)
        return self.lib.orc_gotoh_score_pp(p1, p1.shape[1], p2, p2.shape[1], hfree, vfree, *sc)

    def _run(self, fn, a, m, b, n, hfree, vfree, sc):
        buf = C.create_string_buffer(m + n + 1)
        L = C.c_long(0)
        s = fn(a, m, b, n, hfree, vfree, *sc, buf, C.byref(L))
        return s, buf.raw[: L.value]

    def gotoh_pp(self, p1, p2, hfree, vfree, sc):
        p1, p2 = self._prof(p1), self._prof(p2)
        return self._run(self.lib.orc_gotoh_pp, p1, p1.shape[1], p2, p2.shape[1], hfree, vfree, sc)

    def gotoh_ps(self, p1, seq, hfree, vfree, sc):
        p1 = self._prof(p1)
        return self._run(self.lib.orc_gotoh_ps, p1, p1.shape[1], bytes(seq), len(seq), hfree, vfree, sc)

    def gotoh_ss(self, s1, s2, hfree, vfree, sc):
        return self._run(self.lib.orc_gotoh_ss, bytes(s1), len(s1), bytes(s2), len(s2), hfree, vfree, sc)

    def rows_from_ops(self, a, b, ops):
        """a, b: both float[6][len] profiles or both bytes."""
        L = len(ops)
        r0, r1 = C.create_string_buffer(L + 1), C.create_string_buffer(L + 1)
        if isinstance(a, (bytes, bytearray)):
            ka, kb = C.create_string_buffer(bytes(a), len(a) + 1), C.create_string_buffer(bytes(b), len(b) + 1)
            self.lib.orc_rows_from_ops(1, C.cast(ka, C.c_void_p), len(a), C.cast(kb, C.c_void_p), len(b), bytes(ops), L, r0, r1)
        else:
            a, b = self._prof(a), self._prof(b)
            self.lib.orc_rows_from_ops(0, a.ctypes.data_as(C.c_void_p), a.shape[1], b.ctypes.data_as(C.c_void_p), b.shape[1], bytes(ops), L, r0, r1)
        return r0.raw[:L], r1.raw[:L]

    def onehot(self, seq):
        out = np.zeros((6, len(seq)), np.float32)
        self.lib.orc_onehot_profile(bytes(seq), len(seq), out)
        return out

    def revcomp_profile(self, p):
        p = self._prof(p)
        out = np.empty_like(p)
        self.lib.orc_revcomp_profile(p, p.shape[1], out)
        return out

    def phase(self, pri, sec, r):
        return self.lib.orc_phase_ref_allele(pri, sec, r)

    def decompose_sweep(self, refrow, primary, secondary, vi_end, align_index, var_index, ndel, nins, grid=False):
        fref = np.zeros(max(ndel, 1), np.int32)
        fins = np.zeros(max(nins, 1), np.int32)
        g = np.zeros((max(nins, 1), max(ndel, 1)), np.int32) if grid else None
        self.lib.orc_decompose_sweep(bytes(refrow), len(refrow), bytes(primary), bytes(secondary), vi_end, align_index, var_index,
                                     ndel, nins, fref, fins, g.ctypes.data_as(C.c_void_p) if grid else None)
        return fref[:ndel], fins[:nins], (g[:nins, :ndel] if grid else None)

    def bench_gotoh_ps(self, profs, seqs, m, n, hfree, vfree, sc, with_traceback=True):
        profs = np.ascontiguousarray(profs, np.float32)
        npairs = profs.shape[0]
        scores = np.zeros(npairs, np.int32)
        cells = self.lib.orc_bench_gotoh_ps(profs.reshape(-1), bytes(seqs), npairs, m, n, hfree, vfree, *sc, int(with_traceback), scores)
        return cells, scores


class _Ref:
    """The reference's own code (unmodified headers) behind oracle/ref_bridge.cpp."""

    def __init__(self, path):
        L = self.lib = C.CDLL(path)
        for name, a, b in (("pp", _f32p, _f32p), ("ps", _f32p, C.c_char_p), ("ss", C.c_char_p, C.c_char_p)):
            f = getattr(L, "ref_gotoh_score_" + name)
            f.argtypes = [a, C.c_int, b, C.c_int] + _SC
            f.restype = C.c_int
            g = getattr(L, "ref_gotoh_" + name)
            g.argtypes = [a, C.c_int, b, C.c_int] + _SC + [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
            g.restype = C.c_int
        L.ref_onehot_profile.argtypes = [C.c_char_p, C.c_int, _f32p]
        L.ref_revcomp_profile.argtypes = [_f32p, C.c_int, _f32p]
        L.ref_create_profile.argtypes = [_i32p, C.c_int, _i32p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int]
        L.ref_create_profile.restype = C.c_int
        L.ref_find_breakpoint.argtypes = [_f32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.ref_decompose_alleles.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_uint32, C.c_int, _i32p, C.c_int]
        L.ref_decompose_alleles.restype = C.c_int
        L.ref_bench_gotoh_ps.argtypes = [_f32p, C.c_char_p, C.c_int, C.c_int, C.c_int] + _SC + [C.c_int, _i32p]
        L.ref_bench_gotoh_ps.restype = C.c_longlong
        L.ref_reverse_complement.argtypes = [C.c_char_p, C.c_int]
        L.ref_basecall.argtypes = [_i32p, C.c_int, _i32p, C.c_int, C.c_float, _i32p, C.c_char_p, C.c_char_p, C.c_char_p]
        L.ref_basecall.restype = C.c_int
        L.ref_find_homozygous_breakpoint.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.ref_find_homozygous_breakpoint.restype = C.c_int
        L.ref_generate_secondary_decomposed.argtypes = [_i32p, C.c_int, _i32p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p]
        L.ref_trim_reference_slice.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_int, C.c_int,
                                               C.c_char_p, C.c_int]
        L.ref_trim_reference_slice.restype = C.c_int
        L.ref_profile_from_alignment.argtypes = [C.c_char_p, C.c_int, C.c_int, _f32p]
        _i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        _u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        L.ref_rev_seq_based_on_dist.argtypes = [_f32p, _i64p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.ref_msa.argtypes = [_f32p, _i64p, _i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_char_p, C.c_int,
                              np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS"), _i32p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_msa.restype = C.c_int

        L.ref_fm_build.argtypes = [C.c_char_p, C.c_longlong]
        L.ref_fm_build.restype = C.c_void_p
        L.ref_fm_free.argtypes = [C.c_void_p]
        L.ref_fm_count.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ref_fm_count.restype = C.c_longlong
        L.ref_scan_sequence.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i64p, C.c_longlong,
                                        C.POINTER(C.c_longlong), C.POINTER(C.c_uint)]
        L.ref_scan_sequence.restype = C.c_longlong
        L.ref_set_genome.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_get_reference_slice.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                              C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.c_char_p, C.c_int]
        L.ref_get_reference_slice.restype = C.c_int

    def plot_alignment(self, row0, row1, chr_name, pos, refslice_len, forward, score, key=0, a1a2=(0.0, 0.0), linelimit=60):
        """plotAlignment(...) -> the text it writes (src/fmindex.h:329-420)."""
        import tempfile
        p = tempfile.mktemp()
        self.lib.ref_plot_alignment.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_double, C.c_double, C.c_uint]
        self.lib.ref_plot_alignment(os.fsencode(p), bytes(row0), bytes(row1), len(row0), bytes(chr_name), pos, refslice_len, int(forward), key, score,
                                    float(a1a2[0]), float(a1a2[1]), linelimit)
        out = open(p, "rb").read()
        os.remove(p)
        return out

    def trace_outputs(self, what, acgt, bcpos, qual, primary, secondary, consensus, trim_left=0, trim_right=0, row0=b"", row1=b"", chr_name=b"", pos=0, forward=True):
        """what = "txt": traceTxtOut (src/abif.h:512-534); "json": traceJsonOut (src/json.h:108-117); "align_json":
        alignmentTracePadding + traceAlignJsonOut (src/json.h:383-479, 197-217). Returns the bytes the reference writes."""
        import tempfile
        p = tempfile.mktemp()
        acgt = np.ascontiguousarray(acgt, np.int32)
        bcpos = np.ascontiguousarray(bcpos, np.int32)
        qual = np.ascontiguousarray(qual, np.uint8)
        self.lib.ref_trace_outputs.argtypes = [C.c_char_p, C.c_int, _i32p, C.c_int, _i32p, np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS"), C.c_char_p, C.c_char_p,
                                               C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_uint, C.c_int]
        self.lib.ref_trace_outputs.restype = None
        self.lib.ref_trace_outputs(os.fsencode(p), {"txt": 0, "json": 1, "align_json": 2}[what], acgt.reshape(-1), acgt.shape[1], bcpos, qual, bytes(primary),
                                   bytes(secondary), bytes(consensus), len(bcpos), trim_left, trim_right, bytes(row0), bytes(row1), len(row0), bytes(chr_name),
                                   pos, int(forward))
        out = open(p, "rb").read()
        os.remove(p)
        return out

    def call_variants(self, alignments):
        """callVariants (src/variants.h:56-126) over a sequence of (row0, row1, chr, pos) into ONE variant vector, as indigo() calls
        it for both alleles. Returns the list of (pos, basenum, gt, chr, ref, alt, type) the reference ends up with."""
        L = self.lib
        L.ref_call_variants.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_uint]
        L.ref_variants_dump.argtypes = [C.c_char_p, C.c_int]
        L.ref_variants_dump.restype = C.c_int
        L.ref_variants_reset()
        for row0, row1, chr_name, pos in alignments:
            L.ref_call_variants(bytes(row0), bytes(row1), len(row0), bytes(chr_name), pos)
        buf = C.create_string_buffer(1 << 20)
        n = L.ref_variants_dump(buf, len(buf))
        assert n >= 0
        out = []
        for line in buf.raw[:n].decode("latin-1").splitlines():
            p, b, g, c, r, a, t = line.split("\t")
            out.append((int(p), int(b), int(g), c, r, a, t))
        return out

    def decompose_json(self, cfg, acgt, bcpos, qual, primary, secondary, alignments, allele1, allele2, align3, decomp, indelshift, breakpoint, a1a2, sort=True):
        """traceAlleleAlignJsonOut (src/json.h:260-381) -> the bytes of P.json. alignments: the (row0, row1, chr, pos) sequence whose
        callVariants calls make the variant vector; allele1/allele2: (row0, row1, chr, pos, forward, score); align3: (row0, row1, score)."""
        import tempfile
        L = self.lib
        L.ref_call_variants.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_uint]
        L.ref_variants_reset()
        for row0, row1, chr_name, pos in alignments:
            L.ref_call_variants(bytes(row0), bytes(row1), len(row0), bytes(chr_name), pos)
        prefix = tempfile.mktemp()
        acgt = np.ascontiguousarray(acgt, np.int32)
        bcpos = np.ascontiguousarray(bcpos, np.int32)
        qual = np.ascontiguousarray(qual, np.uint8)
        d = np.ascontiguousarray(decomp, np.int32).reshape(-1)
        u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        al = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_uint, C.c_int, C.c_int]
        L.ref_decompose_json.argtypes = ([C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_char_p, C.c_char_p, _i32p, C.c_int, _i32p, u8, C.c_char_p, C.c_char_p, C.c_int]
                                         + al + al + [C.c_char_p, C.c_char_p, C.c_int, C.c_int, _i32p, C.c_int, C.c_int, C.c_uint, C.c_double, C.c_double, C.c_int])
        L.ref_decompose_json.restype = None
        a1 = [bytes(allele1[0]), bytes(allele1[1]), len(allele1[0]), bytes(allele1[2]), allele1[3], int(allele1[4]), allele1[5]]
        a2 = [bytes(allele2[0]), bytes(allele2[1]), len(allele2[0]), bytes(allele2[2]), allele2[3], int(allele2[4]), allele2[5]]
        L.ref_decompose_json(os.fsencode(prefix), cfg["trim_left"], cfg["trim_right"], cfg["qual_cut"], cfg["pratio"], os.fsencode(cfg["input"]), os.fsencode(cfg["genome"]),
                             acgt.reshape(-1), acgt.shape[1], bcpos, qual, bytes(primary), bytes(secondary), len(bcpos), *a1, *a2,
                             bytes(align3[0]), bytes(align3[1]), len(align3[0]), align3[2], d, len(d) // 2, int(indelshift), breakpoint, float(a1a2[0]), float(a1a2[1]), int(sort))
        out = open(prefix + ".json", "rb").read()
        os.remove(prefix + ".json")
        return out

    def estimate_qualities(self, bcpos, primary, secondary):
        """estimateQualities + findBestTraceSection(bc) (src/abif.h:164-253) -> (estQual uint8[n], best section index)."""
        n = len(bcpos)
        u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        self.lib.ref_estimate_qualities.argtypes = [_i32p, C.c_char_p, C.c_char_p, C.c_int, u8, C.POINTER(C.c_uint32)]
        q = np.zeros(max(n, 1), np.uint8)
        best = C.c_uint32(0)
        self.lib.ref_estimate_qualities(np.ascontiguousarray(bcpos, np.int32), bytes(primary), bytes(secondary), n, q, C.byref(best))
        return q[:n], int(best.value)

    def trim_trace(self, bcpos, secondary, stringency):
        """trimTrace(c, bc, leftTrim, rightTrim) (src/trim.h:35-73) -> (left, right)."""
        self.lib.ref_trim_trace.argtypes = [_i32p, C.c_char_p, C.c_int, C.c_float, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        le, ri = C.c_uint32(0), C.c_uint32(0)
        self.lib.ref_trim_trace(np.ascontiguousarray(bcpos, np.int32), bytes(secondary), len(bcpos), float(stringency), C.byref(le), C.byref(ri))
        return int(le.value), int(ri.value)

    def trim_basecalls(self, nsamples, bcpos, qual, primary, secondary, consensus, trim_left, trim_right):
        """trimTrace(tr, bc, trimLeft, trimRight, nbc) (src/trim.h:76-99) -> (bcpos, qual, primary, secondary, consensus)."""
        n = len(bcpos)
        u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        self.lib.ref_trim_basecalls.argtypes = [C.c_int, _i32p, u8, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_uint, C.c_uint, _i32p, u8, C.c_char_p, C.c_char_p, C.c_char_p]
        self.lib.ref_trim_basecalls.restype = C.c_int
        ob, oq = np.zeros(n, np.int32), np.zeros(n, np.uint8)
        op, os_, oc = C.create_string_buffer(n + 1), C.create_string_buffer(n + 1), C.create_string_buffer(n + 1)
        m = self.lib.ref_trim_basecalls(nsamples, np.ascontiguousarray(bcpos, np.int32), np.ascontiguousarray(qual, np.uint8), bytes(primary), bytes(secondary),
                                        bytes(consensus), n, trim_left, trim_right, ob, oq, op, os_, oc)
        return ob[:m], oq[:m], op.raw[:m], os_.raw[:m], oc.raw[:m]

    def aligned_trace_by_row(self, rows, row, name, forward, isref):
        """alignedTraceByRow (src/json.h:220-246) -> the bytes it writes. rows: uint8[nrow][ncol] character alignment."""
        import tempfile
        p = tempfile.mktemp()
        rows = np.ascontiguousarray(rows, np.uint8)
        self.lib.ref_aligned_trace_by_row.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_uint, C.c_char_p, C.c_int, C.c_int]
        self.lib.ref_aligned_trace_by_row(os.fsencode(p), rows.tobytes(), rows.shape[0], rows.shape[1], row, name.encode(), int(forward), int(isref))
        out = open(p, "rb").read()
        os.remove(p)
        return out

    def reverse_complement_trace(self, acgt, bcpos, qual, primary, secondary, consensus):
        """reverseComplementTrace (src/trim.h:124-151) -> (acgt, bcpos, qual, primary, secondary, consensus)."""
        acgt = np.ascontiguousarray(acgt, np.int32)
        n, ns = len(bcpos), acgt.shape[1]
        u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        self.lib.ref_reverse_complement_trace.argtypes = [_i32p, C.c_int, _i32p, u8, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, _i32p, _i32p, u8, C.c_char_p, C.c_char_p, C.c_char_p]
        self.lib.ref_reverse_complement_trace.restype = C.c_int
        oa, ob, oq = np.zeros(4 * ns, np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
        op, os_, oc = C.create_string_buffer(n + 1), C.create_string_buffer(n + 1), C.create_string_buffer(n + 1)
        m = self.lib.ref_reverse_complement_trace(acgt.reshape(-1), ns, np.ascontiguousarray(bcpos, np.int32), np.ascontiguousarray(qual, np.uint8), bytes(primary),
                                                  bytes(secondary), bytes(consensus), n, oa, ob, oq, op, os_, oc)
        return oa.reshape(4, ns), ob[:m], oq[:m], op.raw[:m], os_.raw[:m], oc.raw[:m]

    def nearest_snp(self, primary, secondary, trim_left, trim_right, rtp):
        """nearestSNP(c, bc, rtp) (src/trim.h:11-33)."""
        self.lib.ref_nearest_snp.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint]
        self.lib.ref_nearest_snp.restype = C.c_uint
        return int(self.lib.ref_nearest_snp(bytes(primary), bytes(secondary), len(primary), trim_left, trim_right, rtp))

    def trace_fastx(self, fastq, otype, trim_left, trim_right, nsamples, bcpos, qual, primary, secondary, consensus):
        """traceFastaOut / traceFastqOut (src/fasta.h:98-158) -> the bytes they write."""
        import tempfile
        p = tempfile.mktemp()
        u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        self.lib.ref_trace_fastx.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, _i32p, u8, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        self.lib.ref_trace_fastx.restype = None
        self.lib.ref_trace_fastx(os.fsencode(p), int(fastq), otype.encode(), trim_left, trim_right, nsamples, np.ascontiguousarray(bcpos, np.int32),
                                 np.ascontiguousarray(qual, np.uint8), bytes(primary), bytes(secondary), bytes(consensus), len(bcpos))
        out = open(p, "rb").read()
        os.remove(p)
        return out

    def write_decomposition(self, pairs):
        import tempfile
        p = tempfile.mktemp()
        a = np.ascontiguousarray(pairs, np.int32).reshape(-1)
        self.lib.ref_write_decomposition.argtypes = [C.c_char_p, _i32p, C.c_int]
        self.lib.ref_write_decomposition(os.fsencode(p), a, len(a) // 2)
        out = open(p, "rb").read()
        os.remove(p)
        return out

    def read_trace(self, path):
        """traceFormat + readab / readscf on a file -> dict(format, ok, samples [list of 4 int32 arrays], basecallpos, qual, basecalls1, basecalls2)."""
        L = self.lib
        L.ref_trace_load.argtypes = [C.c_char_p] + [C.POINTER(C.c_int)] * 6
        L.ref_trace_load.restype = C.c_int
        fmt, nb, n1, n2, nq = (C.c_int(0) for _ in range(5))
        ns4 = (C.c_int * 4)()
        ok = L.ref_trace_load(os.fsencode(path), C.byref(fmt), ns4, C.byref(nb), C.byref(n1), C.byref(n2), C.byref(nq))
        ns = [int(x) for x in ns4]
        samples = np.zeros(max(sum(ns), 1), np.int32)
        ploc = np.zeros(max(nb.value, 1), np.int32)
        qual = np.zeros(max(nq.value, 1), np.uint8)
        b1, b2 = C.create_string_buffer(n1.value + 1), C.create_string_buffer(n2.value + 1)
        L.ref_trace_get.argtypes = [_i32p, _i32p, np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS"), C.c_char_p, C.c_char_p]
        L.ref_trace_get(samples, ploc, qual, b1, b2)
        chans, o = [], 0
        for k in range(4):
            chans.append(samples[o:o + ns[k]].copy()); o += ns[k]
        return dict(format=fmt.value, ok=bool(ok), samples=chans, basecallpos=ploc[:nb.value].copy(), qual=qual[:nq.value].copy(),
                    basecalls1=b1.raw[:n1.value], basecalls2=b2.raw[:n2.value])

    def allelic_fraction(self, acgt, bcpos, primary, secdecompose, trim_left, trim_right):
        """allelicFraction(c, tr, bc), src/decompose.h:412-617 -> (bestI, bestJ)."""
        acgt = np.ascontiguousarray(acgt, np.int32); bcpos = np.ascontiguousarray(bcpos, np.int32)
        a, b = C.c_double(0), C.c_double(0)
        self.lib.ref_allelic_fraction.argtypes = [_i32p, C.c_int, _i32p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        self.lib.ref_allelic_fraction(acgt.reshape(-1), acgt.shape[1], bcpos, bytes(primary), bytes(secdecompose), len(bcpos), trim_left, trim_right, C.byref(a), C.byref(b))
        return a.value, b.value

    # ---- anchoring: the reference's sdsl FM-index and src/fmindex.h:173-326 ----
    def fm_build(self, text):
        """construct_im(csa_wt<>, text, 1) -> opaque handle (free with fm_free)."""
        return self.lib.ref_fm_build(bytes(text), len(text))

    def fm_free(self, h):
        self.lib.ref_fm_free(h)

    def fm_count(self, h, pat):
        return int(self.lib.ref_fm_count(h, bytes(pat), len(pat)))

    def scan_sequence(self, h, cons, trim_left, trim_right, kmer, unique):
        cap = max(1, len(cons)) * (1 if unique else 1000)
        hits = np.zeros(cap, np.int64)
        g, f = C.c_longlong(0), C.c_uint(0)
        n = self.lib.ref_scan_sequence(h, bytes(cons), len(cons), trim_left, trim_right, kmer, int(unique), hits, cap, C.byref(g), C.byref(f))
        return hits[:n].copy(), int(g.value), int(f.value)

    def set_genome(self, names, seqs):
        """Hand the in-memory faidx stand-in its sequences (getReferenceSlice with filetype 0 fetches from it)."""
        self.lib.ref_set_genome(b"\n".join(names), b"\n".join(seqs))

    def get_reference_slice(self, h, filetype, cons, trim_left, trim_right, kmer, maxindel, min_support, refslice=b""):
        """getReferenceSlice -> dict(ok, forward, kmersupport, pos, chr, refslice)."""
        cap = max(len(refslice), len(cons) + 2 * maxindel + 16)
        buf = C.create_string_buffer(bytes(refslice), cap + 1)
        rl, fw, ks, pos = C.c_int(len(refslice)), C.c_int(1), C.c_uint(0), C.c_uint(0)
        chrb = C.create_string_buffer(256)
        ok = self.lib.ref_get_reference_slice(h, filetype, bytes(cons), len(cons), trim_left, trim_right, kmer, maxindel, min_support, buf, cap,
                                              C.byref(rl), C.byref(fw), C.byref(ks), C.byref(pos), chrb, 256)
        return dict(ok=bool(ok), forward=bool(fw.value), kmersupport=int(ks.value), pos=int(pos.value), chr=chrb.value, refslice=buf.raw[: min(rl.value, cap)])

    def basecall(self, acgt, ploc, sigratio=0.33):
        acgt = np.ascontiguousarray(acgt, np.int32); ploc = np.ascontiguousarray(ploc, np.int32)
        n = len(ploc)
        pos = np.zeros(max(n, 1), np.int32)
        p, s, c = (C.create_string_buffer(n + 1) for _ in range(3))
        k = self.lib.ref_basecall(acgt.reshape(-1), acgt.shape[1], ploc, n, sigratio, pos, p, s, c)
        return dict(bcPos=pos[:k].copy(), primary=p.raw[:k], secondary=s.raw[:k], consensus=c.raw[:k])

    def find_homozygous_breakpoint(self, row0, row1):
        a, b, c, d = C.c_int(), C.c_int(), C.c_uint32(), C.c_float()
        ok = self.lib.ref_find_homozygous_breakpoint(bytes(row0), bytes(row1), len(row0), C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return (bool(a.value), bool(b.value), c.value, d.value) if ok else None

    def generate_secondary_decomposed(self, acgt, bcpos, primary, secondary):
        acgt = np.ascontiguousarray(acgt, np.int32); bcpos = np.ascontiguousarray(bcpos, np.int32)
        out = C.create_string_buffer(len(primary) + 1)
        self.lib.ref_generate_secondary_decomposed(acgt.reshape(-1), acgt.shape[1], bcpos, bytes(primary), bytes(secondary), len(primary), out)
        return out.raw[: len(primary)]

    def reverse_complement(self, seq):
        buf = C.create_string_buffer(bytes(seq), len(seq) + 1)
        self.lib.ref_reverse_complement(buf, len(seq))
        return buf.raw[: len(seq)]

    def trim_reference_slice(self, row0, row1, refslice, forward, pos, trim_left, trim_right):
        out = C.create_string_buffer(len(refslice) + 1)
        p = C.c_uint32(pos)
        n = self.lib.ref_trim_reference_slice(bytes(row0), bytes(row1), len(row0), bytes(refslice), len(refslice), int(forward), C.byref(p),
                                              trim_left, trim_right, out, len(refslice) + 1)
        assert n >= 0
        return out.raw[:n], p.value

    def profile_from_alignment(self, rows):
        rows = np.ascontiguousarray(rows, np.uint8)
        out = np.zeros((6, rows.shape[1]), np.float32)
        self.lib.ref_profile_from_alignment(rows.tobytes(), rows.shape[0], rows.shape[1], out)
        return out

    @staticmethod
    def _pack(profiles):
        profiles = [np.ascontiguousarray(p, np.float32) for p in profiles]
        lens = np.array([p.shape[1] for p in profiles], np.int32)
        off = np.concatenate([[0], np.cumsum(6 * lens.astype(np.int64))[:-1]]).astype(np.int64)
        return np.concatenate([p.reshape(-1) for p in profiles]), off, lens

    def rev_seq_based_on_dist(self, profiles, fwd, sc):
        base, off, lens = self._pack(profiles)
        f = np.array([1 if x else 0 for x in fwd], np.uint8)
        self.lib.ref_rev_seq_based_on_dist(base, off, lens, len(profiles), *sc, f)
        return [bool(x) for x in f]

    def msa(self, profiles, sc, fraction_called=0.5):
        base, off, lens = self._pack(profiles)
        n = len(profiles)
        cap = n * int(lens.astype(np.int64).sum()) + 16
        rows = C.create_string_buffer(cap)
        seqidx = np.zeros(n, np.uint32)
        dist = np.zeros(n * n, np.int32)
        ncap = int(lens.astype(np.int64).sum()) + 16
        g, cs, q = C.create_string_buffer(ncap), C.create_string_buffer(ncap), C.create_string_buffer(ncap)
        cl, nr = C.c_int(0), C.c_int(0)
        ncol = self.lib.ref_msa(base, off, lens, n, *sc, fraction_called, rows, cap, seqidx, dist, g, cs, q, C.byref(cl), C.byref(nr))
        assert ncol >= 0
        nrow = nr.value
        align = np.frombuffer(rows.raw[: nrow * ncol], np.uint8).reshape(nrow, ncol).copy()
        return dict(rows=align, seqidx=seqidx[:nrow].copy(), dist=dist.reshape(n, n), gapped=g.raw[:ncol], cons=cs.raw[: cl.value], qual=q.raw[: cl.value])

    def assemble_reference(self, profiles, reference, sc, match_fraction=0.5, fraction_called=0.5, inc_ref=False):
        """The DP sequence of the reference-guided branch of assemble() (src/assemble.h:163-292), composed in ref_bridge.cpp from the
        reference's functions. Returns dict(rows, idx, forward, gapped, cons, qual) -- idx / forward per ranked trace."""
        base, off, lens = self._pack(profiles)
        n = len(profiles)
        width = int(lens.astype(np.int64).sum()) + len(reference) + 16
        cap = (n + 1) * width
        rows = C.create_string_buffer(cap)
        idx, fwd = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.uint8)
        g, cs, q = C.create_string_buffer(width), C.create_string_buffer(width), C.create_string_buffer(width)
        nk, cl = C.c_int(0), C.c_int(0)
        u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        _i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        self.lib.ref_assemble_reference.argtypes = [_f32p, _i64p, _i32p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                                    C.c_char_p, C.c_int, _i32p, u8, C.POINTER(C.c_int), C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]
        self.lib.ref_assemble_reference.restype = C.c_int
        ncol = self.lib.ref_assemble_reference(base, off, lens, n, bytes(reference), len(reference), *sc, match_fraction, fraction_called, int(inc_ref), rows, cap,
                                               idx, fwd, C.byref(nk), g, cs, q, C.byref(cl))
        assert ncol >= 0
        k = nk.value
        nrow = k + 1 if k else 0
        return dict(rows=np.frombuffer(rows.raw[: nrow * ncol], np.uint8).reshape(nrow, ncol).copy() if k else np.zeros((0, 0), np.uint8), idx=[int(x) for x in idx[:k]],
                    forward=[bool(x) for x in fwd[:k]], gapped=g.raw[:ncol], cons=cs.raw[: cl.value], qual=q.raw[: cl.value])

    @staticmethod
    def _conv(x):
        if isinstance(x, (bytes, bytearray)):
            return bytes(x), len(x)
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[0] == 6
        return x, x.shape[1]

    @staticmethod
    def _kind(a, b):
        ka = "s" if isinstance(a, (bytes, bytearray)) else "p"
        kb = "s" if isinstance(b, (bytes, bytearray)) else "p"
        return ka + kb

    def gotoh_score(self, a, b, hfree, vfree, sc):
        f = getattr(self.lib, "ref_gotoh_score_" + self._kind(a, b))
        (a, m), (b, n) = self._conv(a), self._conv(b)
        return f(a, m, b, n, hfree, vfree, *sc)

    def gotoh(self, a, b, hfree, vfree, sc):
        """-> (score, row0, row1) exactly as the reference's gotoh() fills `align`."""
        f = getattr(self.lib, "ref_gotoh_" + self._kind(a, b))
        (a, m), (b, n) = self._conv(a), self._conv(b)
        cap = m + n + 1
        r0, r1 = C.create_string_buffer(cap), C.create_string_buffer(cap)
        L = C.c_int(0)
        s = f(a, m, b, n, hfree, vfree, *sc, r0, r1, cap, C.byref(L))
        assert L.value >= 0
        return s, r0.raw[: L.value], r1.raw[: L.value]

    def onehot(self, seq):
        out = np.zeros((6, len(seq)), np.float32)
        self.lib.ref_onehot_profile(bytes(seq), len(seq), out)
        return out

    def revcomp_profile(self, p):
        p = np.ascontiguousarray(p, np.float32)
        out = np.empty_like(p)
        self.lib.ref_revcomp_profile(p, p.shape[1], out)
        return out

    def create_profile(self, acgt, bcpos, primary, secondary, trimleft=0, trimright=0):
        acgt = np.ascontiguousarray(acgt, np.int32)
        bcpos = np.ascontiguousarray(bcpos, np.int32)
        nbc = len(bcpos)
        out = np.zeros((6, nbc), np.float32)
        n = self.lib.ref_create_profile(acgt.reshape(-1), acgt.shape[1], bcpos, bytes(primary), bytes(secondary), nbc, trimleft, trimright,
                                        out.reshape(-1), nbc)
        assert n >= 0
        return np.ascontiguousarray(out.reshape(-1)[: 6 * n].reshape(6, n))

    def find_breakpoint(self, prof):
        prof = np.ascontiguousarray(prof, np.float32)
        a, b, c, d = C.c_int(), C.c_int(), C.c_uint32(), C.c_float()
        self.lib.ref_find_breakpoint(prof, prof.shape[1], C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return bool(a.value), bool(b.value), c.value, d.value

    def decompose_alleles(self, row0, row1, primary, secondary, trim_left, trim_right, maxindel, madc, breakpoint, refslice_len):
        nbc = len(primary)
        pri = C.create_string_buffer(bytes(primary), nbc + 1)
        sec = C.create_string_buffer(bytes(secondary), nbc + 1)
        dcp = np.zeros(2 * 4096, np.int32)
        n = self.lib.ref_decompose_alleles(bytes(row0), bytes(row1), len(row0), pri, sec, nbc, trim_left, trim_right, maxindel, madc,
                                           breakpoint, refslice_len, dcp, 4096)
        return pri.raw[:nbc], sec.raw[:nbc], dcp[: 2 * n].reshape(n, 2).copy()

    def subcommand(self, name, args):
        """Run the reference's own subcommand entry point files-in -> files-out: name in consensus / align / assemble / decompose
        (src/consensus.h:332, src/sage.h:58, src/assemble.h:57, src/indigo.h:42 -- decompose without -v / -a); args = the command line after the subcommand name. Returns its exit code."""
        what = {"consensus": 0, "align": 1, "assemble": 2, "decompose": 3}[name]
        self.lib.ref_subcommand.argtypes = [C.c_int, C.c_char_p]
        self.lib.ref_subcommand.restype = C.c_int
        return self.lib.ref_subcommand(what, "\n".join([name] + [str(a) for a in args]).encode())

    def gt_letter(self, cl6, use_iupac=False):
        """gtLetter (src/consensus.h:94-171): six weights -> (consensus letter, quality)."""
        self.lib.ref_gt_letter.argtypes = [np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), C.c_int, C.c_char_p]
        self.lib.ref_gt_letter.restype = C.c_uint
        out = C.create_string_buffer(2)
        q = self.lib.ref_gt_letter(np.ascontiguousarray(cl6, np.float64), int(use_iupac), out)
        return out.raw[:1], int(q)

    def pairwise_consensus(self, row0, row1, p1, p2, compute_union=True, use_iupac=False):
        """pairwiseConsensus (src/consensus.h:189-238) -> (consensus bytes, uint32 qualities)."""
        self.lib.ref_pairwise_consensus.argtypes = [C.c_char_p, C.c_char_p, C.c_int, _f32p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                                    np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS"), C.c_int]
        self.lib.ref_pairwise_consensus.restype = C.c_int
        p1 = np.ascontiguousarray(p1, np.float32); p2 = np.ascontiguousarray(p2, np.float32)
        cap = 2 * len(row0) + 8
        out = C.create_string_buffer(cap); q = np.zeros(cap, np.uint32)
        k = self.lib.ref_pairwise_consensus(bytes(row0), bytes(row1), len(row0), p1.reshape(-1), p1.shape[1], p2.reshape(-1), p2.shape[1],
                                            int(compute_union), int(use_iupac), out, q, cap)
        assert k >= 0
        return out.raw[:k], q[:k].copy()

    def bench_gotoh_ps(self, profs, seqs, m, n, hfree, vfree, sc, with_traceback=True):
        profs = np.ascontiguousarray(profs, np.float32)
        npairs = profs.shape[0]
        scores = np.zeros(npairs, np.int32)
        cells = self.lib.ref_bench_gotoh_ps(profs.reshape(-1), bytes(seqs), npairs, m, n, hfree, vfree, *sc, int(with_traceback), scores)
        return cells, scores


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        build()
        _port = _Port(os.path.join(_HERE, "libgotoh_oracle.so"))
    return _port


def ref():
    """The reference build, or None where oracle/_ref was never built (it cannot be rebuilt without /root/reference)."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libtracy_ref.so")
        if not os.path.exists(path):
            return None
        _ref = _Ref(path)
    return _ref


def ops_from_rows(row0, row1):
    """Derive the s/h/v string from the reference's gapped rows (src/align.h:281-291)."""
    out = bytearray()
    for x, y in zip(row0, row1):
        out.append(ord("h") if x == 0x2D else (ord("v") if y == 0x2D else ord("s")))
    return bytes(out)
