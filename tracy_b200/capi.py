"""ctypes binding of include/tracy_b200.h. Loads the in-tree libtracy_b200.so and fails loudly if it is missing:
there is no Python or CPU implementation of the hot path behind this module."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TRACY_B200_LIB") or os.path.join(HERE, "libtracy_b200.so")   # override: kernel variant experiments

TB_OK, TB_ERR_INVALID, TB_ERR_CUDA, TB_ERR_NOMEM, TB_ERR_UNSUPPORTED = range(5)
TB_MEM_HOST, TB_MEM_DEVICE = 0, 1
TB_A1_TRACE_PROFILES = 0x100

# every symbol include/tracy_b200.h declares (tests/test_boundary.py checks the header against this list)
SYMBOLS = [
    "tb_ctx_create", "tb_ctx_destroy", "tb_strerror", "tb_last_error", "tb_host_alloc", "tb_host_free",
    "tb_ctx_set_scratch_limit", "tb_ctx_stats", "tb_ctx_last_kernel_ms", "tb_ctx_last_call_ms", "tb_ctx_last_packed_pairs", "tb_ctx_last_big_pairs", "tb_gotoh_ps", "tb_gotoh_pp", "tb_gotoh_ss",
    "tb_rows_from_ops", "tb_unpack_ops", "tb_decompose_sweep", "tb_version", "tb_create_profile", "tb_revcomp_profile", "tb_trim_reference_slice",
    "tb_find_breakpoint", "tb_basecall", "tb_index_build", "tb_index_destroy", "tb_index_info", "tb_anchor", "tb_ctx_last_anchor_ms",
    "tb_reference_slice", "tb_allelic_fraction", "tb_ctx_last_fraction_ms", "tb_trace_scan", "tb_trace_unpack",
    "tb_trace_set_create", "tb_trace_set_destroy", "tb_trace_set_info",
    "tb_multi_create", "tb_multi_destroy", "tb_multi_size", "tb_multi_ctx", "tb_multi_last_error", "tb_multi_partition", "tb_multi_gotoh",
    "tb_multi_index_build", "tb_multi_anchor",
    "tb_write_trace_txt", "tb_write_align_fasta", "tb_write_plot_alignment", "tb_write_trace_align_json", "tb_write_align_files", "tb_trace_quality",
    "tb_write_decompose_json", "tb_write_decomposition", "tb_write_assemble_files", "tb_pairwise_consensus",
]


class Score(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32)]


class AlignConfig(C.Structure):
    _fields_ = [("h_free", C.c_int32), ("v_free", C.c_int32)]


class Arena(C.Structure):
    _fields_ = [("base", C.c_void_p), ("off", C.c_void_p), ("len", C.c_void_p)]


class Batch(C.Structure):
    _fields_ = [("a1", Arena), ("a2", Arena), ("npairs", C.c_size_t), ("mem", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("scores", C.c_void_p), ("ops", C.c_void_p), ("ops_stride", C.c_int64), ("ops_len", C.c_void_p),
                ("row0", C.c_void_p), ("row1", C.c_void_p), ("rows_stride", C.c_int64), ("ops_packed", C.c_int32)]


class SweepBatch(C.Structure):
    _fields_ = [("refrow", Arena), ("primary", Arena), ("secondary_base", C.c_void_p), ("vi_end", C.c_void_p),
                ("align_index", C.c_void_p), ("var_index", C.c_void_p), ("ndel", C.c_void_p), ("nins", C.c_void_p),
                ("ntraces", C.c_size_t), ("mem", C.c_int32)]


class SweepResult(C.Structure):
    _fields_ = [("fref", C.c_void_p), ("fins", C.c_void_p), ("out_stride", C.c_int32), ("grid", C.c_void_p)]


class ProfileBatch(C.Structure):
    _fields_ = [("trace", Arena), ("bcpos", Arena), ("primary_base", C.c_void_p), ("secondary_base", C.c_void_p),
                ("trim_left", C.c_void_p), ("trim_right", C.c_void_p), ("ntraces", C.c_size_t), ("mem", C.c_int32)]


class BasecallBatch(C.Structure):
    _fields_ = [("trace", Arena), ("ploc", Arena), ("ntraces", C.c_size_t), ("mem", C.c_int32)]


class TraceView(C.Structure):
    _fields_ = [("acgt", C.c_void_p), ("nsamples", C.c_int32), ("bcpos", C.c_void_p), ("qual", C.c_void_p), ("primary", C.c_void_p),
                ("secondary", C.c_void_p), ("consensus", C.c_void_p), ("nbc", C.c_int32)]


class AssembleTrace(C.Structure):
    _fields_ = [("name", C.c_char_p), ("forward", C.c_int32), ("row", C.c_int32), ("trace", TraceView), ("trim_left", C.c_int32), ("trim_right", C.c_int32)]


class DecomposeJson(C.Structure):
    _fields_ = [("trim_left", C.c_int32), ("trim_right", C.c_int32), ("pratio", C.c_float), ("genome", C.c_char_p), ("input", C.c_char_p),
                ("viewport_basecall", C.c_int32),
                ("chr1", C.c_char_p), ("pos1", C.c_uint32), ("alt1", C.c_char_p), ("ref1", C.c_char_p), ("L1", C.c_int32), ("forward1", C.c_int32), ("score1", C.c_int32),
                ("chr2", C.c_char_p), ("pos2", C.c_uint32), ("alt2", C.c_char_p), ("ref2", C.c_char_p), ("L2", C.c_int32), ("forward2", C.c_int32), ("score2", C.c_int32),
                ("a1", C.c_double), ("a2", C.c_double), ("a3row0", C.c_char_p), ("a3row1", C.c_char_p), ("L3", C.c_int32), ("score3", C.c_int32),
                ("hetindel", C.c_int32), ("decomp", C.c_void_p), ("ndecomp", C.c_int32)]


class AnchorConfig(C.Structure):
    _fields_ = [("trim_left", C.c_int32), ("trim_right", C.c_int32), ("kmer", C.c_int32), ("min_kmer_support", C.c_int32)]


class AnchorResult(C.Structure):
    _fields_ = [("anchored", C.c_void_p), ("forward", C.c_void_p), ("kmersupport", C.c_void_p), ("bestpos", C.c_void_p), ("pass_", C.c_void_p)]


class FractionBatch(C.Structure):
    _fields_ = [("trace", Arena), ("bcpos", Arena), ("primary_base", C.c_void_p), ("secdecompose_base", C.c_void_p),
                ("trim_left", C.c_int32), ("trim_right", C.c_int32), ("ntraces", C.c_size_t), ("mem", C.c_int32)]


class TraceInfo(C.Structure):
    _fields_ = [("format", C.c_int32), ("ok", C.c_int32), ("status", C.c_int32), ("nsamples", C.c_int32), ("nbasecalls", C.c_int32)]


class LibraryMissing(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded shared library. Raises LibraryMissing when it has not been built (python -m tracy_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(f"{LIB_PATH} not found: build it with `python -m tracy_b200.build` (nvcc, sm_100a). "
                             "tracy_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.tb_ctx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.tb_ctx_destroy.argtypes = [vp]
    L.tb_ctx_destroy.restype = None
    L.tb_strerror.argtypes = [C.c_int]
    L.tb_strerror.restype = C.c_char_p
    L.tb_last_error.argtypes = [vp]
    L.tb_last_error.restype = C.c_char_p
    L.tb_host_alloc.argtypes = [vp, C.POINTER(vp), C.c_size_t]
    L.tb_host_free.argtypes = [vp, vp]
    L.tb_ctx_set_scratch_limit.argtypes = [vp, C.c_size_t]
    L.tb_trace_quality.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_float, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.tb_write_decompose_json.argtypes = [C.c_char_p, C.POINTER(TraceView), C.POINTER(DecomposeJson)]
    L.tb_write_decomposition.argtypes = [C.c_char_p, C.c_void_p, C.c_int32]
    L.tb_write_assemble_files.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(AssembleTrace), C.c_int32, C.c_char_p, C.c_char_p, C.c_char_p,
                                          C.c_int32, C.c_int32, C.c_int32]
    L.tb_pairwise_consensus.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_char_p,
                                        C.c_void_p, C.POINTER(C.c_int32)]
    L.tb_write_trace_txt.argtypes = [C.c_char_p, C.POINTER(TraceView), C.c_int32, C.c_int32]
    L.tb_write_align_files.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(TraceView), C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p,
                                       C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.tb_write_plot_alignment.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, C.c_uint32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_double, C.c_double, C.c_int32]
    L.tb_write_trace_align_json.argtypes = [C.c_char_p, C.POINTER(TraceView), C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, C.c_uint32, C.c_int32]
    L.tb_write_align_fasta.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]
    L.tb_ctx_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.tb_ctx_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.tb_ctx_last_call_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.tb_ctx_last_packed_pairs.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.tb_ctx_last_big_pairs.argtypes = [vp, C.POINTER(C.c_uint64)]
    for name in ("tb_gotoh_ps", "tb_gotoh_pp", "tb_gotoh_ss"):
        getattr(L, name).argtypes = [vp, C.POINTER(Batch), Score, AlignConfig, C.POINTER(Result)]
    L.tb_rows_from_ops.argtypes = [C.c_int, vp, C.c_int32, vp, C.c_int32, vp, C.c_int32, vp, vp]
    L.tb_unpack_ops.argtypes = [vp, C.c_int32, vp]
    L.tb_decompose_sweep.argtypes = [vp, C.POINTER(SweepBatch), C.POINTER(SweepResult)]
    L.tb_version.restype = C.c_char_p
    L.tb_create_profile.argtypes = [vp, C.POINTER(ProfileBatch), vp, vp, vp]
    L.tb_basecall.argtypes = [vp, C.POINTER(BasecallBatch), C.c_float, vp, vp, vp, vp, vp, vp]
    L.tb_revcomp_profile.argtypes = [vp, C.POINTER(Arena), C.c_size_t, C.c_int32, vp, vp]
    L.tb_trim_reference_slice.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_int32, C.c_int32,
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)]
    L.tb_find_breakpoint.argtypes = [vp, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
    L.tb_index_build.argtypes = [vp, vp, C.c_int64, C.c_int32, C.POINTER(vp)]
    L.tb_index_destroy.argtypes = [vp, vp]
    L.tb_index_info.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.POINTER(vp)]
    L.tb_anchor.argtypes = [vp, vp, C.POINTER(Arena), C.c_size_t, C.c_int32, AnchorConfig, C.POINTER(AnchorResult)]
    L.tb_ctx_last_anchor_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.tb_reference_slice.argtypes = [C.c_int64, vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.tb_allelic_fraction.argtypes = [vp, C.POINTER(FractionBatch), vp, vp]
    L.tb_ctx_last_fraction_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.tb_trace_scan.argtypes = [vp, vp, vp, C.c_size_t, vp]
    L.tb_trace_unpack.argtypes = [vp, vp, vp, vp, C.c_size_t, C.c_int32, vp, vp, vp, vp, vp, vp, vp]
    _lib = L
    return L
