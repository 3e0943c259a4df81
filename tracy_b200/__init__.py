"""tracy_b200 -- B200-native hot path of gear-genomics/tracy (Gotoh profile alignment + decompose sweeps).

The package is a thin host layer over libtracy_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/tracy_b200.h). Importing it never touches the GPU; creating a Context does and raises if no B200 or no
built library is present -- there is no CPU implementation in this package.
"""
from .api import (PS, PP, SS, AlignConfig, Arena, Context, DnaScore, MultiContext, TracyError, default_context, gotoh, gotohScore,
                  find_breakpoint, pack_profiles, pack_seqs, rows_from_ops, trim_reference_slice, unpack_ops, uniform_profiles, uniform_seqs)

__all__ = ["PS", "PP", "SS", "AlignConfig", "Arena", "Context", "MultiContext", "DnaScore", "TracyError", "default_context", "gotoh",
           "gotohScore", "find_breakpoint", "trim_reference_slice", "pack_profiles", "pack_seqs", "rows_from_ops", "unpack_ops", "uniform_profiles", "uniform_seqs"]
