"""awry_b200 -- B200-native batched FM-index search (count / locate), drop-in for the query
side of the Rust crate `awry` (awry::fm_index::FmIndex).

The compute path is the hand-written sm_100a CUDA library awry_b200/libawry_b200.so behind the
C ABI of include/awry_b200.h.  There is no CPU fallback: importing this package without the
built library raises, and every search call fails when no B200 is present.
"""
from .fm_index import (AwryError, FmBuildArgs, FmIndex, LocalizedSequencePosition, SearchRange, Symbol,
                       SymbolAlphabet, library_path, native)

__all__ = ["AwryError", "FmBuildArgs", "FmIndex", "LocalizedSequencePosition", "SearchRange", "Symbol",
           "SymbolAlphabet", "library_path", "native"]
