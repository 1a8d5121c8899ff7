"""bulletproofspp_b200 -- B200-native (sm_100a) Bulletproofs++ argument hot path.

Python is only a thin ctypes binding over the C ABI (include/bppp_b200.h); the product is the
CUDA library built from csrc/.  There is no CPU fallback: importing works anywhere, but every
compute call needs the built library and a CUDA device and fails loudly otherwise.
"""
from .lib import (BpppError, Context, NormLinearArgument, load_library, library_path, int_to_le, le_to_int,
                  point_to_bytes, bytes_to_point, ARG_NL, ARG_IP, RangeProofSetup)

__all__ = ["BpppError", "Context", "NormLinearArgument", "load_library", "library_path", "int_to_le", "le_to_int",
           "point_to_bytes", "bytes_to_point", "ARG_NL", "ARG_IP", "RangeProofSetup"]
