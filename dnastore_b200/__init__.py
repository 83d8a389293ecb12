"""dnastore_b200 -- B200-native batched Viterbi decoder for ihh/dnastore's hot path.

The product is the C-ABI shared library ``libdnastore_b200.so`` (CUDA sm_100a
kernels + C++ host, see ``include/dnastore_b200.h``).  This package is only the
thin ctypes binding tests and ``bench.py`` call it through; it contains no
compute and NO CPU fallback: importing it without the built library raises, and
creating a decoder without a CUDA device raises.
"""
from ._capi import (  # noqa: F401
    DnabError,
    Machine,
    ErrorFlags,
    Compiled,
    Decoder,
    MultiDecoder,
    Tables,
    lib,
    lib_path,
    pack_reads,
    MutatorParams,
    MutatorCounts,
    PairDb,
    pairhmm_fb_batch,
    pairhmm_set_chunk_cells,
    expected_counts,
    baum_welch,
    ExactDecoder,
    exact_decode_bits,
    exact_decode_string,
    exact_decode_fasta,
    pack_decoded_symbols,
    READ_OK,
    READ_NO_DECODING,
    READ_OVERFLOW,
    READ_TRACEBACK_FAILED,
)

__all__ = [
    "DnabError", "Machine", "ErrorFlags", "Compiled", "Decoder", "MultiDecoder", "Tables", "lib", "lib_path", "pack_reads", "MutatorParams", "MutatorCounts", "PairDb",
    "pairhmm_fb_batch", "pairhmm_set_chunk_cells", "expected_counts", "baum_welch",
    "ExactDecoder", "exact_decode_bits", "exact_decode_string", "exact_decode_fasta", "pack_decoded_symbols",
    "READ_OK", "READ_NO_DECODING", "READ_OVERFLOW", "READ_TRACEBACK_FAILED",
]
