"""jpeglibrary_b200 -- B200-native hot path behind yigolden/JpegLibrary's API.

Layout: csrc/ (sm_100a kernels + the C-ABI shim, include/jpegb200.h), host/ (marker walk and
the C++ mirror of the reference API), api.py (Python mirror used by the tests and bench.py).
Importing the package loads the native libraries; there is no fallback implementation.
"""
from . import _native
from ._native import (JB_IN_COEFFICIENTS, JB_IN_GRAY8, JB_IN_RGB24, JB_IN_YCBCR888, JB_OUT_COEFFICIENTS, JB_OUT_PLANAR_I16, JB_OUT_RGB24, JB_OUT_RGBA32,
                      JB_OUT_YCBCR888)
from .api import (ArgumentException, Context, CudaInputReader, CudaOutputWriter, CudaRuntimeError,
                  InvalidDataException, InvalidOperationException, JpegBatchDecoder, JpegBatchEncoder, JpegBatchOptimizer, JpegPipelinedBatchDecoder,
                  JpegBlockInputReader, JpegBlockOutputWriter, JpegDecoder, JpegEncoder, JpegOptimizer, JpegQuantizationTable,
                  JpegStandardQuantizationTable, NotSupportedException, Parsed, decode_coefficients, encode_rgb)

__all__ = [n for n in dir() if not n.startswith("_")]
