"""spectrum_analyzer_b200 -- B200 (sm_100a) implementation of spectrum-analyzer's hot path.

`samples_fft_to_spectrum` (window -> real FFT -> magnitude -> FrequencyLimit -> statistics ->
scaling) as one fused CUDA kernel family behind a C ABI (include/spectrum_analyzer_cuda.h),
plus this thin host-side mirror of the reference crate's API.  No CPU fallback: the shared
library must be built (``__graft_entry__.build()``) and a B200 must be present to compute.
"""
from . import _cabi as cabi
from ._build import LIB_PATH, build
from .api import (BatchedSpectrum, Context, CudaError, FrequencyLimit, FrequencyLimitTooNarrow,
                  FrequencySpectrum, InfinityValuesNotSupported, InvalidFrequencyLimit,
                  NaNValuesNotSupported, NonFiniteResult, SamplesLengthNotAPowerOfTwo, ScalingError,
                  SpectrumAnalyzerError, TooFewSamples, UnsupportedLength, default_context,
                  samples_fft_to_spectrum, samples_fft_to_spectrum_batched, scaling, windows)

__version__ = "0.1.0"
