"""radiorust_b200 -- B200-native (sm_100a) IQ sample chain behind radiorust's block API.

Only the hot path of JanBeh/radiorust named in BASELINE.json lives here:
FreqShifter -> Filter -> Downsampler (+ Upsampler, FmDemod, de-emphasis,
GainControl, Fourier, FmMod, Rechunker, Overlapper), as hand-written CUDA kernels (``csrc/``) behind the C ABI of
``include/radiorust_b200.h``.  This package is the ctypes face of that ABI.
"""
from .chain import (  # noqa: F401
    Chain,
    Context,
    Downsampler,
    Filter,
    FmDemod,
    FmMod,
    Fourier,
    FreqShifter,
    GainControl,
    Overlapper,
    PinnedChunkBufPool,
    Rechunker,
    Upsampler,
    bandwidth,
    kernel_launch_count,
    level,
    rescale_energy,
)
from ._ffi import RadiorustError  # noqa: F401
