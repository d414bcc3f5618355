"""B200-native 3D FFT convolution: drop-in for StephanPreibisch/FourierConvolutionCUDALib's hot path.

The product is the C-ABI shared library built from csrc/ (see include/convolution3Dfft.h); this
package is the thin host-side mirror used by the tests and the benchmark.
"""
from . import _lib  # noqa: F401
from .api import *  # noqa: F401,F403
