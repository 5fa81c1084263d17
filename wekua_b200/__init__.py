"""wekua_b200 -- B200-native (sm_100a) backend for wekua's dense-BLAS hot path.

Layout mirrors src/wekua.zig:1-16: core (Context / CommandQueue / Pipeline), tensor, blas, math, nn.
Everything numerical runs in libwekua_b200.so (hand-written CUDA) through the C ABI in include/wekua_b200.h;
this package is the host-side mirror of the reference's Zig API used by tests, bench.py and examples.
"""
from . import capi  # noqa: F401  (fails loudly when the CUDA library has not been built)
from . import blas, core, math, nn, tensor  # noqa: F401
from .core import CommandQueue, Context, Pipeline  # noqa: F401
from .tensor import Tensor  # noqa: F401
