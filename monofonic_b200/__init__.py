"""monofonic_b200 - B200-native LPT field pipeline of monofonIC (MUSIC2).

Python host mirror of the reference's Grid_FFT / OrszagConvolver / op:: /
particle::lattice_generator interfaces on top of the C ABI of libmfb200.so
(include/mfb200.h).  All numerics run in hand-written sm_100a CUDA kernels;
there is no CPU fallback.
"""
from ._lib import MfbError, load  # noqa: F401
from .grid import (Grid_FFT, Grid_FFT_complex, OrszagConvolver, NaiveConvolver, Particles, init, finalize, kspace_id, rspace_id,  # noqa: F401
                   synchronize, kernel_launches, op, lattice_gradient, particle_type_mask)
from .pipeline import LPTPipeline, lpt_config  # noqa: F401

__all__ = ["Grid_FFT", "Grid_FFT_complex", "OrszagConvolver", "NaiveConvolver", "Particles", "LPTPipeline", "lpt_config", "init", "finalize",
           "synchronize", "kernel_launches", "kspace_id", "rspace_id", "op", "MfbError", "load", "lattice_gradient", "particle_type_mask"]
