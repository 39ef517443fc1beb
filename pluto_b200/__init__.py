"""pluto_b200 -- B200-native unsplit Godunov MHD step behind PLUTO's AdvanceStep seam.

The product is the C-ABI shared library ``pluto_b200/lib/libpluto_gpu.so``
(``include/pluto_gpu.h``).  This package is its thin Python host: a ctypes
binding (no torch types cross the boundary), the time loop of the reference's
``main()`` / ``Integrate()`` / ``NextTimeStep()`` around it, the block
decomposition + halo exchange that replaces ArrayLib, and synthetic initial
conditions of the BASELINE.json shapes for benchmarking.

There is no CPU path: importing works anywhere, but creating a stepper
without the built library or without a CUDA device raises.
"""
from ._lib import load_library, LIB_PATH, PlutoGpuConfig, PlutoGpuStepInfo  # noqa: F401
from .stepper import GpuStepper, MultiGpuStepper, Integrator, StepInfo  # noqa: F401
