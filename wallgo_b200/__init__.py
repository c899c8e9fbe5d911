"""wallgo_b200 -- B200-native (sm_100a) Boltzmann-transport hot path behind the WallGo API.

Public surface:
  BoltzmannSolver            standalone class with the reference's interface (boltzmann.py)
  dropin.install()           subclass of WallGo.BoltzmannSolver for unmodified WallGoManager/EOM
  Grid, Grid3Scales, Particle, BoltzmannBackground, CollisionArray, ETruncationOption
  engine.BoltzmannEngine     thin torch/ctypes wrapper over the C ABI (include/wallgo_b200.h)
"""
import os as _os

# Independent systems share one GPU through one stream (CUDA graph) per slot; the default of 8 hardware
# work queues would serialise more than 8 of them.  Only effective if set before the CUDA context exists.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .boltzmann import BoltzmannSolver  # noqa: E402
from .containers import (BoltzmannBackground, BoltzmannDeltas, BoltzmannResults, CollisionArray,
                         CollisionLoadError, ETruncationOption, Particle, Polynomial)
from .grid import Grid, Grid3Scales

__all__ = ["BoltzmannSolver", "BoltzmannBackground", "BoltzmannDeltas", "BoltzmannResults",
           "CollisionArray", "CollisionLoadError", "ETruncationOption", "Particle", "Polynomial", "Grid", "Grid3Scales"]
__version__ = "0.1.0"
