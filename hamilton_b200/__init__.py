"""hamilton_b200 — B200-native batched Hamiltonian dynamics behind Numeric.Hamilton's API.

The package is a thin mirror of the reference's Haskell module over a C ABI
(include/hamilton_b200.h) implemented in hand-written sm_100a CUDA (hamilton_b200/csrc).
"""
from . import num                                                        # noqa: F401
from ._lib import (AOS, SOA, RK4, RKF45_GSL, FLAG_NOT_SPD, FLAG_NONFINITE, FLAG_STEP_FAILED,     # noqa: F401
                   HamiltonError, NoDeviceError, NumericError)
from .api import (System, Config, Phase, Cfg, Phs, mkSystem, mkSystem_, underlyingPos, pe, momenta, toPhase, keC,   # noqa: F401
                  lagrangian, velocities, fromPhase, keP, hamiltonian, hamEqs, stepHam, evolveHam, evolveHam_,
                  stepHamC, evolveHamC, evolveHamC_)
from . import systems                                                    # noqa: F401
from . import ensemble                                                   # noqa: F401

__all__ = [n for n in dir() if not n.startswith("_")]
