"""
bajes_b200 -- B200-native batched gravitational-wave likelihood: a drop-in for the likelihood hot
path of matteobreschi/bajes (``GWLikelihood.log_like`` with ``Waveform``, ``Detector``, ``Noise``,
``Series``) plus a batched entry for whole ensembles of walkers / live points.

The compute path is hand-written CUDA for sm_100a behind a C ABI (``include/bajes_b200.h``,
``bajes_b200/csrc``); this package is the host-side mirror of the reference's object API.
There is no CPU fallback.
"""
from .detector import Detector
from .engine import EngineError, EngineUnavailable
from .likelihood import GWLikelihood, Likelihood, Posterior
from .noise import Noise
from .series import Series
from .waveform import PolarizationTuple, Waveform

__version__ = "0.1.0"

__all__ = ["Detector", "Noise", "Series", "Waveform", "PolarizationTuple", "GWLikelihood", "Likelihood",
           "Posterior", "EngineError", "EngineUnavailable"]
