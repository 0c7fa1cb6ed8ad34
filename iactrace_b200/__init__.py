"""iactrace_b200 -- B200-native Monte-Carlo ray tracing for Imaging Atmospheric Cherenkov Telescopes.

Drop-in for the hot path of GerritRo/iactrace (``load_telescope``, ``MCIntegrator``,
``Telescope.__call__(sources, values, source_type)``, ``render_response_matrix``, the YAML configs),
backed by hand-written sm_100a CUDA kernels behind a C ABI (``include/iactrace_b200.h``).
There is no CPU fallback.
"""
from .telescope import Telescope, Mirror
from .core import Integrator, MCIntegrator
from .sensors import SquareSensor, HexagonalSensor
from .viz import hexshow, squareshow
from .io import load_telescope
from . import random

__version__ = "0.1.0"

__all__ = ["Telescope", "Mirror", "Integrator", "MCIntegrator", "SquareSensor", "HexagonalSensor",
           "hexshow", "squareshow", "load_telescope", "random"]
