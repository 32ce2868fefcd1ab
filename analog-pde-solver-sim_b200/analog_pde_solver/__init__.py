"""analog_pde_solver -- drop-in B200 backend for danieleschmidt/analog-pde-solver-sim.

Same import surface as the reference package (analog_pde_solver/__init__.py:8-11 upstream):
``AnalogPDESolver``, ``DigitalSolver``, ``EnergyModel`` and ``__version__``.  The relaxation
sweeps run as hand-written sm_100a CUDA kernels in ``libapde.so``; there is no CPU fallback.
"""
from .solver import AnalogPDESolver, DigitalSolver, EnergyModel

__all__ = ["AnalogPDESolver", "DigitalSolver", "EnergyModel"]
__version__ = "1.0.0"
