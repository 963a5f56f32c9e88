"""One-dimensional node families for sparse-grid interpolation and quadrature.

Same classes and behaviour as the reference's ``smolyax.nodes`` package
(/root/reference/src/smolyax/nodes/__init__.py:1-5).  Host-side NumPy only: the tables are small and must
match the reference bit for bit (Gauss-Hermite goes through LAPACK inside ``hermgauss``).
"""
from .core import Generator, Generator1D
from .families import GaussHermite, GaussHermite1D, Leja, Leja1D

__all__ = ["Generator1D", "Generator", "Leja1D", "Leja", "GaussHermite1D", "GaussHermite"]
