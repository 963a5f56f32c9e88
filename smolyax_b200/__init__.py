"""smolyax_b200 — B200-native evaluation of barycentric Smolyak interpolants.

Drop-in for the evaluation path of JoWestermann/smolyax: ``SmolyakBarycentricInterpolator`` with ``__call__``,
``gradient`` and ``integral``, the ``nodes`` generators and the ``indices`` helpers.  The numerical work runs in
hand-written sm_100a CUDA kernels behind the C-ABI of ``include/smolyax_b200.h``; there is no CPU fallback.
"""
from . import indices, nodes  # noqa: F401

__version__ = "0.1.0"
__all__ = ["indices", "nodes", "barycentric", "interpolation", "SmolyakBarycentricInterpolator"]


def __getattr__(name):  # lazy: importing the evaluation modules loads the CUDA library
    if name in ("barycentric", "interpolation"):
        import importlib

        return importlib.import_module(f"{__name__}.{name}")
    if name == "SmolyakBarycentricInterpolator":
        from .interpolation import SmolyakBarycentricInterpolator

        return SmolyakBarycentricInterpolator
    raise AttributeError(name)
