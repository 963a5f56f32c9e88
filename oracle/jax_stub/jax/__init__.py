"""NumPy stand-in for the handful of JAX symbols the reference imports.

TEST INFRASTRUCTURE ONLY.  With this directory placed before /root/reference/src on
PYTHONPATH the reference's files run unmodified.  `jit` is
the identity, `vmap` is a Python loop, arrays are ndarray views.  Reduction order is
NumPy's, not XLA's, so agreement with a real JAX run is up to fp64 rounding only.
"""
import numpy as _np

from . import numpy, random  # noqa: F401  (submodules, as in `import jax.numpy as jnp`)
from .numpy import A as Array  # noqa: F401


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def jit(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


def vmap(f, in_axes=0):
    def g(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(len(a) for a, ax in zip(args, axes) if ax is not None)
        outs = [f(*[a if ax is None else a[i] for a, ax in zip(args, axes)]) for i in range(n)]
        return numpy.A._wrap(_np.stack(outs))

    return g
