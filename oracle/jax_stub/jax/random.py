"""`jax.random` subset (only used by the reference's optional warm-up call)."""
import numpy as _np

from .numpy import A


def PRNGKey(seed):
    return seed


def uniform(key, shape):
    return A._wrap(_np.random.default_rng(key).uniform(size=shape))
