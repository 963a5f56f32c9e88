"""`jax.numpy` subset used by the reference, on top of NumPy (test infrastructure)."""
import numpy as _np

nan = _np.nan


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Setter:
            def set(self, v):
                out = _np.array(arr, copy=True)
                out[idx] = v
                return A._wrap(out)

        return _Setter()


class A(_np.ndarray):
    """ndarray view with the few jax.Array extras the reference touches."""

    @staticmethod
    def _wrap(x):
        return _np.asarray(x).view(A)

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self

    def __iadd__(self, other):  # jax arrays are immutable: `a += b` rebinds and promotes
        return A._wrap(_np.add(_np.asarray(self), _np.asarray(other)))


def _w(fn):
    def g(*a, **k):
        return A._wrap(fn(*a, **k))

    return g


asarray = _w(_np.asarray)
array = _w(_np.array)
zeros = _w(_np.zeros)
arange = _w(_np.arange)
where = _w(_np.where)
prod = _w(_np.prod)
any = _w(_np.any)
sum = _w(_np.sum)
tile = _w(_np.tile)
einsum = _w(_np.einsum)


def divide(a, b):
    with _np.errstate(all="ignore"):
        return A._wrap(_np.divide(a, b))


def broadcast_to(x, shape):
    return A._wrap(_np.array(_np.broadcast_to(x, shape)))
