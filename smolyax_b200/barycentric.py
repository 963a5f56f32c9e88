"""
Barycentric building blocks on the GPU — same names and argument meaning as the reference's
``smolyax.barycentric`` (/root/reference/src/smolyax/barycentric.py:13,34,69,126,158).

Every function runs a CUDA kernel of ``libsmolyax_b200.so`` on the current device; inputs may be NumPy arrays
(copied in, NumPy result) or CUDA ``torch`` tensors (result stays on the device, asynchronous on the current
stream).  The interpolator itself does not call these one summand at a time — it uses the fused handle path — but
they are the reference's own seam (interpolation.py:243-248) and what the parity tests exercise summand by summand.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np
import torch

from . import _lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _to_device(a, dtype=torch.float64):
    """Return (cuda tensor, was_torch_cuda)."""
    if isinstance(a, torch.Tensor):
        if a.is_cuda:
            return a.to(dtype).contiguous(), True
        return a.to(dtype).contiguous().cuda(), False
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype={torch.float64: np.float64, torch.int64: np.int64}[dtype]))
    return t.cuda(), False


def _back(t, keep_on_device):
    return t if keep_on_device else t.cpu().numpy()


def compute_weights(nodes):
    r"""Barycentric weights :math:`w_j = \prod_{i \ne j} 1 / (\xi_i - \xi_j)` of the given nodes, shape ``(n,)``
    (reference barycentric.py:13-31; note the sign convention, it cancels in the normalised basis)."""
    d_nodes, on_dev = _to_device(nodes)
    w = torch.empty_like(d_nodes)
    _lib.check(_lib.lib.smx_compute_weights(d_nodes.data_ptr(), d_nodes.numel(), w.data_ptr(), _stream()), "compute_weights")
    return _back(w, on_dev)


def _basis(x, xi, w, nu_i, derivative):
    d_x, on_dev = _to_device(x)
    d_xi, _ = _to_device(xi)
    d_w, _ = _to_device(w)
    n_points = d_x.numel()
    out = torch.empty((n_points, d_xi.numel()), dtype=torch.float64, device=d_x.device)
    _lib.check(_lib.lib.smx_basis(d_x.data_ptr(), n_points, d_xi.data_ptr(), d_w.data_ptr(), d_xi.numel(), int(nu_i),
                                  derivative, out.data_ptr(), _stream()), "basis")
    return _back(out, on_dev)


def evaluate_basis_unnormalized(x, xi, w, nu_i: int):
    r"""Numerators :math:`w_j / (x - \xi_j)` for ``j <= nu_i`` (other columns zero); rows whose point sits on a
    node become the one-hot pattern of that node.  ``x``: ``(n_points, 1)``; result ``(n_points, m_i)``
    (reference barycentric.py:34-66)."""
    return _basis(x, xi, w, nu_i, 0)


def evaluate_basis_gradient_unnormalized(x, xi, w, nu_i: int):
    r"""Derivative numerators :math:`-w_j / (x - \xi_j)^2` for ``j <= nu_i``; ``NaN`` where the point sits on a
    node (reference barycentric.py:126-155)."""
    return _basis(x, xi, w, nu_i, 1)


def _one_summand(x, F, xi_list, w_list, sorted_dims, sorted_degs, zeta, gradient):
    d_x, on_dev = _to_device(x)
    if d_x.dim() == 1:
        d_x = d_x[None, :]
    n_points, d_in = d_x.shape
    d_F, _ = _to_device(F)
    d_out, n = int(d_F.shape[0]), d_F.dim() - 1
    layout = {
        "F_%d" % n: d_F[None],
        "nodes_%d" % n: _to_device(xi_list)[0][None],
        "weights_%d" % n: _to_device(w_list)[0][None],
        "dims_%d" % n: _to_device(np.asarray(sorted_dims, dtype=np.int64), torch.int64)[0][None],
        "degs_%d" % n: _to_device(np.asarray(sorted_degs, dtype=np.int64), torch.int64)[0][None],
        "zetas_%d" % n: _to_device(np.asarray([zeta], dtype=np.int64), torch.int64)[0],
    }
    assert layout["nodes_%d" % n].shape[-1] == max(d_F.shape[1:]), "xi_list must have length max(F.shape[1:])"
    arr, _, keep = _lib.pack_groups(layout, lambda a, dt: (a, a.data_ptr()))
    if gradient:
        out = torch.empty((n_points, d_out, d_in), dtype=torch.float64, device=d_x.device)
        fn = _lib.lib.smx_group_gradient
    else:
        out = torch.empty((n_points, d_out), dtype=torch.float64, device=d_x.device)
        fn = _lib.lib.smx_group_eval
    _lib.check(fn(d_x.data_ptr(), n_points, d_x.stride(0) if n_points > 1 else d_in, d_in, arr, d_out, out.data_ptr(), 0, _stream()),
               "evaluate_tensor_product")
    del keep
    return _back(out, on_dev)


def evaluate_tensor_product_interpolant(x, F, xi_list, w_list, sorted_dims: Sequence[int], sorted_degs: Sequence[int],
                                        zeta: int):
    """``zeta`` times the tensor-product interpolant of one summand at the points ``x``: ``(n_points, d_out)``.
    ``F``: ``(d_out, mu_1, .., mu_n)`` zero padded; ``xi_list``/``w_list``: ``(n, max mu)``
    (reference barycentric.py:69-123)."""
    return _one_summand(x, F, xi_list, w_list, sorted_dims, sorted_degs, zeta, gradient=False)


def evaluate_tensor_product_gradient(x, F, xi_list, w_list, sorted_dims: Sequence[int], sorted_degs: Sequence[int],
                                     zeta: int):
    """``zeta`` times the gradient of the tensor-product interpolant of one summand: dense ``(n_points, d_out, d_in)``,
    zero outside ``sorted_dims``, ``NaN`` in a dimension whose coordinate sits on a node
    (reference barycentric.py:158-229)."""
    return _one_summand(x, F, xi_list, w_list, sorted_dims, sorted_degs, zeta, gradient=True)
