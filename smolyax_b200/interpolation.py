r"""
``SmolyakBarycentricInterpolator`` — drop-in for the evaluation path of the reference class of the same name
(/root/reference/src/smolyax/interpolation.py:15-390): the Smolyak operator

.. math::  I^{\Lambda}[f] = \sum_{\nu \in \Lambda_{k,t}} \zeta_{\Lambda,\nu}\, I^{\nu}[f]

for vector-valued :math:`f : \mathbb R^{d_{in}} \to \mathbb R^{d_{out}}`, evaluated (``__call__``), differentiated
(``gradient``) and integrated (``integral``) on a B200.

What is the same as the reference: constructor keywords, ``set_f`` and its ``f_evals`` reuse dictionary, the
counters ``n_f_evals`` / ``n_f_evals_new``, argument shapes, result shapes, assertion behaviour, the per-group
tables (``reference_layout()`` returns exactly the six arrays of interpolation.py:230-235, bit for bit).
What is different: the tables go through ``smx_create`` into the device layout of DESIGN.md and every call is a
handful of CUDA kernel launches instead of a Python loop of XLA dispatches.  There is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Sequence

import numpy as np
import torch

from . import indices, nodes  # noqa: F401


class _Native:
    """``libsmolyax_b200.so`` (ctypes binding ``_lib``), loaded at the first use of a device entry point.  The host half of
    ``set_f`` (``_assemble``: tables as NumPy arrays) needs no GPU and no CUDA library - the reference arm of ``bench.py``
    and the golden generator use it that way.  Every compute path goes through here and fails loudly if the library
    cannot be loaded or built: there is no CPU fallback."""

    def __getattr__(self, name):
        from . import _lib as real

        return getattr(real, name)


_lib = _Native()

_CODE = 1 << 20  # (dim, deg) -> dim * _CODE + deg
_GRAD_COLUMN_BLOCK = 1024  # most outputs per set of derivative tables when one handle cannot hold them for every output


def _host_weights(pts: np.ndarray) -> np.ndarray:
    """Barycentric weights on the host, same expression as reference barycentric.py:28-31 (setup only)."""
    diffs = pts[:, None] - pts
    diffs = np.where(diffs == 0, 1, diffs)
    return np.prod(1 / diffs, axis=0)


def write_layout(layout: dict, path, *, d_in: int, d_out: int, k, t: float) -> None:
    """``layout`` (reference or compact form) and the identity of its index set as one uncompressed ``.npz``."""
    arrays = {f"layout/{key}": np.asarray(val) for key, val in layout.items()}
    np.savez(path, **arrays, **{"meta/d_in": np.int64(d_in), "meta/d_out": np.int64(d_out),
                                "meta/k": np.asarray(k, dtype=float), "meta/t": np.float64(t)})


def read_layout(path):
    """Inverse of :func:`write_layout`: ``(layout, meta)`` with the arrays bit for bit as written."""
    with np.load(path) as data:
        layout = {key[len("layout/"):]: data[key] for key in data.files if key.startswith("layout/")}
        meta = {"d_in": int(data["meta/d_in"]), "d_out": int(data["meta/d_out"]), "k": data["meta/k"], "t": float(data["meta/t"])}
    if "compact" in layout:
        layout["compact"] = bool(layout["compact"])
    return layout, meta


class SmolyakBarycentricInterpolator:
    """Smolyak interpolation operator with barycentric tensor-product interpolants, evaluated on the GPU."""

    # ------------------------------------------------------------------ properties (interpolation.py:31-49)
    @property
    def d_in(self) -> int:
        """Input dimension of target function and interpolant"""
        return self._d_in

    @property
    def d_out(self) -> int:
        """Output dimension of target function and interpolant"""
        return self._d_out

    @property
    def n_f_evals(self) -> int:
        """Number of function evaluations (== number of interpolation nodes) used by the interpolator"""
        return self._n_f_evals

    @property
    def n_f_evals_new(self) -> int:
        """Number of function evaluations that were not reused from previous computations"""
        return self._n_f_evals_new

    # ------------------------------------------------------------------ construction (interpolation.py:51-113)
    def __init__(self, node_gen=None, k: Sequence[float] = None, d_out: int = None, t: float = None,
                 f: Callable = None, *, n_inputs: int = None, memory_limit: float = 4.0, method: str = "auto",
                 device: int = None, batched_f: bool = False, nan_at_nodes: bool = True, layout: str = "auto",
                 dense: bool = None) -> None:
        r"""
        Parameters (all accepted as keywords, as in the reference)
        ----------
        node_gen : nodes.Generator
            One 1-D node family per input dimension.
        k : sequence of float
            Increasing anisotropy weights of the multi-index set; ``d_in = len(k)``.
        d_out : int
            Output dimension of the target function.
        t : float
            Threshold of the multi-index set.
        f : callable, optional
            Target function, called with one point ``(d_in,)``; see :meth:`set_f`.
        n_inputs : int, optional
            Expected batch size.  Like the reference (interpolation.py:250-251: a warm-up call on ``n_inputs`` random points
            at the end of ``set_f``), ``set_f`` then sizes the staging buffers of the host pipeline for that batch
            (``smx_prepare``) and runs one evaluation of ``n_inputs`` points, so that the first timed call pays neither.
        memory_limit : float
            Accepted for compatibility.  The fused kernels have no per-summand intermediates, so nothing is batched.
        method : {"auto", "barycentric"}
            "auto": values through the hierarchical fast path, gradient/integral through the per-summand kernels.
            "barycentric": every entry point through the per-summand second-barycentric-form kernels.
        device : int, optional
            CUDA ordinal (default: the current torch device).
        batched_f : bool
            If true ``set_f`` calls ``f`` once with all new nodes ``(n_new, d_in)`` instead of once per node.
        nan_at_nodes : bool
            ``gradient`` returns ``NaN`` in a dimension whose coordinate sits exactly on an interpolation node, as the
            reference does (default).  ``False`` returns the true, finite derivative there (fast path only).
        layout : {"auto", "reference", "compact"}
            "reference": ``set_f`` assembles the reference's zero-padded per-group tables (``reference_layout()``) and
            hands them to ``smx_create``.  "compact": it hands over the exact-shape, node-indexed form instead
            (``smx_create_compact``; one row of function values per interpolation node for nested rules) — the only
            form that fits in memory when ``d_out`` is in the thousands; ``reference_layout()`` and
            ``method="barycentric"`` are then unavailable.  "auto": compact when the padded value tensors would
            exceed 1 GiB.
        dense : bool, optional
            Force (True) or forbid (False) the GEMM-regime form of the value path; default: the library decides
            (``d_out >= 32``).
        """
        assert node_gen is not None and k is not None and d_out is not None and t is not None
        assert method in ("auto", "barycentric")
        assert layout in ("auto", "reference", "compact")
        assert not (layout == "compact" and method == "barycentric"), "the per-summand kernels need the reference layout"
        self._d_in = len(k)
        self._d_out = int(d_out)
        self._node_gen = node_gen
        self._is_nested = node_gen.is_nested
        self._k = k
        self._t = t
        self._method = method
        self._device = (torch.cuda.current_device() if torch.cuda.is_available() else 0) if device is None else int(device)
        self._batched_f = batched_f
        self._nan_at_nodes = nan_at_nodes
        self._memory_limit = memory_limit
        self._n_inputs = n_inputs
        self._layout_mode = layout
        self._dense = dense

        self._layout = None
        self._handle = None
        self._n_f_evals = indices.nodeset_cardinality(k, t, nested=self._is_nested)
        self._n_f_evals_new = 0
        if f is not None:
            self.set_f(f=f)

    def __del__(self):
        self._release()

    def _release(self):
        handle, self._handle = getattr(self, "_handle", None), None
        if handle is None:
            return
        try:  # (module globals are torn down at interpreter exit)
            _lib.lib.smx_destroy(handle)  # puts the caller's current CUDA device back (DeviceGuard in smx_api.cu)
        except Exception:
            pass

    # ------------------------------------------------------------------ set_f (interpolation.py:115-239)
    def set_f(self, *, f: Callable, f_evals: dict = None) -> dict:
        """Evaluate (or reuse from ``f_evals``) the target function at the interpolation nodes and build the device
        tables.  Returns the updated dictionary of evaluations: flat ``{mu_tuple: value}`` for nested rules,
        ``{nu: {mu_tuple: value}}`` otherwise (reference interpolation.py:119,155-163,208-228)."""
        f_evals = {} if f_evals is None else f_evals
        mode = self._layout_mode
        if mode == "auto":
            mode = "compact" if self._method != "barycentric" and self._padded_bytes() > (1 << 30) else "reference"
        layout, f_evals = (self._assemble_compact if mode == "compact" else self._assemble)(f, f_evals)
        self.set_layout(layout)
        return f_evals

    def _padded_bytes(self) -> int:
        """Size of the reference's zero-padded value tensors ``F_n`` (interpolation.py:203) for this index set."""
        offsets, _, degs_all, _ = indices.nonzero_arrays(self._k, self._t)
        lengths = np.diff(offsets)
        total = 0
        for n in set(lengths.tolist()) - {0}:
            sel = np.flatnonzero(lengths == n)
            degs = -np.sort(-degs_all[offsets[sel][:, None] + np.arange(n)[None, :]].astype(np.int64), axis=1)
            total += len(sel) * int(np.prod(degs.max(axis=0) + 1))
        return 8 * total * self._d_out

    def set_layout(self, layout: dict) -> None:
        """Build the device tables from an already assembled reference layout (``reference_layout()`` of another
        instance, e.g. received through ``dist.broadcast_layout``) instead of evaluating ``f`` again."""
        self._layout = layout
        self._release()
        flags = 0 if self._nan_at_nodes else _lib.SMX_GRAD_FINITE_AT_NODES
        if self._dense is not None:
            flags |= _lib.SMX_DENSE_PATH if self._dense else _lib.SMX_NO_DENSE_PATH
        with torch.cuda.device(self._device):
            if layout.get("compact"):
                assert self._method != "barycentric", "the per-summand kernels need the reference layout"
                self._handle = _lib.create_compact(layout, self._d_in, self._d_out, flags, self._device)
            else:
                flags |= _lib.SMX_KEEP_GROUPS | (_lib.SMX_NO_FAST_PATH if self._method == "barycentric" else 0)
                self._handle = _lib.create(layout, self._d_in, self._d_out, flags, self._device)
            if self._n_inputs:  # the reference's warm-up (interpolation.py:250-251): buffers sized, kernels loaded
                n = int(self._n_inputs)
                _lib.check(_lib.lib.smx_prepare(self._handle, n), "smx_prepare")
                warm = torch.rand((n, self._d_in), dtype=torch.float64, device=f"cuda:{self._device}")
                y = torch.empty((n, self._d_out), dtype=torch.float64, device=warm.device)
                _lib.check(_lib.lib.smx_eval(self._handle, warm.data_ptr(), n, self._d_in, y.data_ptr(), self._stream()), "smx_eval")

    def save_layout(self, path) -> None:
        """Write the tables ``set_f`` assembled to ``path`` (``.npz``), so that another process or a later run can build
        the device tables with :meth:`load_layout` without evaluating ``f`` again (the reference keeps no such cache:
        every run of benchmarking/benchmark.py:126-128 re-evaluates the target at all nodes)."""
        assert self._layout is not None, "The operator has not yet been set up for a target function via `set_f`."
        write_layout(self._layout, path, d_in=self._d_in, d_out=self._d_out, k=self._k, t=self._t)

    def load_layout(self, path) -> None:
        """Build the device tables from a file written by :meth:`save_layout`; asserts that the file belongs to this
        index set (``k``, ``t``) and these dimensions."""
        layout, meta = read_layout(path)
        assert meta["d_in"] == self._d_in and meta["d_out"] == self._d_out, \
            f"{path}: tables are for d_in={meta['d_in']}, d_out={meta['d_out']}"
        assert meta["t"] == float(self._t) and np.array_equal(meta["k"], np.asarray(self._k, dtype=float)), \
            f"{path}: tables belong to another index set"
        self.set_layout(layout)

    def _assemble_compact(self, f: Callable, f_evals: dict):
        """Host half of ``set_f`` without the reference's padding: summands in the order of the reference's walk, one
        row of ``values`` per distinct function evaluation (smx_compact_desc in include/smolyax_b200.h).  Same
        ``f_evals`` semantics and counters as :meth:`_assemble`."""
        gen = self._node_gen
        zero = np.array([g(0)[0] for g in gen])
        offsets, dims_all, degs_all, zetas_all = indices.nonzero_arrays(self._k, self._t)
        lengths = np.diff(offsets)
        offset = np.zeros(self._d_out)

        pool_at, pool_nodes, pool_quad, pool_len = {}, [], [], 0  # (generator, degree) -> offset into the pools
        row_of, value_blocks, n_rows = {}, [], 0  # evaluation -> row of `values` (nested rules share rows)
        n_active, slot_dims, slot_degs, slot_nodes, zetas, counts, val_index = [], [], [], [], [], [], []

        # summands in the order of the reference layout (groups by number of active dimensions in order of first
        # appearance, the reference's walk inside a group): both create entries then add up the same numbers in the same
        # order and the two kinds of handle agree bit for bit
        for n in sorted(set(lengths.tolist()), key=lambda v: np.flatnonzero(lengths == v)[0]):
            sel = np.flatnonzero(lengths == n)
            if n == 0:
                store = f_evals if self._is_nested else f_evals.setdefault((), {})
                if () not in store:
                    store[()] = f(zero.copy())
                    self._n_f_evals_new += 1
                offset = np.asarray(int(zetas_all[sel[0]]) * np.asarray(store[()], dtype=float), dtype=float) * np.ones(self._d_out)
                continue
            nn = len(sel)
            gather = offsets[sel][:, None] + np.arange(n)[None, :]
            dims_in, degs_in = dims_all[gather].astype(np.int64), degs_all[gather].astype(np.int64)
            order = np.argsort(-degs_in, axis=1, kind="stable")  # interpolation.py:175; any order is valid for the compact form
            sd, sg = np.take_along_axis(dims_in, order, axis=1), np.take_along_axis(degs_in, order, axis=1)
            node_tab = np.zeros((nn, n, int(sg.max()) + 1))
            node_off = np.empty((nn, n), dtype=np.int64)
            for slot in range(n):
                codes = sd[:, slot] * _CODE + sg[:, slot]
                uniq, first = np.unique(codes, return_index=True)
                for code in uniq[np.argsort(first, kind="stable")].tolist():  # pools grow in the order of the walk
                    dim, deg = code // _CODE, code % _CODE
                    key = (id(gen[dim]), deg)
                    if key not in pool_at:
                        p = np.asarray(gen[dim](deg), dtype=float)
                        q = np.zeros(deg + 1)
                        qw = np.asarray(gen[dim].get_quadrature_weights(deg), dtype=float)
                        q[: len(qw)] = qw
                        pool_at[key] = (pool_len, p)
                        pool_nodes.append(p)
                        pool_quad.append(q)
                        pool_len += deg + 1
                    rows = np.flatnonzero(codes == code)
                    node_off[rows, slot] = pool_at[key][0]
                    node_tab[rows, slot, : deg + 1] = pool_at[key][1]
            n_active.extend([n] * nn)
            slot_dims.append(sd.reshape(-1))
            slot_degs.append(sg.reshape(-1))
            slot_nodes.append(node_off.reshape(-1))
            zetas.append(zetas_all[sel].astype(np.int64))
            counts.append(np.prod(sg + 1, axis=1))

            _, _, keys, owner, inverse = self._walk_group(f, f_evals, zero, dims_in, degs_in, sd, sg, node_tab)
            if self._is_nested:
                row = np.empty(len(keys), dtype=np.int64)
                fresh = []
                for u, key in enumerate(keys):
                    r = row_of.get(key)
                    if r is None:
                        r = row_of[key] = n_rows + len(fresh)
                        fresh.append(f_evals[key])
                    row[u] = r
            else:
                row = n_rows + np.arange(len(keys))
                fresh = [store[key] for store, key in zip(owner, keys)]
            if fresh:
                value_blocks.append(self._value_rows(fresh))
                n_rows += len(fresh)
            val_index.append(row[inverse])

        values = np.concatenate(value_blocks) if value_blocks else np.empty((0, self._d_out))
        cat = lambda parts: np.concatenate(parts).astype(np.int64) if parts else np.zeros(0, dtype=np.int64)
        slot_dims, slot_degs, slot_nodes, zetas, val_index = (cat(v) for v in (slot_dims, slot_degs, slot_nodes, zetas, val_index))
        val_off = np.concatenate([[0], np.cumsum(cat(counts))]).astype(np.int64)
        layout = {
            "compact": True, "offset": offset, "n_active": np.asarray(n_active, dtype=np.int32),
            "slot_off": np.concatenate([[0], np.cumsum(n_active, dtype=np.int64)]).astype(np.int64),
            "dims": slot_dims, "degs": slot_degs, "node_off": slot_nodes,
            "node_pool": np.concatenate(pool_nodes) if pool_nodes else np.zeros(1),
            "quad_pool": np.concatenate(pool_quad) if pool_quad else np.zeros(1),
            "zetas": zetas, "val_off": val_off, "val_index": val_index, "values": np.ascontiguousarray(values),
        }
        return layout, f_evals

    def _assemble(self, f: Callable, f_evals: dict):
        """Host half of ``set_f``: the reference's per-group tables as NumPy arrays (no GPU needed)."""
        gen = self._node_gen
        zero = np.array([g(0)[0] for g in gen])
        offsets, dims_all, degs_all, zetas_all = indices.nonzero_arrays(self._k, self._t)
        lengths = np.diff(offsets)
        layout = {"offset": np.zeros(self._d_out)}

        for n in sorted(set(lengths.tolist()), key=lambda v: np.flatnonzero(lengths == v)[0]):
            sel = np.flatnonzero(lengths == n)  # summands of this group, in the order of the reference's walk
            zetas = zetas_all[sel].astype(np.int64)
            if n == 0:
                assert len(sel) == 1
                store = f_evals if self._is_nested else f_evals.get((), {})
                if () not in store:
                    store[()] = f(zero.copy())
                    self._n_f_evals_new += 1
                if not self._is_nested:
                    f_evals[()] = store
                layout["offset"] = np.asarray(int(zetas[0]) * np.asarray(store[()], dtype=float), dtype=float) * np.ones(self._d_out)
                continue

            nn = len(sel)
            gather = offsets[sel][:, None] + np.arange(n)[None, :]
            dims_in, degs_in = dims_all[gather].astype(np.int64), degs_all[gather].astype(np.int64)
            # active dimensions by degree, descending; ties keep ascending dimension (interpolation.py:175)
            order = np.argsort(-degs_in, axis=1, kind="stable")
            sorted_dims = np.take_along_axis(dims_in, order, axis=1)
            sorted_degs = np.take_along_axis(degs_in, order, axis=1)

            tau = tuple(int(v) for v in sorted_degs.max(axis=0))
            tw = max(tau) + 1
            node_tab = np.zeros((nn, n, tw))
            weight_tab = np.zeros((nn, n, tw))
            quad_tab = np.zeros((nn, n, tw))
            cache = {}
            for slot in range(n):
                codes = sorted_dims[:, slot] * _CODE + sorted_degs[:, slot]
                for code in np.unique(codes):
                    dim, deg = int(code // _CODE), int(code % _CODE)
                    key = (id(gen[dim]), deg)
                    if key not in cache:
                        pts = np.asarray(gen[dim](deg), dtype=float)
                        cache[key] = (pts, _host_weights(pts), np.asarray(gen[dim].get_quadrature_weights(deg), dtype=float))
                    pts, wts, qts = cache[key]
                    rows = np.flatnonzero(codes == code)
                    node_tab[rows, slot, : deg + 1] = pts
                    weight_tab[rows, slot, : deg + 1] = wts
                    quad_tab[rows, slot, : len(qts)] = qts

            F = np.zeros((nn,) + tuple(v + 1 for v in tau) + (self._d_out,))
            self._fill_values(F, f, f_evals, zero, dims_in, degs_in, sorted_dims, sorted_degs, node_tab)

            layout[f"F_{n}"] = np.ascontiguousarray(np.moveaxis(F, -1, 1))
            layout[f"nodes_{n}"] = node_tab
            layout[f"weights_{n}"] = weight_tab
            layout[f"dims_{n}"] = sorted_dims
            layout[f"degs_{n}"] = sorted_degs
            layout[f"zetas_{n}"] = zetas
            layout[f"quad_{n}"] = quad_tab

        return layout, f_evals

    def _walk_group(self, f, f_evals, zero, dims_in, degs_in, sorted_dims, sorted_degs, node_tab):
        """The reference's walk over the grids of one group (interpolation.py:208-228: summands in order, grid points
        ``mu`` in C order over the sorted slots, ``f`` called at the first occurrence of a node, ``f_evals`` keyed by
        the ``(dim, mu)`` pairs with ``mu > 0`` in ascending dimension), with the grids enumerated and de-duplicated by
        NumPy: Python touches only the distinct nodes (nested rules: ``n_f_evals`` of them instead of every grid point).

        Returns ``(si, mu, keys, owner, inverse)``: summand and grid index of every grid point in walk order, the keys
        of the distinct nodes in the order the walk meets them, the dictionary holding each of them, and the distinct
        node of every grid point."""
        nn, n = sorted_dims.shape
        d_out = self._d_out
        m = sorted_degs + 1
        stride = np.ones_like(m)
        for j in range(n - 2, -1, -1):
            stride[:, j] = stride[:, j + 1] * m[:, j + 1]
        count = stride[:, 0] * m[:, 0]
        start = np.concatenate(([0], np.cumsum(count)))
        total = int(start[-1])
        si = np.repeat(np.arange(nn), count)
        local = np.arange(total) - start[si]
        mu = (local[:, None] // stride[si]) % m[si]  # (total, n): grid index per sorted slot
        sentinel = np.iinfo(np.int64).max
        rows = np.sort(np.where(mu > 0, sorted_dims[si] * _CODE + mu, sentinel), axis=1)  # key pairs by ascending dim

        if self._is_nested:
            _, first, inverse = np.unique(rows, axis=0, return_index=True, return_inverse=True)
            walk = np.argsort(first, kind="stable")  # distinct nodes in the order the walk meets them
            rank = np.empty_like(walk)
            rank[walk] = np.arange(len(walk))
            rep, inverse = first[walk], rank[inverse.reshape(-1)]
            owner = [f_evals] * len(rep)
        else:
            rep, inverse = np.arange(total), np.arange(total)
            stores = []
            for i in range(nn):
                nu = tuple(zip(dims_in[i].tolist(), degs_in[i].tolist()))
                stores.append(f_evals.setdefault(nu, {}))
            owner = [stores[i] for i in si.tolist()]

        rep_rows = rows[rep]
        nnz = (rep_rows != sentinel).sum(axis=1)
        keys = [None] * len(rep)
        for r in np.unique(nnz).tolist():  # keys of r pairs at a time: the tuples are built by zip, not in Python code
            where = np.flatnonzero(nnz == r)
            sub = rep_rows[where, :r]
            for u, a, b in zip(where.tolist(), (sub // _CODE).tolist(), (sub % _CODE).tolist()):
                keys[u] = tuple(zip(a, b))
        missing = [u for u, key in enumerate(keys) if key not in owner[u]]
        self._n_f_evals_new += len(missing)
        chunk = max(1, (1 << 23) // len(zero))  # <= 64 MB of points at a time
        for c0 in range(0, len(missing), chunk):
            part = missing[c0:c0 + chunk]
            g = rep[part]
            X = np.tile(zero, (len(part), 1)) if zero.any() else np.zeros((len(part), len(zero)))
            X[np.arange(len(part))[:, None], sorted_dims[si[g]]] = node_tab[si[g][:, None], np.arange(n)[None, :], mu[g]]
            if self._batched_f:
                vals = np.asarray(f(X), dtype=float).reshape(len(part), -1)
                vals = [v if d_out > 1 else (v[0] if v.size == 1 else v) for v in vals]
            else:
                vals = [f(x) for x in X]
            for u, v in zip(part, vals):
                owner[u][keys[u]] = v
        return si, mu, keys, owner, inverse

    def _value_rows(self, found):
        """``(len(found), d_out)`` array of the function values in ``found`` (scalars are broadcast)."""
        try:
            table = np.asarray(found, dtype=float).reshape(len(found), -1)
            return np.broadcast_to(table, (len(found), self._d_out))
        except ValueError:  # values of mixed shapes (scalars from the caller's f_evals beside arrays)
            table = np.empty((len(found), self._d_out))
            for u, v in enumerate(found):
                table[u] = np.asarray(v, dtype=float).reshape(-1)
            return table

    def _fill_values(self, F, f, f_evals, zero, dims_in, degs_in, sorted_dims, sorted_degs, node_tab):
        """Function values of one group into ``F (nn, tau_1+1, .., tau_n+1, d_out)`` (interpolation.py:203-228)."""
        n = sorted_dims.shape[1]
        si, mu, keys, owner, inverse = self._walk_group(f, f_evals, zero, dims_in, degs_in, sorted_dims, sorted_degs, node_tab)
        f_stride = np.ones(n, dtype=np.int64)
        for j in range(n - 2, -1, -1):
            f_stride[j] = f_stride[j + 1] * F.shape[j + 2]
        pos = si * int(np.prod(F.shape[1:-1])) + mu @ f_stride
        table = self._value_rows([store[key] for store, key in zip(owner, keys)])
        F.reshape(-1, self._d_out)[pos] = table[inverse]

    # ------------------------------------------------------------------ helpers
    def reference_layout(self) -> dict:
        """The per-group tables exactly as the reference keeps them after ``set_f`` (interpolation.py:230-235), as
        NumPy arrays keyed ``offset``, ``F_n``, ``nodes_n``, ``weights_n``, ``dims_n``, ``degs_n``, ``zetas_n``
        (+ ``quad_n``, the tables of interpolation.py:361-379)."""
        assert self._layout is not None, "The operator has not yet been set up for a target function via `set_f`."
        assert not self._layout.get("compact"), "set_f used the compact layout; construct with layout='reference'"
        return self._layout

    def device_info(self) -> dict:
        """Sizes of the device layout (smx_get_info)."""
        assert self._handle is not None, "The operator has not yet been set up for a target function via `set_f`."
        return _lib.info(self._handle)

    def _validate_input(self, x):
        """interpolation.py:253-262: set-up check, ``(d_in,)`` -> ``(1, d_in)``, column-count assertion."""
        assert self._handle is not None, "The operator has not yet been set up for a target function via `set_f`."
        if isinstance(x, torch.Tensor):
            if x.dim() == 1 and x.shape[0] == self._d_in:
                x = x[None, :]
            assert x.dim() == 2 and x.shape[1] == self._d_in, f"{tuple(x.shape)[-1]} != {self._d_in}"
            x = x.to(torch.float64)
            if x.stride(1) != 1:
                x = x.contiguous()
            kind = "cuda" if x.is_cuda else "torch_cpu"
        else:
            x = np.asarray(x, dtype=np.float64)
            if x.shape == (self._d_in,):
                x = x[None, :]
            assert x.ndim == 2 and x.shape[1] == self._d_in, f"{x.shape[-1]} != {self._d_in}"
            x = np.ascontiguousarray(x)
            kind = "numpy"
        self._n_inputs = x.shape[0]
        return x, kind

    def _ldx(self, x) -> int:
        """Row pitch in elements (a single row may carry an arbitrary stride on its size-1 axis)."""
        return int(x.stride(0)) if x.shape[0] > 1 else self._d_in

    @staticmethod
    def _stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ------------------------------------------------------------------ __call__ (interpolation.py:264-304)
    def __call__(self, x, out=None):
        """Evaluate the interpolant at ``x`` of shape ``(n_points, d_in)`` or ``(d_in,)``; result ``(n_points, d_out)``.

        A CUDA ``torch`` tensor is evaluated in place on the current stream and a CUDA tensor is returned without
        synchronising (like the reference's un-synchronised ``jax.Array``).  Host input (NumPy, or a CPU tensor —
        pinned for full speed) goes through the pipelined host path and returns a host array of the same kind.

        ``out`` (not in the reference): a C-contiguous float64 buffer of shape ``(n_points, d_out)`` of the same kind as
        ``x`` to write the result into.  Worth it for wide outputs from host memory: a fresh page-locked result buffer
        costs about a second per 8 GB, more than the evaluation.
        """
        x, kind = self._validate_input(x)
        n_points = x.shape[0]
        lib = _lib.lib
        if out is not None:
            ok = (isinstance(out, torch.Tensor) and out.dtype == torch.float64 and out.is_contiguous() and
                  (out.is_cuda and out.device == x.device if kind == "cuda" else not out.is_cuda)) if kind != "numpy" else \
                 (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous and out.flags.writeable)
            assert ok and tuple(out.shape) == (n_points, self._d_out), "out: C-contiguous float64 (n_points, d_out) of the same kind as x"
        with torch.cuda.device(self._device):
            if kind == "cuda":
                y = torch.empty((n_points, self._d_out), dtype=torch.float64, device=x.device) if out is None else out
                _lib.check(lib.smx_eval(self._handle, x.data_ptr(), n_points, self._ldx(x), y.data_ptr(), self._stream()), "smx_eval")
                return y
            if kind == "torch_cpu":
                y = torch.empty((n_points, self._d_out), dtype=torch.float64, pin_memory=x.is_pinned()) if out is None else out
                _lib.check(lib.smx_eval_host(self._handle, x.data_ptr(), n_points, self._ldx(x), y.data_ptr(), 0), "smx_eval_host")
                return y
            y = np.empty((n_points, self._d_out)) if out is None else out
            _lib.check(lib.smx_eval_host(self._handle, x.ctypes.data, n_points, x.shape[1], y.ctypes.data, 0), "smx_eval_host")
            return y

    # ------------------------------------------------------------------ gradient (interpolation.py:306-345)
    def gradient(self, x):
        """Gradient of the interpolant at ``x``: ``(n_points, d_out, d_in)``.  As in the reference, a coordinate that
        sits exactly on an interpolation node of its dimension yields ``NaN`` in that dimension."""
        x, kind = self._validate_input(x)
        n_points = x.shape[0]
        with torch.cuda.device(self._device):
            xd = x if kind == "cuda" else (x.cuda() if kind == "torch_cpu" else torch.from_numpy(x).cuda())
            J = torch.empty((n_points, self._d_out, self._d_in), dtype=torch.float64, device=xd.device)
            status = _lib.lib.smx_gradient(self._handle, xd.data_ptr(), n_points, self._ldx(xd), J.data_ptr(), self._stream())
            if status == _lib.SMX_ERR_UNSUPPORTED and self._layout.get("compact") and self._d_out > 1:
                self._gradient_by_columns(xd, J)  # derivative sets of all outputs at once were not built (d_out too large)
            else:
                _lib.check(status, "smx_gradient")
            if kind == "cuda":
                return J
            return J.cpu() if kind == "torch_cpu" else J.cpu().numpy()

    def _gradient_by_columns(self, xd, J) -> None:
        """Gradient of a handle too wide for one set of derivative tables (compact layout, many outputs): the interpolant is
        linear in ``f``, so ``J[:, lo:hi, :]`` is the gradient of the interpolant of outputs ``lo:hi`` alone.  One
        temporary handle per block of outputs (its derivative sets are built, used once and freed: the tables of all blocks
        together are what did not fit).  The block width starts at ``_GRAD_COLUMN_BLOCK`` and is halved until the
        library accepts it (the derivative tables grow with terms x hot dimensions x outputs)."""
        from .dist import column_slice

        width = min(_GRAD_COLUMN_BLOCK, max(1, self._d_out // 2))
        lo = 0
        while lo < self._d_out:
            hi = min(lo + width, self._d_out)
            block = SmolyakBarycentricInterpolator(node_gen=self._node_gen, k=self._k, t=self._t, d_out=hi - lo,
                                                   device=self._device, nan_at_nodes=self._nan_at_nodes, layout="compact")
            block.set_layout(column_slice(self._layout, lo, hi))
            Jb = torch.empty((xd.shape[0], hi - lo, self._d_in), dtype=torch.float64, device=xd.device)
            status = _lib.lib.smx_gradient(block._handle, xd.data_ptr(), xd.shape[0], self._ldx(xd), Jb.data_ptr(), self._stream())
            if status == _lib.SMX_ERR_UNSUPPORTED and width > 1:
                block._release()
                width = max(1, width // 2)
                continue
            _lib.check(status, "smx_gradient")
            J[:, lo:hi, :] = Jb
            torch.cuda.current_stream().synchronize()  # the block's tables are freed next
            block._release()
            lo = hi

    # ------------------------------------------------------------------ integral (interpolation.py:347-390)
    def integral(self):
        """Integral of the interpolant w.r.t. the probability measure of the node family (= Smolyak quadrature of
        ``f``): shape ``(d_out,)``, synchronised like the reference's ``block_until_ready``."""
        assert self._handle is not None, "The operator has not yet been set up for a target function via `set_f`."
        with torch.cuda.device(self._device):
            q = torch.empty(self._d_out, dtype=torch.float64, device="cuda")
            _lib.check(_lib.lib.smx_integral(self._handle, q.data_ptr(), self._stream()), "smx_integral")
            return q.cpu().numpy()
