"""``UpdateMatrix``: the logical ``U[N, d]`` stack of client updates.

The reference materialises ``U`` on the driver CPU with ``torch.stack``
(/root/reference/src/blades/aggregators/mean.py:21-28) and every aggregator is
written against that dense tensor.  Here aggregators are written against a
small set of *matrix primitives* instead, so the same aggregator code runs on

* ``LocalMatrix``    -- one dense tensor (CPU oracle, or one GPU with our kernels)
* ``ShardedMatrix``  -- rows spread over G trainer shards in NVLink-addressable
  symmetric memory (``comm.symm``); primitives run as pull-mode fused kernels
  and ``U`` is never gathered on one device.

Primitives (everything the 11 reference aggregators need):

=================  =========================================================
``mean()``          column mean                                  (K2)
``combine(w)``      ``sum_i w_i U[i]``                            (K2)
``trimmed_mean(b)`` coordinate-wise trimmed mean                 (K3)
``median()``        coordinate-wise median (avg of middle two)   (K4)
``gram(extra)``     ``[U;extra][U;extra]^T`` as float64 ``[N+e,N+e]`` (K5)
``rows()``          dense ``[N,d]`` (gathers; escape hatch for custom aggs)
=================  =========================================================

*Virtual rows* (SURVEY 7.2.3): a ``VirtualRows`` spec replaces the first ``f``
rows by a coordinate-wise closed form of the honest rows (ALIE / IPM), so the
fused kernels never store f identical malicious rows.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

__all__ = ["VirtualRows", "UpdateMatrix", "LocalMatrix", "as_matrix"]


@dataclass
class VirtualRows:
    """``count`` identical rows ``value(honest rows)`` replacing rows ``replaced``.

    kind='alie': value = mean - z * std_unbiased   (reference alieclient.py:32-36)
    kind='ipm' : value = -eps * mean               (reference ipmclient.py:16)
    Honest rows = all rows whose index is not in ``byzantine`` (all Byzantine rows,
    not only the replaced ones, are excluded from the statistics).
    """
    kind: str
    param: float
    replaced: Sequence[int]
    byzantine: Sequence[int] = field(default_factory=list)

    def __post_init__(self):
        if self.kind not in ("alie", "ipm"):
            raise ValueError(self.kind)
        if not self.byzantine:
            self.byzantine = list(self.replaced)

    @property
    def count(self) -> int:
        return len(self.replaced)

    def value(self, honest: torch.Tensor) -> torch.Tensor:
        mu = honest.mean(dim=0)
        if self.kind == "alie":
            return mu - self.param * honest.std(dim=0, unbiased=True)
        return -self.param * mu


class UpdateMatrix:
    """Interface; see module docstring."""

    n_rows: int
    n_cols: int
    device: torch.device
    virtual: Optional[VirtualRows] = None
    #: (lr,) -> the final primitive also applies ``theta += lr * agg`` on every replica (K8 fusion)
    server_step = None
    step_applied = False
    #: pipelined aggregation (engine/round.py): this matrix object only covers coordinates ``[window[0], window[1])``
    #: -- chunk number ``chunk`` of the round, ``last_chunk`` closes it.  Coordinate-wise primitives (mean / combine,
    #: trimmed mean, median, virtual-row materialisation) honour it; ``out_buffer`` is the round's shared result vector.
    window = None
    chunk = 0
    last_chunk = True
    out_buffer = None

    def _cols(self):
        return (0, self.n_cols) if self.window is None else (int(self.window[0]), int(self.window[1]))

    # coordinate-wise -----------------------------------------------------------
    def mean(self) -> torch.Tensor:
        w = torch.full((self.n_rows,), 1.0 / self.n_rows, dtype=torch.float64)
        return self.combine(w)

    def combine(self, weights, extra: Optional[torch.Tensor] = None, extra_weight: float = 0.0) -> torch.Tensor:
        """``sum_i w_i U[i] (+ extra_weight * extra)``."""
        raise NotImplementedError
    def trimmed_mean(self, b: int) -> torch.Tensor: raise NotImplementedError
    def median(self) -> torch.Tensor: raise NotImplementedError
    # geometry ------------------------------------------------------------------
    def gram(self, extra: Optional[torch.Tensor] = None) -> np.ndarray: raise NotImplementedError

    def gram_device(self, extra: Optional[torch.Tensor] = None):
        """The Gram matrix left ON THE DEVICE (``ops.gram_solve.DeviceGram``) for the on-device solvers, or ``None``
        when this matrix has no kernel path (CPU oracle): callers then fall back to ``gram()`` + the host solvers."""
        return None
    # escape hatch ----------------------------------------------------------------
    def rows(self) -> torch.Tensor: raise NotImplementedError

    def __len__(self) -> int:
        return self.n_rows


def _sanitize(u: torch.Tensor) -> torch.Tensor:
    return u


class LocalMatrix(UpdateMatrix):
    """Dense ``[N, d]`` tensor on one device.

    On CUDA the primitives dispatch to the hand-written sm_100a kernels in
    ``blades_b200.ops`` (coordinate-select, row-combine, tcgen05 Gram); on CPU they
    are the pure-torch oracle used by the tests.
    """

    def __init__(self, data: torch.Tensor, virtual: Optional[VirtualRows] = None,
                 use_kernels: Optional[bool] = None, theta: Optional[torch.Tensor] = None):
        assert data.dim() == 2, data.shape
        self.data = data
        self.theta = theta               # flat parameter vector for the fused server step
        self.n_rows, self.n_cols = data.shape
        self.device = data.device
        self.virtual = virtual
        if use_kernels is None:
            use_kernels = data.is_cuda and data.dtype == torch.float32
        self.use_kernels = use_kernels

    # -- helpers ------------------------------------------------------------------
    def _honest(self) -> torch.Tensor:
        v = self.virtual
        keep = [i for i in range(self.n_rows) if i not in set(v.byzantine)]
        return self.data[keep]

    def materialize_virtual(self) -> torch.Tensor:
        """Write the virtual rows into ``data`` (what the reference's callbacks do)."""
        v = self.virtual
        if v is not None and v.count:
            if self.use_kernels:
                from ..ops import attack as _a, select as _s
                byz = set(v.byzantine)
                honest = [i for i in range(self.n_rows) if i not in byz]
                c0, c1 = self._cols()
                _a.attack_rows(_s.row_pointers(self.data, honest), _s.row_pointers(self.data, list(v.replaced)),
                               v.kind, v.param, c0, c1, self.data.device)
            else:
                val = v.value(self._honest())
                self.data[list(v.replaced)] = val
            self.virtual = None
        return self.data

    def rows(self) -> torch.Tensor:
        return self.materialize_virtual()

    # -- primitives -----------------------------------------------------------------
    def _out(self, device) -> torch.Tensor:
        if self.out_buffer is not None:
            return self.out_buffer
        return torch.empty(self.n_cols, device=device, dtype=torch.float32)

    def _windowed(self, full: torch.Tensor) -> torch.Tensor:
        """CPU oracle of a windowed primitive: the full-width result restricted to this object's window, written into
        the round's shared result vector (coordinates outside the window are left to the other chunks)."""
        if self.window is None:
            return full
        c0, c1 = self._cols()
        if self.out_buffer is None:
            self.out_buffer = torch.zeros_like(full)
        self.out_buffer[c0:c1] = full[c0:c1]
        return self.out_buffer

    def _epilogue(self, out: torch.Tensor):
        from ..ops import select as _s
        if self.server_step is not None and self.theta is not None:
            self.step_applied = True
            return _s.make_epilogue([out.data_ptr()], [self.theta.data_ptr()], self.theta.data_ptr(),
                                    float(self.server_step[0]))
        return _s.make_epilogue([out.data_ptr()])

    def combine(self, weights, extra: Optional[torch.Tensor] = None, extra_weight: float = 0.0) -> torch.Tensor:
        if self.use_kernels and torch.is_tensor(weights) and weights.is_cuda:
            # device-resident weights (on-device Gram solvers); with ``extra`` the last weight belongs to it
            from ..ops import combine as _k, select as _s
            data = self.rows()
            rows = _s.row_pointers(data)
            keep = None
            if extra is not None:
                keep = extra.to(data.device, torch.float32).contiguous()
                rows = rows + [keep.data_ptr()]
            out = self._out(data.device)
            c0, c1 = self._cols()
            _k.launch_combine(rows, weights.to(torch.float32).contiguous(), c0, c1, self._epilogue(out), data.device)
            return out
        w = torch.as_tensor(np.asarray(weights, dtype=np.float64) if not torch.is_tensor(weights) else weights)
        if self.use_kernels:
            from ..ops import combine as _k, select as _s
            data = self.rows()
            rows = _s.row_pointers(data)
            wl = [float(x) for x in w.tolist()]
            if extra is not None and extra_weight != 0.0:
                extra = extra.contiguous()
                rows, wl = rows + [extra.data_ptr()], wl + [float(extra_weight)]
            out = self._out(data.device)
            c0, c1 = self._cols()
            _k.launch_combine(rows, wl, c0, c1, self._epilogue(out), data.device)
            return out
        data = self.rows()
        res = (w.to(data.device, torch.float64)[:, None] * data.double()).sum(0)
        if extra is not None and extra_weight != 0.0:
            res = res + extra_weight * extra.to(res.device, torch.float64)
        return self._windowed(res.to(data.dtype))

    def mean(self) -> torch.Tensor:
        if self.use_kernels:
            return super().mean()
        return self._windowed(self.rows().mean(dim=0))

    def trimmed_mean(self, b: int) -> torch.Tensor:
        n = self.n_rows
        if n - 2 * b <= 0:
            raise ValueError(f"trim {b} too large for {n} rows")
        if self.use_kernels:
            return self._select_kernel(0, b)
        data = self.rows()
        if b == 0:
            return self._windowed(data.mean(0))
        srt = data.sort(dim=0).values
        return self._windowed(srt[b: n - b].mean(dim=0))

    def _select_kernel(self, mode: int, b: int) -> torch.Tensor:
        from ..ops import select as _s
        data, v, n = self.data, self.virtual, self.n_rows
        out = self._out(data.device)
        c0, c1 = self._cols()
        if v is not None and v.count:
            byz, rep = set(v.byzantine), set(v.replaced)
            stat = _s.row_pointers(data, [i for i in range(n) if i not in byz])
            other = _s.row_pointers(data, [i for i in range(n) if i in byz and i not in rep])
            _s.launch_select(stat, other, v.count, v.kind, v.param, mode, b, c0, c1,
                             self._epilogue(out), data.device)
        else:
            _s.launch_select(_s.row_pointers(data), [], 0, None, 0.0, mode, b, c0, c1,
                             self._epilogue(out), data.device)
        return out

    def median(self) -> torch.Tensor:
        if self.use_kernels:
            return self._select_kernel(1, 0)
        data = self.rows()
        n = self.n_rows
        srt = data.sort(dim=0).values
        return self._windowed((srt[(n - 1) // 2] + srt[n // 2]) * 0.5)

    def gram(self, extra: Optional[torch.Tensor] = None) -> np.ndarray:
        data = self.rows()
        if extra is not None:
            extra = extra.reshape(-1, self.n_cols).to(data.device, data.dtype)
        if self.use_kernels:
            from ..ops import gram as _k
            return _k.gram(data, extra)
        full = data if extra is None else torch.cat([data, extra], 0)
        full = full.double()
        return (full @ full.T).cpu().numpy()


def _local_gram_device(self, extra: Optional[torch.Tensor] = None):
    if not self.use_kernels:
        return None
    from ..ops import _gram_impl, _loader, gram as _k, gram_solve
    if not gram_solve.enabled() or _k.PRECISION == "fp32":
        return None
    data = self.rows()
    if extra is not None:
        extra = extra.reshape(-1, self.n_cols).to(data.device, data.dtype)
    r = _gram_impl.gram_padded(_loader.cuda_lib(), data, extra, _k.PRECISION)
    if r is None:
        return None
    out, idx = r
    return gram_solve.DeviceGram(out, gram_solve.index_tensor(idx, data.device), len(idx))


LocalMatrix.gram_device = _local_gram_device


def as_matrix(inputs, virtual: Optional[VirtualRows] = None) -> UpdateMatrix:
    """Coerce the reference's three input conventions (mean.py:21-28) to a matrix."""
    from ..client import BladesClient
    if isinstance(inputs, UpdateMatrix):
        return inputs
    if torch.is_tensor(inputs):
        return LocalMatrix(inputs if inputs.dim() == 2 else inputs.reshape(1, -1), virtual)
    inputs = list(inputs)
    if len(inputs) and all(isinstance(e, BladesClient) for e in inputs):
        # virtual clients bound to consecutive rows of one matrix -> zero-copy
        slots = [c._slot for c in inputs]
        if all(s is not None for s in slots) and all(s.matrix is slots[0].matrix for s in slots) \
                and [s.row for s in slots] == list(range(len(slots))) \
                and slots[0].matrix.shape[0] == len(slots):
            mat = slots[0].matrix
            torch.nan_to_num_(mat)
            return LocalMatrix(mat, virtual)
        return LocalMatrix(torch.stack([c.get_update() for c in inputs]), virtual)
    if len(inputs) and all(torch.is_tensor(e) for e in inputs):
        return LocalMatrix(torch.stack(inputs, dim=0), virtual)
    if len(inputs) and all(callable(getattr(e, "get_update", None)) for e in inputs):
        # duck-typed clients (the reference only ever calls ``.get_update()``, mean.py:23)
        return LocalMatrix(torch.stack([e.get_update() for e in inputs]), virtual)
    raise TypeError("aggregator inputs must be clients, tensors or an UpdateMatrix")
