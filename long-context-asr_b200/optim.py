"""Device-side optimizer step for the training loop: drop-ins for ``lcasr/optim/madgrad.py`` (``MADGRAD``,
same constructor, param groups, ``state`` layout: ``grad_sum_sq`` / ``s`` / ``x0`` per parameter and the step counter
``state['k']``) and for the ``torch.nn.utils.clip_grad_norm_`` call of ``exp/train.py:55``.

The reference runs ~12 elementwise torch kernels per parameter tensor per step; here one multi-tensor reduction
(`lcasr_grad_sumsq`) and one multi-tensor update per parameter group (`lcasr_madgrad_step`) do the whole step.  With
``optimizer.max_grad_norm = clip`` the clipping coefficient is applied inside the update (read from device memory: the
step never synchronises with the host)."""
from __future__ import annotations

import math
from typing import Iterable, Optional

import numpy as np
import torch

from . import _lib as L


class _Table:
    """device table of lcasr_opt_tensor + chunk lists for a list of parameters"""

    def __init__(self, params, device):
        self.params = params
        E = int(L.lib.lcasr_opt_chunk_elems())
        ct, ci = [], []
        for t, p in enumerate(params):
            n_chunks = (p.numel() + E - 1) // E
            ct += [t] * n_chunks
            ci += list(range(n_chunks))
        self.n_chunks = len(ct)
        self.chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=device)
        self.chunk_index = torch.tensor(ci, dtype=torch.int32, device=device)
        self.host = torch.zeros(len(params), 6, dtype=torch.int64).pin_memory()
        self.dev = torch.zeros(len(params), 6, dtype=torch.int64, device=device)
        self._copied = None  # event after the last H2D copy of the pinned table

    def fill(self, state=None):
        if self._copied is not None:  # the previous step's copy may still be reading the pinned table
            self._copied.synchronize()
        a = self.host.numpy()
        for t, p in enumerate(self.params):
            g = p.grad
            if g is not None and (g.dtype != torch.float32 or not g.is_contiguous()):
                raise RuntimeError("lcasr_b200.optim needs contiguous fp32 gradients")
            a[t, 0] = p.data_ptr()
            a[t, 1] = g.data_ptr() if g is not None else 0
            if state is not None and g is not None:
                st = state[p]
                a[t, 2], a[t, 3] = st["grad_sum_sq"].data_ptr(), st["s"].data_ptr()
                a[t, 4] = st["x0"].data_ptr() if "x0" in st else 0
            a[t, 5] = p.numel()
        self.dev.copy_(self.host, non_blocking=True)
        if self._copied is None:
            self._copied = torch.cuda.Event()
        self._copied.record()
        return self.dev


def _check(params):
    for p in params:
        if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
            raise RuntimeError("lcasr_b200.optim works on contiguous fp32 CUDA parameters (there is no CPU fallback)")


_CLIP_TABLES = {}


def clip_grad_norm_(parameters: Iterable[torch.Tensor], max_norm: float, norm_type: float = 2.0) -> torch.Tensor:
    """torch.nn.utils.clip_grad_norm_ (L2 norm): scales the gradients in place, returns the total norm (device tensor)."""
    if norm_type != 2.0:
        raise NotImplementedError("only the L2 norm is used by the reference (exp/train.py:55)")
    params = [p for p in (parameters if not isinstance(parameters, torch.Tensor) else [parameters]) if p.grad is not None]
    if not params:
        return torch.tensor(0.0)
    _check(params)
    key = tuple(id(p) for p in params)
    if key not in _CLIP_TABLES:
        _CLIP_TABLES.clear()
        _CLIP_TABLES[key] = _Table(params, params[0].device)
    tab = _CLIP_TABLES[key]
    dev = tab.fill()
    sumsq = torch.empty(1, dtype=torch.float32, device=params[0].device)
    st = L.current_stream()
    L.call("lcasr_grad_sumsq", L.ptr(dev), L.ptr(tab.chunk_tensor), L.ptr(tab.chunk_index), tab.n_chunks, L.ptr(sumsq), st)
    L.call("lcasr_grad_scale", L.ptr(dev), L.ptr(tab.chunk_tensor), L.ptr(tab.chunk_index), tab.n_chunks, L.ptr(sumsq),
           float(max_norm), st)
    return sumsq.sqrt()[0]


class MADGRAD(torch.optim.Optimizer):
    """lcasr/optim/madgrad.py:18-212 (dense fp32 parameters).  Extra, optional: ``max_grad_norm`` (attribute or
    ``step(max_grad_norm=...)``) fuses exp/train.py:55's global-norm clipping into the update."""

    def __init__(self, params, lr: float = 1e-2, momentum: float = 0.9, weight_decay: float = 0, eps: float = 1e-6,
                 decouple_decay=False):
        if momentum < 0 or momentum >= 1:
            raise ValueError(f"Momentum {momentum} must be in the range [0,1)")
        if lr < 0:
            raise ValueError(f"Learning rate {lr} must be non-negative")
        if weight_decay < 0:
            raise ValueError(f"Weight decay {weight_decay} must be non-negative")
        if eps < 0:
            raise ValueError("Eps must be non-negative")
        super().__init__(params, dict(lr=lr, eps=eps, momentum=momentum, weight_decay=weight_decay, decouple_decay=decouple_decay))
        self.max_grad_norm: float = 0.0
        self._tables = {}
        self._all = None
        self.last_grad_sumsq: Optional[torch.Tensor] = None

    @property
    def supports_memory_efficient_fp16(self) -> bool:
        return False

    @property
    def supports_flat_params(self) -> bool:
        return True

    def _table(self, gi, params, device):
        key = (gi, tuple(id(p) for p in params))
        if self._tables.get(gi, (None, None))[0] != key:
            self._tables[gi] = (key, _Table(params, device))
        return self._tables[gi][1]

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: Optional[float] = None):
        loss = closure() if closure is not None else None
        if "k" not in self.state:
            self.state["k"] = torch.tensor([0], dtype=torch.long)
        if not hasattr(self, "_k_host"):
            self._k_host = int(self.state["k"].item())  # once (e.g. after load_state_dict); then mirrored on the host
        k = self._k_host
        clip = self.max_grad_norm if max_grad_norm is None else max_grad_norm
        st = L.current_stream()
        sumsq = None
        if clip and clip > 0:  # global norm over every parameter of every group (clip_grad_norm_(model.parameters()))
            allp = [p for g in self.param_groups for p in g["params"]]
            _check(allp)
            tab = self._table(-1, allp, allp[0].device)
            dev = tab.fill()
            sumsq = torch.empty(1, dtype=torch.float32, device=allp[0].device)
            L.call("lcasr_grad_sumsq", L.ptr(dev), L.ptr(tab.chunk_tensor), L.ptr(tab.chunk_index), tab.n_chunks, L.ptr(sumsq), st)
            self.last_grad_sumsq = sumsq
        for gi, group in enumerate(self.param_groups):
            eps, lr = group["eps"], group["lr"]
            if lr != 0.0:
                lr = lr + eps  # "For stability" (madgrad.py:104)
            momentum = group["momentum"]
            lamb = lr * math.pow(k + 1, 0.5)
            params = list(group["params"])
            if not params:
                continue
            _check(params)
            for p in params:
                if p.grad is None:
                    continue
                state = self.state[p]
                if "grad_sum_sq" not in state:
                    state["grad_sum_sq"] = torch.zeros_like(p.data)
                    state["s"] = torch.zeros_like(p.data)
                    if momentum != 0:
                        state["x0"] = torch.clone(p.data).detach()
            tab = self._table(gi, params, params[0].device)
            dev = tab.fill(self.state)
            L.call("lcasr_madgrad_step", L.ptr(dev), L.ptr(tab.chunk_tensor), L.ptr(tab.chunk_index), tab.n_chunks, L.ptr(sumsq),
                   float(clip or 0.0), float(lr), float(lamb), float(eps), float(group["weight_decay"]), float(momentum),
                   int(bool(group.get("decouple_decay", False))), st)
        self.state["k"] += 1
        self._k_host += 1
        L.bump_weight_epoch()  # parameters were written through raw pointers: packed eval copies are stale
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        if hasattr(self, "_k_host"):
            del self._k_host
        self._tables = {}
