"""`torch.library` registration of the propagation operator — the "thin C-ABI torch custom-op layer" of
BASELINE.json's north_star.

    out = torch.ops.lgcn_b200.propagate(weight, handle)        # mean_k A_hat^k weight  (model/lgcn.py:78-86)

`handle` is an integer key of a registered `LightGCN` (custom ops take tensors and scalars, not Python
objects; the model owns the CSR graph and the ping-pong buffers).  The op is a plain wrapper over
`lgcn_propagate_layer` xK through ctypes; `register_autograd` supplies the backward — the same operator
in Horner form, because A_hat is symmetric — and `register_fake` the shape rule, so the
op composes with autograd, `torch.compile` graphs and `torch.library.opcheck`.  `LightGCN.computer()`
routes through it whenever a gradient is required.
"""
from __future__ import annotations

import weakref
from typing import Dict

import torch

_MODELS: Dict[int, "weakref.ReferenceType"] = {}
_NEXT = [1]


def register_model(model) -> int:
    """Returns the integer handle `torch.ops.lgcn_b200.propagate` takes for this model."""
    h = getattr(model, "_op_handle", None)
    if h is None:
        h = _NEXT[0]
        _NEXT[0] += 1
        _MODELS[h] = weakref.ref(model)
        model._op_handle = h
    return h


def _model(handle: int):
    ref = _MODELS.get(int(handle))
    m = ref() if ref is not None else None
    if m is None:
        raise RuntimeError(f"lgcn_b200.propagate: unknown or dead model handle {handle}")
    return m


@torch.library.custom_op("lgcn_b200::propagate", mutates_args=())
def propagate(weight: torch.Tensor, handle: int) -> torch.Tensor:
    m = _model(handle)
    out = torch.empty_like(weight)
    m._propagate_into(weight.detach(), out)
    m._op_drop_bwd = m._drop_bwd     # the dropout weights of THIS pass (model/MF.py:158-192), for its backward
    return out


@propagate.register_fake
def _(weight, handle):
    return torch.empty_like(weight)


@torch.library.custom_op("lgcn_b200::propagate_backward", mutates_args=())
def propagate_backward(grad_out: torch.Tensor, handle: int) -> torch.Tensor:
    m = _model(handle)
    grad = torch.empty_like(grad_out)
    m._horner_into(grad_out.contiguous(), grad_mode=1, reg_coef=0.0, grad=grad, cnt=m._zero_cnt(),
                   edge_w=getattr(m, "_op_drop_bwd", None))
    return grad


@propagate_backward.register_fake
def _(grad_out, handle):
    return torch.empty_like(grad_out)


def _setup_context(ctx, inputs, output):
    ctx.handle = inputs[1]


def _backward(ctx, grad_out):
    return propagate_backward(grad_out, ctx.handle), None


propagate.register_autograd(_backward, setup_context=_setup_context)
