"""torch.autograd bindings of the interaction kernels: forward and backward both run in
``libmatchmaker_b200.so``; autograd only routes tensors."""
from __future__ import annotations

from typing import Optional

import torch

from . import interaction


class _MaxSim(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, d, q_mask, d_mask, docs_per_query):
        need_grad = q.requires_grad or d.requires_grad
        if need_grad:
            out, argmax = interaction.maxsim(q, d, q_mask, d_mask, docs_per_query=docs_per_query, return_argmax=True)
            ctx.save_for_backward(q, d, argmax)
            ctx.docs_per_query = docs_per_query
        else:
            out = interaction.maxsim(q, d, q_mask, d_mask, docs_per_query=docs_per_query)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        q, d, argmax = ctx.saved_tensors
        gq, gd = interaction.maxsim_bwd(q, d, grad_out, argmax, ctx.docs_per_query)
        return gq.to(q.dtype), gd.to(d.dtype), None, None, None


def maxsim(q: torch.Tensor, d: torch.Tensor, q_mask: Optional[torch.Tensor] = None,
           d_mask: Optional[torch.Tensor] = None, docs_per_query: int = 1) -> torch.Tensor:
    """Differentiable ColBERT max-sim (pairs mode); see :func:`matchmaker_b200.interaction.maxsim`."""
    return _MaxSim.apply(q, d, q_mask, d_mask, docs_per_query)
