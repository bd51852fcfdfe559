"""Forward-only guard shared by the reference-named entry points.

The kernels build no autograd graph ("Autograd": the north-star is the forward pass).  The
reference differentiates through several of these functions in ``train.py``; silently returning tensors without
history there would train on wrong gradients, so a call that autograd would have to record -- grad mode on and an
input that requires grad -- fails loudly instead.  Under ``torch.no_grad()`` / ``model.eval()`` inference, or with
detached inputs, the functions run as usual.
"""
from __future__ import annotations

import functools

import torch


def forward_only(fn):
    """Run ``fn`` without autograd; refuse calls whose inputs would need a backward pass."""

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if torch.is_grad_enabled():
            for a in list(args) + list(kwargs.values()):
                if isinstance(a, torch.Tensor) and a.requires_grad:
                    raise RuntimeError(
                        f"ogmm_b200.{fn.__qualname__} is forward-only: it got an input that requires grad while autograd "
                        "is recording, and its kernels have no backward.  Call it under torch.no_grad() (inference) or "
                        "detach the input; for training keep the reference function for this call.")
        with torch.no_grad():
            return fn(*args, **kwargs)

    return wrapper
