"""The AdvMix per-pixel convex mix (lib/core/function.py:137-146) as a fused CUDA op with
autograd to the weights / logits (the G step back-propagates through it, :158-164)."""
import ctypes as C

import torch

from . import _lib


def _xptrs(inputs):
    arr = (C.c_void_p * len(inputs))(*[t.data_ptr() for t in inputs])
    return arr


def _check_inputs(inputs, w):
    x0 = inputs[0]
    B, Cc, H, W = x0.shape
    K = len(inputs)
    for x in inputs:
        if x.shape != x0.shape or x.dtype != x0.dtype or not x.is_cuda:
            raise ValueError("mix: all chain tensors must be CUDA tensors of one shape/dtype")
    if tuple(w.shape) != (B, K, H, W):
        raise ValueError("mix: weights must be [B,K,H,W]=%s, got %s" % ((B, K, H, W), tuple(w.shape)))
    return B, K, Cc, H, W


class _Mix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w_or_logits, apply_softmax, *inputs):
        lib = _lib.load()
        inputs = [x.contiguous() for x in inputs]
        wl = w_or_logits.to(torch.float32).contiguous()
        B, K, Cc, H, W = _check_inputs(inputs, wl)
        out = torch.empty_like(inputs[0])
        w_out = torch.empty_like(wl) if apply_softmax else None
        _lib.check(lib.advmix_mix_fwd(_xptrs(inputs), _lib.ptr(wl), int(apply_softmax), _lib.ptr(out),
                                      _lib.ptr(w_out), B, K, Cc, H, W, _lib.dtype_code(out.dtype),
                                      _lib.stream_ptr()), "advmix_mix_fwd")
        ctx.apply_softmax = bool(apply_softmax)
        ctx.save_for_backward(w_out if apply_softmax else wl, *inputs)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        w, *inputs = ctx.saved_tensors
        B, K, Cc, H, W = _check_inputs(inputs, w)
        go = grad_out.to(inputs[0].dtype).contiguous()
        gw = torch.empty_like(w)
        _lib.check(lib.advmix_mix_bwd(_xptrs(inputs), _lib.ptr(w), _lib.ptr(go), _lib.ptr(gw),
                                      int(ctx.apply_softmax), B, K, Cc, H, W, _lib.dtype_code(go.dtype),
                                      _lib.stream_ptr()), "advmix_mix_bwd")
        return (gw, None) + (None,) * len(inputs)


def mix(inputs, mix_weight):
    """tmp = sum_k inputs[k] * mix_weight[:, k].unsqueeze(1)  (function.py:142-144).
    Differentiable w.r.t. mix_weight; the chain tensors come from the loader (no grad)."""
    return _Mix.apply(mix_weight, False, *inputs)


def mix_from_logits(inputs, logits):
    """Fuses F.softmax(G_out, dim=1) (function.py:138) with the mix."""
    return _Mix.apply(logits, True, *inputs)
