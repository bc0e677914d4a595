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


# ---- fused chain + mix: the chains are recomputed from the uint8 crop inside the mix kernel ----------
_plan_ws = {}


def autoaug_plan(crop_u8, ops, mags):
    """Per-image autoaug plans (histogram + LUT planning half of chains.autoaug) for the fused kernels:
    uint8 [B, advmix_autoaug_plan_bytes(1)].  ops int32 [B,2] / mags float32 [B,2] as in chains.autoaug."""
    lib = _lib.load()
    crop_u8 = crop_u8.contiguous()
    B, H, W, _ = crop_u8.shape
    dev = crop_u8.device
    ops = ops if torch.is_tensor(ops) else torch.as_tensor(ops)
    mags = mags if torch.is_tensor(mags) else torch.as_tensor(mags)
    ops = ops.to(dev, torch.int32).contiguous()
    mags = mags.to(dev, torch.float32).contiguous()
    plans = torch.empty((B, int(lib.advmix_autoaug_plan_bytes(1))), dtype=torch.uint8, device=dev)
    need = B * 768 * 4
    ws = _plan_ws.get(str(dev))
    if ws is None or ws.numel() < need:
        ws = _plan_ws[str(dev)] = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    _lib.check(lib.advmix_autoaug_plan_u8c3(_lib.ptr(crop_u8), _lib.ptr(ops), _lib.ptr(mags), _lib.ptr(plans), B, H, W,
                                            _lib.ptr(ws), need, _lib.stream_ptr()), "advmix_autoaug_plan_u8c3")
    return plans


def _gm(gm_params, dev):
    if gm_params is None:
        return None
    g = gm_params if torch.is_tensor(gm_params) else torch.as_tensor(gm_params)
    return g.to(dev, torch.int32).contiguous()


def _lut(lut, dev):
    from .transforms import normalize_lut
    return normalize_lut(device=dev) if lut is None else lut


def chains_g_input(crop_u8, plans, gm_params, dtype=torch.float32, lut=None):
    """G_input = torch.cat([clean, autoaug, gridmask], dim=1) (function.py:137) straight from the uint8 crop: [B,9,H,W]."""
    lib = _lib.load()
    crop_u8 = crop_u8.contiguous()
    B, H, W, _ = crop_u8.shape
    out = torch.empty((B, 9, H, W), dtype=dtype, device=crop_u8.device)
    _lib.check(lib.advmix_chains_emit_u8c3(_lib.ptr(crop_u8), _lib.ptr(plans), _lib.ptr(_gm(gm_params, crop_u8.device)),
                                           _lib.ptr(_lut(lut, crop_u8.device)), _lib.ptr(out), B, H, W, _lib.dtype_code(dtype),
                                           _lib.stream_ptr()), "advmix_chains_emit_u8c3")
    return out


class _ChainMix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w_or_logits, apply_softmax, crop_u8, plans, gm, lut, out_dtype):
        lib = _lib.load()
        wl = w_or_logits.contiguous()
        if wl.dtype not in (torch.float32, torch.bfloat16):
            wl = wl.to(torch.float32)
        B, H, W, _ = crop_u8.shape
        if tuple(wl.shape) != (B, 3, H, W):
            raise ValueError("chain_mix: weights must be [B,3,H,W]=%s, got %s" % ((B, 3, H, W), tuple(wl.shape)))
        out = torch.empty((B, 3, H, W), dtype=out_dtype, device=crop_u8.device)
        _lib.check(lib.advmix_chainmix_fwd(_lib.ptr(crop_u8), _lib.ptr(plans), _lib.ptr(gm), _lib.ptr(lut), _lib.ptr(wl),
                                           _lib.dtype_code(wl.dtype), int(apply_softmax), _lib.ptr(out), _lib.dtype_code(out_dtype),
                                           None, B, H, W, _lib.stream_ptr()), "advmix_chainmix_fwd")
        ctx.apply_softmax = bool(apply_softmax)
        ctx.in_dtype = w_or_logits.dtype
        ctx.save_for_backward(wl, crop_u8, lut, *([plans] if plans is not None else []), *([gm] if gm is not None else []))
        ctx.has = (plans is not None, gm is not None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        saved = list(ctx.saved_tensors)
        wl, crop_u8, lut = saved[:3]
        rest = saved[3:]
        plans = rest.pop(0) if ctx.has[0] else None
        gm = rest.pop(0) if ctx.has[1] else None
        B, H, W, _ = crop_u8.shape
        go = grad_out.contiguous()
        if go.dtype not in (torch.float32, torch.bfloat16):
            go = go.to(torch.float32)
        gw = torch.empty((B, 3, H, W), dtype=torch.float32, device=crop_u8.device)
        _lib.check(lib.advmix_chainmix_bwd(_lib.ptr(crop_u8), _lib.ptr(plans), _lib.ptr(gm), _lib.ptr(lut), _lib.ptr(wl),
                                           _lib.dtype_code(wl.dtype), int(ctx.apply_softmax), _lib.ptr(go), _lib.dtype_code(go.dtype),
                                           _lib.ptr(gw), B, H, W, _lib.stream_ptr()), "advmix_chainmix_bwd")
        return (gw.to(ctx.in_dtype), None, None, None, None, None, None)


def chain_mix(crop_u8, plans, gm_params, mix_weight, out_dtype=torch.float32, lut=None):
    """tmp of function.py:142-144 with the three reference-actual chains recomputed from the uint8 crop (no chain tensors
    are read): bit-identical to mix([clean, autoaug, gridmask], mix_weight).  Differentiable w.r.t. mix_weight."""
    crop_u8 = crop_u8.contiguous()
    return _ChainMix.apply(mix_weight, False, crop_u8, plans, _gm(gm_params, crop_u8.device), _lut(lut, crop_u8.device), out_dtype)


def chain_mix_from_logits(crop_u8, plans, gm_params, logits, out_dtype=torch.float32, lut=None):
    """Fuses F.softmax(G_out, dim=1) (function.py:138) as well; the backward pass recomputes the softmax from the logits."""
    crop_u8 = crop_u8.contiguous()
    return _ChainMix.apply(logits, True, crop_u8, plans, _gm(gm_params, crop_u8.device), _lut(lut, crop_u8.device), out_dtype)


class _MixU8(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w_or_logits, apply_softmax, lut, out_dtype, *chains):
        lib = _lib.load()
        chains = [c.contiguous() for c in chains]
        wl = w_or_logits.contiguous()
        if wl.dtype not in (torch.float32, torch.bfloat16):
            wl = wl.to(torch.float32)
        B, H, W, _ = chains[0].shape
        K = len(chains)
        for c in chains:
            if c.shape != chains[0].shape or c.dtype != torch.uint8 or not c.is_cuda:
                raise ValueError("mix_u8: all chains must be CUDA uint8 tensors [B,H,W,3] of one shape")
        if tuple(wl.shape) != (B, K, H, W):
            raise ValueError("mix_u8: weights must be [B,K,H,W]=%s, got %s" % ((B, K, H, W), tuple(wl.shape)))
        out = torch.empty((B, 3, H, W), dtype=out_dtype, device=chains[0].device)
        _lib.check(lib.advmix_mix_u8_fwd(_xptrs(chains), _lib.ptr(lut), _lib.ptr(wl), _lib.dtype_code(wl.dtype), int(apply_softmax),
                                         _lib.ptr(out), _lib.dtype_code(out_dtype), None, B, K, H, W, _lib.stream_ptr()),
                   "advmix_mix_u8_fwd")
        ctx.apply_softmax = bool(apply_softmax)
        ctx.in_dtype = w_or_logits.dtype
        ctx.save_for_backward(wl, lut, *chains)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        wl, lut, *chains = ctx.saved_tensors
        B, H, W, _ = chains[0].shape
        K = len(chains)
        go = grad_out.contiguous()
        if go.dtype not in (torch.float32, torch.bfloat16):
            go = go.to(torch.float32)
        gw = torch.empty((B, K, H, W), dtype=torch.float32, device=go.device)
        _lib.check(lib.advmix_mix_u8_bwd(_xptrs(chains), _lib.ptr(lut), _lib.ptr(wl), _lib.dtype_code(wl.dtype), int(ctx.apply_softmax),
                                         _lib.ptr(go), _lib.dtype_code(go.dtype), _lib.ptr(gw), B, K, H, W, _lib.stream_ptr()),
                   "advmix_mix_u8_bwd")
        return (gw.to(ctx.in_dtype), None, None, None) + (None,) * K


def mix_u8(chains_u8, mix_weight, out_dtype=torch.float32, lut=None):
    """The mix over K <= 4 uint8 HWC chain images (e.g. corruption chains of the 15x5 set), normalised in registers:
    equals mix([to_tensor_normalize(c) for c in chains_u8], mix_weight) while reading a quarter of the chain bytes."""
    return _MixU8.apply(mix_weight, False, _lut(lut, chains_u8[0].device), out_dtype, *chains_u8)


def mix_u8_from_logits(chains_u8, logits, out_dtype=torch.float32, lut=None):
    return _MixU8.apply(logits, True, _lut(lut, chains_u8[0].device), out_dtype, *chains_u8)
