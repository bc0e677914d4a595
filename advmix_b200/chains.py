"""Host mirror of lib/dataset/advaug.py: MixCombine's three chains ('clean', 'autoaug',
'gridmask') for a device-resident batch."""
import random

import numpy as np
import torch

from . import _lib
from . import transforms as T

OP_CODE = {"none": 0, "equalize": 1, "posterize": 2, "solarize": 3, "invert": 4, "sharpness": 5}

# (p1, op1, magnitude_idx1, p2, op2, magnitude_idx2)  - advaug.py:22-35
POLICIES = [
    (0.8, "equalize", 8, 0.6, "equalize", 3), (0.6, "posterize", 7, 0.6, "posterize", 6),
    (0.4, "equalize", 7, 0.2, "solarize", 4), (0.6, "solarize", 3, 0.6, "equalize", 7),
    (0.8, "posterize", 5, 1.0, "equalize", 2), (0.6, "equalize", 8, 0.4, "posterize", 6),
    (0.0, "equalize", 7, 0.8, "equalize", 8), (0.6, "invert", 4, 1.0, "equalize", 8),
    (0.4, "sharpness", 7, 0.6, "invert", 8), (0.4, "equalize", 7, 0.2, "solarize", 4),
    (0.6, "invert", 4, 1.0, "equalize", 8), (0.8, "equalize", 8, 0.6, "equalize", 3),
]
_RANGES = {  # advaug.py:48-63
    "posterize": np.round(np.linspace(8, 4, 10), 0).astype(int),
    "solarize": np.linspace(256, 0, 10),
    "sharpness": np.linspace(0.0, 0.9, 10),
    "equalize": [0] * 10,
    "invert": [0] * 10,
}


def _stage(op, mag_idx, sign):
    m = _RANGES[op][mag_idx]
    if op == "sharpness":
        m = 1 + m * sign
    return OP_CODE[op], float(m)


def plan_autoaug(policy_idx, coin1, coin2, sign1=1, sign2=1):
    """(ops[2], mags[2]) for one sample after the probability coin flips of SubPolicy.__call__
    (advaug.py:104-107)."""
    p1, op1, m1, p2, op2, m2 = POLICIES[policy_idx]
    o1, g1 = _stage(op1, m1, sign1) if coin1 < p1 else (0, 0.0)
    o2, g2 = _stage(op2, m2, sign2) if coin2 < p2 else (0, 0.0)
    return (o1, o2), (g1, g2)


def sample_autoaug(B, rng=random):
    """Draws exactly as ImageNetPolicy.__call__ / SubPolicy.__call__ do, from Python's `random`:
    policy_idx = randint(0, 11); coin1 = random(); [sign1 = choice([-1,1]) only if sharpness runs]; ..."""
    ops = np.zeros((B, 2), np.int32)
    mags = np.zeros((B, 2), np.float32)
    for b in range(B):
        pi = rng.randint(0, len(POLICIES) - 1)
        p1, op1, m1, p2, op2, m2 = POLICIES[pi]
        if rng.random() < p1:
            s = rng.choice([-1, 1]) if op1 == "sharpness" else 1
            ops[b, 0], mags[b, 0] = _stage(op1, m1, s)
        if rng.random() < p2:
            s = rng.choice([-1, 1]) if op2 == "sharpness" else 1
            ops[b, 1], mags[b, 1] = _stage(op2, m2, s)
    return ops, mags


def _policy_tables():
    """Per sub-policy and stage: probability, op code, magnitude for sign +1 / -1 (vectorised sampling)."""
    P = len(POLICIES)
    prob = np.zeros((P, 2)); code = np.zeros((P, 2), np.int32); mag = np.zeros((P, 2, 2), np.float32)
    for i, (p1, op1, m1, p2, op2, m2) in enumerate(POLICIES):
        for st, (p_, op, m) in enumerate(((p1, op1, m1), (p2, op2, m2))):
            prob[i, st] = p_
            code[i, st] = OP_CODE[op]
            mag[i, st, 0] = _stage(op, m, 1)[1]
            mag[i, st, 1] = _stage(op, m, -1)[1]
    return prob, code, mag


_TABLES = None


def sample_autoaug_batch(B, g):
    """Same distribution as sample_autoaug, drawn for the whole batch at once from a numpy Generator (the per-sample
    loop over Python's `random` costs ~1 ms per 256 samples on the host, more than the whole device step)."""
    global _TABLES
    if _TABLES is None:
        _TABLES = _policy_tables()
    prob, code, mag = _TABLES
    pi = g.integers(0, len(POLICIES), B)
    on = g.random((B, 2)) < prob[pi]
    sign = g.integers(0, 2, (B, 2))                                   # 0: +1, 1: -1 (only sharpness looks at it)
    ops = np.where(on, code[pi], 0).astype(np.int32)
    mags = np.where(on, mag[pi[:, None], np.arange(2)[None, :], sign], 0.0).astype(np.float32)
    return ops, mags


def sample_gridmask_batch(B, H, W, g, prob=0.7):
    """Same distribution as sample_gridmask for the whole batch at once (numpy Generator)."""
    on = ~(g.random(B) > prob)
    d = g.integers(2, min(H, W), B)
    params = np.zeros((B, 4), np.int32)
    params[:, 0] = on
    params[:, 1] = np.where(on, d, 0)
    params[:, 2] = np.where(on, (g.random(B) * d).astype(np.int64), 0)
    params[:, 3] = np.where(on, (g.random(B) * d).astype(np.int64), 0)
    return params


_ws = {}


def autoaug(images_u8, ops, mags, norm_dtype=None, want_u8=True, lut=None):
    """images uint8 [B,H,W,3] (CUDA) -> (uint8 result or None, normalised [B,3,H,W] or None)."""
    lib = _lib.load()
    images_u8 = images_u8.contiguous()
    B, H, W, _ = images_u8.shape
    dev = images_u8.device
    ops_t = (torch.as_tensor(np.ascontiguousarray(ops, dtype=np.int32)) if not torch.is_tensor(ops) else ops).to(dev, torch.int32).contiguous()
    mags_t = (torch.as_tensor(np.ascontiguousarray(mags, dtype=np.float32)) if not torch.is_tensor(mags) else mags).to(dev, torch.float32).contiguous()
    out = torch.empty_like(images_u8) if want_u8 else None
    out_n, code = None, _lib.F32
    if norm_dtype is not None:
        code = _lib.dtype_code(norm_dtype)
        out_n = torch.empty((B, 3, H, W), dtype=norm_dtype, device=dev)
        if lut is None:
            lut = T.normalize_lut(device=dev)
    nbytes = int(lib.advmix_autoaug_workspace_bytes(B, H, W))
    ws = _ws.get(str(dev))
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _ws[str(dev)] = ws
    _lib.check(lib.advmix_autoaug_u8c3(_lib.ptr(images_u8), _lib.ptr(out), _lib.ptr(out_n), _lib.ptr(lut),
                                       _lib.ptr(ops_t), _lib.ptr(mags_t), B, H, W, code, _lib.ptr(ws), nbytes,
                                       _lib.stream_ptr()), "advmix_autoaug_u8c3")
    return out, out_n


def sample_gridmask(B, H, W, prob=0.7, rng=np.random):
    """Draws of grid_aug (advaug.py:112-125) from np.random, in the reference's order."""
    params = np.zeros((B, 4), np.int32)
    for b in range(B):
        if rng.rand() > prob:
            continue
        d = rng.randint(2, min(H, W))
        params[b] = (1, d, rng.randint(d), rng.randint(d))
        rng.randint(1)  # r = np.random.randint(rotate) with rotate=1 (always 0, but consumes a draw)
    return params


def gridmask(img, params, joints=None, joints_vis=None):
    """grid_aug(mode=1) on normalised [B,3,H,W]; returns (img_out, joints_vis_out)."""
    lib = _lib.load()
    img = img.contiguous()
    B, _, H, W = img.shape
    dev = img.device
    params_t = (torch.as_tensor(np.ascontiguousarray(params, dtype=np.int32)) if not torch.is_tensor(params) else params).to(dev, torch.int32).contiguous()
    out = torch.empty_like(img)
    J, vo = 0, None
    if joints is not None:
        joints = joints.to(torch.float64).contiguous()
        joints_vis = joints_vis.to(torch.float64).contiguous()
        J = joints.shape[1]
        vo = torch.empty_like(joints_vis)
    _lib.check(lib.advmix_gridmask(_lib.ptr(img), _lib.ptr(out), _lib.ptr(params_t), _lib.ptr(joints),
                                   _lib.ptr(joints_vis), _lib.ptr(vo), B, H, W, J, _lib.dtype_code(img.dtype),
                                   _lib.stream_ptr()), "advmix_gridmask")
    return out, vo
