"""Device JPEG decode of the source images: the replacement for
`cv2.imread(image_file, cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION)` at
lib/dataset/JointsDataset.py:148.  Encoded files in, a transforms.SourceBatch (decoded uint8 HWC images
resident in HBM, ready for warp_affine) out.  Bit-identical to cv2.imdecode for the baseline files it
accepts; anything else raises (no CPU fallback).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .transforms import SourceBatch

STATUS = {0: "ok", 1: "corrupt or truncated", 2: "progressive JPEG", 3: "unsupported JPEG flavour"}


class EncodedBatch:
    """B encoded files packed into ONE pinned host buffer (16-byte aligned offsets)."""

    def __init__(self, files):
        lens = np.array([len(f) for f in files], np.int64)
        offs = np.zeros(len(files), np.int64)
        if len(files):
            offs[1:] = np.cumsum((lens[:-1] + 15) & ~15)
        total = int(offs[-1] + ((lens[-1] + 15) & ~15)) if len(files) else 0
        self.host = torch.empty(max(total, 16), dtype=torch.uint8).pin_memory()
        hv = self.host.numpy()
        for f, o in zip(files, offs):
            hv[o:o + len(f)] = np.frombuffer(f, np.uint8)
        self.offsets, self.lengths, self.nbytes = offs, lens, total

    def __len__(self):
        return len(self.offsets)


_ws = {}


def _workspace(nbytes, device):
    key = str(device)
    b = _ws.get(key)
    if b is None or b.numel() < nbytes:
        b = _ws[key] = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
    return b


def decode_batch(files, color="bgr", device="cuda", out=None):
    """files: list of bytes objects (whole JPEG files) or an EncodedBatch.
    color: 'bgr' (cv2.imread order) or 'rgb'.  Returns a transforms.SourceBatch on `device`."""
    lib = _lib.load()
    enc = files if isinstance(files, EncodedBatch) else EncodedBatch(files)
    B = len(enc)
    stride = int(lib.advmix_jpeg_plan_stride())
    plans_h = torch.empty((max(B, 1), stride), dtype=torch.uint8).pin_memory()
    totals = np.zeros(3, np.int64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.advmix_jpeg_plan_h(C.c_void_p(enc.host.data_ptr()), vp(enc.offsets), vp(enc.lengths), B,
                                C.c_void_p(plans_h.data_ptr()), vp(totals[0:]), vp(totals[1:]), vp(totals[2:]))
    pv = plans_h.numpy()[:B]
    if rc != 0:
        status = pv[:, 156:160].copy().view(np.int32)[:, 0]
        bad = [(int(i), STATUS.get(int(s), "?")) for i, s in enumerate(status) if s != 0]
        raise _lib.AdvmixError("advmix_jpeg_plan_h: files that cannot be decoded on the device: %s" % bad[:8])
    f64 = lambda o: pv[:, o:o + 8].copy().view(np.int64)[:, 0]
    f32 = lambda o: pv[:, o:o + 4].copy().view(np.int32)[:, 0]
    out_off, out_pitch, widths, heights = f64(16), f64(24), f32(32), f32(36)
    plane_w = pv[:, 224:240].copy().view(np.int32)
    plane_h = pv[:, 240:256].copy().view(np.int32)
    max_blocks = int(((plane_w // 8) * (plane_h // 8)).sum(1).max()) if B else 0
    max_pixels = int((widths.astype(np.int64) * heights).max()) if B else 0
    out_bytes, coef_elems, plane_bytes = (int(t) for t in totals)
    dev = torch.device(device)
    files_d = torch.empty(enc.host.numel(), dtype=torch.uint8, device=dev)
    files_d.copy_(enc.host, non_blocking=True)
    plans_d = torch.empty(plans_h.shape, dtype=torch.uint8, device=dev)
    plans_d.copy_(plans_h, non_blocking=True)
    if out is None:
        out = torch.empty(max(out_bytes, 16), dtype=torch.uint8, device=dev)
    assert out.numel() >= out_bytes
    files_bytes = int(enc.host.numel())
    ws_bytes = ((coef_elems * 2 + 255) & ~255) + ((plane_bytes + 255) & ~255) + files_bytes + 64
    any_restart = int((f32(52) != 0).any()) if B else 0
    ws = _workspace(ws_bytes, dev)
    _lib.check(lib.advmix_jpeg_decode(_lib.ptr(files_d), _lib.ptr(plans_d), B, max_blocks, max_pixels, _lib.ptr(out),
                                      _lib.ptr(ws), ws_bytes, coef_elems, plane_bytes, files_bytes, any_restart,
                                      int(color == "bgr"),
                                      _lib.stream_ptr()), "advmix_jpeg_decode")
    decode_batch.last_h2d_bytes = int(enc.nbytes + plans_h.numel())
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
    return SourceBatch(out, t(out_off, torch.int64), t(heights, torch.int32), t(widths, torch.int32), t(out_pitch, torch.int64))
