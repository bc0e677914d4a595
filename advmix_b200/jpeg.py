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


class _PinnedRing:
    """Reusable pinned host buffers.  A buffer is handed out again only after the copy that read it completed; a buffer
    that was handed out but not uploaded yet (a PlannedBatch still waiting for to_device()) is never reused - the ring
    grows instead.  Thread-safe."""
    _PENDING = "pending"

    def __init__(self, depth=4):
        import threading
        self.slots, self.next, self.depth, self.lock = [], 0, depth, threading.Lock()

    def get(self, nbytes):
        with self.lock:
            slot = None
            if len(self.slots) >= self.depth:
                for _ in range(len(self.slots)):
                    cand = self.slots[self.next]
                    self.next = (self.next + 1) % len(self.slots)
                    if cand[1] is not self._PENDING:
                        slot = cand
                        break
            if slot is None:                      # ring not full yet, or every buffer is waiting for its upload
                slot = [torch.empty(max(int(nbytes), 1), dtype=torch.uint8).pin_memory(), None]
                self.slots.append(slot)
            ev = slot[1]
            slot[1] = self._PENDING
        if ev is not None and ev is not self._PENDING:
            ev.synchronize()
        if slot[0].numel() < nbytes:
            slot[0] = torch.empty(int(nbytes), dtype=torch.uint8).pin_memory()
        return slot

    @staticmethod
    def mark(slot):
        ev = torch.cuda.Event()
        ev.record()
        slot[1] = ev


_plan_ring = _PinnedRing()


class PlannedBatch:
    """Host-side parse of an EncodedBatch (advmix_jpeg_plan_h): geometry, tables and buffer layout."""

    def __init__(self, enc):
        lib = _lib.load()
        self.enc = enc
        B = self.B = len(enc)
        stride = int(lib.advmix_jpeg_plan_stride())
        self._slot = _plan_ring.get(max(B, 1) * stride)
        self.plans_h = self._slot[0][:max(B, 1) * stride].view(max(B, 1), stride)
        totals = np.zeros(3, np.int64)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = lib.advmix_jpeg_plan_h(C.c_void_p(enc.host.data_ptr()), vp(enc.offsets), vp(enc.lengths), B,
                                    C.c_void_p(self.plans_h.data_ptr()), vp(totals[0:]), vp(totals[1:]), vp(totals[2:]))
        pv = self.plans_h.numpy()[:B]
        f64 = lambda o: pv[:, o:o + 8].copy().view(np.int64)[:, 0]
        f32 = lambda o: pv[:, o:o + 4].copy().view(np.int32)[:, 0]
        if rc != 0:
            bad = [(int(i), STATUS.get(int(st), "?")) for i, st in enumerate(f32(156)) if st != 0]
            raise _lib.AdvmixError("advmix_jpeg_plan_h: files that cannot be decoded on the device: %s" % bad[:8])
        self.out_off, self.out_pitch, self.widths, self.heights = f64(16), f64(24), f32(32), f32(36)
        plane_w = pv[:, 224:240].copy().view(np.int32)
        plane_h = pv[:, 240:256].copy().view(np.int32)
        self.max_blocks = int(((plane_w // 8) * (plane_h // 8)).sum(1).max()) if B else 0
        self.max_pixels = int((self.widths.astype(np.int64) * self.heights).max()) if B else 0
        self.out_bytes, self.coef_elems, self.plane_bytes = (int(t) for t in totals)
        self.any_restart = int((f32(52) != 0).any()) if B else 0
        self.files_bytes = int(enc.host.numel())
        self.ws_bytes = ((self.coef_elems * 2 + 255) & ~255) + ((self.plane_bytes + 255) & ~255) + self.files_bytes + 64

    def to_device(self, device="cuda"):
        """(files, plans) on the device: the only bytes of the images that cross PCIe."""
        dev = torch.device(device)
        files_d = torch.empty(self.enc.host.numel(), dtype=torch.uint8, device=dev)
        files_d.copy_(self.enc.host, non_blocking=True)
        plans_d = torch.empty(self.plans_h.shape, dtype=torch.uint8, device=dev)
        plans_d.copy_(self.plans_h, non_blocking=True)
        _PinnedRing.mark(self._slot)
        return files_d, plans_d

    @property
    def h2d_bytes(self):
        return int(self.enc.nbytes + self.plans_h.numel())


def decode_planned(pb, files_d, plans_d, color="bgr", out=None):
    """Device part of the decode: encoded bytes resident in HBM -> transforms.SourceBatch."""
    lib = _lib.load()
    dev = files_d.device
    if out is None:
        out = torch.empty(max(pb.out_bytes, 16), dtype=torch.uint8, device=dev)
    assert out.numel() >= pb.out_bytes
    ws = _workspace(pb.ws_bytes, dev)
    _lib.check(lib.advmix_jpeg_decode(_lib.ptr(files_d), _lib.ptr(plans_d), pb.B, pb.max_blocks, pb.max_pixels, _lib.ptr(out),
                                      _lib.ptr(ws), pb.ws_bytes, pb.coef_elems, pb.plane_bytes, pb.files_bytes, pb.any_restart,
                                      int(color == "bgr"), _lib.stream_ptr()), "advmix_jpeg_decode")
    # the SourceBatch geometry is sliced out of the plan records already on the device (no extra H2D)
    col = lambda o, n, dt: plans_d[:pb.B, o:o + n].contiguous().view(dt).view(-1)
    geom = (col(16, 8, torch.int64), col(36, 4, torch.int32), col(32, 4, torch.int32), col(24, 8, torch.int64))
    return SourceBatch(out, *geom)


def decode_batch(files, color="bgr", device="cuda", out=None):
    """files: list of bytes objects (whole JPEG files) or an EncodedBatch.
    color: 'bgr' (cv2.imread order) or 'rgb'.  Returns a transforms.SourceBatch on `device`."""
    enc = files if isinstance(files, EncodedBatch) else EncodedBatch(files)
    pb = PlannedBatch(enc)
    files_d, plans_d = pb.to_device(device)
    decode_batch.last_h2d_bytes = pb.h2d_bytes
    return decode_planned(pb, files_d, plans_d, color, out)


def encode_batch_device(images, quality=75, stride=None, out=None):
    """images: uint8 [n, H, W, 3] RGB on the device -> (files uint8 [n, stride] on the device, lengths int32 [n]).
    File i is files[i, :lengths[i]], byte-identical to PIL's Image.fromarray(images[i]).save(f, "JPEG", quality=quality)
    (tools/make_datasets.py:45 uses PIL's default quality 75).  lengths[i] == -1: file i does not fit `stride` bytes
    (default 1.5 bytes per pixel, ~10 x a typical file; encode_batch retries those with the hard upper bound).
    out: optional (files, lengths) tensors to write into (asynchronous pipelines that read the lengths back later)."""
    lib = _lib.load()
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] != 3 or not images.is_cuda:
        raise TypeError("encode_batch_device expects a uint8 CUDA tensor [n, H, W, 3]")
    images = images.contiguous()
    n, H, W = int(images.shape[0]), int(images.shape[1]), int(images.shape[2])
    if out is not None:
        files, lengths = out
        if (files.dtype != torch.uint8 or lengths.dtype != torch.int32 or files.dim() != 2 or files.shape[0] != n or
                lengths.numel() != n or not files.is_contiguous() or not lengths.is_contiguous() or files.device != images.device):
            raise TypeError("encode_batch_device: out = (uint8 [n, stride], int32 [n]) contiguous tensors on the images' device")
        stride = int(files.shape[1])
    else:
        if stride is None:
            stride = H * W * 3 // 2 + 4096
        stride = (int(stride) + 15) & ~15
        files = torch.empty((n, stride), dtype=torch.uint8, device=images.device)
        lengths = torch.empty(n, dtype=torch.int32, device=images.device)
    ws_bytes = int(lib.advmix_jpeg_encode_workspace_bytes(n, H, W))
    ws = _workspace(ws_bytes, images.device)
    _lib.check(lib.advmix_jpeg_encode_u8c3(_lib.ptr(images), n, H, W, int(quality), _lib.ptr(files), stride, _lib.ptr(lengths),
                                           _lib.ptr(ws), ws_bytes, _lib.stream_ptr()), "advmix_jpeg_encode_u8c3")
    return files, lengths


def encode_batch(images, quality=75):
    """Like encode_batch_device, returning the n files as `bytes` objects (only the encoded bytes cross PCIe)."""
    files, lengths = encode_batch_device(images, quality)
    ln = lengths.cpu().numpy()
    if len(ln) == 0:
        return []
    if (ln < 0).any():
        # dense noise at quality ~100: retry with the hard bound (4 bytes per coefficient, every byte stuffed)
        H, W = int(images.shape[1]), int(images.shape[2])
        samples = ((H + 15) // 16 * 16) * ((W + 15) // 16 * 16) * 3 // 2
        files, lengths = encode_batch_device(images, quality, stride=623 + 8 * samples + 64)
        ln = lengths.cpu().numpy()
        if (ln < 0).any():
            raise _lib.AdvmixError("advmix_jpeg_encode_u8c3 failed for images %s" % np.nonzero(ln < 0)[0][:8].tolist())
    packed, offsets = pack_files(files, lengths)
    off = offsets.cpu().numpy()
    host = packed[:int(off[-1])].cpu().numpy()
    return [host[off[i]:off[i + 1]].tobytes() for i in range(len(ln))]


def pack_files(files, lengths, out=None):
    """(files uint8 [n, stride], lengths int32 [n]) on the device -> (packed uint8 [capacity], offsets int64 [n + 1]) on the device:
    file i is packed[offsets[i]:offsets[i + 1]]; offsets[n] is the number of bytes worth copying to the host.
    out: optional (packed, offsets) tensors to write into (packed.numel() >= the sum of the lengths)."""
    lib = _lib.load()
    n, stride = int(files.shape[0]), int(files.shape[1])
    if out is not None:
        packed, offsets = out
    else:
        packed = torch.empty(max(n * stride, 16), dtype=torch.uint8, device=files.device)
        offsets = torch.empty(n + 1, dtype=torch.int64, device=files.device)
    _lib.check(lib.advmix_pack_files(_lib.ptr(files), stride, _lib.ptr(lengths), n, _lib.ptr(packed), _lib.ptr(offsets), _lib.stream_ptr()),
               "advmix_pack_files")
    return packed, offsets
