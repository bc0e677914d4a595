"""Host mirror of lib/utils/transforms.py + the crop of lib/dataset/JointsDataset.py for
batches resident on the GPU.  Same names and argument meaning as the reference where a
reference function exists; every function dispatches to the C ABI (no torch math).
"""
import numpy as np
import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

_lut_cache = {}


def normalize_lut(mean=IMAGENET_MEAN, std=IMAGENET_STD, device="cuda"):
    """float32 [3,256] table equal to torchvision ToTensor()+Normalize(mean,std) per byte value
    (tools/train.py:116-126).  Built with the same float32 torch ops, so it is bit-identical."""
    key = (tuple(mean), tuple(std), str(device))
    if key not in _lut_cache:
        v = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)
        m = torch.tensor(mean, dtype=torch.float32)[:, None]
        s = torch.tensor(std, dtype=torch.float32)[:, None]
        _lut_cache[key] = ((v[None, :] - m) / s).contiguous().to(device)
    return _lut_cache[key]


def get_affine_transform(center, scale, rot, output_size):
    """Batched lib/utils/transforms.py:69-101 (inv=0, shift=0) on the device.

    center: float32 [B,2]; scale: float32 or float64 [B,2] (the dtype selects numpy's promotion
    of `scale * 200.0`, see include/advmix_b200.h); rot: float64 [B] degrees; output_size (w, h).
    Returns float64 [B,2,3] forward matrices."""
    lib = _lib.load()
    center = center.to(torch.float32).contiguous()
    scale_f32 = scale.dtype != torch.float64
    scale = scale.to(torch.float64).contiguous()
    rot = rot.to(torch.float64).contiguous()
    B = center.shape[0]
    M = torch.empty((B, 2, 3), dtype=torch.float64, device=center.device)
    _lib.check(lib.advmix_affine_matrices(_lib.ptr(center), _lib.ptr(scale), int(scale_f32), _lib.ptr(rot),
                                          _lib.ptr(M), B, int(output_size[0]), int(output_size[1]),
                                          _lib.stream_ptr()), "advmix_affine_matrices")
    return M


class SourceBatch:
    """B source images (uint8 HWC, any sizes) packed in one device buffer."""

    def __init__(self, buffer, offsets, heights, widths, pitches):
        self.buffer, self.offsets, self.heights, self.widths, self.pitches = buffer, offsets, heights, widths, pitches

    @classmethod
    def from_numpy(cls, images, device="cuda", bgr_to_rgb=False):
        """Packs the images with 16-byte aligned rows (pitch = 3*W rounded up to 16) and offsets, which is what the
        crop kernel's staged fast path needs; unaligned sources still work through its direct path.
        bgr_to_rgb: reverse the channel order while packing (cfg.DATASET.COLOR_RGB: cv2.cvtColor(BGR2RGB) at
        JointsDataset.py:151-152; device-decoded sources ask the decoder for RGB instead)."""
        offs, hs, ws, ps, total = [], [], [], [], 0
        for im in images:
            assert im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3
            pitch = (im.shape[1] * 3 + 15) // 16 * 16
            offs.append(total)
            hs.append(im.shape[0]); ws.append(im.shape[1]); ps.append(pitch)
            total += (im.shape[0] * pitch + 255) // 256 * 256
        host = torch.zeros(max(total, 16), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        hv = host.numpy()
        for im, o, p in zip(images, offs, ps):
            H, W = im.shape[0], im.shape[1]
            hv[o:o + H * p].reshape(H, p)[:, :3 * W] = np.ascontiguousarray(im[:, :, ::-1] if bgr_to_rgb else im).reshape(H, 3 * W)
        dev = torch.device(device)
        return cls(host.to(dev, non_blocking=True), torch.tensor(offs, dtype=torch.int64, device=dev),
                   torch.tensor(hs, dtype=torch.int32, device=dev), torch.tensor(ws, dtype=torch.int32, device=dev),
                   torch.tensor(ps, dtype=torch.int64, device=dev))

    @classmethod
    def from_tensor(cls, images):
        """images: uint8 [B,H,W,3] device tensor (all the same size)."""
        B, H, W, _ = images.shape
        images = images.contiguous()
        dev = images.device
        return cls(images.view(-1), torch.arange(B, dtype=torch.int64, device=dev) * (H * W * 3),
                   torch.full((B,), H, dtype=torch.int32, device=dev), torch.full((B,), W, dtype=torch.int32, device=dev),
                   torch.full((B,), W * 3, dtype=torch.int64, device=dev))

    def __len__(self):
        return self.offsets.shape[0]


def source_row_ranges(center, scale, rot, flip, heights, output_size, margin=3):
    """Host-side (numpy, vectorised) conservative range of source rows each crop reads: the destination
    rectangle mapped back through the same 3-point similarity as get_affine_transform, +- `margin` rows.
    Flips are horizontal, so they do not change the rows.  Returns int32 (row_lo, row_hi)."""
    center = np.asarray(center, np.float64); scale = np.asarray(scale, np.float64)
    rot = np.asarray(rot, np.float64); heights = np.asarray(heights)
    w, h = float(output_size[0]), float(output_size[1])
    src_w = scale[:, 0] * 200.0
    k = src_w / w                                    # source pixels per destination pixel
    sn, cs = np.sin(np.pi * rot / 180), np.cos(np.pi * rot / 180)
    # source = center + R(rot) * k * (dst - dst_centre); rows = y component
    dx = np.array([-w / 2, w / 2, w / 2, -w / 2])[None, :]
    dy = np.array([-h / 2, -h / 2, h / 2, h / 2])[None, :]
    ys = center[:, 1:2] + k[:, None] * (sn[:, None] * dx + cs[:, None] * dy)
    lo = np.floor(ys.min(1)).astype(np.int64) - margin
    hi = np.ceil(ys.max(1)).astype(np.int64) + margin + 2
    lo = np.clip(lo, 0, heights); hi = np.clip(hi, 0, heights)
    return lo.astype(np.int32), np.maximum(hi, lo).astype(np.int32)


def source_boxes(center, scale, rot, flip, heights, widths, pitches, output_size, margin=3):
    """Host-side (numpy, vectorised) conservative box of source bytes each crop reads: the destination
    rectangle mapped back through the same 3-point similarity as get_affine_transform, +- `margin` pixels,
    mirrored for flipped samples (the crop samples the `[:, ::-1]` view), columns widened to 16-byte chunks.
    Returns int32 (row_lo, row_hi, byte_lo, byte_hi) and the float32 source quadrilateral [B,4,2] (x, y)."""
    center = np.asarray(center, np.float64); scale = np.asarray(scale, np.float64)
    rot = np.asarray(rot, np.float64); heights = np.asarray(heights); widths = np.asarray(widths)
    pitches = np.asarray(pitches, np.int64)
    w, h = float(output_size[0]), float(output_size[1])
    k = scale[:, 0] * 200.0 / w                      # source pixels per destination pixel
    sn, cs = np.sin(np.pi * rot / 180), np.cos(np.pi * rot / 180)
    dx = np.array([-w / 2, w / 2, w / 2, -w / 2])[None, :]
    dy = np.array([-h / 2, -h / 2, h / 2, h / 2])[None, :]
    xs = center[:, 0:1] + k[:, None] * (cs[:, None] * dx - sn[:, None] * dy)
    ys = center[:, 1:2] + k[:, None] * (sn[:, None] * dx + cs[:, None] * dy)
    lo = np.clip(np.floor(ys.min(1)).astype(np.int64) - margin, 0, heights)
    hi = np.clip(np.ceil(ys.max(1)).astype(np.int64) + margin + 2, 0, heights)
    xlo = np.floor(xs.min(1)).astype(np.int64) - margin
    xhi = np.ceil(xs.max(1)).astype(np.int64) + margin + 2
    f = np.asarray(flip).astype(bool)
    xlo, xhi = np.where(f, widths - xhi, xlo), np.where(f, widths - xlo, xhi)      # view column v = source column W-1-v
    quad = np.stack([np.where(f[:, None], widths[:, None] - 1 - xs, xs), ys], -1).astype(np.float32)
    xlo = np.clip(xlo, 0, widths); xhi = np.clip(xhi, 0, widths)
    blo = (3 * xlo) & ~15
    bhi = np.minimum((3 * xhi + 15) & ~15, pitches & ~15)
    blo = np.minimum(blo, bhi)
    return lo.astype(np.int32), np.maximum(hi, lo).astype(np.int32), blo.astype(np.int32), bhi.astype(np.int32), quad


class HostSourceBatch:
    """Decoded uint8 sources in ONE pinned host buffer mirroring a device SourceBatch layout; `upload_rows`
    sends only the rows the crops of this step read (advmix_h2d_source_rows)."""

    def __init__(self, host_buffer, device_batch, offsets_h, pitches_h):
        self.host, self.dev = host_buffer, device_batch
        self.offsets_h = np.ascontiguousarray(offsets_h, np.int64)
        self.pitches_h = np.ascontiguousarray(pitches_h, np.int64)

    @classmethod
    def from_tensor(cls, images_host_pinned, device="cuda"):
        """images_host_pinned: pinned uint8 [B,H,W,3] host tensor."""
        B, H, W, _ = images_host_pinned.shape
        dev = torch.empty(images_host_pinned.shape, dtype=torch.uint8, device=device)
        return cls(images_host_pinned.view(-1), SourceBatch.from_tensor(dev), np.arange(B, dtype=np.int64) * (H * W * 3),
                   np.full(B, W * 3, np.int64))

    def upload_rows(self, row_lo, row_hi):
        import ctypes as C
        lib = _lib.load()
        lo = np.ascontiguousarray(row_lo, np.int32); hi = np.ascontiguousarray(row_hi, np.int32)
        nbytes = int(((hi - lo).astype(np.int64) * self.pitches_h).sum())
        if nbytes >= 0.7 * self.host.numel():
            # the crops touch most rows anyway: one big copy beats B small ones
            self.dev.buffer.copy_(self.host, non_blocking=True)
            return int(self.host.numel())
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(lib.advmix_h2d_source_rows(C.c_void_p(self.host.data_ptr()), _lib.ptr(self.dev.buffer), vp(self.offsets_h),
                                              vp(self.pitches_h), vp(lo), vp(hi), len(lo), _lib.stream_ptr()),
                   "advmix_h2d_source_rows")
        return nbytes

    def upload_boxes(self, row_lo, row_hi, byte_lo, byte_hi, quad):
        """Send only the source bytes the crops of this step read (per row: the extent of the crop's source
        quadrilateral, 16-byte chunks), in one launch (advmix_h2d_source_boxes: the device gathers them out of
        the pinned buffer).  Falls back to one bulk copy when the boxes cover most of the buffer or rows are
        not 16-byte aligned.  Returns an upper bound of the bytes sent (the bounding boxes); the exact count
        accumulates in `self.bytes_sent` (device uint64)."""
        lib = _lib.load()
        nbytes = int(((row_hi - row_lo).astype(np.int64) * (byte_hi - byte_lo)).sum())
        aligned = bool(np.all(self.pitches_h % 16 == 0) and np.all(self.offsets_h % 16 == 0))
        if nbytes >= 0.85 * self.host.numel() or not aligned or not self.host.is_pinned():
            self.dev.buffer.copy_(self.host, non_blocking=True)
            return int(self.host.numel())
        B = len(row_lo)
        # descriptor ring: the device reads the (pinned) descriptors when the launch executes, which can be after
        # this call returned, so a buffer is reused only once the launch that read it has completed
        ring = getattr(self, "_box_ring", None)
        if ring is None or ring[0][0].shape[0] < B:
            ring = self._box_ring = [[torch.empty((B, 8), dtype=torch.int64).pin_memory(), None] for _ in range(4)]
            self.bytes_sent = torch.zeros(1, dtype=torch.int64, device=self.dev.buffer.device)
            self._box_next = 0
        slot = ring[self._box_next]
        self._box_next = (self._box_next + 1) % len(ring)
        if slot[1] is not None:
            slot[1].synchronize()
        boxes = slot[0]
        bx = boxes.numpy()
        bx[:B, 0] = self.offsets_h; bx[:B, 1] = self.pitches_h
        v32 = bx[:B, 2:4].view(np.int32)                # row_lo, row_hi, byte_lo, byte_hi
        v32[:, 0] = row_lo; v32[:, 1] = row_hi; v32[:, 2] = byte_lo; v32[:, 3] = byte_hi
        q = bx[:B, 4:].view(np.float32)                 # qx[4], qy[4]
        q[:, :4] = quad[:, :, 0]; q[:, 4:] = quad[:, :, 1]
        import ctypes as C
        _lib.check(lib.advmix_h2d_source_boxes(C.c_void_p(self.host.data_ptr()), _lib.ptr(self.dev.buffer),
                                               C.c_void_p(boxes.data_ptr()), B, int((row_hi - row_lo).max()),
                                               _lib.ptr(self.bytes_sent), _lib.stream_ptr()), "advmix_h2d_source_boxes")
        slot[1] = torch.cuda.Event()
        slot[1].record()
        return nbytes


def warp_affine(src, trans, output_size, flip=None, want_u8=True, norm_dtype=None, lut=None):
    """cv2.warpAffine(img, trans, (w,h), flags=INTER_LINEAR) for a SourceBatch
    (lib/dataset/JointsDataset.py:190-195), optionally on the `[:, ::-1, :]` flipped view
    (:184-188) and fused with ToTensor()+Normalize() (:331-332).

    Returns (u8 [B,h,w,3] or None, normalised [B,3,h,w] or None)."""
    lib = _lib.load()
    B = len(src)
    w, h = int(output_size[0]), int(output_size[1])
    dev = src.buffer.device
    trans = trans.to(torch.float64).contiguous()
    assert trans.shape == (B, 2, 3)
    flip_t = None if flip is None else flip.to(torch.uint8).contiguous()
    out_u8 = torch.empty((B, h, w, 3), dtype=torch.uint8, device=dev) if want_u8 else None
    out_n = None
    code = _lib.F32
    if norm_dtype is not None:
        code = _lib.dtype_code(norm_dtype)
        out_n = torch.empty((B, 3, h, w), dtype=norm_dtype, device=dev)
        if lut is None:
            lut = normalize_lut(device=dev)
    _lib.check(lib.advmix_warp_affine_u8c3(_lib.ptr(src.buffer), _lib.ptr(src.offsets), _lib.ptr(src.heights),
                                           _lib.ptr(src.widths), _lib.ptr(src.pitches), _lib.ptr(flip_t),
                                           _lib.ptr(trans), _lib.ptr(out_u8), _lib.ptr(out_n), _lib.ptr(lut), B, w, h,
                                           code, _lib.stream_ptr()), "advmix_warp_affine_u8c3")
    return out_u8, out_n


def flip_perm(num_joints, flip_pairs, device="cuda"):
    perm = list(range(num_joints))
    for a, b in flip_pairs:
        perm[a], perm[b] = b, a
    return torch.tensor(perm, dtype=torch.int32, device=device)


def fliplr_affine_joints(joints, joints_vis, trans, flip=None, widths=None, perm=None):
    """fliplr_joints (transforms.py:44-58) where flip[b], then affine_transform on every joint
    with vis > 0 (JointsDataset.py:197-199).  float64 [B,J,3] in / out."""
    lib = _lib.load()
    joints = joints.to(torch.float64).contiguous()
    joints_vis = joints_vis.to(torch.float64).contiguous()
    B, J, _ = joints.shape
    jo, vo = torch.empty_like(joints), torch.empty_like(joints_vis)
    flip_t = None if flip is None else flip.to(torch.uint8).contiguous()
    widths_t = None if widths is None else widths.to(torch.int32).contiguous()
    _lib.check(lib.advmix_joints_flip_affine(_lib.ptr(joints), _lib.ptr(joints_vis), _lib.ptr(flip_t),
                                             _lib.ptr(widths_t), _lib.ptr(perm), _lib.ptr(trans.contiguous()),
                                             _lib.ptr(jo), _lib.ptr(vo), B, J, _lib.stream_ptr()),
               "advmix_joints_flip_affine")
    return jo, vo


def _csr(center, scale, rot):
    center = center.to(torch.float32).contiguous()
    scale_f32 = scale.dtype != torch.float64
    return center, scale.to(torch.float64).contiguous(), int(scale_f32), rot.to(torch.float64).contiguous()


def crop_csr(src, center, scale, rot, output_size, flip=None, want_u8=True, norm_dtype=None, lut=None):
    """get_affine_transform + cv2.warpAffine (+ transform) in one launch (JointsDataset.py:189-195,331-332):
    same results as `warp_affine(src, get_affine_transform(center, scale, rot, size), ...)`."""
    lib = _lib.load()
    B = len(src)
    w, h = int(output_size[0]), int(output_size[1])
    dev = src.buffer.device
    center, scale, scale_f32, rot = _csr(center, scale, rot)
    flip_t = None if flip is None else flip.to(torch.uint8).contiguous()
    out_u8 = torch.empty((B, h, w, 3), dtype=torch.uint8, device=dev) if want_u8 else None
    out_n = None
    code = _lib.F32
    if norm_dtype is not None:
        code = _lib.dtype_code(norm_dtype)
        out_n = torch.empty((B, 3, h, w), dtype=norm_dtype, device=dev)
        if lut is None:
            lut = normalize_lut(device=dev)
    _lib.check(lib.advmix_crop_csr_u8c3(_lib.ptr(src.buffer), _lib.ptr(src.offsets), _lib.ptr(src.heights),
                                        _lib.ptr(src.widths), _lib.ptr(src.pitches), _lib.ptr(flip_t), _lib.ptr(center),
                                        _lib.ptr(scale), scale_f32, _lib.ptr(rot), _lib.ptr(out_u8), _lib.ptr(out_n),
                                        _lib.ptr(lut), B, w, h, code, _lib.stream_ptr()), "advmix_crop_csr_u8c3")
    return out_u8, out_n


def joints_csr(joints, joints_vis, center, scale, rot, output_size, flip=None, widths=None, perm=None, want_trans=False):
    """get_affine_transform + fliplr_joints + affine_transform in one launch (JointsDataset.py:180-199).
    Returns (joints, joints_vis[, trans])."""
    lib = _lib.load()
    joints = joints.to(torch.float64).contiguous()
    joints_vis = joints_vis.to(torch.float64).contiguous()
    B, J, _ = joints.shape
    center, scale, scale_f32, rot = _csr(center, scale, rot)
    jo, vo = torch.empty_like(joints), torch.empty_like(joints_vis)
    M = torch.empty((B, 2, 3), dtype=torch.float64, device=joints.device) if want_trans else None
    flip_t = None if flip is None else flip.to(torch.uint8).contiguous()
    widths_t = None if widths is None else widths.to(torch.int32).contiguous()
    _lib.check(lib.advmix_joints_csr(_lib.ptr(joints), _lib.ptr(joints_vis), _lib.ptr(flip_t), _lib.ptr(widths_t),
                                     _lib.ptr(perm), _lib.ptr(center), _lib.ptr(scale), scale_f32, _lib.ptr(rot),
                                     _lib.ptr(M), _lib.ptr(jo), _lib.ptr(vo), B, J, int(output_size[0]),
                                     int(output_size[1]), _lib.stream_ptr()), "advmix_joints_csr")
    return (jo, vo, M) if want_trans else (jo, vo)


def to_tensor_normalize(images_u8, dtype=torch.float32, lut=None):
    """ToTensor()+Normalize() on uint8 [B,H,W,3] -> [B,3,H,W]."""
    lib = _lib.load()
    images_u8 = images_u8.contiguous()
    B, H, W, _ = images_u8.shape
    if lut is None:
        lut = normalize_lut(device=images_u8.device)
    out = torch.empty((B, 3, H, W), dtype=dtype, device=images_u8.device)
    _lib.check(lib.advmix_normalize_u8c3(_lib.ptr(images_u8), _lib.ptr(out), _lib.ptr(lut), B, H, W,
                                         _lib.dtype_code(dtype), _lib.stream_ptr()), "advmix_normalize_u8c3")
    return out
