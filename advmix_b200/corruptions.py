"""Drop-in for the `imagecorruptions` package API used by the reference
(tools/make_datasets.py:38-41, lib/dataset/JointsDataset.py:259-264,286), running on the GPU.

`corrupt(image, severity, corruption_name, corruption_number)` and
`get_corruption_names(subset)` keep the package's signature, validation and exception
types; `corrupt_batch` is the device-resident batched form the hot path uses.
"""
import ctypes

import numpy as np
import torch

from . import _lib

CORRUPTIONS = _lib.OPS
_OP_INDEX = {n: i for i, n in enumerate(_lib.OPS)}


def get_corruption_names(subset="common"):
    if subset == "common":
        return list(CORRUPTIONS[:15])
    if subset == "validation":
        return list(CORRUPTIONS[15:])
    if subset == "all":
        return list(CORRUPTIONS)
    if subset == "noise":
        return list(CORRUPTIONS[0:3])
    if subset == "blur":
        return list(CORRUPTIONS[3:7])
    if subset == "weather":
        return list(CORRUPTIONS[7:11])
    if subset == "digital":
        return list(CORRUPTIONS[11:15])
    raise ValueError("subset must be one of ['common', 'validation', 'all']")


def op_index(name):
    if name not in _OP_INDEX:
        raise KeyError(name)
    return _OP_INDEX[name]


class _Workspace:
    """Grow-only per-device scratch buffer (the library itself never allocates)."""

    def __init__(self):
        self.buf = {}

    def get(self, nbytes, device):
        key = str(device)
        b = self.buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)
            self.buf[key] = b
        return b


_ws = _Workspace()

_frost_bank = {}            # device -> registered textures (as given)
_frost_fit = {}             # (device, H, W) -> textures large enough for H x W images
_frost_warned = [False]


def set_frost_bank(bank, device="cuda"):
    """Register the frost textures: uint8 [N,fh,fw,3] RGB.  (The package ships frost1-3.png / frost4-6.jpg; they are not
    redistributable here, so the caller loads them - or any texture set - once.)  Textures smaller than an image are
    up-scaled for that image size the way the package does (frost(): scale = max(H/fh, W/fw) * 1.1, cv2.resize
    INTER_CUBIC), once per image size."""
    t = torch.as_tensor(np.ascontiguousarray(bank)) if not torch.is_tensor(bank) else bank
    assert t.dtype == torch.uint8 and t.ndim == 4 and t.shape[3] == 3
    key = str(torch.device(device))
    _frost_bank[key] = t.to(device).contiguous()
    for k in [k for k in _frost_fit if k[0] == key]:
        del _frost_fit[k]


def default_frost_bank(fh=384, fw=384, n=5, seed=7):
    """Deterministic synthetic stand-in textures (smooth bright blobs) for tests and benchmarks.  NOT the package's
    frost photographs: COCO-C / MPII-C frost splits built with it are not comparable with the published benchmark."""
    import cv2
    rng = np.random.default_rng(seed)
    bank = np.empty((n, fh, fw, 3), np.uint8)
    for i in range(n):
        low = rng.random((fh // 16 + 1, fw // 16 + 1, 3)).astype(np.float32)
        t = cv2.resize(low, (fw, fh), interpolation=cv2.INTER_CUBIC)
        t = t + 0.15 * rng.random((fh, fw, 3)).astype(np.float32)
        bank[i] = np.clip(t * 200 + 40, 0, 255).astype(np.uint8)
    return bank


def _get_frost(device, H, W):
    key = str(torch.device(device))
    b = _frost_bank.get(key)
    if b is None:
        if not _frost_warned[0]:
            import warnings
            warnings.warn("advmix_b200: 'frost' is running on SYNTHETIC stand-in textures because no frost bank was registered "
                          "(set_frost_bank(<the imagecorruptions frost1-6 images>)); the result is a frost-like corruption, not the "
                          "package's, and is not comparable with the published COCO-C / MPII-C frost numbers", stacklevel=3)
            _frost_warned[0] = True
        set_frost_bank(default_frost_bank(max(384, H + 32), max(384, W + 32)), device)
        b = _frost_bank[key]
    if b.shape[1] >= H and b.shape[2] >= W:
        return b
    fit = _frost_fit.get((key, H, W))
    if fit is None:
        # the package's rule for a texture smaller than the image (one-time table preparation, cached per size)
        import cv2
        scale = max(H / b.shape[1], W / b.shape[2]) * 1.1
        fh, fw = int(np.ceil(b.shape[1] * scale)), int(np.ceil(b.shape[2] * scale))
        up = np.stack([cv2.resize(t, (fw, fh), interpolation=cv2.INTER_CUBIC) for t in b.cpu().numpy()])
        fit = _frost_fit[(key, H, W)] = torch.from_numpy(up).to(device).contiguous()
    return fit


def rand_field_bytes(name, severity, H, W):
    return int(_lib.load().advmix_corrupt_rand_field_bytes(op_index(name), severity, H, W))


def fill_rand(name, severity, n, H, W, seed, sample_base=0, idx=None, device="cuda", frost_shape=None):
    """Materialise the draws perf mode would consume: (rand_field uint8 [n, field_bytes] or None,
    rand_param float64 [n,4])."""
    lib = _lib.load()
    op = op_index(name)
    fb = rand_field_bytes(name, severity, H, W)
    field = torch.empty((n, fb), dtype=torch.uint8, device=device) if fb else None
    param = torch.zeros((n, 4), dtype=torch.float64, device=device)
    fn, fh, fw = frost_shape if frost_shape is not None else (0, 0, 0)
    _lib.check(lib.advmix_corrupt_fill_rand(op, severity, n, H, W, int(seed), int(sample_base), _lib.ptr(idx),
                                            _lib.ptr(field), _lib.ptr(param), fn, fh, fw, _lib.stream_ptr()),
               "advmix_corrupt_fill_rand")
    return field, param


def corrupt_batch(images, corruption_name, severity=1, seed=0, sample_base=0, idx=None, out=None,
                  rand_field=None, rand_param=None, frost_bank=None, fast=False):
    """images: uint8 [B,H,W,3] CUDA tensor -> corrupted uint8 [B,H,W,3].

    idx (int32 device tensor, optional) restricts the call to those batch entries (the others
    of `out` are left untouched).  rand_field / rand_param inject the random draws (layouts in
    include/advmix_b200.h); otherwise they are generated in-register from (seed, sample_base+i).
    fast=True selects float32 arithmetic where an op has such a kernel (<= 1 LSB from the exact path)."""
    lib = _lib.load()
    if not (torch.is_tensor(images) and images.is_cuda and images.dtype == torch.uint8 and images.ndim == 4
            and images.shape[3] == 3):
        raise TypeError("corrupt_batch expects a CUDA uint8 tensor [B,H,W,3]")
    if severity not in (1, 2, 3, 4, 5):
        raise AttributeError("Severity must be an integer in [1, 5]")
    images = images.contiguous()
    B, H, W, _ = images.shape
    op = op_index(corruption_name)
    n = B if idx is None else int(idx.numel())
    if out is None:
        out = torch.empty_like(images) if idx is None else images.clone()
    ws_bytes = int(lib.advmix_corrupt_workspace_bytes(op, severity, n, H, W))
    ws = _ws.get(ws_bytes, images.device) if ws_bytes else None
    fb, fn, fh, fw = None, 0, 0, 0
    if corruption_name == "frost":
        fb = frost_bank if frost_bank is not None else _get_frost(images.device, H, W)
        fn, fh, fw = int(fb.shape[0]), int(fb.shape[1]), int(fb.shape[2])
    _lib.check(lib.advmix_corrupt_u8c3(op | (_lib.CORRUPT_FAST if fast else 0), severity, _lib.ptr(images), _lib.ptr(out), n, _lib.ptr(idx), H, W,
                                       _lib.ptr(rand_field), _lib.ptr(rand_param), int(seed), int(sample_base),
                                       _lib.ptr(fb), fn, fh, fw, _lib.ptr(ws), ws_bytes, _lib.stream_ptr()),
               "advmix_corrupt_u8c3(%s)" % corruption_name)
    return out


def corrupt_sweep(images, corruption_name, seed=0, sample_base=0, idx=None, out=None, frost_bank=None, fast=False):
    """All five severities of one corruption: uint8 [B,H,W,3] CUDA tensor -> uint8 [5,B,H,W,3] (out[s-1] == corrupt_batch(...,
    severity=s) for the same seed).  The reference's dataset builder (tools/make_datasets.py:38-45) runs the severities
    innermost over the same image; the crops are read once and what does not depend on the severity is computed once."""
    lib = _lib.load()
    if not (torch.is_tensor(images) and images.is_cuda and images.dtype == torch.uint8 and images.ndim == 4
            and images.shape[3] == 3):
        raise TypeError("corrupt_sweep expects a CUDA uint8 tensor [B,H,W,3]")
    images = images.contiguous()
    B, H, W, _ = images.shape
    op = op_index(corruption_name)
    n = B if idx is None else int(idx.numel())
    if out is None:
        out = torch.empty((5,) + tuple(images.shape), dtype=torch.uint8, device=images.device)
        if idx is not None:
            out[:] = images
    elif not (out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (5,) + tuple(images.shape) and out.is_contiguous()):
        raise TypeError("corrupt_sweep: out must be a contiguous CUDA uint8 tensor [5,B,H,W,3]")
    ws_bytes = int(lib.advmix_corrupt_sweep_workspace_bytes(op, n, H, W))
    ws = _ws.get(ws_bytes, images.device) if ws_bytes else None
    fb, fn, fh, fw = None, 0, 0, 0
    if corruption_name == "frost":
        fb = frost_bank if frost_bank is not None else _get_frost(images.device, H, W)
        fn, fh, fw = int(fb.shape[0]), int(fb.shape[1]), int(fb.shape[2])
    outs = (ctypes.c_void_p * 5)(*[out[s].data_ptr() for s in range(5)])
    _lib.check(lib.advmix_corrupt_sweep_u8c3(op | (_lib.CORRUPT_FAST if fast else 0), _lib.ptr(images), outs, n, _lib.ptr(idx), H, W,
                                             int(seed), int(sample_base), _lib.ptr(fb), fn, fh, fw, _lib.ptr(ws), ws_bytes,
                                             _lib.stream_ptr()),
               "advmix_corrupt_sweep_u8c3(%s)" % corruption_name)
    return out


def corrupt(image, severity=1, corruption_name=None, corruption_number=-1):
    """imagecorruptions.corrupt: numpy uint8 [H,W], [H,W,1] or [H,W,3] -> uint8 [H,W,3].
    Random draws are keyed by a seed taken from the global np.random (so np.random.seed(1)
    before each call, as tools/make_datasets.py:40 does, makes it reproducible)."""
    if not isinstance(image, np.ndarray):
        raise AttributeError('Expecting type(image) to be numpy.ndarray')
    if not (image.dtype.type is np.uint8):
        raise AttributeError('Expecting image.dtype.type to be numpy.uint8')
    if not (image.ndim in [2, 3]):
        raise AttributeError('Expecting image.shape to be either (height x width) or (height x width x channels)')
    if image.ndim == 2:
        image = np.stack((image,) * 3, axis=-1)
    height, width, channels = image.shape
    if height < 32 or width < 32:
        raise AttributeError('Image width and height must be at least 32 pixels')
    if not (channels in [1, 3]):
        raise AttributeError('Expecting image to have either 1 or 3 channels (last dimension)')
    if channels == 1:
        image = np.stack((np.squeeze(image),) * 3, axis=-1)
    if not (severity in [1, 2, 3, 4, 5]):
        raise AttributeError('Severity must be an integer in [1, 5]')
    if corruption_name is None and corruption_number == -1:
        raise ValueError("Either corruption_name or corruption_number must be passed")
    name = corruption_name if corruption_name is not None else CORRUPTIONS[corruption_number]
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    dev = torch.device("cuda", torch.cuda.current_device())
    t = torch.from_numpy(np.ascontiguousarray(image)).to(dev)[None]
    return corrupt_batch(t, name, severity, seed=seed)[0].cpu().numpy()
