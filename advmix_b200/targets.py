"""Host mirror of JointsDataset.generate_target (lib/dataset/JointsDataset.py:412-491)."""
import numpy as np
import torch

from . import _lib

_tab_cache = {}


def gaussian_table(sigma, device="cuda"):
    """The un-normalised (6*sigma+1)^2 float32 patch, built with the reference's own numpy
    expression (JointsDataset.py:470-476) so the device copy is bit-identical."""
    key = (int(sigma), str(device))
    if key not in _tab_cache:
        tmp_size = sigma * 3
        size = 2 * tmp_size + 1
        x = np.arange(0, size, 1, np.float32)
        y = x[:, np.newaxis]
        x0 = y0 = size // 2
        g = np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))
        _tab_cache[key] = torch.from_numpy(np.ascontiguousarray(g, dtype=np.float32)).to(device)
    return _tab_cache[key]


def generate_target(joints, joints_vis, image_size=(192, 256), heatmap_size=(48, 64), sigma=2,
                    joints_weight=None):
    """Batched generate_target.  joints, joints_vis: float64 [B,J,3] device tensors.
    Returns ([heatmap f32 [B,J,Hh,Wh], mu f32 [B,J,2]], target_weight f32 [B,J,1]) -
    the `target` list and `target_weight` of the reference, with a leading batch dim."""
    lib = _lib.load()
    joints = joints.to(torch.float64).contiguous()
    joints_vis = joints_vis.to(torch.float64).contiguous()
    B, J, _ = joints.shape
    dev = joints.device
    Wh, Hh = int(heatmap_size[0]), int(heatmap_size[1])
    hm = torch.empty((B, J, Hh, Wh), dtype=torch.float32, device=dev)
    mu = torch.empty((B, J, 2), dtype=torch.float32, device=dev)
    tw = torch.empty((B, J, 1), dtype=torch.float32, device=dev)
    jw = None
    if joints_weight is not None:
        jw = torch.as_tensor(np.asarray(joints_weight, dtype=np.float32).reshape(-1)).to(dev)
        assert jw.numel() == J
    tab = gaussian_table(sigma, dev)
    _lib.check(lib.advmix_heatmap_targets(_lib.ptr(joints), _lib.ptr(joints_vis), _lib.ptr(tab), _lib.ptr(jw),
                                          _lib.ptr(hm), _lib.ptr(mu), _lib.ptr(tw), B, J, Hh, Wh, int(image_size[0]),
                                          int(image_size[1]), int(sigma), _lib.stream_ptr()),
               "advmix_heatmap_targets")
    return [hm, mu], tw
