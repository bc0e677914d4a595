"""Device mirror of the reference's heat-map consumers.

  get_max_preds / get_final_preds   <- lib/core/inference.py:22-95
  flip_back / flip_merge            <- lib/utils/transforms.py:16-41, lib/core/function.py:241-261

Inputs and outputs are CUDA tensors; nothing is copied to the host.
"""
import torch

from . import _lib


def _decode(batch_heatmaps, center, scale, post_process, want_preds):
    lib = _lib.load()
    if not (torch.is_tensor(batch_heatmaps) and batch_heatmaps.is_cuda):
        raise TypeError("batch_heatmaps should be a CUDA tensor")
    assert batch_heatmaps.ndim == 4, "batch_images should be 4-ndim"
    hm = batch_heatmaps.to(torch.float32).contiguous()
    B, J, H, W = hm.shape
    dev = hm.device
    maxvals = torch.empty((B, J, 1), dtype=torch.float32, device=dev)
    coords = torch.empty((B, J, 2), dtype=torch.float32, device=dev)
    preds = c = s = None
    s_f32 = 1
    if want_preds:
        preds = torch.empty((B, J, 2), dtype=torch.float32, device=dev)
        c = torch.as_tensor(center).to(dev, torch.float32).contiguous()
        s = torch.as_tensor(scale)
        s_f32 = int(s.dtype != torch.float64)
        s = s.to(dev, torch.float64).contiguous()
    _lib.check(lib.advmix_heatmap_decode(_lib.ptr(hm), _lib.ptr(c), _lib.ptr(s), s_f32, int(bool(post_process)),
                                         _lib.ptr(preds), _lib.ptr(maxvals), _lib.ptr(coords), B, J, H, W,
                                         _lib.stream_ptr()), "advmix_heatmap_decode")
    return preds, maxvals, coords


def get_max_preds(batch_heatmaps):
    """lib/core/inference.py:22-49: (preds float32 [B,J,2] in heat-map pixels, maxvals float32 [B,J,1])."""
    _, maxvals, coords = _decode(batch_heatmaps, None, None, False, False)
    return coords, maxvals


def get_final_preds(batch_heatmaps, center, scale, post_process=True):
    """lib/core/inference.py:52-95 with cal_hm_coord=True, coord=None (the heat-map path):
    (preds float32 [B,J,2] in source-image pixels, maxvals float32 [B,J,1]).
    post_process = config.TEST.POST_PROCESS; center [B,2], scale [B,2] as in the batch meta."""
    preds, maxvals, _ = _decode(batch_heatmaps, center, scale, post_process, True)
    return preds, maxvals


def flip_merge(output, output_flipped, flip_perm=None, shift_heatmap=True, out=None):
    """(output + shift(flip_back(output_flipped))) * 0.5 (lib/core/function.py:241-261) in one pass.
    flip_perm: int32 device tensor from transforms.flip_perm(num_joints, flip_pairs)."""
    lib = _lib.load()
    f = output_flipped.to(torch.float32).contiguous()
    a = None if output is None else output.to(torch.float32).contiguous()
    assert f.ndim == 4 and (a is None or a.shape == f.shape)
    B, J, H, W = f.shape
    if out is None:
        out = torch.empty_like(f)
    _lib.check(lib.advmix_flip_merge(_lib.ptr(a), _lib.ptr(f), _lib.ptr(flip_perm), int(bool(shift_heatmap)),
                                     _lib.ptr(out), B, J, H, W, _lib.stream_ptr()), "advmix_flip_merge")
    return out


def flip_back(output_flipped, flip_perm=None):
    """lib/utils/transforms.py:16-41 for 4-D heat maps (reverse x, swap matched joints)."""
    return flip_merge(None, output_flipped, flip_perm, shift_heatmap=False)
