"""Batched device versions of the per-record helpers of the reference's dataset classes:
  xywh2cs              <- COCODataset._xywh2cs             lib/dataset/coco.py:205-220
  half_body_transform  <- JointsDataset.half_body_transform lib/dataset/JointsDataset.py:69-111
  select_data          <- JointsDataset.select_data         lib/dataset/JointsDataset.py:366-399
"""
import torch

from . import _lib


def xywh2cs(boxes, aspect_ratio, pixel_std=200):
    """boxes: float64 [B,4] (x, y, w, h) on the device -> (center float32 [B,2], scale float32 [B,2])."""
    lib = _lib.load()
    boxes = boxes.to(torch.float64).contiguous()
    B = boxes.shape[0]
    c = torch.empty((B, 2), dtype=torch.float32, device=boxes.device)
    s = torch.empty_like(c)
    _lib.check(lib.advmix_xywh2cs(_lib.ptr(boxes), _lib.ptr(c), _lib.ptr(s), B, float(aspect_ratio), float(pixel_std),
                                  _lib.stream_ptr()), "advmix_xywh2cs")
    return c, s


def half_body_transform(joints, joints_vis, upper_body_ids, randn_draw, aspect_ratio, pixel_std=200):
    """joints / joints_vis: float64 [B,J,3]; randn_draw: float64 [B] (the np.random.randn() of the reference).
    Returns (center [B,2], scale [B,2], valid bool [B]); rows with valid == False are the reference's (None, None)."""
    lib = _lib.load()
    joints = joints.to(torch.float64).contiguous()
    joints_vis = joints_vis.to(torch.float64).contiguous()
    B, J, _ = joints.shape
    dev = joints.device
    mask = torch.zeros(J, dtype=torch.uint8)
    mask[list(upper_body_ids)] = 1
    mask = mask.to(dev)
    draw = randn_draw.to(dev, torch.float64).contiguous()
    c = torch.empty((B, 2), dtype=torch.float32, device=dev)
    s = torch.empty_like(c)
    valid = torch.empty(B, dtype=torch.uint8, device=dev)
    _lib.check(lib.advmix_half_body_cs(_lib.ptr(joints), _lib.ptr(joints_vis), _lib.ptr(mask), _lib.ptr(draw), _lib.ptr(c),
                                       _lib.ptr(s), _lib.ptr(valid), B, J, float(aspect_ratio), float(pixel_std),
                                       _lib.stream_ptr()), "advmix_half_body_cs")
    return c, s, valid.bool()


def select_data(joints, joints_vis, center, scale, pixel_std=200):
    """Boolean mask [B] of the records JointsDataset.select_data keeps."""
    lib = _lib.load()
    joints = joints.to(torch.float64).contiguous()
    joints_vis = joints_vis.to(torch.float64).contiguous()
    B, J, _ = joints.shape
    center = center.to(torch.float32).contiguous()
    scale = scale.to(torch.float32).contiguous()
    keep = torch.empty(B, dtype=torch.uint8, device=joints.device)
    _lib.check(lib.advmix_select_data(_lib.ptr(joints), _lib.ptr(joints_vis), _lib.ptr(center), _lib.ptr(scale), _lib.ptr(keep),
                                      B, J, float(pixel_std), _lib.stream_ptr()), "advmix_select_data")
    return keep.bool()
