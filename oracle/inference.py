"""Oracle for SURVEY row f3: heat-map decode and flip-test merge.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Pinned: tests/golden/inference.npz holds outputs of the real
reference functions (lib/core/inference.py, lib/utils/transforms.py, imported by oracle/make_golden.py).
"""
import math

import numpy as np

from .affine import get_affine_transform


def get_max_preds(batch_heatmaps):
    """lib/core/inference.py:22-49."""
    B, J, H, W = batch_heatmaps.shape
    flat = batch_heatmaps.reshape((B, J, -1))
    idx = np.argmax(flat, 2).reshape((B, J, 1))
    maxvals = np.amax(flat, 2).reshape((B, J, 1))
    preds = np.tile(idx, (1, 1, 2)).astype(np.float32)
    preds[:, :, 0] = preds[:, :, 0] % W
    preds[:, :, 1] = np.floor(preds[:, :, 1] / W)
    preds *= np.tile(np.greater(maxvals, 0.0), (1, 1, 2)).astype(np.float32)
    return preds, maxvals


def transform_preds(coords, center, scale, output_size):
    """lib/utils/transforms.py:61-66 (inverse matrix of get_affine_transform, rot 0)."""
    out = np.zeros(coords.shape)
    trans = get_affine_transform(center, scale, 0, output_size, inv=1)
    for p in range(coords.shape[0]):
        out[p, 0:2] = np.dot(trans, np.array([coords[p, 0], coords[p, 1], 1.]).T)[:2]
    return out


def get_final_preds(batch_heatmaps, center, scale, post_process=True):
    """lib/core/inference.py:52-95 with cal_hm_coord=True, coord=None."""
    coords, maxvals = get_max_preds(batch_heatmaps)
    H, W = batch_heatmaps.shape[2], batch_heatmaps.shape[3]
    if post_process:
        for n in range(coords.shape[0]):
            for p in range(coords.shape[1]):
                hm = batch_heatmaps[n][p]
                px = int(math.floor(coords[n][p][0] + 0.5))
                py = int(math.floor(coords[n][p][1] + 0.5))
                if 1 < px < W - 1 and 1 < py < H - 1:
                    diff = np.array([hm[py][px + 1] - hm[py][px - 1], hm[py + 1][px] - hm[py - 1][px]])
                    coords[n][p] += np.sign(diff) * .25
    preds = coords.copy()
    for i in range(coords.shape[0]):
        preds[i] = transform_preds(coords[i], center[i], scale[i], [W, H])
    return preds, maxvals, coords


def flip_back(output_flipped, matched_parts):
    """lib/utils/transforms.py:16-41, 4-D branch."""
    out = output_flipped[..., ::-1].copy()
    for a, b in matched_parts:
        tmp = out[:, a, ...].copy()
        out[:, a, ...] = out[:, b, ...]
        out[:, b, ...] = tmp
    return out


def flip_merge(output, output_flipped, matched_parts, shift_heatmap=True):
    """lib/core/function.py:241-261 (float32 like the torch tensors there)."""
    f = flip_back(output_flipped, matched_parts)
    if shift_heatmap:
        f[:, :, :, 1:] = f.copy()[:, :, :, 0:-1]
    return ((output + f) * np.float32(0.5)).astype(np.float32)
