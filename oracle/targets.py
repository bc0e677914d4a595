"""Oracle for SURVEY row a4: generate_target Gaussian heatmaps.

TEST INFRASTRUCTURE - see oracle/__init__.py.
Restates lib/dataset/JointsDataset.py:412-491 (live branch :454-486) and the
argmax decode lib/core/inference.py:22-49 (get_max_preds, used as parity checker).
"""
import numpy as np


def gaussian_table(sigma):
    """The (6*sigma+1)^2 float32 patch of JointsDataset.py:470-476 (centre == 1)."""
    tmp_size = sigma * 3
    size = 2 * tmp_size + 1
    x = np.arange(0, size, 1, np.float32)
    y = x[:, np.newaxis]
    x0 = y0 = size // 2
    return np.exp(-((x - x0) ** 2 + (y - y0) ** 2) / (2 * sigma ** 2))


def generate_target(joints, joints_vis, image_size=(192, 256), heatmap_size=(48, 64),
                    sigma=2, joints_weight=None):
    """-> ([heatmap f32 [J,Hh,Wh], mu f32 [J,2]], target_weight f32 [J,1])."""
    image_size = np.array(image_size)
    heatmap_size = np.array(heatmap_size)
    J = joints.shape[0]
    target_weight = np.ones((J, 1), dtype=np.float32)
    target_weight[:, 0] = joints_vis[:, 0]
    target = [np.zeros((J, heatmap_size[1], heatmap_size[0]), dtype=np.float32),
              np.zeros((J, 2), dtype=np.float32)]
    tmp_size = sigma * 3
    g = gaussian_table(sigma)
    for j in range(J):
        feat_stride = image_size / heatmap_size
        mu_x = int(joints[j][0] / feat_stride[0] + 0.5)
        mu_y = int(joints[j][1] / feat_stride[1] + 0.5)
        ul = [int(mu_x - tmp_size), int(mu_y - tmp_size)]
        br = [int(mu_x + tmp_size + 1), int(mu_y + tmp_size + 1)]
        if ul[0] >= heatmap_size[0] or ul[1] >= heatmap_size[1] or br[0] < 0 or br[1] < 0:
            target_weight[j] = 0
            continue
        g_x = max(0, -ul[0]), min(br[0], heatmap_size[0]) - ul[0]
        g_y = max(0, -ul[1]), min(br[1], heatmap_size[1]) - ul[1]
        img_x = max(0, ul[0]), min(br[0], heatmap_size[0])
        img_y = max(0, ul[1]), min(br[1], heatmap_size[1])
        if target_weight[j] > 0.5:
            target[0][j][img_y[0]:img_y[1], img_x[0]:img_x[1]] = g[g_y[0]:g_y[1], g_x[0]:g_x[1]]
            target[1][j] = np.array([mu_x, mu_y], dtype=np.float32)
    if joints_weight is not None:
        target_weight = np.multiply(target_weight, joints_weight)
    return target, target_weight


def get_max_preds(batch_heatmaps):
    """lib/core/inference.py:22-49."""
    B, J, _, width = batch_heatmaps.shape
    r = batch_heatmaps.reshape((B, J, -1))
    idx = np.argmax(r, 2).reshape((B, J, 1))
    maxvals = np.amax(r, 2).reshape((B, J, 1))
    preds = np.tile(idx, (1, 1, 2)).astype(np.float32)
    preds[:, :, 0] = preds[:, :, 0] % width
    preds[:, :, 1] = np.floor(preds[:, :, 1] / width)
    preds *= np.tile(np.greater(maxvals, 0.0), (1, 1, 2)).astype(np.float32)
    return preds, maxvals
