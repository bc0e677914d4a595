"""Oracle for SURVEY row f4: the per-record helpers of the dataset classes.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Pinned: tests/golden/records.npz holds outputs of the real
JointsDataset.half_body_transform / select_data (imported by oracle/make_golden.py).  _xywh2cs lives in
lib/dataset/coco.py, which needs pycocotools to import; it is restated here from coco.py:205-220.
"""
import numpy as np


def xywh2cs(x, y, w, h, aspect_ratio, pixel_std=200):
    """lib/dataset/coco.py:205-220."""
    center = np.zeros((2), dtype=np.float32)
    center[0] = x + w * 0.5
    center[1] = y + h * 0.5
    if w > aspect_ratio * h:
        h = w * 1.0 / aspect_ratio
    elif w < aspect_ratio * h:
        w = h * aspect_ratio
    scale = np.array([w * 1.0 / pixel_std, h * 1.0 / pixel_std], dtype=np.float32)
    if center[0] != -1:
        scale = scale * 1.25
    return center, scale


def half_body_transform(joints, joints_vis, upper_body_ids, randn_draw, aspect_ratio, pixel_std=200):
    """lib/dataset/JointsDataset.py:69-111 with the np.random.randn() of :80 passed in."""
    upper_joints, lower_joints = [], []
    for joint_id in range(joints.shape[0]):
        if joints_vis[joint_id][0] > 0:
            (upper_joints if joint_id in upper_body_ids else lower_joints).append(joints[joint_id])
    if randn_draw < 0.5 and len(upper_joints) > 2:
        selected_joints = upper_joints
    else:
        selected_joints = lower_joints if len(lower_joints) > 2 else upper_joints
    if len(selected_joints) < 2:
        return None, None
    selected_joints = np.array(selected_joints, dtype=np.float32)
    center = selected_joints.mean(axis=0)[:2]
    left_top = np.amin(selected_joints, axis=0)
    right_bottom = np.amax(selected_joints, axis=0)
    w = right_bottom[0] - left_top[0]
    h = right_bottom[1] - left_top[1]
    if w > aspect_ratio * h:
        h = w * 1.0 / aspect_ratio
    elif w < aspect_ratio * h:
        w = h * aspect_ratio
    scale = np.array([w * 1.0 / pixel_std, h * 1.0 / pixel_std], dtype=np.float32)
    return center, scale * 1.5


def select_data_mask(db, pixel_std=200):
    """lib/dataset/JointsDataset.py:366-399 as a keep-mask over the records."""
    keep = []
    for rec in db:
        num_vis, joints_x, joints_y = 0, 0.0, 0.0
        for joint, joint_vis in zip(rec['joints_3d'], rec['joints_3d_vis']):
            if joint_vis[0] <= 0:
                continue
            num_vis += 1
            joints_x += joint[0]
            joints_y += joint[1]
        if num_vis == 0:
            keep.append(False)
            continue
        joints_x, joints_y = joints_x / num_vis, joints_y / num_vis
        area = rec['scale'][0] * rec['scale'][1] * (pixel_std ** 2)
        joints_center = np.array([joints_x, joints_y])
        bbox_center = np.array(rec['center'])
        diff_norm2 = np.linalg.norm((joints_center - bbox_center), 2)
        ks = np.exp(-1.0 * (diff_norm2 ** 2) / ((0.2) ** 2 * 2.0 * area))
        metric = (0.2 / 16) * num_vis + 0.45 - 0.2 / 16
        keep.append(bool(ks > metric))
    return np.array(keep)
