"""Oracle for SURVEY row a1: top-down affine crop.

TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates
  * lib/utils/transforms.py:69-101  get_affine_transform (+ :110-122 helpers)
  * lib/utils/transforms.py:104-107 affine_transform
  * lib/utils/transforms.py:44-58   fliplr_joints
  * lib/dataset/JointsDataset.py:167-199 (draws, flip view, warpAffine, joints)
  * cv2.warpAffine(uint8 C3, INTER_LINEAR, BORDER_CONSTANT 0) as called at
    lib/dataset/JointsDataset.py:190-195 / :324-329 - OpenCV's fixed-point path
    (AB_BITS=10, INTER_BITS=5, INTER_REMAP_COEF_BITS=15).
"""
import numpy as np

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
REMAP_COEF_BITS = 15
ROUND_DELTA = AB_SCALE // INTER_TAB_SIZE // 2  # 16 for INTER_LINEAR


def get_dir(src_point, rot_rad):
    # transforms.py:115-122
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    return [src_point[0] * cs - src_point[1] * sn,
            src_point[0] * sn + src_point[1] * cs]


def get_3rd_point(a, b):
    # transforms.py:110-112
    direct = a - b
    return b + np.array([-direct[1], direct[0]], dtype=np.float32)


def affine_points(center, scale, rot, output_size, shift=(0.0, 0.0)):
    """The two float32 point triples of transforms.py:76-93 (src, dst)."""
    scale = np.asarray(scale)
    if scale.ndim == 0:
        scale = np.array([scale, scale])
    shift = np.array(shift, dtype=np.float32)
    scale_tmp = scale * 200.0
    src_w = scale_tmp[0]
    dst_w, dst_h = output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    src_dir = get_dir([0, src_w * -0.5], rot_rad)
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_tmp * shift
    src[1, :] = center + src_dir + scale_tmp * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    src[2:, :] = get_3rd_point(src[0, :], src[1, :])
    dst[2:, :] = get_3rd_point(dst[0, :], dst[1, :])
    return src, dst


def solve_affine_3pt(src, dst):
    """float64 2x3 M with M @ [x,y,1] = dst for three float32 point pairs.

    Restates cv2.getAffineTransform (transforms.py:95-99): the 6x6 system solved by cv::solve(DECOMP_LU), i.e.
    OpenCV's generic LUImpl<double> (modules/core/src/matrix_decomp.cpp: partial pivoting, d = -1/pivot,
    separate multiply and add, back substitution).  Bit-identical to cv2 4.13 on every case of
    tests/test_oracle_pinning.py::test_affine_lu_matches_cv2.
    """
    src = np.asarray(src, np.float32).astype(np.float64)
    dst = np.asarray(dst, np.float32).astype(np.float64)
    A = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        A[2 * i, 0:3] = (src[i, 0], src[i, 1], 1.0)
        A[2 * i + 1, 3:6] = (src[i, 0], src[i, 1], 1.0)
        b[2 * i], b[2 * i + 1] = dst[i, 0], dst[i, 1]
    eps = np.finfo(np.float64).eps * 100
    for i in range(6):
        k = i
        for j in range(i + 1, 6):
            if abs(A[j, i]) > abs(A[k, i]):
                k = j
        if abs(A[k, i]) < eps:
            return np.zeros((2, 3))
        if k != i:
            A[[i, k], i:] = A[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1.0 / A[i, i]
        for j in range(i + 1, 6):
            alpha = A[j, i] * d
            for c in range(i + 1, 6):
                A[j, c] = A[j, c] + alpha * A[i, c]
            b[j] = b[j] + alpha * b[i]
    for i in range(5, -1, -1):
        s = b[i]
        for c in range(i + 1, 6):
            s = s - A[i, c] * b[c]
        b[i] = s / A[i, i]
    return b.reshape(2, 3).copy()


def get_affine_transform(center, scale, rot, output_size, shift=(0.0, 0.0), inv=0,
                         use_cv2=True):
    """transforms.py:69-101.  use_cv2=True calls the real cv2.getAffineTransform."""
    src, dst = affine_points(center, scale, rot, output_size, shift)
    if inv:
        src, dst = dst, src
    if use_cv2:
        import cv2
        return cv2.getAffineTransform(np.float32(src), np.float32(dst))
    return solve_affine_3pt(src, dst)


def affine_transform(pt, t):
    # transforms.py:104-107
    new_pt = np.array([pt[0], pt[1], 1.]).T
    return np.dot(t, new_pt)[:2]


def fliplr_joints(joints, joints_vis, width, matched_parts):
    # transforms.py:44-58 (operates on copies)
    joints = joints.copy()
    joints_vis = joints_vis.copy()
    joints[:, 0] = width - joints[:, 0] - 1
    for a, b in matched_parts:
        joints[[a, b], :] = joints[[b, a], :]
        joints_vis[[a, b], :] = joints_vis[[b, a], :]
    return joints * joints_vis, joints_vis


def transform_joints(joints, joints_vis, trans):
    # JointsDataset.py:197-199
    joints = joints.copy()
    for i in range(joints.shape[0]):
        if joints_vis[i, 0] > 0.0:
            joints[i, 0:2] = affine_transform(joints[i, 0:2], trans)
    return joints


def invert_affine(M):
    """cv::warpAffine's in-place inversion of the forward 2x3 matrix (float64)."""
    M = np.array(M, dtype=np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11 = M[4] * D
    A22 = M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2] = b1
    M[5] = b2
    return M


def warp_affine_fixedpoint(src, M_fwd, dsize):
    """Numpy restatement of cv2.warpAffine(src u8 [H,W,C], M, (dw,dh), INTER_LINEAR).

    Border constant 0.  Coordinates: 10-bit fixed point, sampled on a 1/32 grid;
    weights 15-bit; out = (sum + 2^14) >> 15.
    """
    src = np.asarray(src)
    assert src.dtype == np.uint8 and src.ndim == 3
    H, W, C = src.shape
    dw, dh = int(dsize[0]), int(dsize[1])
    M = invert_affine(M_fwd)
    xs = np.arange(dw, dtype=np.float64)
    ys = np.arange(dh, dtype=np.float64)
    adelta = np.rint(M[0] * xs * AB_SCALE).astype(np.int64)
    bdelta = np.rint(M[3] * xs * AB_SCALE).astype(np.int64)
    X0 = np.rint((M[1] * ys + M[2]) * AB_SCALE).astype(np.int64) + ROUND_DELTA
    Y0 = np.rint((M[4] * ys + M[5]) * AB_SCALE).astype(np.int64) + ROUND_DELTA
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    fx = X & (INTER_TAB_SIZE - 1)
    fy = Y & (INTER_TAB_SIZE - 1)
    w00 = (32 - fx) * (32 - fy) * 32
    w01 = fx * (32 - fy) * 32
    w10 = (32 - fx) * fy * 32
    w11 = fx * fy * 32

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64)
        return v * ok[..., None]

    acc = (tap(sy, sx) * w00[..., None] + tap(sy, sx + 1) * w01[..., None] +
           tap(sy + 1, sx) * w10[..., None] + tap(sy + 1, sx + 1) * w11[..., None])
    out = (acc + (1 << (REMAP_COEF_BITS - 1))) >> REMAP_COEF_BITS
    return np.clip(out, 0, 255).astype(np.uint8)


def warp_affine_cv2(src, M_fwd, dsize):
    """The reference's actual call (JointsDataset.py:190-195)."""
    import cv2
    return cv2.warpAffine(src, np.asarray(M_fwd, np.float64), (int(dsize[0]), int(dsize[1])),
                          flags=cv2.INTER_LINEAR)


def xywh2cs(x, y, w, h, aspect_ratio=0.75, pixel_std=200):
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


def normalize_lut(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """[3,256] float32 table equal to torchvision ToTensor()+Normalize() per value
    (tools/train.py:116-126): ((v/255) - mean)/std in float32, torch op order."""
    import torch
    v = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)
    m = torch.tensor(mean, dtype=torch.float32)[:, None]
    s = torch.tensor(std, dtype=torch.float32)[:, None]
    return ((v[None, :] - m) / s).numpy().copy()


def to_tensor_normalize(img_u8, lut=None):
    """u8 HWC -> f32 CHW, bit-identical to ToTensor()+Normalize()."""
    if lut is None:
        lut = normalize_lut()
    img_u8 = np.asarray(img_u8)
    return np.stack([lut[c][img_u8[:, :, c]] for c in range(3)], axis=0)
