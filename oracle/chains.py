"""Oracle for SURVEY rows a5/a6: the reference-actual AdvMix chains.

TEST INFRASTRUCTURE - see oracle/__init__.py.
Restates lib/dataset/advaug.py:
  * :10-42  ImageNetPolicy (12 sub-policies), :46-107 SubPolicy.  Only five PIL ops
    are reachable from the 12 policies: equalize, posterize, solarize, invert,
    sharpness.  Each is restated in numpy AND checked against PIL itself
    (tests/test_oracle_pinning.py); ``*_pil`` variants call PIL directly.
  * :111-170 grid_aug (GridMask, mode=1, rotate=1, ratio=0.5, prob=0.7 as hard-wired
    at advaug.py:192-203).
"""
import numpy as np

# (p1, op1, magnitude_idx1, p2, op2, magnitude_idx2)  advaug.py:22-35
POLICIES = [
    (0.8, "equalize", 8, 0.6, "equalize", 3),
    (0.6, "posterize", 7, 0.6, "posterize", 6),
    (0.4, "equalize", 7, 0.2, "solarize", 4),
    (0.6, "solarize", 3, 0.6, "equalize", 7),
    (0.8, "posterize", 5, 1.0, "equalize", 2),
    (0.6, "equalize", 8, 0.4, "posterize", 6),
    (0.0, "equalize", 7, 0.8, "equalize", 8),
    (0.6, "invert", 4, 1.0, "equalize", 8),
    (0.4, "sharpness", 7, 0.6, "invert", 8),
    (0.4, "equalize", 7, 0.2, "solarize", 4),
    (0.6, "invert", 4, 1.0, "equalize", 8),
    (0.8, "equalize", 8, 0.6, "equalize", 3),
]

RANGES = {  # advaug.py:48-63 (only the reachable ops)
    "posterize": np.round(np.linspace(8, 4, 10), 0).astype(int),
    "solarize": np.linspace(256, 0, 10),
    "sharpness": np.linspace(0.0, 0.9, 10),
    "equalize": [0] * 10,
    "invert": [0] * 10,
}


def magnitude(op, idx):
    return RANGES[op][idx]


# ---- numpy restatements of the PIL ops (PIL 12.2 semantics) ---------------------
def equalize(img):
    """ImageOps.equalize: per-band histogram LUT."""
    out = np.empty_like(img)
    for c in range(img.shape[2]):
        h = np.bincount(img[:, :, c].ravel(), minlength=256)
        nz = h[h > 0]
        if len(nz) <= 1:
            out[:, :, c] = img[:, :, c]
            continue
        step = (int(h.sum()) - int(nz[-1])) // 255
        if not step:
            out[:, :, c] = img[:, :, c]
            continue
        n = step // 2
        lut = np.empty(256, np.uint8)
        for i in range(256):
            lut[i] = min(n // step, 255)
            n += int(h[i])
        out[:, :, c] = lut[img[:, :, c]]
    return out


def posterize(img, bits):
    mask = ~(2 ** (8 - int(bits)) - 1) & 0xFF
    return img & np.uint8(mask)


def solarize(img, threshold):
    return np.where(img < threshold, img, 255 - img).astype(np.uint8)


def invert(img):
    return (255 - img).astype(np.uint8)


def smooth_filter(img):
    """ImageFilter.SMOOTH: 3x3 (1,1,1,1,5,1,1,1,1)/13, float32 accumulate + 0.5,
    truncate+clip; the 1-pixel border is copied unchanged."""
    k = np.array([[1, 1, 1], [1, 5, 1], [1, 1, 1]], np.float32) / np.float32(13)
    H, W, _ = img.shape
    out = img.copy()
    f = img.astype(np.float32)
    acc = np.full((H - 2, W - 2, img.shape[2]), 0.5, np.float32)
    # PIL ImagingFilter3x3 order: ss = offset+0.5; then for rows y+1, y, y-1:
    # ss += (in[x-1]*k0 + in[x]*k1 + in[x+1]*k2), all float32, no FMA.
    for ky in (2, 1, 0):
        row = (f[ky:ky + H - 2, 0:W - 2] * k[2 - ky, 0] + f[ky:ky + H - 2, 1:W - 1] * k[2 - ky, 1]) \
            + f[ky:ky + H - 2, 2:W] * k[2 - ky, 2]
        acc = acc + row
    out[1:-1, 1:-1] = np.clip(acc, 0, 255).astype(np.uint8)
    return out


def blend(im1, im2, alpha):
    """Image.blend(im1, im2, alpha) for uint8: im1 + alpha*(im2-im1), float32,
    clipped then truncated (PIL ImagingBlend)."""
    a = np.float32(alpha)
    t = im1.astype(np.float32) + a * (im2.astype(np.float32) - im1.astype(np.float32))
    if 0.0 <= alpha <= 1.0:
        return t.astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t)).astype(np.uint8)


def sharpness(img, factor):
    """ImageEnhance.Sharpness(img).enhance(factor) = blend(SMOOTH(img), img, factor)."""
    return blend(smooth_filter(img), img, factor)


def apply_op(img, op, mag, sign=1):
    if op == "equalize":
        return equalize(img)
    if op == "posterize":
        return posterize(img, mag)
    if op == "solarize":
        return solarize(img, mag)
    if op == "invert":
        return invert(img)
    if op == "sharpness":
        return sharpness(img, 1 + mag * sign)
    raise ValueError(op)


def apply_op_pil(img, op, mag, sign=1):
    """The real PIL calls of advaug.py:82-96."""
    from PIL import Image, ImageOps, ImageEnhance
    im = Image.fromarray(img)
    if op == "equalize":
        im = ImageOps.equalize(im)
    elif op == "posterize":
        im = ImageOps.posterize(im, int(mag))
    elif op == "solarize":
        im = ImageOps.solarize(im, mag)
    elif op == "invert":
        im = ImageOps.invert(im)
    elif op == "sharpness":
        im = ImageEnhance.Sharpness(im).enhance(1 + mag * sign)
    else:
        raise ValueError(op)
    return np.array(im)


def autoaug(img, policy_idx, coin1, coin2, sign1=1, sign2=1, use_pil=False):
    """advaug.py:37-39 + :104-107 with the draws made explicit:
    policy_idx = random.randint(0, 11); coin1/coin2 = random.random() values;
    sign = random.choice([-1, 1]) (only consumed by sharpness)."""
    p1, op1, m1, p2, op2, m2 = POLICIES[policy_idx]
    f = apply_op_pil if use_pil else apply_op
    if coin1 < p1:
        img = f(img, op1, magnitude(op1, m1), sign1)
    if coin2 < p2:
        img = f(img, op2, magnitude(op2, m2), sign2)
    return img


def gridmask_mask(h, w, d, st_h, st_w, ratio=0.5, mode=1):
    """The {0,1} float mask of advaug.py:114-151 (rotate=1 -> angle 0, use_h=use_w=True)."""
    hh, ww = int(1.5 * h), int(1.5 * w)
    l = min(max(int(d * ratio + 0.5), 1), d - 1)
    mask = np.ones((hh, ww), np.float32)
    for i in range(hh // d):
        s = d * i + st_h
        t = min(s + l, hh)
        mask[s:t, :] *= 0
    for i in range(ww // d):
        s = d * i + st_w
        t = min(s + l, ww)
        mask[:, s:t] *= 0
    mask = mask[(hh - h) // 2:(hh - h) // 2 + h, (ww - w) // 2:(ww - w) // 2 + w]
    if mode == 1:
        mask = 1 - mask
    return mask


def gridmask(img_chw, joints, joints_vis, apply, d, st_h, st_w, joints_num=None):
    """advaug.py:111-170 with draws explicit: apply = not (np.random.rand() > 0.7);
    d = randint(2, min(h,w)); st_h, st_w = randint(d).  img_chw: normalised f32 [3,H,W].
    Returns (img, joints_vis) - joints_vis[j,0:2] zeroed where the joint hits mask 0."""
    joints_vis = joints_vis.copy()
    if not apply:
        return img_chw.copy(), joints_vis
    _, h, w = img_chw.shape
    mask = gridmask_mask(h, w, d, st_h, st_w)
    out = img_chw * mask[None]
    J = joints.shape[0] if joints_num is None else joints_num
    for j in range(J):
        tx = max(min(int(joints[j][0]), w - 1), 0)
        ty = max(min(int(joints[j][1]), h - 1), 0)
        if mask[ty, tx] == 0:
            joints_vis[j][0] = 0
            joints_vis[j][1] = 0
    return out, joints_vis
