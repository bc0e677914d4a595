"""Oracle for SURVEY row a2 (+ row f2): the 15 common and 4 validation `imagecorruptions` operators x 5 severities.

TEST INFRASTRUCTURE - see oracle/__init__.py.

**PARITY UNPINNED.**  The algorithm lives in the third-party PyPI package
``imagecorruptions`` (bethgelab; un-pinned at /root/reference/requirements.txt:12,
latest known 1.1.2), which is absent from /root/reference and cannot be installed
offline; the reference has no tests or golden vectors for it.  This module restates
the package's published algorithm from the *same* primitives
the package calls - scipy.ndimage.{zoom, map_coordinates, gaussian_filter},
cv2.{filter2D, GaussianBlur, cvtColor}, PIL resize / JPEG - with skimage's thin
helpers (gaussian, rgb2hsv, hsv2rgb, random_noise 's&p') restated.  Parity is
anchored on the reference's call sites: tools/make_datasets.py:38-41
(names 'all', severity+1, np.random.seed(1) before each call) and
lib/dataset/JointsDataset.py:259-264,286.

Every random draw is an explicit argument (`draws`), so that the CUDA path can be
fed identical values (north_star: "random draws injected identically from a shared
seed buffer").  `make_draws` produces them from a numpy Generator in the layout the
C ABI documents (include/advmix_b200.h); `corrupt()` keeps the package's signature
and draws from the global ``np.random`` like the package does.
"""
import math
from io import BytesIO

import numpy as np

CORRUPTIONS = (
    "gaussian_noise", "shot_noise", "impulse_noise", "defocus_blur", "glass_blur",
    "motion_blur", "zoom_blur", "snow", "frost", "fog", "brightness", "contrast",
    "elastic_transform", "pixelate", "jpeg_compression",
    "speckle_noise", "gaussian_blur", "spatter", "saturate",
)
N_COMMON = 15

SEVERITY = {
    "gaussian_noise": [0.08, 0.12, 0.18, 0.26, 0.38],
    "shot_noise": [60, 25, 12, 5, 3],
    "impulse_noise": [0.03, 0.06, 0.09, 0.17, 0.27],
    "defocus_blur": [(3, 0.1), (4, 0.5), (6, 0.5), (8, 0.5), (10, 0.5)],
    "glass_blur": [(0.7, 1, 2), (0.9, 2, 1), (1, 2, 3), (1.1, 3, 2), (1.5, 4, 2)],
    "motion_blur": [(10, 3), (15, 5), (15, 8), (15, 12), (20, 15)],
    "zoom_blur": [np.arange(1, 1.11, 0.01), np.arange(1, 1.16, 0.01), np.arange(1, 1.21, 0.02),
                  np.arange(1, 1.26, 0.02), np.arange(1, 1.33, 0.03)],
    "snow": [(0.1, 0.3, 3, 0.5, 10, 4, 0.8), (0.2, 0.3, 2, 0.5, 12, 4, 0.7),
             (0.55, 0.3, 4, 0.9, 12, 8, 0.7), (0.55, 0.3, 4.5, 0.85, 12, 8, 0.65),
             (0.55, 0.3, 2.5, 0.85, 12, 12, 0.55)],
    "frost": [(1, 0.4), (0.8, 0.6), (0.7, 0.7), (0.65, 0.7), (0.6, 0.75)],
    "fog": [(1.5, 2), (2., 2), (2.5, 1.7), (2.5, 1.5), (3., 1.4)],
    "brightness": [.1, .2, .3, .4, .5],
    "contrast": [0.4, .3, .2, .1, .05],
    "elastic_transform": [250 * 0.05, 250 * 0.065, 250 * 0.085, 250 * 0.1, 250 * 0.12],
    "pixelate": [0.6, 0.5, 0.4, 0.3, 0.25],
    "jpeg_compression": [25, 18, 15, 10, 7],
    "speckle_noise": [.15, .2, 0.35, 0.45, 0.6],
    "gaussian_blur": [1, 2, 3, 4, 6],
    "spatter": [(0.65, 0.3, 4, 0.69, 0.6, 0), (0.65, 0.3, 3, 0.68, 0.6, 0), (0.65, 0.3, 2, 0.68, 0.5, 0),
                (0.65, 0.3, 1, 0.65, 1.5, 1), (0.67, 0.4, 1, 0.65, 1.5, 1)],
    "saturate": [(0.3, 0), (0.1, 0), (2, 0), (5, 0.1), (20, 0.2)],
}


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


# --------------------------------------------------------------------------- helpers
def sk_gaussian(image, sigma, mode="nearest", truncate=4.0, multichannel=True, keep_dtype=False):
    """skimage.filters.gaussian for float input = scipy gaussian_filter.  float64 images stay float64;
    with keep_dtype a float32 image stays float32 like skimage does (scipy then filters each line in
    float64 and rounds to float32 after every axis pass)."""
    from scipy import ndimage as ndi
    image = np.asarray(image) if keep_dtype else np.asarray(image, dtype=np.float64)
    if multichannel:
        sig = (sigma, sigma, 0) if np.isscalar(sigma) else tuple(sigma) + (0,)
    else:
        sig = sigma
    return ndi.gaussian_filter(image, sig, mode=mode, truncate=truncate)


def rgb2hsv(rgb):
    """skimage.color.rgb2hsv (float64)."""
    arr = np.asarray(rgb, dtype=np.float64)
    out = np.empty_like(arr)
    out_v = arr.max(-1)
    delta = np.ptp(arr, -1)
    old = np.seterr(invalid="ignore", divide="ignore")
    out_s = delta / out_v
    out_s[delta == 0.0] = 0.0
    # red is max
    idx = arr[..., 0] == out_v
    out[idx, 0] = (arr[idx, 1] - arr[idx, 2]) / delta[idx]
    # green is max
    idx = arr[..., 1] == out_v
    out[idx, 0] = 2.0 + (arr[idx, 2] - arr[idx, 0]) / delta[idx]
    # blue is max
    idx = arr[..., 2] == out_v
    out[idx, 0] = 4.0 + (arr[idx, 0] - arr[idx, 1]) / delta[idx]
    out_h = (out[..., 0] / 6.0) % 1.0
    out_h[delta == 0.0] = 0.0
    np.seterr(**old)
    out[..., 0] = out_h
    out[..., 1] = out_s
    out[..., 2] = out_v
    out[np.isnan(out)] = 0
    return out


def hsv2rgb(hsv):
    """skimage.color.hsv2rgb (float64)."""
    arr = np.asarray(hsv, dtype=np.float64)
    hi = np.floor(arr[..., 0] * 6)
    f = arr[..., 0] * 6 - hi
    p = arr[..., 2] * (1 - arr[..., 1])
    q = arr[..., 2] * (1 - f * arr[..., 1])
    t = arr[..., 2] * (1 - (1 - f) * arr[..., 1])
    v = arr[..., 2]
    hi = np.stack([hi, hi, hi], axis=-1).astype(np.uint8) % 6
    return np.choose(hi, np.stack([np.stack((v, t, p), axis=-1), np.stack((q, v, p), axis=-1),
                                   np.stack((p, v, t), axis=-1), np.stack((p, q, v), axis=-1),
                                   np.stack((t, p, v), axis=-1), np.stack((v, p, q), axis=-1)]))


def disk(radius, alias_blur=0.1, dtype=np.float32):
    import cv2
    if radius <= 8:
        L = np.arange(-8, 8 + 1)
        ksize = (3, 3)
    else:
        L = np.arange(-radius, radius + 1)
        ksize = (5, 5)
    X, Y = np.meshgrid(L, L)
    aliased_disk = np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=dtype)
    aliased_disk /= np.sum(aliased_disk)
    return cv2.GaussianBlur(aliased_disk, ksize=ksize, sigmaX=alias_blur)


def clipped_zoom(img, zoom_factor):
    from scipy.ndimage import zoom as scizoom
    ch0 = int(np.ceil(img.shape[0] / float(zoom_factor)))
    top0 = (img.shape[0] - ch0) // 2
    ch1 = int(np.ceil(img.shape[1] / float(zoom_factor)))
    top1 = (img.shape[1] - ch1) // 2
    return scizoom(img[top0:top0 + ch0, top1:top1 + ch1], (zoom_factor, zoom_factor, 1), order=1)


def motion_kernel(radius, sigma):
    width = radius * 2 + 1
    x = np.arange(width)
    k = (np.exp(-x ** 2 / (2 * (sigma ** 2)))) / (np.sqrt(2 * np.pi) * sigma)
    return k / np.sum(k)


def motion_offsets(radius, angle, H, W):
    """(dy_i, dx_i) list of _motion_blur, truncated at the package's `break`."""
    width = radius * 2 + 1
    point = (width * np.sin(np.deg2rad(angle)), width * np.cos(np.deg2rad(angle)))
    hypot = math.hypot(point[0], point[1])
    offs = []
    for i in range(width):
        dy = -math.ceil(((i * point[0]) / hypot) - 0.5)
        dx = -math.ceil(((i * point[1]) / hypot) - 0.5)
        if abs(dy) >= H or abs(dx) >= W:
            break
        offs.append((dy, dx))
    return offs


def _shift(image, dx, dy):
    # roll + replicate-edge fill == clamp-to-edge sampling
    H, W = image.shape[:2]
    ys = np.clip(np.arange(H) - dy, 0, H - 1)
    xs = np.clip(np.arange(W) - dx, 0, W - 1)
    return image[ys][:, xs]


def _motion_blur(x, radius, sigma, angle):
    kernel = motion_kernel(radius, sigma)
    blurred = np.zeros_like(x, dtype=np.float32)
    for i, (dy, dx) in enumerate(motion_offsets(radius, angle, x.shape[0], x.shape[1])):
        blurred = blurred + kernel[i] * _shift(x, dx, dy)
    return blurred


def plasma_from_uniforms(mapsize, wibbledecay, U):
    """Diamond-square plasma fractal of the package's fog().  U[i,j] in [0,1) is the
    uniform consumed for map cell (i,j) (every cell but (0,0) consumes exactly one):
    np.random.uniform(-w, w) == -w + 2w*U."""
    assert mapsize & (mapsize - 1) == 0
    maparray = np.empty((mapsize, mapsize), dtype=np.float64)
    maparray[0, 0] = 0
    stepsize = mapsize
    wibble = 100.0

    def wibbledmean(array, u):
        return array / 4 + wibble * (-wibble + (wibble - (-wibble)) * u)

    while stepsize >= 2:
        h = stepsize // 2
        cornerref = maparray[0:mapsize:stepsize, 0:mapsize:stepsize]
        squareaccum = cornerref + np.roll(cornerref, shift=-1, axis=0)
        squareaccum = squareaccum + np.roll(squareaccum, shift=-1, axis=1)
        maparray[h:mapsize:stepsize, h:mapsize:stepsize] = wibbledmean(
            squareaccum, U[h:mapsize:stepsize, h:mapsize:stepsize])
        drgrid = maparray[h:mapsize:stepsize, h:mapsize:stepsize]
        ulgrid = maparray[0:mapsize:stepsize, 0:mapsize:stepsize]
        ldrsum = drgrid + np.roll(drgrid, 1, axis=0)
        lulsum = ulgrid + np.roll(ulgrid, -1, axis=1)
        ltsum = ldrsum + lulsum
        maparray[0:mapsize:stepsize, h:mapsize:stepsize] = wibbledmean(
            ltsum, U[0:mapsize:stepsize, h:mapsize:stepsize])
        tdrsum = drgrid + np.roll(drgrid, 1, axis=1)
        tulsum = ulgrid + np.roll(ulgrid, -1, axis=0)
        ttsum = tdrsum + tulsum
        maparray[h:mapsize:stepsize, 0:mapsize:stepsize] = wibbledmean(
            ttsum, U[h:mapsize:stepsize, 0:mapsize:stepsize])
        stepsize //= 2
        wibble /= wibbledecay
    maparray -= maparray.min()
    return maparray / maparray.max()


def next_power_of_2(x):
    return 1 if x == 0 else 2 ** (x - 1).bit_length()


POISSON_KMAX = 128


def poisson_cdf_table(c):
    """[256, POISSON_KMAX] float64 CDF rows for lam = (v/255.)*c, v = 0..255.
    Recurrence (fixed op order): p0 = exp(-lam); p_k = p_{k-1}*lam/k; cdf_k = cdf_{k-1}+p_k."""
    lam = (np.arange(256) / 255.) * c
    tab = np.empty((256, POISSON_KMAX), np.float64)
    p = np.exp(-lam)
    cdf = p.copy()
    tab[:, 0] = cdf
    for k in range(1, POISSON_KMAX):
        p = p * lam / k
        cdf = cdf + p
        tab[:, k] = cdf
    return tab


def poisson_from_uniform(v_u8, c, u):
    """Inverse-CDF Poisson(lam=(v/255)*c): smallest k with u < cdf[k] (capped)."""
    tab = poisson_cdf_table(c)
    rows = tab[v_u8]                                  # [..., K]
    k = (rows <= np.asarray(u, np.float64)[..., None]).sum(-1)
    return np.minimum(k, POISSON_KMAX - 1)


def glass_shuffle(x, delta, iters, offs):
    """The package's _shuffle_pixels loop.  offs int [iters,H,W,2] = (dx,dy) for every
    (h,w); only the visited cells are consumed.  For 3-channel arrays the tuple 'swap'
    of numpy views is a one-way copy x[h,w] <- x[h+dy,w+dx] (SURVEY Hard part 3)."""
    x = x.copy()
    H, W = x.shape[:2]
    for i in range(iters):
        for h in range(H - delta, delta, -1):
            for w in range(W - delta, delta, -1):
                dx, dy = int(offs[i, h, w, 0]), int(offs[i, h, w, 1])
                x[h, w] = x[h + dy, w + dx]
    return x


def glass_shuffle_gather(x, delta, iters, offs):
    """Same result, iteration-parallel form used to validate the CUDA formulation:
    out[p] = in[root(p)], following references only through already-visited cells."""
    H, W = x.shape[:2]
    for i in range(iters):
        out = x.copy()
        for h in range(delta + 1, H - delta + 1):
            for w in range(delta + 1, W - delta + 1):
                ch, cw = h, w
                while True:
                    nh = ch + int(offs[i, ch, cw, 1])
                    nw = cw + int(offs[i, ch, cw, 0])
                    visited = (delta < nh <= H - delta) and (delta < nw <= W - delta) and \
                        (nh > ch or (nh == ch and nw > cw))
                    ch, cw = nh, nw
                    if not visited:
                        break
                out[h, w] = x[ch, cw]
        x = out
    return x


# --------------------------------------------------------------------------- draws
def make_draws(name, severity, H, W, rng, frost_bank_shape=None):
    """Random draws for one image, as float32/int arrays (C-ABI layout, see header).
    rng: np.random.Generator."""
    c = SEVERITY.get(name)
    c = c[severity - 1] if c is not None else None
    f32 = np.float32
    if name == "gaussian_noise":
        return {"field": rng.standard_normal((H, W, 3)).astype(f32)}
    if name == "shot_noise":
        return {"field": rng.random((H, W, 3), dtype=f32)}
    if name == "impulse_noise":
        return {"field": rng.random((2, H, W, 3), dtype=f32)}
    if name == "glass_blur":
        d = c[1]
        return {"field": rng.integers(-d, d, size=(c[2], H, W, 2)).astype(np.int8)}
    if name == "motion_blur":
        return {"param": np.array([rng.uniform(-45, 45), 0, 0, 0], np.float64)}
    if name == "snow":
        return {"field": rng.standard_normal((H, W)).astype(f32),
                "param": np.array([rng.uniform(-135, -45), 0, 0, 0], np.float64)}
    if name == "frost":
        n, fh, fw = frost_bank_shape[:3]
        idx = int(rng.integers(min(5, n)))
        xs = int(rng.integers(0, fh - H)) if fh > H else 0
        ys = int(rng.integers(0, fw - W)) if fw > W else 0
        return {"param": np.array([idx, xs, ys, 0], np.float64)}
    if name == "fog":
        m = next_power_of_2(int(max(H, W)))
        return {"field": rng.random((m, m), dtype=f32)}
    if name == "elastic_transform":
        return {"field": rng.random((2, H, W), dtype=f32)}
    if name == "speckle_noise":
        return {"field": rng.standard_normal((H, W, 3)).astype(f32)}
    if name == "spatter":
        return {"field": rng.standard_normal((H, W)).astype(f32)}
    return {}


# --------------------------------------------------------------------------- the ops
def _f64img(img):
    return np.array(img) / 255.


def gaussian_noise(img, severity, draws):
    c = SEVERITY["gaussian_noise"][severity - 1]
    x = _f64img(img)
    n = draws["field"].astype(np.float64) * c          # np.random.normal(scale=c) = c * N(0,1)
    return np.clip(x + n, 0, 1) * 255


def shot_noise(img, severity, draws):
    c = SEVERITY["shot_noise"][severity - 1]
    k = poisson_from_uniform(np.asarray(img), c, draws["field"])
    return np.clip(k / float(c), 0, 1) * 255


def impulse_noise(img, severity, draws):
    c = SEVERITY["impulse_noise"][severity - 1]
    x = _f64img(img).copy()
    u = draws["field"].astype(np.float64)
    flipped = u[0] < c                                  # skimage random_noise 's&p', amount=c
    salted = u[1] < 0.5                                 # salt_vs_pepper = 0.5
    x[flipped & salted] = 1
    x[flipped & ~salted] = 0
    return np.clip(x, 0, 1) * 255


def defocus_blur(img, severity, draws=None):
    import cv2
    c = SEVERITY["defocus_blur"][severity - 1]
    x = _f64img(img)
    kernel = disk(radius=c[0], alias_blur=c[1])
    channels = [cv2.filter2D(x[:, :, d], -1, kernel) for d in range(3)]
    channels = np.array(channels).transpose((1, 2, 0))
    return np.clip(channels, 0, 1) * 255


def glass_blur(img, severity, draws, gather=False):
    c = SEVERITY["glass_blur"][severity - 1]
    x = np.uint8(sk_gaussian(_f64img(img), sigma=c[0]) * 255)
    f = glass_shuffle_gather if gather else glass_shuffle
    x = f(x, c[1], c[2], draws["field"])
    return np.clip(sk_gaussian(x / 255., sigma=c[0]), 0, 1) * 255


def motion_blur(img, severity, draws):
    c = SEVERITY["motion_blur"][severity - 1]
    x = np.array(img)
    x = _motion_blur(x, radius=c[0], sigma=c[1], angle=float(draws["param"][0]))
    return np.clip(x, 0, 255)


def zoom_blur(img, severity, draws=None):
    c = SEVERITY["zoom_blur"][severity - 1]
    x = (np.array(img) / 255.).astype(np.float32)
    out = np.zeros_like(x)
    for zoom_factor in c:
        zoom_layer = clipped_zoom(x, zoom_factor)
        zoom_layer = zoom_layer[:x.shape[0], :x.shape[1], :]
        out += zoom_layer
    x = (x + out) / (len(c) + 1)
    return np.clip(x, 0, 1) * 255


def snow(img, severity, draws):
    import cv2
    c = SEVERITY["snow"][severity - 1]
    x = np.array(img, dtype=np.float32) / 255.
    snow_layer = draws["field"].astype(np.float64) * c[1] + c[0]   # normal(loc=c0, scale=c1)
    snow_layer = clipped_zoom(snow_layer[..., np.newaxis], c[2])
    snow_layer[snow_layer < c[3]] = 0
    snow_layer = np.clip(snow_layer.squeeze(), 0, 1)
    snow_layer = _motion_blur(snow_layer, radius=c[4], sigma=c[5], angle=float(draws["param"][0]))
    snow_layer = np.round(snow_layer * 255).astype(np.uint8) / 255.
    snow_layer = snow_layer[..., np.newaxis]
    snow_layer = snow_layer[:x.shape[0], :x.shape[1], :]
    gray = cv2.cvtColor(x, cv2.COLOR_RGB2GRAY).reshape(x.shape[0], x.shape[1], 1)
    x = c[6] * x + (1 - c[6]) * np.maximum(x, gray * 1.5 + 0.5)
    return np.clip(x + snow_layer + np.rot90(snow_layer, k=2), 0, 1) * 255


def frost(img, severity, draws, frost_bank):
    """frost_bank: uint8 [N,fh,fw,3] RGB textures, already >= image size (the package's
    cv2.imread + optional INTER_CUBIC upscale + BGR->RGB happen on the host side)."""
    c = SEVERITY["frost"][severity - 1]
    idx, xs, ys = (int(v) for v in draws["param"][:3])
    H, W = np.array(img).shape[:2]
    fr = frost_bank[idx][xs:xs + H, ys:ys + W]
    return np.clip(c[0] * np.array(img) + c[1] * fr, 0, 255)


def fog(img, severity, draws):
    c = SEVERITY["fog"][severity - 1]
    shape = np.array(img).shape
    map_size = next_power_of_2(int(np.max(shape)))
    x = _f64img(img)
    max_val = x.max()
    pf = plasma_from_uniforms(map_size, c[1], draws["field"].astype(np.float64))
    x = x + c[0] * pf[:shape[0], :shape[1]][..., np.newaxis]
    return np.clip(x * max_val / (max_val + c[0]), 0, 1) * 255


def brightness(img, severity, draws=None):
    c = SEVERITY["brightness"][severity - 1]
    x = rgb2hsv(_f64img(img))
    x[:, :, 2] = np.clip(x[:, :, 2] + c, 0, 1)
    x = hsv2rgb(x)
    return np.clip(x, 0, 1) * 255


def contrast(img, severity, draws=None):
    c = SEVERITY["contrast"][severity - 1]
    x = _f64img(img)
    means = np.mean(x, axis=(0, 1), keepdims=True)
    return np.clip((x - means) * c + means, 0, 1) * 255


def elastic_transform(img, severity, draws):
    from scipy.ndimage import map_coordinates
    image = np.array(img, dtype=np.float32) / 255.
    shape = image.shape
    sigma = np.array(shape[:2]) * 0.01
    alpha = SEVERITY["elastic_transform"][severity - 1]
    max_dx = shape[0] * 0.005
    max_dy = shape[0] * 0.005
    u = draws["field"].astype(np.float64)
    ux = -max_dx + (max_dx - (-max_dx)) * u[0]          # np.random.uniform(-max, max)
    uy = -max_dy + (max_dy - (-max_dy)) * u[1]
    dx = (sk_gaussian(ux, sigma, mode="reflect", truncate=3, multichannel=False) * alpha).astype(np.float32)
    dy = (sk_gaussian(uy, sigma, mode="reflect", truncate=3, multichannel=False) * alpha).astype(np.float32)
    dx, dy = dx[..., np.newaxis], dy[..., np.newaxis]
    x, y, z = np.meshgrid(np.arange(shape[1]), np.arange(shape[0]), np.arange(shape[2]))
    indices = np.reshape(y + dy, (-1, 1)), np.reshape(x + dx, (-1, 1)), np.reshape(z, (-1, 1))
    return np.clip(map_coordinates(image, indices, order=1, mode="reflect").reshape(shape), 0, 1) * 255


def pixelate(img, severity, draws=None):
    from PIL import Image
    c = SEVERITY["pixelate"][severity - 1]
    im = Image.fromarray(np.asarray(img))
    H, W = np.asarray(img).shape[:2]
    im = im.resize((int(W * c), int(H * c)), Image.BOX)
    im = im.resize((W, H), Image.NEAREST)
    return np.array(im)


def jpeg_compression(img, severity, draws=None):
    from PIL import Image
    c = SEVERITY["jpeg_compression"][severity - 1]
    out = BytesIO()
    Image.fromarray(np.asarray(img)).save(out, "JPEG", quality=c)
    return np.array(Image.open(out))


# --------------------------------------------------------------------------- the 4 'validation' ops
def speckle_noise(img, severity, draws):
    c = SEVERITY["speckle_noise"][severity - 1]
    x = _f64img(img)
    n = draws["field"].astype(np.float64) * c          # np.random.normal(size, scale=c)
    return np.clip(x + x * n, 0, 1) * 255


def gaussian_blur(img, severity, draws=None):
    c = SEVERITY["gaussian_blur"][severity - 1]
    x = sk_gaussian(_f64img(img), sigma=c)
    return np.clip(x, 0, 1) * 255


def saturate(img, severity, draws=None):
    c = SEVERITY["saturate"][severity - 1]
    x = rgb2hsv(_f64img(img))
    x[:, :, 1] = np.clip(x[:, :, 1] * c[0] + c[1], 0, 1)
    x = hsv2rgb(x)
    return np.clip(x, 0, 1) * 255


def spatter_water_mask(liquid_u8):
    """The cv2 chain of spatter severities 1-3 (Canny -> chamfer distance -> blur -> equalise -> emboss
    -> blur), calling cv2 itself: uint8 [H,W] -> float32 [H,W]."""
    import cv2
    dist = 255 - cv2.Canny(liquid_u8, 50, 150)
    dist = cv2.distanceTransform(dist, cv2.DIST_L2, 5)
    _, dist = cv2.threshold(dist, 20, 20, cv2.THRESH_TRUNC)
    dist = cv2.blur(dist, (3, 3)).astype(np.uint8)
    dist = cv2.equalizeHist(dist)
    ker = np.array([[-2, -1, 0], [-1, 1, 1], [0, 1, 2]])
    dist = cv2.filter2D(dist, cv2.CV_8U, ker)
    return cv2.blur(dist, (3, 3)).astype(np.float32)


def spatter(img, severity, draws):
    import cv2
    c = SEVERITY["spatter"][severity - 1]
    x = np.array(img, dtype=np.float32) / 255.
    liquid_layer = c[0] + c[1] * draws["field"].astype(np.float64)      # np.random.normal(loc, scale)
    liquid_layer = sk_gaussian(liquid_layer, sigma=c[2], multichannel=False)
    liquid_layer[liquid_layer < c[3]] = 0
    if c[5] == 0:
        liquid_layer = (liquid_layer * 255).astype(np.uint8)
        dist = spatter_water_mask(liquid_layer)
        m = cv2.cvtColor(liquid_layer * dist, cv2.COLOR_GRAY2BGRA)
        m /= np.max(m, axis=(0, 1))
        m *= c[4]
        # water is pale turquoise
        color = np.concatenate((175 / 255. * np.ones_like(m[..., :1]), 238 / 255. * np.ones_like(m[..., :1]),
                                238 / 255. * np.ones_like(m[..., :1])), axis=2)
        color = cv2.cvtColor(color, cv2.COLOR_BGR2BGRA)
        x = cv2.cvtColor(x, cv2.COLOR_BGR2BGRA)
        return cv2.cvtColor(np.clip(x + m * color, 0, 1), cv2.COLOR_BGRA2BGR) * 255
    m = np.where(liquid_layer > c[3], 1, 0)
    m = sk_gaussian(m.astype(np.float32), sigma=c[4], multichannel=False, keep_dtype=True)
    m[m < 0.8] = 0
    # mud brown; np.ones_like of the uint8 image -> the colour planes are float64
    ones = np.ones_like(np.asarray(img)[..., :1])
    color = np.concatenate((63 / 255. * ones, 42 / 255. * ones, 20 / 255. * ones), axis=2)
    color *= m[..., np.newaxis]
    x *= (1 - m[..., np.newaxis])
    return np.clip(x + color, 0, 1) * 255


_OPS = {
    "gaussian_noise": gaussian_noise, "shot_noise": shot_noise, "impulse_noise": impulse_noise,
    "defocus_blur": defocus_blur, "glass_blur": glass_blur, "motion_blur": motion_blur,
    "zoom_blur": zoom_blur, "snow": snow, "frost": frost, "fog": fog, "brightness": brightness,
    "contrast": contrast, "elastic_transform": elastic_transform, "pixelate": pixelate,
    "jpeg_compression": jpeg_compression,
    "speckle_noise": speckle_noise, "gaussian_blur": gaussian_blur, "spatter": spatter, "saturate": saturate,
}


def synthetic_frost_bank(n=5, fh=384, fw=384, seed=7):
    """Stand-in for the package's frost1-6 textures (unavailable): smooth bright blobs."""
    import cv2
    rng = np.random.default_rng(seed)
    bank = np.empty((n, fh, fw, 3), np.uint8)
    for i in range(n):
        low = rng.random((fh // 16 + 1, fw // 16 + 1, 3)).astype(np.float32)
        t = cv2.resize(low, (fw, fh), interpolation=cv2.INTER_CUBIC)
        t = t + 0.15 * rng.random((fh, fw, 3)).astype(np.float32)
        bank[i] = np.clip(t * 200 + 40, 0, 255).astype(np.uint8)
    return bank


def corrupt_with_draws(image, severity, name, draws, frost_bank=None):
    """uint8 [H,W,3] -> uint8 [H,W,3]; final np.uint8() is C truncation like the package."""
    if name == "frost":
        r = frost(image, severity, draws, frost_bank)
    else:
        r = _OPS[name](image, severity, draws)
    return np.uint8(r)


def corrupt(image, severity=1, corruption_name=None, corruption_number=-1, frost_bank=None):
    """Package-signature entry point; draws from the global np.random (make_datasets.py:40
    seeds it with 1 before every call)."""
    if not isinstance(image, np.ndarray):
        raise AttributeError("Expecting type(image) to be numpy.ndarray")
    if not (image.dtype.type is np.uint8):
        raise AttributeError("Expecting image.dtype.type to be numpy.uint8")
    if not (image.ndim in [2, 3]):
        raise AttributeError("Expecting image.shape to be either (height x width) or (height x width x channels)")
    if image.ndim == 2:
        image = np.stack((image,) * 3, axis=-1)
    height, width, channels = image.shape
    if height < 32 or width < 32:
        raise AttributeError("Image width and height must be at least 32 pixels")
    if not (channels in [1, 3]):
        raise AttributeError("Expecting image to have either 1 or 3 channels (last dimension)")
    if channels == 1:
        image = np.stack((np.squeeze(image),) * 3, axis=-1)
    if not (severity in [1, 2, 3, 4, 5]):
        raise AttributeError("Severity must be an integer in [1, 5]")
    if corruption_name is None and corruption_number == -1:
        raise ValueError("Either corruption_name or corruption_number must be passed")
    name = corruption_name if corruption_name is not None else CORRUPTIONS[corruption_number]
    if name not in _OPS:
        raise ValueError("unknown corruption %r" % name)
    rng = np.random.default_rng(np.random.randint(0, 2 ** 31 - 1))
    if name == "frost" and frost_bank is None:
        frost_bank = synthetic_frost_bank(fh=max(384, height + 32), fw=max(384, width + 32))
    draws = make_draws(name, severity, height, width, rng,
                       frost_bank.shape if frost_bank is not None else None)
    return corrupt_with_draws(image, severity, name, draws, frost_bank)
