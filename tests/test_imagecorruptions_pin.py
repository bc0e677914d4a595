"""Conditional pin of SURVEY rows a2 / f2: `imagecorruptions` (requirements.txt:12 of the reference) is not
vendored, not installed and not installable offline, so oracle/corruptions.py is a restatement of the package's
published algorithm ("parity unpinned").  The day the real package is importable (site-packages or a copy under
baseline/_ref/) these tests compare the restatement with it; until then they skip.

Draw-free ops are compared directly.  Ops that draw from the global np.random are compared by feeding the oracle
the very draws the package consumed: np.random.seed(1) before the package call (tools/make_datasets.py:40), then
the same seed and the same sequence of np.random calls to build the oracle's explicit draw arrays."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for extra in (os.path.join(ROOT, "baseline", "_ref"), os.path.join(ROOT, "oracle", "_ref", "site-packages")):
    if os.path.isdir(extra) and extra not in sys.path:
        sys.path.append(extra)

IC = pytest.importorskip("imagecorruptions", reason="the real imagecorruptions package is not available (parity unpinned)")
if not hasattr(IC, "corruptions"):        # the harness stub oracle/ref_harness.py installs is not the package
    pytest.skip("imagecorruptions in sys.modules is the harness stub", allow_module_level=True)

from oracle import corruptions as OK      # noqa: E402


def _image(H=256, W=192, seed=0):
    import cv2
    rng = np.random.default_rng(seed)
    low = rng.random((H // 16 + 2, W // 16 + 2, 3)).astype(np.float32)
    img = cv2.resize(low, (W, H), interpolation=cv2.INTER_CUBIC) * 255 + rng.normal(0, 8, (H, W, 3))
    img = np.clip(img, 0, 255).astype(np.uint8)
    img[:24, :24] = 255; img[-24:, -24:] = 0
    return img


def test_names_and_order():
    for subset in ("common", "validation", "all", "noise", "blur", "weather", "digital"):
        assert OK.get_corruption_names(subset) == IC.get_corruption_names(subset)


@pytest.mark.parametrize("name", ["defocus_blur", "zoom_blur", "brightness", "contrast", "pixelate", "jpeg_compression",
                                  "saturate", "gaussian_blur"])
@pytest.mark.parametrize("severity", [1, 2, 3, 4, 5])
def test_draw_free_ops_equal_package(name, severity):
    img = _image()
    exp = IC.corrupt(img, severity=severity, corruption_name=name)
    got = OK.corrupt_with_draws(img, severity, name, {})
    assert np.array_equal(got, exp), (name, severity, int(np.abs(got.astype(int) - exp.astype(int)).max()))


@pytest.mark.parametrize("severity", [1, 3, 5])
def test_gaussian_and_speckle_noise_with_package_draws(severity):
    img = _image()
    for name in ("gaussian_noise", "speckle_noise"):
        np.random.seed(1)
        exp = IC.corrupt(img, severity=severity, corruption_name=name)
        c = OK.SEVERITY[name][severity - 1]
        np.random.seed(1)
        field = np.random.normal(size=img.shape, scale=c) / c          # the N(0,1) field the package scaled by c
        got = OK.corrupt_with_draws(img, severity, name, {"field": field.astype(np.float64)})
        diff = np.abs(got.astype(int) - exp.astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() < 2e-3, (name, severity, int(diff.max()))


@pytest.mark.parametrize("severity", [1, 3, 5])
def test_motion_blur_and_snow_with_package_draws(severity):
    img = _image()
    np.random.seed(1)
    exp = IC.corrupt(img, severity=severity, corruption_name="motion_blur")
    np.random.seed(1)
    angle = np.random.uniform(-45, 45)
    got = OK.corrupt_with_draws(img, severity, "motion_blur", {"param": np.array([angle, 0, 0, 0.0])})
    assert np.abs(got.astype(int) - exp.astype(int)).max() <= 1
    np.random.seed(1)
    exp = IC.corrupt(img, severity=severity, corruption_name="snow")
    c = OK.SEVERITY["snow"][severity - 1]
    np.random.seed(1)
    layer = (np.random.normal(size=img.shape[:2], loc=c[0], scale=c[1]) - c[0]) / c[1]
    angle = np.random.uniform(-135, -45)
    got = OK.corrupt_with_draws(img, severity, "snow", {"field": layer.astype(np.float32), "param": np.array([angle, 0, 0, 0.0])})
    assert np.abs(got.astype(int) - exp.astype(int)).max() <= 1


@pytest.mark.parametrize("severity", [1, 3, 5])
def test_elastic_and_fog_with_package_draws(severity):
    img = _image()
    H, W = img.shape[:2]
    np.random.seed(1)
    exp = IC.corrupt(img, severity=severity, corruption_name="elastic_transform")
    np.random.seed(1)
    m = H * 0.01 * 0.5
    u0 = (np.random.uniform(-m, m, size=(H, W)) + m) / (2 * m)
    u1 = (np.random.uniform(-m, m, size=(H, W)) + m) / (2 * m)
    got = OK.corrupt_with_draws(img, severity, "elastic_transform", {"field": np.stack([u0, u1]).astype(np.float32)})
    diff = np.abs(got.astype(int) - exp.astype(int))
    assert diff.max() <= 1 or (diff > 1).mean() < 1e-3       # float32 uniforms vs the package's float64 ones
