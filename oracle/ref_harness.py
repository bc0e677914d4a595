"""Import the REAL reference modules for pinning: from /root/reference (read-only) where it
exists (the build container), otherwise from the byte-compiled copy oracle/build_ref.py leaves under
oracle/_ref/lib (the GPU box).

TEST INFRASTRUCTURE - see oracle/__init__.py.  ``available()`` says whether either is there.  Used by
oracle/make_golden.py to produce tests/golden/*.npz, by the pinning tests, by the train_advmix drive
test and by bench.py's reference arm.  Three harness-side shims, none of
which touch the reference tree:
  1. a stub ``imagecorruptions`` module (hard import at JointsDataset.py:23);
  2. ``np.int`` / ``np.float`` aliases (advaug.py:55, coco.py:174 on numpy>=1.24);
  3. a synthetic ``dataset`` package that bypasses lib/dataset/__init__.py
     (pycocotools / json_tricks are absent).
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("ADVMIX_REFERENCE", "/root/reference")
_PYC_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_cache = {}


def _lib_dir():
    """<root>/lib of the live reference tree, else of the byte-compiled oracle/_ref, else None."""
    if os.path.isfile(os.path.join(REF_ROOT, "lib", "dataset", "JointsDataset.py")):
        return os.path.join(REF_ROOT, "lib")
    tag = os.path.join(_PYC_ROOT, "lib", "PYTHON_TAG")
    if os.path.isfile(os.path.join(_PYC_ROOT, "lib", "dataset", "JointsDataset.refc")) and os.path.isfile(tag):
        if open(tag).read().strip() == sys.implementation.cache_tag:
            return os.path.join(_PYC_ROOT, "lib")
    return None


def available():
    return _lib_dir() is not None


def kind():
    """'source' (live tree), 'pyc' (oracle/_ref) or None."""
    d = _lib_dir()
    if d is None:
        return None
    return "source" if d.startswith(REF_ROOT) else "pyc"


class _RefcFinder(importlib.abc.MetaPathFinder):
    """Finds the reference's `utils.*`, `core.*`, `dataset.*` modules among the byte-compiled `.refc` files of oracle/_ref/lib."""

    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        parts = fullname.split(".")
        if parts[0] not in ("utils", "core", "dataset"):
            return None
        base = os.path.join(self.root, *parts)
        if os.path.isfile(base + ".refc"):
            loader = importlib.machinery.SourcelessFileLoader(fullname, base + ".refc")
            return importlib.util.spec_from_file_location(fullname, base + ".refc", loader=loader)
        init = os.path.join(base, "__init__.refc")
        if os.path.isfile(init):
            loader = importlib.machinery.SourcelessFileLoader(fullname, init)
            return importlib.util.spec_from_file_location(fullname, init, loader=loader, submodule_search_locations=[base])
        if os.path.isdir(base):                       # core/, dataset/: packages without an __init__ here
            spec = importlib.machinery.ModuleSpec(fullname, None, is_package=True)
            spec.submodule_search_locations = [base]
            return spec
        return None


def _install_shims(corrupt=None):
    lib = _lib_dir()
    if kind() == "pyc":
        if not any(isinstance(f, _RefcFinder) for f in sys.meta_path):
            sys.meta_path.insert(0, _RefcFinder(lib))
    elif lib not in sys.path:
        sys.path.insert(0, lib)
    if "imagecorruptions" not in sys.modules:
        stub = types.ModuleType("imagecorruptions")

        def _corrupt(image, severity=1, corruption_name=None, corruption_number=-1):
            from oracle import corruptions
            return corruptions.corrupt(image, severity, corruption_name, corruption_number)

        def _names(subset="common"):
            from oracle import corruptions
            return corruptions.get_corruption_names(subset)

        stub.corrupt = corrupt or _corrupt
        stub.get_corruption_names = _names
        sys.modules["imagecorruptions"] = stub
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    if "dataset" not in sys.modules:
        pkg = types.ModuleType("dataset")
        pkg.__path__ = [os.path.join(lib, "dataset")]
        sys.modules["dataset"] = pkg


def load():
    """-> namespace with the reference's transforms, JointsDataset, advaug modules."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference not present at %s nor byte-compiled under %s" % (REF_ROOT, _PYC_ROOT))
    _install_shims()
    ns = types.SimpleNamespace()
    ns.transforms = importlib.import_module("utils.transforms")
    ns.JointsDataset = importlib.import_module("dataset.JointsDataset")
    ns.advaug = importlib.import_module("dataset.advaug")
    ns.inference = importlib.import_module("core.inference")
    ns.function = importlib.import_module("core.function")
    ns.loss = importlib.import_module("core.loss")
    _cache["ns"] = ns
    return ns


COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]
COCO_UPPER = (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10)
COCO_JOINTS_WEIGHT = np.array([1., 1., 1., 1., 1., 1., 1., 1.2, 1.2, 1.5, 1.5, 1., 1., 1.2, 1.2,
                               1.5, 1.5], dtype=np.float32).reshape((17, 1))


def make_cfg(image_size=(192, 256), heatmap_size=(48, 64), sigma=2, scale_factor=0.3,
             rot_factor=40, flip=True, color_rgb=False, prob_half_body=0.0,
             num_joints_half_body=8, use_different_joints_weight=False):
    N = types.SimpleNamespace
    return N(OUTPUT_DIR="", DATASET=N(DATA_FORMAT="jpg", SCALE_FACTOR=scale_factor,
                                      ROT_FACTOR=rot_factor, FLIP=flip,
                                      NUM_JOINTS_HALF_BODY=num_joints_half_body,
                                      PROB_HALF_BODY=prob_half_body, COLOR_RGB=color_rgb),
             MODEL=N(TARGET_TYPE="gaussian", IMAGE_SIZE=list(image_size),
                     HEATMAP_SIZE=list(heatmap_size), SIGMA=sigma),
             LOSS=N(USE_DIFFERENT_JOINTS_WEIGHT=use_different_joints_weight))


def make_dataset(db, is_train=True, sample_times=3, num_joints=17, transform="default", **cfg_kw):
    """A real reference JointsDataset wired like COCODataset would (coco.py:42,70-83)."""
    ns = load()
    from torchvision import transforms as T
    if transform == "default":
        transform = T.Compose([T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406],
                                                         std=[0.229, 0.224, 0.225])])
    cfg = make_cfg(**cfg_kw)
    args = types.SimpleNamespace(sample_times=sample_times, random_corruption=False,
                                 sp_style=False, joints_num=num_joints)
    ds = ns.JointsDataset.JointsDataset(cfg, args, "", "train", is_train, transform)
    ds.num_joints = num_joints
    ds.flip_pairs = COCO_FLIP_PAIRS if num_joints == 17 else []
    ds.upper_body_ids = COCO_UPPER
    w, h = cfg.MODEL.IMAGE_SIZE
    ds.aspect_ratio = w * 1.0 / h
    ds.joints_weight = COCO_JOINTS_WEIGHT if num_joints == 17 else 1
    ds.db = db
    return ds
