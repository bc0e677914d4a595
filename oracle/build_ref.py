"""Recipe for oracle/_ref: the reference's own modules for this path, byte-compiled.

TEST INFRASTRUCTURE - see oracle/__init__.py.

The reference is Python, so "compiling it from the sources where they lie" means byte-compiling:
`python -m oracle.build_ref` runs py_compile on the handful of reference modules the hot path lives in
(read in place under /root/reference, never copied) and writes only the resulting sourceless byte-code
files under oracle/_ref/lib/ - with the extension `.refc`, because snapshot tools (gpurun included) drop `*.pyc`.  oracle/_ref/ is git-ignored (stays out of history) but travels to the
GPU box with the snapshot, like the built libadvmix_b200.so; the reference tree itself does not exist
there.  oracle/ref_harness.py imports the real modules from /root/reference when it exists and from
oracle/_ref/lib otherwise, so on the GPU box
  * `bench.py --impl reference` / `cpu_baseline` time the REAL JointsDataset.__getitem__
    (cpu_baseline.kind = "reference"), and
  * tests/test_gpu_train_advmix.py drives the REAL lib/core/function.py:train_advmix
with no reference source in the repository.  The .pyc files are only valid for the interpreter that
wrote them (same image on both boxes); ref_harness falls back to "unavailable" when they do not load.
"""
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "lib")
REF = os.environ.get("ADVMIX_REFERENCE", "/root/reference")

# module (relative to <reference>/lib) -> why it is needed
MODULES = {
    "utils/__init__.py": "package marker",
    "utils/transforms.py": "get_affine_transform, affine_transform, fliplr_joints, flip_back, transform_preds",
    "utils/vis.py": "imported by core/function.py (save_debug_images)",
    "core/inference.py": "get_max_preds / get_final_preds (argmax parity checker)",
    "core/evaluate.py": "accuracy (called by train_advmix)",
    "core/loss.py": "JointsMSELoss (criterion of the train_advmix drive test)",
    "core/function.py": "train_advmix (lines 107-197): the consumer of the a8 contract and the a3 mix",
    "dataset/JointsDataset.py": "__getitem__ / get_base / get_var / get_clean / generate_target",
    "dataset/advaug.py": "MixCombine, ImageNetPolicy, grid_aug",
}


def build(verbose=True):
    """Byte-compile the listed reference modules into oracle/_ref/lib.  Returns True if it was (re)built,
    False if the reference tree is absent (the prebuilt files, if any, are left alone)."""
    lib = os.path.join(REF, "lib")
    if not os.path.isdir(lib):
        if verbose:
            print("oracle/_ref: reference tree not present at %s - keeping prebuilt files" % REF)
        return False
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    for rel in MODULES:
        src = os.path.join(lib, rel)
        dst = os.path.join(OUT, rel[:-3] + ".refc")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, doraise=True)
    with open(os.path.join(OUT, "PYTHON_TAG"), "w") as f:
        f.write(sys.implementation.cache_tag + "\n")
    if verbose:
        print("oracle/_ref: byte-compiled %d reference modules into %s" % (len(MODULES), OUT))
    return True


if __name__ == "__main__":
    build()
