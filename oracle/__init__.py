"""CPU oracle for the AdvMix augmentation + target hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / reported CPU
baseline.  ``advmix_b200`` never imports this package and has no CPU fallback.

Every function cites the reference file:line it restates (``/root/reference`` =
AIprogrammer/AdvMix).  Pinning status per row of SURVEY.md section 8:

* a1 warp / a4 targets / a5 autoaug / a6 gridmask / a7 normalise / a3 mix: pinned.
  The restatements are checked bit-for-bit against the *real* reference functions
  (imported through ``oracle/ref_harness.py`` in the build container) and against
  the real third-party calls the reference makes (``cv2.warpAffine``, PIL,
  torchvision), and the resulting vectors are committed under ``tests/golden/``
  (generator: ``oracle/make_golden.py``).
* a2 ``imagecorruptions``: **parity unpinned**.  The package (PyPI
  ``imagecorruptions``, un-pinned in the reference's requirements.txt:12, latest
  known 1.1.2) is not vendored in the reference, not installed and not
  installable offline, and the reference has no tests or golden vectors.
  ``oracle/corruptions.py`` restates its published algorithm from the same
  scipy / cv2 / PIL primitives the package calls, anchored on the reference's
  call sites (tools/make_datasets.py:38-41, lib/dataset/JointsDataset.py:259-286).
  The same holds for the 4 validation operators (speckle_noise, gaussian_blur, spatter,
  saturate); the cv2 stages of spatter call cv2 itself, so that part is pinned to cv2.
* f1 JPEG decode: the oracle is cv2.imdecode / PIL themselves (libjpeg-turbo), called in
  the tests - pinned by construction.  JPEG encode (tools/make_datasets.py:45): the oracle is PIL's
  ``Image.save`` itself, the device files must equal its bytes - pinned by construction.
* f3 heat-map consumers (``oracle/inference.py``) and f4 record helpers
  (``oracle/records.py``): pinned to fixtures produced by the real
  ``get_max_preds`` / ``get_final_preds`` / ``flip_back`` / ``half_body_transform`` /
  ``select_data`` (``tests/golden/inference.npz``, ``records.npz``); ``_xywh2cs`` is
  restated from lib/dataset/coco.py:205-220 (that module needs pycocotools to import).
"""
