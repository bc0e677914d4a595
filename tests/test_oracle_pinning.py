"""not-gpu: pin the oracle.  a1/a3/a4/a5/a6/a7 against fixtures produced by the REAL reference
(oracle/make_golden.py) and against the real third-party calls (cv2, PIL, torchvision); when the
reference tree is present (build container) also against the reference functions directly.
a2 (imagecorruptions) is *parity unpinned* (package absent): only known-answer and internal
consistency checks are possible, see oracle/corruptions.py."""
import numpy as np
import pytest

from oracle import affine as OA
from oracle import chains as OC
from oracle import corruptions as OK
from oracle import mix as OM
from oracle import ref_harness
from oracle import targets as OT


def test_warp_restatement_equals_cv2_and_golden(golden):
    g = golden("warp")
    for i in range(int(g["n"])):
        src = g["src%d" % i]
        view = src[:, ::-1, :] if int(g["flip%d" % i]) else src
        ds = tuple(int(v) for v in g["dsize%d" % i])
        out = OA.warp_affine_fixedpoint(view, g["trans%d" % i], ds)
        assert np.array_equal(out, g["dst%d" % i])
        assert np.array_equal(OA.warp_affine_cv2(view, g["trans%d" % i], ds), g["dst%d" % i])
        M = OA.get_affine_transform(g["center%d" % i], g["scale%d" % i], float(g["rot%d" % i]), ds)
        assert np.array_equal(M, g["trans%d" % i])
        M2 = OA.get_affine_transform(g["center%d" % i], g["scale%d" % i], float(g["rot%d" % i]), ds, use_cv2=False)
        np.testing.assert_allclose(M2, g["trans%d" % i], rtol=0, atol=1e-9)
        j2, v2 = g["joints%d" % i].copy(), g["vis%d" % i].copy()
        if int(g["flip%d" % i]):
            j2, v2 = OA.fliplr_joints(j2, v2, src.shape[1], ref_harness.COCO_FLIP_PAIRS)
        j2 = OA.transform_joints(j2, v2, g["trans%d" % i])
        assert np.array_equal(j2, g["joints_out%d" % i]) and np.array_equal(v2, g["vis_out%d" % i])
    assert np.array_equal(OA.normalize_lut(), g["norm_lut"])


def test_warp_restatement_random_vs_cv2():
    rng = np.random.default_rng(0)
    for it in range(40):
        H, W = int(rng.integers(60, 300)), int(rng.integers(60, 300))
        src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        c = np.array([rng.uniform(-0.2, 1.2) * W, rng.uniform(-0.2, 1.2) * H], np.float32)
        s = np.array([rng.uniform(0.2, 2.5), rng.uniform(0.2, 2.5)], np.float32)
        ds = [(192, 256), (64, 64), (50, 37)][it % 3]
        M = OA.get_affine_transform(c, s, float(rng.uniform(-80, 80)), ds)
        assert np.array_equal(OA.warp_affine_fixedpoint(src, M, ds), OA.warp_affine_cv2(src, M, ds))


@pytest.mark.parametrize("tag,img,hm,jw", [("coco", (192, 256), (48, 64), False), ("cocow", (192, 256), (48, 64), True),
                                           ("mpii", (256, 256), (64, 64), False), ("big", (512, 512), (128, 128), False)])
def test_targets_oracle_equals_reference_fixture(golden, tag, img, hm, jw):
    g = golden("targets")
    jwv = ref_harness.COCO_JOINTS_WEIGHT if jw else None
    for i in range(len(g[tag + "_joints"])):
        t, w = OT.generate_target(g[tag + "_joints"][i], g[tag + "_vis"][i], img, hm, 2, jwv)
        assert np.array_equal(t[0], g[tag + "_hm"][i])
        assert np.array_equal(t[1], g[tag + "_mu"][i])
        assert np.array_equal(w, g[tag + "_tw"][i])
    preds, maxvals = OT.get_max_preds(g[tag + "_hm"])
    assert np.array_equal(preds, g[tag + "_preds"]) and np.array_equal(maxvals, g[tag + "_maxvals"])


def test_mix_oracle_equals_reference_fixture(golden):
    import torch
    g = golden("mix")
    inputs = [torch.from_numpy(x) for x in g["inputs"]]
    tmp, w = OM.mix_from_logits(inputs, torch.from_numpy(g["logits"]))
    assert np.array_equal(tmp.numpy(), g["tmp"]) and np.array_equal(w.numpy(), g["weights"])
    gl = OM.mix_backward(inputs, torch.from_numpy(g["logits"]), torch.from_numpy(g["grad_out"]))
    assert np.array_equal(gl.numpy(), g["grad_logits"])


def test_chain_oracles_equal_reference_fixture(golden):
    g = golden("chains")
    for i, (pidx, c1, c2, s1, s2) in enumerate(g["aa_plan"]):
        out = OC.autoaug(g["aa_in"][i], int(pidx), c1, c2, int(s1), int(s2))
        assert np.array_equal(out, g["aa_out"][i]), i
        assert np.array_equal(OC.autoaug(g["aa_in"][i], int(pidx), c1, c2, int(s1), int(s2), use_pil=True), g["aa_out"][i])
    for i in range(len(g["gm_in"])):
        p = g["gm_params"][i]
        img, vis = OC.gridmask(g["gm_in"][i], g["gm_joints"][i], g["gm_vis"][i], bool(p[0]), int(p[1]), int(p[2]), int(p[3]))
        assert np.array_equal(img, g["gm_out"][i]) and np.array_equal(vis, g["gm_vis_out"][i])


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present (GPU box)")
def test_oracle_against_live_reference():
    ns = ref_harness.load()
    ds = ref_harness.make_dataset([], is_train=True)
    rng = np.random.default_rng(5)
    for it in range(50):
        j = np.zeros((17, 3)); j[:, :2] = rng.uniform(-40, 300, (17, 2))
        v = np.zeros((17, 3)); vv = (rng.random(17) < 0.8).astype(float); v[:, 0] = vv; v[:, 1] = vv
        t, w = ds.generate_target(j.copy(), v.copy())
        t2, w2 = OT.generate_target(j, v)
        assert np.array_equal(t[0], t2[0]) and np.array_equal(t[1], t2[1]) and np.array_equal(w, w2)
        c = rng.uniform(0, 500, 2).astype(np.float32); s = rng.uniform(0.3, 3, 2).astype(np.float32); r = rng.uniform(-80, 80)
        assert np.array_equal(ns.transforms.get_affine_transform(c, s, r, [192, 256]), OA.get_affine_transform(c, s, r, [192, 256]))


# ---- a2: imagecorruptions restatement (parity unpinned) -----------------------------------------
def _img(rng, H=64, W=48):
    import cv2
    low = rng.random((H // 16 + 2, W // 16 + 2, 3)).astype(np.float32)
    return np.clip(cv2.resize(low, (W, H), interpolation=cv2.INTER_CUBIC) * 255 + rng.normal(0, 8, (H, W, 3)), 0, 255).astype(np.uint8)


def test_corruption_api_surface_and_errors():
    assert OK.get_corruption_names() == OK.get_corruption_names("common") == list(OK.CORRUPTIONS[:15])
    assert len(OK.get_corruption_names("all")) == 19 and OK.get_corruption_names("validation") == list(OK.CORRUPTIONS[15:])
    assert OK.get_corruption_names("noise") == ["gaussian_noise", "shot_noise", "impulse_noise"]
    assert OK.get_corruption_names("blur") == ["defocus_blur", "glass_blur", "motion_blur", "zoom_blur"]
    assert OK.get_corruption_names("weather") == ["snow", "frost", "fog", "brightness"]
    assert OK.get_corruption_names("digital") == ["contrast", "elastic_transform", "pixelate", "jpeg_compression"]
    with pytest.raises(ValueError):
        OK.get_corruption_names("nope")
    img = np.zeros((64, 48, 3), np.uint8)
    with pytest.raises(AttributeError):
        OK.corrupt(img.astype(np.float32), 1, "contrast")
    with pytest.raises(AttributeError):
        OK.corrupt(img[:31], 1, "contrast")
    with pytest.raises(AttributeError):
        OK.corrupt(img, 6, "contrast")
    with pytest.raises(ValueError):
        OK.corrupt(img, 1)
    np.random.seed(1)
    a = OK.corrupt(img + 100, 2, "gaussian_noise")
    np.random.seed(1)
    b = OK.corrupt(img + 100, 2, corruption_number=0)
    assert a.dtype == np.uint8 and a.shape == (64, 48, 3) and np.array_equal(a, b)


def test_corruption_known_answers_and_consistency():
    rng = np.random.default_rng(1)
    img = _img(rng)
    # (v/255.)*255 truncates back to v: untouched impulse-noise pixels keep their value
    v = np.arange(256)
    assert np.array_equal(np.uint8((v / 255.) * 255), v)
    d = OK.make_draws("impulse_noise", 3, 64, 48, rng)
    out = OK.corrupt_with_draws(img, 3, "impulse_noise", d)
    keep = d["field"][0] >= 0.09
    assert np.array_equal(out[keep], img[keep]) and set(np.unique(out[~keep])) <= {0, 255}
    # glass blur: sequential in-place scan == iteration-parallel root-following gather
    for sev in (1, 4, 5):
        dg = OK.make_draws("glass_blur", sev, 40, 36, rng)
        assert np.array_equal(OK.glass_blur(img[:40, :36], sev, dg), OK.glass_blur(img[:40, :36], sev, dg, gather=True))
    # plasma fractal is normalised to [0, 1]
    pf = OK.plasma_from_uniforms(64, 2.0, rng.random((64, 64)))
    assert pf.min() == 0.0 and pf.max() == 1.0
    # inverse-CDF Poisson: mean / variance ~ lambda
    u = rng.random(200000)
    k = OK.poisson_from_uniform(np.full(200000, 128, np.uint8), 60, u)
    lam = 128 / 255. * 60
    assert abs(k.mean() - lam) < 0.1 and abs(k.var() - lam) < 0.5
    # zoom factor counts with this numpy (float arange end points)
    assert [len(c) for c in OK.SEVERITY["zoom_blur"]] == [11, 16, 11, 13, 11] or [len(c) for c in OK.SEVERITY["zoom_blur"]] == [12, 16, 11, 13, 12]
    # every op: uint8 HxWx3 out, deterministic given the draws, and actually changes the image
    bank = OK.synthetic_frost_bank(fh=128, fw=96)
    for name in OK.get_corruption_names():
        d = OK.make_draws(name, 3, 64, 48, rng, bank.shape)
        a = OK.corrupt_with_draws(img, 3, name, d, bank)
        b = OK.corrupt_with_draws(img, 3, name, d, bank)
        assert a.dtype == np.uint8 and a.shape == img.shape and np.array_equal(a, b) and not np.array_equal(a, img), name


def test_chamfer_table_reproduces_cv2_distance_transform():
    """spatter (row f2): the CUDA path evaluates cv2.distanceTransform(DIST_L2, 5), truncated at 20, as the
    minimum over zero pixels of a per-displacement cost table taken from cv2 (csrc/const_tables.inc).  Pin
    both the table and the min-over-sources model against cv2 itself."""
    import os
    import re
    import cv2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inc = open(os.path.join(root, "advmix_b200", "csrc", "const_tables.inc")).read()
    body = inc[inc.index("CHAMFER_L2_5["):]
    body = body[body.index("{") + 1:body.index("};")]
    tab = np.array([float.fromhex(t.rstrip("f")) for t in re.findall(r"[-0-9a-fx.p+]+f", body)], np.float32)
    R = 20
    assert tab.size == (2 * R + 1) ** 2
    T = tab.reshape(2 * R + 1, 2 * R + 1)
    S = 2 * R + 41
    one = np.full((S, S), 255, np.uint8)
    one[S // 2, S // 2] = 0
    ref = cv2.distanceTransform(one, cv2.DIST_L2, 5)[S // 2 - R:S // 2 + R + 1, S // 2 - R:S // 2 + R + 1]
    assert np.array_equal(T, ref)
    rng = np.random.default_rng(3)
    for (H, W, p) in ((64, 48, 0.01), (96, 130, 0.002), (50, 50, 0.05)):
        src = rng.random((H, W)) < p
        src[0, 0] = src[H - 1, W - 3] = True            # sources on the frame
        exp = np.minimum(cv2.distanceTransform(np.where(src, 0, 255).astype(np.uint8), cv2.DIST_L2, 5), 20)
        best = np.full((H, W), np.float32(20))
        P = np.pad(src, R)
        for dy in range(-R, R + 1):
            for dx in range(-R, R + 1):
                sh = P[R - dy:R - dy + H, R - dx:R - dx + W]
                best = np.where(sh, np.minimum(best, T[dy + R, dx + R]), best)
        assert np.array_equal(best, exp)


def test_inference_oracle_equals_reference_fixture(golden):
    """Row f3: oracle/inference.py against outputs of the real get_max_preds / get_final_preds / flip_back."""
    import hashlib
    from oracle import inference as OI
    g = golden("inference")
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    hm = g["heatmaps"]
    p, m = OI.get_max_preds(hm)
    assert np.array_equal(p, g["max_preds"]) and np.array_equal(m, g["maxvals"])
    for pp in (0, 1):
        preds, mv, _ = OI.get_final_preds(hm, g["center"], g["scale"], bool(pp))
        assert np.array_equal(preds, g["final_preds_pp%d" % pp]) and np.array_equal(mv, g["maxvals"])
    pairs = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]
    hf = (np.random.default_rng(77).standard_normal(hm.shape) * 0.3).astype(np.float32)
    assert sha(hf) == str(g["flipped_in_sha"])
    assert sha(OI.flip_back(hf, pairs)) == str(g["flip_back_sha"])
    assert sha(OI.flip_merge(hm, hf, pairs, False)) == str(g["merged_shift0_sha"])
    assert np.array_equal(OI.flip_merge(hm, hf, pairs, True), g["merged_shift1"])


def test_records_oracle_equals_reference_fixture(golden):
    """Row f4: oracle/records.py against the real JointsDataset.half_body_transform / select_data outputs."""
    from oracle import records as OR
    g = golden("records")
    B = len(g["randn"])
    for b in range(B):
        c, s = OR.half_body_transform(g["joints"][b], g["vis"][b], tuple(g["upper"]), g["randn"][b], float(g["aspect"]))
        if c is None:
            assert not g["hb_valid"][b]
        else:
            assert g["hb_valid"][b] and np.array_equal(c, g["hb_center"][b]) and np.array_equal(s, g["hb_scale"][b])
    db = [{"joints_3d": g["joints"][b], "joints_3d_vis": g["vis"][b], "center": g["center"][b], "scale": g["scale"][b]}
          for b in range(B)]
    assert np.array_equal(OR.select_data_mask(db), g["keep"])
