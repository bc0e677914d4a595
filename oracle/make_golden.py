"""Generate tests/golden/*.npz from the REAL reference (imported via oracle/ref_harness.py)
and the real third-party calls it makes (cv2, PIL, torchvision, torch).

TEST INFRASTRUCTURE - see oracle/__init__.py.  Run in the build container only:
    python -m oracle.make_golden
The fixtures travel to the GPU box; /root/reference does not.
"""
import hashlib
import os
import random

import numpy as np
import torch

from . import affine, chains, corruptions, ref_harness, targets

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def natural_image(rng, H, W):
    """'natural-like' synthetic image: 4-octave bilinear noise + 8-LSB iid noise."""
    import cv2
    acc = np.zeros((H, W, 3), np.float32)
    for o in range(4):
        s = 2 ** (o + 2)
        low = rng.random((H // s + 2, W // s + 2, 3)).astype(np.float32)
        acc += cv2.resize(low, (W, H), interpolation=cv2.INTER_LINEAR) / (o + 1)
    acc = acc / acc.max() * 255
    acc += rng.integers(-8, 9, acc.shape)
    return np.clip(acc, 0, 255).astype(np.uint8)


def gen_targets(ns):
    rng = np.random.default_rng(20261017)
    ds = ref_harness.make_dataset([], is_train=True)
    ds_w = ref_harness.make_dataset([], is_train=True, use_different_joints_weight=True)
    ds_mpii = ref_harness.make_dataset([], is_train=True, num_joints=16, image_size=(256, 256), heatmap_size=(64, 64))
    ds_big = ref_harness.make_dataset([], is_train=True, image_size=(512, 512), heatmap_size=(128, 128))
    out = {}
    for tag, d, J, (w, h), n in (("coco", ds, 17, (192, 256), 24), ("cocow", ds_w, 17, (192, 256), 8),
                                ("mpii", ds_mpii, 16, (256, 256), 8), ("big", ds_big, 17, (512, 512), 4)):
        joints = np.zeros((n, J, 3))
        vis = np.zeros((n, J, 3))
        joints[:, :, 0] = rng.uniform(-40, w + 40, (n, J))
        joints[:, :, 1] = rng.uniform(-40, h + 40, (n, J))
        joints[n // 2:, :, :2] = np.round(joints[n // 2:, :, :2] * 2) / 2      # exact .5 ties
        v = (rng.random((n, J)) < 0.8).astype(np.float64)
        vis[:, :, 0] = v
        vis[:, :, 1] = v
        hm, mu, tw = [], [], []
        for i in range(n):
            t, w_ = d.generate_target(joints[i].copy(), vis[i].copy())
            hm.append(t[0]); mu.append(t[1]); tw.append(w_)
        hm = np.stack(hm)
        preds, maxvals = ns.inference.get_max_preds(hm)
        out.update({tag + "_joints": joints, tag + "_vis": vis, tag + "_hm": hm, tag + "_mu": np.stack(mu),
                    tag + "_tw": np.stack(tw), tag + "_preds": preds, tag + "_maxvals": maxvals})
    np.savez_compressed(os.path.join(OUT, "targets.npz"), **out)


def gen_warp(ns):
    import cv2
    rng = np.random.default_rng(20261018)
    out = {}
    cases = [((96, 128), (48, 64)), ((120, 90), (48, 64)), ((64, 64), (64, 64)), ((150, 200), (192, 256)),
             ((97, 131), (50, 70)), ((80, 100), (48, 64))]
    for i, ((sh, sw), ds) in enumerate(cases):
        src = natural_image(rng, sh, sw) if i % 2 else rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        c = np.array([rng.uniform(0.2, 0.8) * sw, rng.uniform(0.2, 0.8) * sh], np.float32)
        s = np.array([rng.uniform(0.2, 0.9), rng.uniform(0.2, 0.9)], np.float32)
        r = float(rng.uniform(-80, 80)) if i % 3 else 0.0
        flip = bool(i % 2)
        view = src[:, ::-1, :] if flip else src
        trans = ns.transforms.get_affine_transform(c, s, r, ds)
        dst = cv2.warpAffine(view, trans, (int(ds[0]), int(ds[1])), flags=cv2.INTER_LINEAR)
        joints = np.zeros((17, 3)); joints[:, :2] = rng.uniform(0, [sw, sh], (17, 2))
        vis = np.repeat((rng.random((17, 1)) < 0.8).astype(np.float64), 3, 1); vis[:, 2] = 0
        j2, v2 = joints.copy(), vis.copy()
        if flip:
            j2, v2 = ns.transforms.fliplr_joints(j2, v2, sw, ref_harness.COCO_FLIP_PAIRS)
        for k in range(17):
            if v2[k, 0] > 0.0:
                j2[k, 0:2] = ns.transforms.affine_transform(j2[k, 0:2], trans)
        out.update({"src%d" % i: src, "center%d" % i: c, "scale%d" % i: s, "rot%d" % i: np.float64(r),
                    "flip%d" % i: np.uint8(flip), "dsize%d" % i: np.array(ds), "trans%d" % i: trans,
                    "dst%d" % i: dst, "joints%d" % i: joints, "vis%d" % i: vis, "joints_out%d" % i: j2,
                    "vis_out%d" % i: v2})
    out["n"] = np.int64(len(cases))
    # ToTensor + Normalize LUT from real torchvision
    from torchvision import transforms as T
    tf = T.Compose([T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    ramp = np.tile(np.arange(256, dtype=np.uint8)[:, None, None], (1, 1, 3))
    out["norm_lut"] = tf(ramp).numpy()[:, :, 0]          # [3,256]
    np.savez_compressed(os.path.join(OUT, "warp.npz"), **out)


def gen_chains(ns):
    """autoaug via the real ImageNetPolicy/SubPolicy (PIL) and gridmask via the real grid_aug."""
    from PIL import Image
    rng = np.random.default_rng(20261019)
    out = {}
    H, W = 64, 48
    pol = ns.advaug.ImageNetPolicy()
    imgs, res, plans = [], [], []
    for i in range(len(chains.POLICIES) * 2):
        img = natural_image(rng, H, W)
        pidx = i % len(chains.POLICIES)
        # replay SubPolicy.__call__ with recorded draws
        random.seed(1000 + i)
        sp = pol.policies[pidx]
        c1 = random.random()
        state = random.getstate()
        random.setstate(state)
        random.seed(1000 + i)
        out_img = np.array(sp(Image.fromarray(img)))
        # recover draws: re-run the same stream
        random.seed(1000 + i)
        p1, op1, m1, p2, op2, m2 = chains.POLICIES[pidx]
        coin1 = random.random()
        sign1 = random.choice([-1, 1]) if (coin1 < p1 and op1 == "sharpness") else 1
        coin2 = random.random()
        sign2 = random.choice([-1, 1]) if (coin2 < p2 and op2 == "sharpness") else 1
        chk = chains.autoaug(img, pidx, coin1, coin2, sign1, sign2)
        assert np.array_equal(chk, out_img), "oracle autoaug != real SubPolicy for policy %d" % pidx
        imgs.append(img); res.append(out_img); plans.append([pidx, coin1, coin2, sign1, sign2])
    out["aa_in"] = np.stack(imgs); out["aa_out"] = np.stack(res); out["aa_plan"] = np.array(plans, np.float64)
    # gridmask
    args = ref_harness.types.SimpleNamespace(joints_num=17)
    g_in, g_out, g_par, g_j, g_v, g_vo = [], [], [], [], [], []
    for i in range(8):
        img = torch.from_numpy(rng.standard_normal((3, H, W)).astype(np.float32))
        joints = np.zeros((17, 3)); joints[:, :2] = rng.uniform(-5, [W + 5, H + 5], (17, 2))
        vis = np.ones((17, 3)); vis[:, 2] = 0
        np.random.seed(500 + i)
        coin = np.random.rand()
        apply = not (coin > 0.7)
        d = st_h = st_w = 0
        if apply:
            d = np.random.randint(2, min(H, W)); st_h = np.random.randint(d); st_w = np.random.randint(d)
        np.random.seed(500 + i)
        o_img, _, o_vis, _ = ns.advaug.grid_aug(args, img.clone(), joints.copy(), vis.copy(), True, True, 1, False, 0.5, 1, 0.7, {})
        chk_img, chk_vis = chains.gridmask(img.numpy(), joints, vis, apply, d, st_h, st_w)
        assert np.array_equal(chk_img, o_img.numpy()) and np.array_equal(chk_vis, o_vis)
        g_in.append(img.numpy()); g_out.append(o_img.numpy()); g_par.append([int(apply), d, st_h, st_w])
        g_j.append(joints); g_v.append(vis); g_vo.append(o_vis)
    out.update({"gm_in": np.stack(g_in), "gm_out": np.stack(g_out), "gm_params": np.array(g_par, np.int32),
                "gm_joints": np.stack(g_j), "gm_vis": np.stack(g_v), "gm_vis_out": np.stack(g_vo)})
    np.savez_compressed(os.path.join(OUT, "chains.npz"), **out)


def gen_mix():
    """The mix expression of the real train_advmix loop is 4 lines of torch (function.py:138-144);
    run them verbatim on CPU."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(20261020)
    B, K, C, H, W = 2, 3, 3, 16, 12
    inputs = [torch.randn(B, C, H, W, generator=g) for _ in range(K)]
    logits = torch.randn(B, K, H, W, generator=g, requires_grad=True)
    mix_weight = F.softmax(logits, dim=1)
    tmp = inputs[0] * mix_weight[:, 0, ...].unsqueeze(dim=1)
    for list_index in range(1, len(inputs)):
        tmp += inputs[list_index] * mix_weight[:, list_index].unsqueeze(dim=1)
    go = torch.randn(B, C, H, W, generator=g)
    tmp.backward(go)
    np.savez_compressed(os.path.join(OUT, "mix.npz"), inputs=torch.stack(inputs).numpy(), logits=logits.detach().numpy(),
                        weights=mix_weight.detach().numpy(), tmp=tmp.detach().numpy(), grad_out=go.numpy(),
                        grad_logits=logits.grad.numpy())


def gen_getitem(ns):
    """Full reference __getitem__ (sample_times=3 and =1) on synthetic in-memory images."""
    import cv2
    rng = np.random.default_rng(20261021)
    H0, W0 = 120, 160
    imgs = [natural_image(rng, H0, W0) for _ in range(3)]
    db = []
    for i, im in enumerate(imgs):
        x, y, w, h = rng.uniform(10, 40), rng.uniform(5, 30), rng.uniform(40, 100), rng.uniform(50, 80)
        c, s = affine.xywh2cs(x, y, w, h)
        j = np.zeros((17, 3)); j[:, 0] = rng.uniform(x, x + w, 17); j[:, 1] = rng.uniform(y, y + h, 17)
        v = np.zeros((17, 3)); vv = (rng.random(17) < 0.85).astype(np.float64); v[:, 0] = vv; v[:, 1] = vv
        db.append({"image": "mem:%d" % i, "center": c, "scale": s, "joints_3d": j, "joints_3d_vis": v,
                   "filename": "", "imgnum": 0})
    real_imread = cv2.imread
    cv2.imread = lambda path, flags=None: imgs[int(path.split(":")[1])].copy()
    out = {"images": np.stack(imgs), "centers": np.stack([d["center"] for d in db]),
           "scales": np.stack([d["scale"] for d in db]), "joints": np.stack([d["joints_3d"] for d in db]),
           "vis": np.stack([d["joints_3d_vis"] for d in db])}
    try:
        ds = ref_harness.make_dataset(db, is_train=True, sample_times=3)
        for i in range(3):
            np.random.seed(700 + i); random.seed(700 + i)
            inputs, tgts, tws, metas = ds[i]
            m = metas[0]
            out.update({"k3_in%d" % i: np.stack([t.numpy() for t in inputs]),
                        "k3_hm%d" % i: np.stack([t.numpy() for t in tgts]),
                        "k3_tw%d" % i: np.stack([t.numpy() for t in tws]),
                        "k3_center%d" % i: np.asarray(m["center"]), "k3_scale%d" % i: np.asarray(m["scale"]),
                        "k3_rot%d" % i: np.float64(m["rotation"]), "k3_joints%d" % i: m["joints"],
                        "k3_vis%d" % i: m["joints_vis"]})
        ds1 = ref_harness.make_dataset(db, is_train=False, sample_times=1)
        for i in range(3):
            inp, tgt, tw, m = ds1[i]
            out.update({"k1_in%d" % i: inp.numpy(), "k1_hm%d" % i: tgt[0].numpy(), "k1_mu%d" % i: tgt[1].numpy(),
                        "k1_tw%d" % i: tw.numpy(), "k1_joints%d" % i: m["joints"]})
    finally:
        cv2.imread = real_imread
    np.savez_compressed(os.path.join(OUT, "getitem.npz"), **out)


def gen_inference(ns):
    """Row f3: real get_max_preds / get_final_preds (core/inference.py) and flip_back (utils/transforms.py),
    plus the flip-test merge expression of core/function.py:241-261 evaluated with torch like the reference."""
    from types import SimpleNamespace
    rng = np.random.default_rng(20261017)
    B, J, H, W = 3, 17, 64, 48
    hm = (rng.standard_normal((B, J, H, W)) * 0.02).astype(np.float32)
    g = targets.gaussian_table(2)
    for b in range(B):
        for j in range(J):
            cx, cy = int(rng.integers(-3, W + 3)), int(rng.integers(-3, H + 3))      # some peaks on / over the frame
            for dy in range(-6, 7):
                for dx in range(-6, 7):
                    y, x = cy + dy, cx + dx
                    if 0 <= y < H and 0 <= x < W:
                        hm[b, j, y, x] += g[dy + 6, dx + 6] * rng.uniform(0.3, 1.0)
    hm[0, 0] = -np.abs(hm[0, 0]) - 0.01            # nothing positive: coordinates are zeroed
    hm[0, 1] = 0.0                                 # all equal: first index wins, max not > 0
    hm[0, 2] = 0.0; hm[0, 2, 10, 7] = 0.5; hm[0, 2, 30, 20] = 0.5     # tie: first maximum
    hm[0, 3] = 0.0; hm[0, 3, 1, 1] = 1.0           # too close to the frame for the quarter-pixel step
    hm[0, 4] = 0.0; hm[0, 4, 2, 2] = 1.0; hm[0, 4, 2, 3] = 0.4; hm[0, 4, 3, 2] = 0.1; hm[0, 4, 1, 2] = 0.3
    center = np.stack([rng.uniform(50, 400, B), rng.uniform(50, 300, B)], 1).astype(np.float32)
    scale = np.stack([rng.uniform(0.5, 3.0, B), rng.uniform(0.6, 4.0, B)], 1).astype(np.float32)
    out = {"heatmaps": hm, "center": center, "scale": scale}
    p, m = ns.inference.get_max_preds(hm.copy())
    out["max_preds"], out["maxvals"] = p, m
    for pp in (0, 1):
        cfg = SimpleNamespace(TEST=SimpleNamespace(POST_PROCESS=bool(pp)), MODEL=SimpleNamespace(IMAGE_SIZE=[192, 256]))
        preds, mv = ns.inference.get_final_preds(cfg, None, hm.copy(), center, scale)
        out["final_preds_pp%d" % pp] = preds
        assert np.array_equal(mv, m)
    pairs = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]
    # the flipped-input maps are re-derived in the test from this seed (keeps the fixture small); outputs
    # other than the shifted merge are stored as SHA-256 of their float32 bytes
    hf = (np.random.default_rng(77).standard_normal((B, J, H, W)) * 0.3).astype(np.float32)
    out["flipped_in_sha"] = sha(hf)
    fb = ns.transforms.flip_back(hf.copy(), pairs)
    out["flip_back_sha"] = sha(fb)
    for shift in (0, 1):
        output = torch.from_numpy(hm.copy())
        output_flipped = torch.from_numpy(fb.copy())
        if shift:
            output_flipped[:, :, :, 1:] = output_flipped.clone()[:, :, :, 0:-1]
        merged = ((output + output_flipped) * 0.5).numpy()
        if shift:
            out["merged_shift1"] = merged
        else:
            out["merged_shift0_sha"] = sha(merged)
    np.savez_compressed(os.path.join(OUT, "inference.npz"), **out)


def gen_records(ns):
    """Row f4: the real JointsDataset.half_body_transform and select_data on synthetic records."""
    rng = np.random.default_rng(4242)
    B, J = 96, 17
    db = []
    for b in range(B):
        cx, cy = rng.uniform(80, 560), rng.uniform(80, 400)
        bw, bh = rng.uniform(40, 300), rng.uniform(60, 360)
        joints = np.zeros((J, 3))
        joints[:, 0] = rng.normal(cx, bw / 3, J); joints[:, 1] = rng.normal(cy, bh / 3, J)
        nvis = int(rng.integers(0, J + 1)) if b % 7 else int(rng.integers(0, 3))
        vis = np.zeros((J, 3)); vis[rng.permutation(J)[:nvis], :2] = 1
        # box centre sometimes far from the joints (select_data drops those)
        off = rng.uniform(-1, 1, 2) * (bw if b % 3 == 0 else 5)
        c, s = affine_cs(cx + off[0], cy + off[1], bw, bh)
        db.append({"joints_3d": joints, "joints_3d_vis": vis, "center": c, "scale": s})
    ds = ref_harness.make_dataset(db, is_train=True, sample_times=1)
    draws = np.zeros(B); hb_c = np.zeros((B, 2), np.float32); hb_s = np.zeros((B, 2), np.float32); hb_ok = np.zeros(B, bool)
    for b, rec in enumerate(db):
        np.random.seed(1000 + b)
        draws[b] = np.random.randn()
        np.random.seed(1000 + b)
        c, s = ds.half_body_transform(rec["joints_3d"], rec["joints_3d_vis"])
        if c is not None:
            hb_c[b], hb_s[b], hb_ok[b] = c, s, True
    kept = ds.select_data(db)
    keep = np.array([any(k is r for k in kept) for r in db])
    np.savez_compressed(os.path.join(OUT, "records.npz"), joints=np.stack([r["joints_3d"] for r in db]),
                        vis=np.stack([r["joints_3d_vis"] for r in db]), center=np.stack([r["center"] for r in db]),
                        scale=np.stack([r["scale"] for r in db]), randn=draws, hb_center=hb_c, hb_scale=hb_s, hb_valid=hb_ok,
                        keep=keep, upper=np.array(ref_harness.COCO_UPPER), aspect=np.float64(ds.aspect_ratio))


def gen_replay(ns):
    """Round-2 replay fixture: 64 samples of the REAL __getitem__ at sample_times=3 with PROB_HALF_BODY=0.3, and 32
    samples of the K=1 training path (get_clean) with COLOR_RGB=True.  Every output tensor is stored as SHA-256 of its
    bytes (the comparison is bit-exact) together with the meta the reference produced, from which a failing test can
    recompute the expected crop with cv2."""
    import cv2
    rng = np.random.default_rng(20261022)
    H0, W0 = 120, 160
    imgs = [natural_image(rng, H0, W0) for _ in range(8)]
    N3, N1 = 64, 32
    db = []
    for i in range(N3):
        x, y, w, h = rng.uniform(5, 50), rng.uniform(5, 30), rng.uniform(40, 100), rng.uniform(50, 85)
        c, s = affine.xywh2cs(x, y, w, h)
        j = np.zeros((17, 3)); j[:, 0] = rng.uniform(x, x + w, 17); j[:, 1] = rng.uniform(y, y + h, 17)
        v = np.zeros((17, 3)); vv = (rng.random(17) < (0.9 if i % 4 else 0.5)).astype(np.float64); v[:, 0] = vv; v[:, 1] = vv
        db.append({"image": "mem:%d" % (i % len(imgs)), "center": c, "scale": s, "joints_3d": j, "joints_3d_vis": v,
                   "filename": "", "imgnum": 0})
    real_imread = cv2.imread
    cv2.imread = lambda path, flags=None: imgs[int(path.split(":")[1])].copy()
    out = {"images": np.stack(imgs), "image_index": np.array([i % len(imgs) for i in range(N3)]),
           "centers": np.stack([d["center"] for d in db]), "scales": np.stack([d["scale"] for d in db]),
           "joints": np.stack([d["joints_3d"] for d in db]), "vis": np.stack([d["joints_3d_vis"] for d in db])}
    try:
        ds = ref_harness.make_dataset(db, is_train=True, sample_times=3, prob_half_body=0.3)
        k3 = {k: [] for k in ("center", "scale", "rot", "joints", "vis", "in_sha", "hm_sha", "tw_sha")}
        for i in range(N3):
            np.random.seed(9000 + i); random.seed(9000 + i)
            inputs, tgts, tws, metas = ds[i]
            m = metas[0]
            k3["center"].append(np.asarray(m["center"], np.float32)); k3["scale"].append(np.asarray(m["scale"], np.float64))
            k3["rot"].append(np.float64(m["rotation"])); k3["joints"].append(m["joints"]); k3["vis"].append(m["joints_vis"])
            k3["in_sha"].append([sha(t.numpy()) for t in inputs]); k3["hm_sha"].append([sha(t.numpy()) for t in tgts])
            k3["tw_sha"].append([sha(t.numpy()) for t in tws])
        out.update({"k3_" + k: np.array(v) for k, v in k3.items()})
        ds1 = ref_harness.make_dataset(db[:N1], is_train=True, sample_times=1, prob_half_body=0.3, color_rgb=True)
        k1 = {k: [] for k in ("center", "scale", "rot", "joints", "in_sha", "hm_sha", "mu_sha", "tw_sha")}
        for i in range(N1):
            np.random.seed(9500 + i); random.seed(9500 + i)
            inp, tgt, tw, m = ds1[i]
            k1["center"].append(np.asarray(m["center"], np.float32)); k1["scale"].append(np.asarray(m["scale"], np.float64))
            k1["rot"].append(np.float64(m["rotation"])); k1["joints"].append(m["joints"])
            k1["in_sha"].append(sha(inp.numpy())); k1["hm_sha"].append(sha(tgt[0].numpy())); k1["mu_sha"].append(sha(tgt[1].numpy()))
            k1["tw_sha"].append(sha(tw.numpy()))
        out.update({"k1_" + k: np.array(v) for k, v in k1.items()})
    finally:
        cv2.imread = real_imread
    np.savez_compressed(os.path.join(OUT, "replay.npz"), **out)


def gen_targets_hr(ns):
    """BASELINE configs[3]: generate_target at IMAGE_SIZE 512x512 with HEATMAP_SIZE 256x256 (feat_stride 2) next to the
    128x128 case of targets.npz (it is size-generic, JointsDataset.py:455-486).  Heat maps stored as SHA-256 + arg-max."""
    rng = np.random.default_rng(20261023)
    d = ref_harness.make_dataset([], is_train=True, image_size=(512, 512), heatmap_size=(256, 256))
    n, J = 4, 17
    joints = np.zeros((n, J, 3)); vis = np.zeros((n, J, 3))
    joints[:, :, 0] = rng.uniform(-30, 542, (n, J)); joints[:, :, 1] = rng.uniform(-30, 542, (n, J))
    joints[n // 2:, :, :2] = np.round(joints[n // 2:, :, :2])                    # exact .5 ties at stride 2
    v = (rng.random((n, J)) < 0.8).astype(np.float64); vis[:, :, 0] = v; vis[:, :, 1] = v
    hm, mu, tw = [], [], []
    for i in range(n):
        t, w_ = d.generate_target(joints[i].copy(), vis[i].copy())
        hm.append(t[0]); mu.append(t[1]); tw.append(w_)
    hm = np.stack(hm)
    preds, maxvals = ns.inference.get_max_preds(hm)
    np.savez_compressed(os.path.join(OUT, "targets_hr.npz"), joints=joints, vis=vis, hm_sha=np.array([sha(h) for h in hm]),
                        hm_sum=hm.reshape(n, J, -1).sum(-1), mu=np.stack(mu), tw=np.stack(tw), preds=preds, maxvals=maxvals)


def affine_cs(cx, cy, w, h, aspect=0.75, pixel_std=200):
    from . import records
    return records.xywh2cs(cx - w * 0.5, cy - h * 0.5, w, h, aspect, pixel_std)


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_harness.load()
    gen_targets(ns)
    gen_warp(ns)
    gen_chains(ns)
    gen_mix()
    gen_getitem(ns)
    gen_inference(ns)
    gen_records(ns)
    gen_replay(ns)
    gen_targets_hr(ns)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
