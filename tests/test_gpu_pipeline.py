"""GPU parity of the batched JointsDataset mirror against fixtures produced by the REAL
reference __getitem__ (tests/golden/getitem.npz, generator oracle/make_golden.py)."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def records(g):
    return [{"image": g["images"][i], "center": g["centers"][i], "scale": g["scales"][i],
             "joints_3d": g["joints"][i], "joints_3d_vis": g["vis"][i]} for i in range(len(g["images"]))]


def close_images(got, exp, what):
    """Normalised tensors come from a 256-entry LUT, so they are equal wherever the uint8 crop is.  The crop is
    bit-exact given the matrix, and the device-side get_affine_transform restates cv2.getAffineTransform's LU
    operation by operation, so the whole replay is bit-exact."""
    neq = (got != exp)
    assert not neq.any(), "%s: %.4f%% of values differ, max |d| = %.3f u8 LSB" % (
        what, 100 * neq.mean(), np.abs(got - exp).max() * 255 * 0.224)


def _sha(t):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()).hexdigest()


def replay_records(g, n):
    return [{"image": g["images"][int(g["image_index"][i])], "center": g["centers"][i], "scale": g["scales"][i],
             "joints_3d": g["joints"][i], "joints_3d_vis": g["vis"][i]} for i in range(n)]


def _expected_clean(g, i, prefix, rgb):
    """The reference's crop of sample i recomputed with cv2 from the meta it produced (diagnostics for a hash mismatch)."""
    from oracle import affine as OA
    img = g["images"][int(g["image_index"][i])]
    if rgb:
        img = img[:, :, ::-1]
    c, s, r = g[prefix + "center"][i], g[prefix + "scale"][i], float(g[prefix + "rot"][i])
    flipped = c[0] != g["centers"][i][0] and abs((img.shape[1] - g["centers"][i][0] - 1) - c[0]) < 1e-3
    # (half-body samples change the centre too; the flag is only used for the diagnostic image)
    M = OA.get_affine_transform(c, s, r, (192, 256))
    return OA.to_tensor_normalize(OA.warp_affine_cv2(img[:, ::-1, :] if flipped else img, M, (192, 256)))


def test_replay_k3_64_samples_half_body_bit_exact(built_library, golden):
    """64 samples of the REAL __getitem__ (sample_times=3, PROB_HALF_BODY=0.3) replayed from the same np.random / random
    seeds: draws, half-body boxes, matrices, crops, all three chains, heat maps and weights bit-exact (SHA-256)."""
    from advmix_b200.dataset import AdvMixBatchPipeline
    g = golden("replay")
    N = len(g["k3_rot"])
    recs = replay_records(g, N)
    pipe = AdvMixBatchPipeline(sample_times=3, is_train=True, draw_mode="reference", prob_half_body=0.3)
    n_hb = 0
    for i in range(N):
        np.random.seed(9000 + i); random.seed(9000 + i)
        inputs, tgts, tws, metas = pipe([recs[i]])
        m = metas[0]
        assert np.array_equal(m["center"][0].cpu().numpy(), g["k3_center"][i]), i
        assert np.array_equal(m["scale"][0].cpu().numpy(), g["k3_scale"][i]), i
        assert float(m["rotation"][0]) == float(g["k3_rot"][i]), i
        n_hb += int(not np.allclose(g["k3_scale"][i] / g["scales"][i], (g["k3_scale"][i] / g["scales"][i])[0]) or
                    abs(g["k3_center"][i][1] - g["centers"][i][1]) > 1e-6)
        np.testing.assert_allclose(m["joints"][0].cpu().numpy(), g["k3_joints"][i], rtol=0, atol=1e-8)
        assert np.array_equal(m["joints_vis"][0].cpu().numpy(), g["k3_vis"][i]), i
        for k in range(3):
            if _sha(inputs[k][0]) != g["k3_in_sha"][i][k] and k == 0:
                close_images(inputs[0][0].cpu().numpy(), _expected_clean(g, i, "k3_", False), "sample %d clean chain" % i)
            assert _sha(inputs[k][0]) == g["k3_in_sha"][i][k], "sample %d chain %d input" % (i, k)
            assert _sha(tgts[k][0]) == g["k3_hm_sha"][i][k], "sample %d chain %d heat map" % (i, k)
            assert _sha(tws[k][0]) == g["k3_tw_sha"][i][k], "sample %d chain %d target_weight" % (i, k)
    assert n_hb >= 5, "the fixture should exercise the half-body branch (%d)" % n_hb


def test_replay_k1_train_color_rgb_bit_exact(built_library, golden):
    """32 samples of the K=1 training path (get_clean) with COLOR_RGB=True and PROB_HALF_BODY=0.3."""
    from advmix_b200.dataset import AdvMixBatchPipeline
    g = golden("replay")
    N = len(g["k1_rot"])
    recs = replay_records(g, N)
    pipe = AdvMixBatchPipeline(sample_times=1, is_train=True, draw_mode="reference", prob_half_body=0.3, color_rgb=True)
    for i in range(N):
        np.random.seed(9500 + i); random.seed(9500 + i)
        inp, target, tw, m = pipe([recs[i]])
        assert np.array_equal(m["center"][0].cpu().numpy(), g["k1_center"][i]), i
        assert np.array_equal(m["scale"][0].cpu().numpy(), g["k1_scale"][i]), i
        assert float(m["rotation"][0]) == float(g["k1_rot"][i]), i
        if _sha(inp[0]) != g["k1_in_sha"][i]:
            close_images(inp[0].cpu().numpy(), _expected_clean(g, i, "k1_", True), "sample %d" % i)
        assert _sha(inp[0]) == g["k1_in_sha"][i], i
        assert _sha(target[0][0]) == g["k1_hm_sha"][i] and _sha(target[1][0]) == g["k1_mu_sha"][i], i
        assert _sha(tw[0]) == g["k1_tw_sha"][i], i


def test_random_corruption_branch(built_library, golden):
    """get_clean's --random_corruption (JointsDataset.py:284-286): (name, severity) drawn from `random` before the other
    draws, corrupt() applied to the FULL source image before the crop."""
    import advmix_b200 as A
    from advmix_b200 import transforms as TF
    from advmix_b200.dataset import AdvMixBatchPipeline, RANDOM_CORRUPTIONS
    g = golden("replay")
    recs = replay_records(g, 6)
    A.corruptions.set_frost_bank(A.corruptions.default_frost_bank(192, 192))
    pipe = AdvMixBatchPipeline(sample_times=1, is_train=False, random_corruption=True, draw_mode="reference", seed=3)
    random.seed(11)
    inp, target, tw, meta = pipe(recs)
    random.seed(11)
    exp = [(random.choice(RANDOM_CORRUPTIONS), random.randint(1, 5)) for _ in recs]
    assert meta["random_corruption"] == exp
    plain = AdvMixBatchPipeline(sample_times=1, is_train=False)
    for b, (name, sev) in enumerate(exp):
        src = torch.from_numpy(np.ascontiguousarray(recs[b]["image"])).cuda()[None]
        cor = A.corrupt_batch(src, name, sev, seed=(3 << 20) ^ 0, sample_base=meta["random_corruption"].index((name, sev)))
        rec = dict(recs[b]); rec["image"] = cor[0].cpu().numpy()
        e_inp, e_t, e_tw, _ = plain([rec])
        same_group_first = [i for i, e in enumerate(exp) if e == (name, sev)][0]
        if same_group_first == b:           # draws are keyed by the first member of the (name, severity) group
            assert torch.equal(inp[b], e_inp[0]), (b, name, sev)
        assert torch.equal(target[0][b], e_t[0][0]) and torch.equal(tw[b], e_tw[0])
    clean_inp, _, _, _ = plain(recs)
    assert not torch.equal(inp, clean_inp)


def test_getitem_k3_reference_draw_replay(built_library, golden):
    from advmix_b200.dataset import AdvMixBatchPipeline
    g = golden("getitem")
    recs = records(g)
    pipe = AdvMixBatchPipeline(sample_times=3, is_train=True, draw_mode="reference")
    for i in range(3):
        np.random.seed(700 + i); random.seed(700 + i)
        inputs, tgts, tws, metas = pipe([recs[i]])
        assert len(inputs) == len(tgts) == len(tws) == len(metas) == 3
        m = metas[0]
        assert np.array_equal(m["center"][0].cpu().numpy(), g["k3_center%d" % i])
        np.testing.assert_array_equal(m["scale"][0].cpu().numpy(), g["k3_scale%d" % i])
        assert float(m["rotation"][0]) == float(g["k3_rot%d" % i])
        np.testing.assert_allclose(m["joints"][0].cpu().numpy(), g["k3_joints%d" % i], rtol=0, atol=1e-8)
        assert np.array_equal(m["joints_vis"][0].cpu().numpy(), g["k3_vis%d" % i])
        for k in range(3):
            assert inputs[k].shape == (1, 3, 256, 192) and inputs[k].dtype == torch.float32
            close_images(inputs[k][0].cpu().numpy(), g["k3_in%d" % i][k], "sample %d chain %d" % (i, k))
            assert np.array_equal(tgts[k][0].cpu().numpy(), g["k3_hm%d" % i][k]), (i, k)
            assert np.array_equal(tws[k][0].cpu().numpy(), g["k3_tw%d" % i][k]), (i, k)


def test_getitem_k1_eval_path(built_library, golden):
    from advmix_b200.dataset import AdvMixBatchPipeline
    g = golden("getitem")
    recs = records(g)
    pipe = AdvMixBatchPipeline(sample_times=1, is_train=False)
    inp, target, tw, meta = pipe(recs)                      # whole batch at once
    assert isinstance(target, list) and len(target) == 2
    for i in range(3):
        close_images(inp[i].cpu().numpy(), g["k1_in%d" % i], "eval sample %d" % i)
        assert np.array_equal(target[0][i].cpu().numpy(), g["k1_hm%d" % i])
        assert np.array_equal(target[1][i].cpu().numpy(), g["k1_mu%d" % i])
        assert np.array_equal(tw[i].cpu().numpy(), g["k1_tw%d" % i])
        np.testing.assert_allclose(meta["joints"][i].cpu().numpy(), g["k1_joints%d" % i], rtol=0, atol=1e-8)


def test_batched_mode_and_corruption_chains(built_library, golden):
    """Vectorised draws + the BASELINE config-3 target workload (chains drawn from the 15x5 set)."""
    import advmix_b200 as A
    from advmix_b200.dataset import AdvMixBatchPipeline, corruption_chains
    g = golden("getitem")
    recs = records(g) * 6
    pipe = AdvMixBatchPipeline(sample_times=3, is_train=True, draw_mode="batched", seed=3, norm_dtype=torch.bfloat16)
    random.seed(0); np.random.seed(0)
    inputs, tgts, tws, metas = pipe(recs)
    B = len(recs)
    assert inputs[0].shape == (B, 3, 256, 192) and inputs[0].dtype == torch.bfloat16
    assert tgts[2].shape == (B, 17, 64, 48) and tws[0].shape == (B, 17, 1)
    assert torch.all(metas[0]["joints_vis"][:, :, 0] <= torch.from_numpy(np.stack([r["joints_3d_vis"] for r in recs]))[:, :, 0].max().item())
    # corruption chains on a uint8 crop batch
    rng = np.random.default_rng(0)
    crop = torch.from_numpy(rng.integers(0, 256, (B, 256, 192, 3), dtype=np.uint8)).cuda()
    names = [A.get_corruption_names()[i % 15] for i in range(B)]
    sevs = [1 + i % 5 for i in range(B)]
    u8, nrm = corruption_chains(crop, names, sevs, seed=11)
    for b in (0, 7, 13):
        single = A.corrupt_batch(crop, names[b], sevs[b], seed=11, idx=torch.tensor([b], dtype=torch.int32, device="cuda"))
        assert torch.equal(single[b], u8[b])
    logits = torch.randn(B, 3, 256, 192, device="cuda")
    mixed = A.mix_from_logits([A.to_tensor_normalize(crop), nrm, nrm], logits)
    assert mixed.shape == (B, 3, 256, 192) and torch.isfinite(mixed).all()


def test_box_cropped_h2d_equals_full_upload(built_library):
    """advmix_h2d_source_boxes gathers only the boxes the crops read (rotations, flips); the result must not
    depend on the rest of the device buffer."""
    from advmix_b200 import transforms as TF
    from advmix_b200.dataset import AdvMixBatchPipeline
    rng = np.random.default_rng(12)
    B, H, W = 24, 240, 320
    imgs = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    recs = []
    for b in range(B):
        c = np.array([rng.uniform(0.2, 0.8) * W, rng.uniform(0.2, 0.8) * H], np.float32)
        sc = np.array([rng.uniform(0.1, 0.4), rng.uniform(0.15, 0.5)], np.float32)   # small boxes: few rows touched
        j = np.zeros((17, 3)); j[:, :2] = rng.uniform(0, [W, H], (17, 2))
        recs.append({"center": c, "scale": sc, "joints_3d": j, "joints_3d_vis": np.ones((17, 3)), "width": W, "height": H})
    c = np.stack([r["center"] for r in recs]); s = np.stack([r["scale"] for r in recs]).astype(np.float64)
    rot = rng.uniform(-80, 80, B); flip = rng.random(B) < 0.5
    draws = (c, s, rot, flip)
    pipe = AdvMixBatchPipeline(sample_times=1, is_train=True)
    full = TF.SourceBatch.from_tensor(torch.from_numpy(imgs).cuda())
    ref_inp, ref_t, ref_tw, _ = pipe(recs, sources=full, draws=draws)
    host = torch.from_numpy(imgs).pin_memory()
    hsb = TF.HostSourceBatch.from_tensor(host)
    hsb.dev.buffer.fill_(173)                                  # stale garbage everywhere that is not uploaded
    inp, t, tw, _ = pipe(recs, draws=draws, host_sources=hsb)
    assert 0 < pipe.last_h2d_bytes < imgs.size
    assert torch.equal(inp, ref_inp) and torch.equal(t[0], ref_t[0]) and torch.equal(tw, ref_tw)


def test_source_gather_at_bench_scale(built_library):
    """Property at the bench's size and draw distribution (256 samples, 640x480, scale/rotation/flip draws of
    SURVEY 8d config 2): crops from the zero-copy gather (only the bytes the crops read) equal crops from fully
    uploaded sources, and the gather moves well under the full buffer."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from advmix_b200 import transforms as TF
    from advmix_b200.dataset import AdvMixBatchPipeline
    B = 256
    rng = np.random.default_rng(77)
    recs = bench.synth_records(B, rng)
    draws = bench.synth_draws(recs, rng)
    for r in recs:
        r["width"], r["height"] = bench.SRC_W, bench.SRC_H
    imgs = torch.randint(0, 256, (B, bench.SRC_H, bench.SRC_W, 3), dtype=torch.uint8)
    pipe = AdvMixBatchPipeline(sample_times=1, is_train=True)
    ref_inp, ref_t, ref_tw, _ = pipe(recs, sources=TF.SourceBatch.from_tensor(imgs.cuda()), draws=draws)
    hsb = TF.HostSourceBatch.from_tensor(imgs.pin_memory())
    hsb.dev.buffer.fill_(91)
    inp, t, tw, _ = pipe(recs, draws=draws, host_sources=hsb)
    torch.cuda.synchronize()
    assert torch.equal(inp, ref_inp) and torch.equal(t[0], ref_t[0]) and torch.equal(tw, ref_tw)
    sent = int(hsb.bytes_sent.item())
    assert 0 < sent < 0.6 * imgs.numel(), sent
