"""GPU parity: affine crop (a1), normalise (a7), heatmap targets (a4), mix (a3) - through the
Python host mirror -> C ABI -> CUDA, against the oracle and the golden fixtures produced by the
real reference."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import affine as OA          # noqa: E402
from oracle import mix as OM             # noqa: E402
from oracle import targets as OT         # noqa: E402

COCO_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]


def dev():
    return torch.device("cuda:0")


def test_warp_golden_bit_exact(built_library, golden):
    import advmix_b200 as A
    from advmix_b200 import transforms as T
    g = golden("warp")
    for i in range(int(g["n"])):
        src = g["src%d" % i]
        ds = tuple(int(v) for v in g["dsize%d" % i])
        sb = A.SourceBatch.from_numpy([src])
        trans = torch.from_numpy(g["trans%d" % i])[None].to(dev())
        flip = torch.tensor([int(g["flip%d" % i])], dtype=torch.uint8, device=dev())
        u8, nrm = A.warp_affine(sb, trans, ds, flip=flip, norm_dtype=torch.float32)
        assert np.array_equal(u8[0].cpu().numpy(), g["dst%d" % i]), "case %d" % i
        exp_n = OA.to_tensor_normalize(g["dst%d" % i], g["norm_lut"])
        assert np.array_equal(nrm[0].cpu().numpy(), exp_n)
        # joints
        perm = T.flip_perm(17, COCO_PAIRS)
        jo, vo = A.fliplr_affine_joints(torch.from_numpy(g["joints%d" % i])[None].to(dev()),
                                        torch.from_numpy(g["vis%d" % i])[None].to(dev()), trans, flip=flip,
                                        widths=torch.tensor([src.shape[1]], dtype=torch.int32, device=dev()), perm=perm)
        np.testing.assert_allclose(jo[0].cpu().numpy(), g["joints_out%d" % i], rtol=0, atol=1e-9)
        assert np.array_equal(vo[0].cpu().numpy(), g["vis_out%d" % i])
        # device-side get_affine_transform agrees with cv2.getAffineTransform to ~1e-9
        M = A.get_affine_transform(torch.from_numpy(g["center%d" % i])[None].to(dev()),
                                   torch.from_numpy(g["scale%d" % i])[None].to(dev()),
                                   torch.tensor([float(g["rot%d" % i])], dtype=torch.float64, device=dev()), ds)
        np.testing.assert_allclose(M[0].cpu().numpy(), g["trans%d" % i], rtol=1e-9, atol=1e-7)


@pytest.mark.parametrize("dsize", [(192, 256), (256, 256), (512, 512), (50, 37)])
def test_warp_random_batch_vs_cv2(built_library, dsize):
    """Ragged source sizes, heavy borders, flips: bit-exact against cv2.warpAffine itself."""
    import advmix_b200 as A
    rng = np.random.default_rng(hash(dsize) % 1000)
    srcs, Ms, flips, exp = [], [], [], []
    for i in range(12):
        H, W = int(rng.integers(40, 400)), int(rng.integers(40, 400))
        src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        c = np.array([rng.uniform(-0.2, 1.2) * W, rng.uniform(-0.2, 1.2) * H], np.float32)
        s = np.array([rng.uniform(0.2, 2.5), rng.uniform(0.2, 2.5)], np.float32)
        r = float(rng.uniform(-80, 80)) if rng.random() < 0.6 else 0.0
        f = bool(rng.random() < 0.5)
        M = OA.get_affine_transform(c, s, r, dsize)
        srcs.append(src); Ms.append(M); flips.append(f)
        exp.append(OA.warp_affine_cv2(src[:, ::-1, :] if f else src, M, dsize))
    sb = A.SourceBatch.from_numpy(srcs)
    u8, nb = A.warp_affine(sb, torch.from_numpy(np.stack(Ms)).to(dev()), dsize,
                           flip=torch.tensor(flips, dtype=torch.uint8, device=dev()), norm_dtype=torch.bfloat16)
    got = u8.cpu().numpy()
    for i in range(12):
        assert np.array_equal(got[i], exp[i]), "sample %d" % i
    lut = OA.normalize_lut()
    ref_b = torch.from_numpy(OA.to_tensor_normalize(exp[0], lut)).to(torch.bfloat16)
    assert torch.equal(nb[0].cpu(), ref_b)


@pytest.mark.parametrize("scale_dtype", [np.float32, np.float64])
def test_csr_fused_equals_two_step(built_library, scale_dtype):
    """advmix_crop_csr_u8c3 / advmix_joints_csr (matrix evaluated in-kernel) are bit-identical to
    advmix_affine_matrices followed by advmix_warp_affine_u8c3 / advmix_joints_flip_affine."""
    import advmix_b200 as A
    rng = np.random.default_rng(5)
    B, J, ds = 24, 17, (192, 256)
    srcs = [rng.integers(0, 256, (int(rng.integers(60, 300)), int(rng.integers(60, 300)), 3), dtype=np.uint8) for _ in range(B)]
    sb = A.SourceBatch.from_numpy(srcs)
    c = torch.from_numpy(rng.uniform(20, 250, (B, 2)).astype(np.float32)).to(dev())
    s = torch.from_numpy(rng.uniform(0.3, 2.0, (B, 2)).astype(scale_dtype)).to(dev())
    r = torch.from_numpy(np.where(rng.random(B) < 0.6, rng.uniform(-80, 80, B), 0.0)).to(dev())
    f = torch.from_numpy((rng.random(B) < 0.5).astype(np.uint8)).to(dev())
    jt = torch.from_numpy(np.concatenate([rng.uniform(0, 300, (B, J, 2)), np.zeros((B, J, 1))], -1)).to(dev())
    v = torch.from_numpy(np.repeat((rng.random((B, J, 1)) < 0.8).astype(np.float64), 3, -1)).to(dev())
    perm = A.transforms.flip_perm(J, [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]], dev())
    M = A.get_affine_transform(c, s, r, ds)
    u8a, na = A.warp_affine(sb, M, ds, flip=f, norm_dtype=torch.float32)
    ja, va = A.fliplr_affine_joints(jt, v, M, flip=f, widths=sb.widths, perm=perm)
    u8b, nb = A.crop_csr(sb, c, s, r, ds, flip=f, norm_dtype=torch.float32)
    jb, vb, Mb = A.joints_csr(jt, v, c, s, r, ds, flip=f, widths=sb.widths, perm=perm, want_trans=True)
    assert torch.equal(M, Mb)
    assert torch.equal(u8a, u8b) and torch.equal(na, nb)
    assert torch.equal(ja, jb) and torch.equal(va, vb)
    jc, vc = A.joints_csr(jt, v, c, s, r, ds, flip=f, widths=sb.widths, perm=perm)
    assert torch.equal(ja, jc) and torch.equal(va, vc)


def test_warp_unaligned_sources_take_the_direct_path(built_library):
    """Sources whose rows are not 16-byte aligned (pitch = 3*W, odd offsets) cannot be staged with 16-byte
    copies; the kernel samples them from global memory with the same arithmetic: still bit-exact vs cv2."""
    import advmix_b200 as A
    rng = np.random.default_rng(21)
    srcs, Ms, flips, exp, offs, total = [], [], [], [], [], 3
    for i in range(6):
        H, W = int(rng.integers(40, 200)), int(rng.integers(41, 200)) | 1          # odd widths: 3*W % 16 != 0
        src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        c = np.array([rng.uniform(0, 1) * W, rng.uniform(0, 1) * H], np.float32)
        s = np.array([rng.uniform(0.2, 1.5), rng.uniform(0.2, 1.5)], np.float32)
        M = OA.get_affine_transform(c, s, float(rng.uniform(-60, 60)), (192, 256))
        f = bool(i % 2)
        srcs.append(src); Ms.append(M); flips.append(f); offs.append(total)
        total += src.size + 5
        exp.append(OA.warp_affine_cv2(src[:, ::-1, :] if f else src, M, (192, 256)))
    buf = np.zeros(total, np.uint8)
    for src, o in zip(srcs, offs):
        buf[o:o + src.size] = src.reshape(-1)
    d = dev()
    sb = A.SourceBatch(torch.from_numpy(buf).to(d), torch.tensor(offs, dtype=torch.int64, device=d),
                       torch.tensor([x.shape[0] for x in srcs], dtype=torch.int32, device=d),
                       torch.tensor([x.shape[1] for x in srcs], dtype=torch.int32, device=d),
                       torch.tensor([x.shape[1] * 3 for x in srcs], dtype=torch.int64, device=d))
    u8, _ = A.warp_affine(sb, torch.from_numpy(np.stack(Ms)).to(d), (192, 256), flip=torch.tensor(flips, dtype=torch.uint8, device=d))
    for i in range(6):
        assert np.array_equal(u8[i].cpu().numpy(), exp[i]), i


def test_warp_empty_and_errors(built_library):
    import advmix_b200 as A
    sb = A.SourceBatch.from_tensor(torch.zeros((0, 8, 8, 3), dtype=torch.uint8, device=dev()))
    u8, _ = A.warp_affine(sb, torch.zeros((0, 2, 3), dtype=torch.float64, device=dev()), (16, 16))
    assert u8.shape == (0, 16, 16, 3)
    sb = A.SourceBatch.from_tensor(torch.zeros((1, 8, 8, 3), dtype=torch.uint8, device=dev()))
    with pytest.raises(A.AdvmixError):
        A.warp_affine(sb, torch.zeros((1, 2, 3), dtype=torch.float64, device=dev()), (16, 16), want_u8=False)


def test_normalize_matches_torchvision_lut(built_library, golden):
    import advmix_b200 as A
    g = golden("warp")
    rng = np.random.default_rng(5)
    for shape in [(3, 256, 192, 3), (2, 31, 33, 3)]:
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        out = A.to_tensor_normalize(torch.from_numpy(img).to(dev()))
        for b in range(shape[0]):
            assert np.array_equal(out[b].cpu().numpy(), OA.to_tensor_normalize(img[b], g["norm_lut"]))


@pytest.mark.parametrize("tag,img,hm,sigma,jw", [("coco", (192, 256), (48, 64), 2, False), ("cocow", (192, 256), (48, 64), 2, True),
                                                 ("mpii", (256, 256), (64, 64), 2, False), ("big", (512, 512), (128, 128), 2, False)])
def test_heatmap_golden(built_library, golden, tag, img, hm, sigma, jw):
    import advmix_b200 as A
    g = golden("targets")
    jwv = np.array([1., 1., 1., 1., 1., 1., 1., 1.2, 1.2, 1.5, 1.5, 1., 1., 1.2, 1.2, 1.5, 1.5], np.float32) if jw else None
    (h, mu), tw = A.generate_target(torch.from_numpy(g[tag + "_joints"]).to(dev()), torch.from_numpy(g[tag + "_vis"]).to(dev()),
                                    image_size=img, heatmap_size=hm, sigma=sigma, joints_weight=jwv)
    hn = h.cpu().numpy()
    assert np.abs(hn - g[tag + "_hm"]).max() == 0.0            # L_inf bar is 1e-5; we get 0
    assert np.array_equal(mu.cpu().numpy(), g[tag + "_mu"])
    assert np.array_equal(tw.cpu().numpy(), g[tag + "_tw"])
    preds, maxvals = OT.get_max_preds(hn)
    assert np.array_equal(preds, g[tag + "_preds"]) and np.array_equal(maxvals, g[tag + "_maxvals"])


def test_heatmap_256x256_configs3(built_library, golden):
    """BASELINE configs[3]: IMAGE_SIZE 512x512 with HEATMAP_SIZE 256x256 (stride 2) next to 128x128 above; fixture from
    the real generate_target (oracle/make_golden.py:gen_targets_hr), heat maps compared by SHA-256 and arg-max."""
    import hashlib
    import advmix_b200 as A
    g = golden("targets_hr")
    (h, mu), tw = A.generate_target(torch.from_numpy(g["joints"]).to(dev()), torch.from_numpy(g["vis"]).to(dev()),
                                    image_size=(512, 512), heatmap_size=(256, 256), sigma=2)
    hn = h.cpu().numpy()
    assert hn.shape == (4, 17, 256, 256)
    for i in range(4):
        assert hashlib.sha256(np.ascontiguousarray(hn[i]).tobytes()).hexdigest() == str(g["hm_sha"][i]), i
    assert np.array_equal(mu.cpu().numpy(), g["mu"]) and np.array_equal(tw.cpu().numpy(), g["tw"])
    preds, maxvals = OT.get_max_preds(hn)
    assert np.array_equal(preds, g["preds"]) and np.array_equal(maxvals, g["maxvals"])
    # and against the oracle restatement for a batch of 32 (row-chunked launch)
    rng = np.random.default_rng(5)
    joints = np.zeros((32, 17, 3)); joints[:, :, :2] = rng.uniform(-20, 530, (32, 17, 2))
    vis = np.zeros((32, 17, 3)); vis[:, :, :2] = (rng.random((32, 17, 1)) < 0.8)
    for hs in ((128, 128), (256, 256)):
        (h, mu), tw = A.generate_target(torch.from_numpy(joints).to(dev()), torch.from_numpy(vis).to(dev()),
                                        image_size=(512, 512), heatmap_size=hs, sigma=2)
        for b in (0, 13, 31):
            t, w = OT.generate_target(joints[b], vis[b], image_size=(512, 512), heatmap_size=hs, sigma=2)
            assert np.array_equal(h[b].cpu().numpy(), t[0]) and np.array_equal(mu[b].cpu().numpy(), t[1])
            assert np.array_equal(tw[b].cpu().numpy(), w)


def test_heatmap_wide_map(built_library):
    """Heat maps wider than 4 * 256 columns take the scalar store path (the vectorised one would cover no row)."""
    import advmix_b200 as A
    joints = np.zeros((2, 3, 3)); joints[:, :, 0] = [[10.0, 2000.0, 4090.0]] * 2; joints[:, :, 1] = 9.0
    vis = np.ones((2, 3, 3))
    (h, mu), tw = A.generate_target(torch.from_numpy(joints).to(dev()), torch.from_numpy(vis).to(dev()),
                                    image_size=(4100, 16), heatmap_size=(2052, 8), sigma=1)
    for b in range(2):
        t, w = OT.generate_target(joints[b], vis[b], image_size=(4100, 16), heatmap_size=(2052, 8), sigma=1)
        assert np.array_equal(h[b].cpu().numpy(), t[0]) and np.array_equal(tw[b].cpu().numpy(), w)


def test_heatmap_random_vs_oracle_edge_cases(built_library):
    import advmix_b200 as A
    rng = np.random.default_rng(11)
    B, J = 37, 17
    joints = np.zeros((B, J, 3)); vis = np.zeros((B, J, 3))
    joints[:, :, :2] = rng.uniform(-60, 320, (B, J, 2))
    joints[0, :, :2] = [[-28.0, 10.0]] * J            # window fully left of the map
    joints[1, :, :2] = [[-26.0, 10.0]] * J            # br == 0 quirk: weight kept, nothing pasted
    joints[2, :, :2] = [[191.9, 255.9]] * J
    joints[3, :, :2] = [[-1.9, -1.9]] * J             # int() truncation toward zero
    joints[4, :, :2] = [[1e7, -1e7]] * J
    v = (rng.random((B, J)) < 0.7).astype(np.float64); v[:5] = 1
    vis[:, :, 0] = v; vis[:, :, 1] = v
    vis[5, :, 0] = 0.5                                  # v > 0.5 is strict
    (h, mu), tw = A.generate_target(torch.from_numpy(joints).to(dev()), torch.from_numpy(vis).to(dev()))
    for b in range(B):
        t, w = OT.generate_target(joints[b], vis[b])
        assert np.array_equal(h[b].cpu().numpy(), t[0]), b
        assert np.array_equal(mu[b].cpu().numpy(), t[1]), b
        assert np.array_equal(tw[b].cpu().numpy(), w), b
    (h0, _), tw0 = A.generate_target(torch.zeros((0, 17, 3), dtype=torch.float64, device=dev()),
                                     torch.zeros((0, 17, 3), dtype=torch.float64, device=dev()))
    assert h0.shape == (0, 17, 64, 48)
    # odd heatmap width exercises the scalar store path
    (h1, _), _ = A.generate_target(torch.from_numpy(joints[:4]).to(dev()), torch.from_numpy(vis[:4]).to(dev()),
                                   image_size=(188, 252), heatmap_size=(47, 63), sigma=3)
    for b in range(4):
        t, _ = OT.generate_target(joints[b], vis[b], image_size=(188, 252), heatmap_size=(47, 63), sigma=3)
        assert np.array_equal(h1[b].cpu().numpy(), t[0])


def test_heatmap_full_size_property(built_library):
    """BASELINE batch 256: every pasted plane peaks at exactly 1.0 on mu; others are all-zero."""
    import advmix_b200 as A
    g = torch.Generator(device="cpu").manual_seed(3)
    joints = torch.zeros((256, 17, 3), dtype=torch.float64)
    joints[:, :, 0] = torch.rand((256, 17), generator=g, dtype=torch.float64) * 189
    joints[:, :, 1] = torch.rand((256, 17), generator=g, dtype=torch.float64) * 253
    vis = torch.ones((256, 17, 3), dtype=torch.float64)
    (h, mu), tw = A.generate_target(joints.to(dev()), vis.to(dev()))
    flat = h.view(256, 17, -1)
    mx, am = flat.max(dim=2)
    assert torch.all(mx == 1.0) and torch.all(tw == 1.0)
    assert torch.equal((am % 48).float(), mu[:, :, 0]) and torch.equal((am // 48).float(), mu[:, :, 1])


def test_mix_golden_and_oracle(built_library, golden):
    import advmix_b200 as A
    g = golden("mix")
    inputs = [torch.from_numpy(x).to(dev()) for x in g["inputs"]]
    w = torch.from_numpy(g["weights"]).to(dev())
    out = A.mix(inputs, w)
    assert np.array_equal(out.cpu().numpy(), g["tmp"])          # bit-exact given the weights
    logits = torch.from_numpy(g["logits"]).to(dev()).requires_grad_(True)
    out2 = A.mix_from_logits(inputs, logits)
    np.testing.assert_allclose(out2.detach().cpu().numpy(), g["tmp"], rtol=0, atol=2e-6)   # fused softmax: fp32 tolerance
    out2.backward(torch.from_numpy(g["grad_out"]).to(dev()))
    np.testing.assert_allclose(logits.grad.cpu().numpy(), g["grad_logits"], rtol=0, atol=5e-6)


@pytest.mark.parametrize("K,dtype", [(1, torch.float32), (2, torch.float32), (3, torch.float32), (3, torch.bfloat16), (8, torch.float32)])
def test_mix_shapes_dtypes(built_library, K, dtype):
    import advmix_b200 as A
    g = torch.Generator().manual_seed(K)
    B, C, H, W = 3, 3, 64, 48
    xs = [torch.randn(B, C, H, W, generator=g).to(dtype) for _ in range(K)]
    logits = torch.randn(B, K, H, W, generator=g)
    w = torch.softmax(logits, 1)
    ref = OM.mix_from_weights([x.float() for x in xs], w)
    out = A.mix([x.to(dev()) for x in xs], w.to(dev()))
    if dtype == torch.float32:
        assert torch.equal(out.cpu(), ref)
    else:
        assert torch.equal(out.cpu(), ref.to(torch.bfloat16))
    # backward w.r.t. weights
    wd = w.to(dev()).requires_grad_(True)
    go = torch.randn(B, C, H, W, generator=g)
    A.mix([x.to(dev()) for x in xs], wd).backward(go.to(dev()).to(dtype))
    ref_g = OM.mix_backward([x.float() for x in xs], w, go.to(dtype).float(), through_softmax=False)
    torch.testing.assert_close(wd.grad.cpu(), ref_g, rtol=1e-5, atol=1e-5)


def test_mix_linearity_full_size(built_library):
    """BASELINE config 3 size (32x3x256x192): convexity and linearity properties."""
    import advmix_b200 as A
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = [torch.randn(32, 3, 256, 192, device=dev(), generator=g) for _ in range(3)]
    logits = torch.randn(32, 3, 256, 192, device=dev(), generator=g)
    out = A.mix_from_logits(xs, logits)
    lo = torch.minimum(torch.minimum(xs[0], xs[1]), xs[2]); hi = torch.maximum(torch.maximum(xs[0], xs[1]), xs[2])
    assert torch.all(out >= lo - 1e-5) and torch.all(out <= hi + 1e-5)
    one_hot = torch.zeros_like(logits); one_hot[:, 1] = 1
    assert torch.equal(A.mix(xs, one_hot), xs[1])
    same = A.mix_from_logits([xs[0]] * 3, logits)
    torch.testing.assert_close(same, xs[0], rtol=1e-5, atol=1e-5)
