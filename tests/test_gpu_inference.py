"""GPU parity for SURVEY row f3: heat-map decode (get_max_preds / get_final_preds) and the flip-test
merge, through the C ABI, against the reference fixture and the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import inference as OI       # noqa: E402

COCO_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]


def dev():
    return torch.device("cuda:0")


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def test_decode_golden(built_library, golden):
    import advmix_b200 as A
    g = golden("inference")
    hm = t(g["heatmaps"])
    p, m = A.get_max_preds(hm)
    assert np.array_equal(p.cpu().numpy(), g["max_preds"])
    assert np.array_equal(m.cpu().numpy(), g["maxvals"])
    for pp in (0, 1):
        preds, mv = A.get_final_preds(hm, t(g["center"]), t(g["scale"]), post_process=bool(pp))
        assert np.array_equal(mv.cpu().numpy(), g["maxvals"])
        # the 2x3 inverse comes from a closed-form float64 solve (cv2 uses LU): equal after the float32 store
        # except where the float64 value sits on a float32 rounding boundary
        got, exp = preds.cpu().numpy(), g["final_preds_pp%d" % pp]
        np.testing.assert_allclose(got, exp, rtol=2e-7, atol=0)
        assert (got != exp).mean() < 0.02


def test_flip_merge_golden(built_library, golden):
    import advmix_b200 as A
    from advmix_b200.transforms import flip_perm
    g = golden("inference")
    hm = g["heatmaps"]
    hf = (np.random.default_rng(77).standard_normal(hm.shape) * 0.3).astype(np.float32)
    perm = flip_perm(17, COCO_PAIRS, dev())
    assert np.array_equal(A.flip_merge(t(hm), t(hf), perm, True).cpu().numpy(), g["merged_shift1"])
    assert np.array_equal(A.flip_merge(t(hm), t(hf), perm, False).cpu().numpy(), OI.flip_merge(hm, hf, COCO_PAIRS, False))
    assert np.array_equal(A.flip_back(t(hf), perm).cpu().numpy(), OI.flip_back(hf, COCO_PAIRS))
    # in place on `output`
    a = t(hm)
    A.flip_merge(a, t(hf), perm, True, out=a)
    assert np.array_equal(a.cpu().numpy(), g["merged_shift1"])


@pytest.mark.parametrize("shape", [(5, 16, 64, 64), (2, 17, 128, 96), (3, 4, 33, 31), (1, 1, 8, 5)])
def test_decode_random_vs_oracle(built_library, shape):
    """Other map sizes (MPII 64x64, 512-input 128x96, odd sizes that miss the float4 path), ties, peaks on
    the frame, negative planes."""
    import advmix_b200 as A
    rng = np.random.default_rng(sum(shape))
    B, J, H, W = shape
    hm = (rng.standard_normal(shape) * 0.05).astype(np.float32)
    for b in range(B):
        for j in range(J):
            y, x = int(rng.integers(0, H)), int(rng.integers(0, W))
            hm[b, j, y, x] += 1.0
            if rng.random() < 0.3:
                y2, x2 = int(rng.integers(0, H)), int(rng.integers(0, W))
                hm[b, j, y2, x2] = hm[b, j, y, x]            # exact tie
    hm[0, 0] = -1.0
    center = np.stack([rng.uniform(50, 400, B), rng.uniform(50, 300, B)], 1).astype(np.float32)
    scale = np.stack([rng.uniform(0.5, 3.0, B), rng.uniform(0.6, 4.0, B)], 1).astype(np.float32)
    for pp in (False, True):
        e_preds, e_max, e_coords = OI.get_final_preds(hm.copy(), center, scale, pp)
        lib_preds, lib_max, lib_coords = A.inference._decode(t(hm), t(center), t(scale), pp, True)
        assert np.array_equal(lib_max.cpu().numpy(), e_max)
        assert np.array_equal(lib_coords.cpu().numpy(), e_coords)
        np.testing.assert_allclose(lib_preds.cpu().numpy(), e_preds, rtol=2e-7, atol=0)


def test_decode_of_generated_targets_full_batch(built_library):
    """Size-independent property at the bench batch (B=256): decoding the heat maps generate_target wrote
    returns mu wherever the joint is visible and on the map."""
    import advmix_b200 as A
    rng = np.random.default_rng(11)
    B, J = 256, 17
    joints = np.zeros((B, J, 3)); joints[..., 0] = rng.uniform(0, 192, (B, J)); joints[..., 1] = rng.uniform(0, 256, (B, J))
    vis = np.repeat((rng.random((B, J, 1)) < 0.8).astype(np.float64), 3, -1)
    (hm, mu), tw = A.generate_target(t(joints), t(vis), (192, 256), (48, 64), 2)
    coords, maxvals = A.get_max_preds(hm)
    on = (tw[..., 0] > 0) & (mu[..., 0] >= 0) & (mu[..., 0] < 48) & (mu[..., 1] >= 0) & (mu[..., 1] < 64)
    assert on.float().mean() > 0.6
    assert torch.equal(coords[on], mu[on])
    assert torch.all(maxvals[on] == 1.0)
    assert torch.all(coords[tw[..., 0] == 0] == 0)
    # flip_back twice is the identity
    from advmix_b200.transforms import flip_perm
    perm = flip_perm(J, COCO_PAIRS, dev())
    assert torch.equal(A.flip_back(A.flip_back(hm, perm), perm), hm)


def test_decode_empty_and_errors(built_library):
    import advmix_b200 as A
    p, m = A.get_max_preds(torch.zeros((0, 17, 64, 48), device=dev()))
    assert p.shape == (0, 17, 2) and m.shape == (0, 17, 1)
    with pytest.raises(TypeError):
        A.get_max_preds(torch.zeros((1, 17, 64, 48)))
    with pytest.raises(AssertionError):
        A.get_max_preds(torch.zeros((17, 64, 48), device=dev()))
