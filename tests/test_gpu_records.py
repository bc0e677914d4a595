"""GPU parity for SURVEY row f4: _xywh2cs, half_body_transform, select_data batched on the device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import records as OR       # noqa: E402


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_half_body_and_select_data_golden(built_library, golden):
    from advmix_b200 import records as R
    g = golden("records")
    c, s, valid = R.half_body_transform(t(g["joints"]), t(g["vis"]), tuple(int(u) for u in g["upper"]), t(g["randn"]),
                                        float(g["aspect"]))
    assert np.array_equal(valid.cpu().numpy(), g["hb_valid"])
    ok = g["hb_valid"]
    assert np.array_equal(c.cpu().numpy()[ok], g["hb_center"][ok])
    assert np.array_equal(s.cpu().numpy()[ok], g["hb_scale"][ok])
    keep = R.select_data(t(g["joints"]), t(g["vis"]), t(g["center"]), t(g["scale"]))
    assert np.array_equal(keep.cpu().numpy(), g["keep"])


@pytest.mark.parametrize("aspect", [192 / 256, 1.0, 288 / 384, 3 / 7])
def test_xywh2cs_vs_oracle(built_library, aspect):
    from advmix_b200 import records as R
    rng = np.random.default_rng(3)
    boxes = np.stack([rng.uniform(-5, 600, 500), rng.uniform(-5, 400, 500), rng.uniform(1, 400, 500), rng.uniform(1, 400, 500)], 1)
    boxes[0] = [10, 20, 75, 100]                       # w == aspect*h exactly for 0.75
    boxes[1] = [-41, 5, 80, 80]                        # centre x == -1: no 1.25 padding
    c, s = R.xywh2cs(t(boxes), aspect)
    for b in range(len(boxes)):
        ec, es = OR.xywh2cs(*[float(v) for v in boxes[b]], aspect)
        assert np.array_equal(c[b].cpu().numpy(), ec) and np.array_equal(s[b].cpu().numpy(), es), b


def test_half_body_random_vs_oracle(built_library):
    from advmix_b200 import records as R
    rng = np.random.default_rng(8)
    B, J = 400, 16
    upper = (7, 8, 9, 10, 11, 12, 13, 14, 15)           # MPII upper body ids
    joints = np.zeros((B, J, 3)); joints[..., :2] = rng.uniform(0, 500, (B, J, 2))
    vis = np.zeros((B, J, 3)); vis[..., 0] = vis[..., 1] = rng.random((B, J)) < rng.random((B, 1))
    draw = rng.standard_normal(B)
    c, s, valid = R.half_body_transform(t(joints), t(vis), upper, t(draw), 1.0)
    for b in range(B):
        ec, es = OR.half_body_transform(joints[b], vis[b], upper, draw[b], 1.0)
        if ec is None:
            assert not bool(valid[b])
        else:
            assert bool(valid[b]) and np.array_equal(c[b].cpu().numpy(), ec) and np.array_equal(s[b].cpu().numpy(), es), b
    assert R.xywh2cs(torch.zeros((0, 4), dtype=torch.float64, device="cuda"), 0.75)[0].shape == (0, 2)
