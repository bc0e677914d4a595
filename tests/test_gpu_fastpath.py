"""advmix_b200.fastpath: the one-call K = 1 step (advmix_crop_targets_step) and the HBM source cache must give exactly what
AdvMixBatchPipeline gives for the same draws (which the replay tests pin to the reference's __getitem__)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _records(B, rng, H, W):
    recs = []
    for _ in range(B):
        x, y, w, h = rng.uniform(5, W / 3), rng.uniform(5, H / 3), rng.uniform(W / 4, W / 2), rng.uniform(H / 4, H / 2)
        j = np.zeros((17, 3)); j[:, 0] = rng.uniform(x, x + w, 17); j[:, 1] = rng.uniform(y, y + h, 17)
        v = np.zeros((17, 3)); vv = (rng.random(17) < 0.8).astype(np.float64); v[:, 0] = vv; v[:, 1] = vv
        ar = 192 / 256
        if w > ar * h:
            h = w / ar
        else:
            w = h * ar
        recs.append({"center": np.array([x + w / 2, y + h / 2], np.float32), "scale": np.array([w / 200, h / 200], np.float32) * 1.25,
                     "joints_3d": j, "joints_3d_vis": v, "width": W, "height": H})
    return recs


@pytest.mark.parametrize("rows,out_ring", [("device", 0), ("host", 0), ("device", 2), ("graph", 3), ("graph2", 4)])
@pytest.mark.parametrize("size", [(120, 160), (97, 131)])
def test_step_equals_pipeline(built_library, size, rows, out_ring):
    import advmix_b200 as A
    from advmix_b200 import fastpath as F
    from advmix_b200.dataset import AdvMixBatchPipeline
    dev = torch.device("cuda:0")
    H, W = size
    N, B = 40, 16
    rng = np.random.default_rng(H)
    images = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(N)]
    recs = _records(N, rng, H, W)
    table = F.RecordTable.from_records(recs)
    pinned = [torch.from_numpy(im).pin_memory() for im in images]
    cache = F.SourceCache(N * ((3 * W + 15) // 16 * 16) * H + N * 256, N, dev)
    step = (F.CropTargetsStep(B, device=dev, seed=5, out_ring=out_ring, ring=out_ring, graph=True, prefetch_streams=2 if rows == "graph2" else 1)
            if rows.startswith("graph") else
            F.CropTargetsStep(B, device=dev, seed=5, record_rows=rows, out_ring=out_ring))
    pipe = AdvMixBatchPipeline(sample_times=1, is_train=True, device=dev)
    for it in range(10 if rows.startswith("graph") else 4):          # graph mode: every ring entry is captured once, then replayed
        ids = rng.permutation(N)[:B]
        before = cache.uploaded_bytes
        off, pitch, hh, ww = cache.ensure(ids, lambda i: pinned[i])
        if it == 2:
            ids2 = ids.copy()
            cache.ensure(ids2, lambda i: pinned[i])
            assert cache.uploaded_bytes >= before                       # (re-)ensuring resident images uploads nothing more
        c, s, rot, flip = step.draw(table.centers[ids], table.scales[ids], table.widths[ids])
        inp, (hm, mu), tw, meta = step(table, ids, cache.buffer, off, pitch, hh, ww, draws=(c, s, rot, flip), after=cache.take_upload_event())
        sb = A.SourceBatch.from_numpy([images[i] for i in ids], dev)
        e_inp, (e_hm, e_mu), e_tw, e_meta = pipe([recs[i] for i in ids], sources=sb, draws=(c, s, rot, flip))
        assert torch.equal(inp, e_inp), "crop differs (iteration %d)" % it
        assert torch.equal(hm, e_hm) and torch.equal(mu, e_mu) and torch.equal(tw, e_tw)
        assert torch.equal(meta["joints"], e_meta["joints"]) and torch.equal(meta["joints_vis"], e_meta["joints_vis"])
    total = cache.uploaded_bytes
    ids = np.arange(B)
    cache.ensure(ids, lambda i: pinned[i]); t1 = cache.uploaded_bytes
    cache.ensure(ids, lambda i: pinned[i])
    assert cache.uploaded_bytes == t1 and t1 >= total


def test_chunked_draws_follow_the_reference_distributions(built_library):
    """JointsDataset.py:177-188: s *= clip(randn*sf + 1, 1-sf, 1+sf); r = clip(randn*rf, -2rf, 2rf) with probability 0.6 else 0;
    flip with probability 0.5.  The step draws DRAW_CHUNK steps per numpy call; every step must get fresh values."""
    from advmix_b200 import fastpath as F
    B = 256
    step = F.CropTargetsStep(B, device=torch.device("cuda:0"), seed=9)
    centers = np.tile(np.array([[100.0, 50.0]], np.float32), (B, 1)); scales = np.ones((B, 2), np.float32); widths = np.full(B, 640)
    S, R, Fl = [], [], []
    for _ in range(F.CropTargetsStep.DRAW_CHUNK + 8):              # crosses a refill
        c, s, rot, flip = step.draw(centers, scales, widths)
        assert np.array_equal(c[:, 0], np.where(flip, 640 - 100.0 - 1, 100.0).astype(np.float32)) and np.all(c[:, 1] == 50.0)
        assert np.array_equal(s[:, 0], s[:, 1])
        S.append(s[:, 0].copy()); R.append(rot.copy()); Fl.append(flip.copy())
    S, R, Fl = np.concatenate(S), np.concatenate(R), np.concatenate(Fl)
    assert S.min() >= 0.7 and S.max() <= 1.3 and abs(S.mean() - 1.0) < 0.01 and 0.2 < S.std() < 0.3
    assert abs((R == 0).mean() - 0.4) < 0.02 and np.abs(R).max() <= 80.0 and 30 < R[R != 0].std() < 45
    assert abs(Fl.mean() - 0.5) < 0.02
    assert len(np.unique(S)) > 0.6 * len(S)                         # no chunk row handed out twice


def test_step_in_cuda_graph(built_library):
    """advmix_crop_targets_step forks onto a library-owned side stream with events: it must be capturable."""
    from advmix_b200 import fastpath as F
    dev = torch.device("cuda:0")
    H, W, N, B = 120, 160, 16, 16
    rng = np.random.default_rng(1)
    images = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(N)]
    table = F.RecordTable.from_records(_records(N, rng, H, W))
    pinned = [torch.from_numpy(im).pin_memory() for im in images]
    cache = F.SourceCache(N * W * 3 * H + N * 256, N, dev)
    step = F.CropTargetsStep(B, device=dev, seed=2)
    ids = np.arange(B)
    off, pitch, hh, ww = cache.ensure(ids, lambda i: pinned[i])
    draws = step.draw(table.centers[ids], table.scales[ids], table.widths[ids])
    ref = step(table, ids, cache.buffer, off, pitch, hh, ww, draws=draws)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step(table, ids, cache.buffer, off, pitch, hh, ww, draws=draws)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[0], ref[0]) and torch.equal(out[1][0], ref[1][0]) and torch.equal(out[2], ref[2])


@pytest.mark.parametrize("mode", ["eager", "ring_graph_prefetch"])
def test_advmix_step_equals_pipeline_k3(built_library, mode):
    """The K = 3 one-call step: crop, targets of the clean / gridmask chains and the materialised `inputs` list must equal
    AdvMixBatchPipeline(sample_times=3) for the same draws, and batch.mix must equal the mix of those inputs - eagerly with fresh
    outputs, and as CUDA-graph replays into an output ring on two prefetch streams (every ring entry captured, then replayed)."""
    import advmix_b200 as A
    from advmix_b200 import fastpath as F
    from advmix_b200.dataset import AdvMixBatchPipeline
    dev = torch.device("cuda:0")
    H, W, N, B = 120, 160, 24, 8
    rng = np.random.default_rng(3)
    images = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(N)]
    recs = _records(N, rng, H, W)
    table = F.RecordTable.from_records(recs)
    pinned = [torch.from_numpy(im).pin_memory() for im in images]
    cache = F.SourceCache(N * W * 3 * H + N * 256, N, dev)
    step = F.AdvMixStep(B, device=dev, seed=7) if mode == "eager" else \
        F.AdvMixStep(B, device=dev, seed=7, ring=4, out_ring=4, graph=True, prefetch_streams=2)
    pipe = AdvMixBatchPipeline(sample_times=3, is_train=True, device=dev, draw_mode="batched", seed=11)
    for it in range(1 if mode == "eager" else 9):
        ids = rng.permutation(N)[:B]
        off, pitch, hh, ww = cache.ensure(ids, lambda i: pinned[i])
        draws = step.draw(table.centers[ids], table.scales[ids], table.widths[ids])
        sb = A.SourceBatch.from_numpy([images[i] for i in ids], dev)
        inputs, tgts, tws, metas = pipe([recs[i] for i in ids], sources=sb, draws=draws)
        cp = pipe.last_chain_params                                   # the chain draws the pipeline made: feed the same ones to the step
        aa = (cp["autoaug"][0].cpu().numpy(), cp["autoaug"][1].cpu().numpy())
        gm = cp["gridmask"].cpu().numpy()
        batch = step(table, ids, cache.buffer, off, pitch, hh, ww, draws=draws, chain_draws=(aa, gm), after=cache.take_upload_event())
        assert torch.equal(batch.crop_u8, cp["crop_u8"]), it
        got = batch.inputs()
        for k in range(3):
            assert torch.equal(got[k], inputs[k]), "chain %d differs (iteration %d)" % (k, it)
        assert torch.equal(batch.target, tgts[0]) and torch.equal(batch.target_weight, tws[0])
        assert torch.equal(batch.target_gridmask, tgts[2]) and torch.equal(batch.target_weight_gridmask, tws[2])
        assert torch.equal(batch.joints_vis_gridmask, metas[2]["joints_vis"])
        logits = torch.randn(B, 3, 256, 192, device=dev, requires_grad=True)
        tmp = batch.mix(logits)
        ref = A.mix_from_logits([t.contiguous() for t in inputs], logits.detach())
        assert torch.equal(tmp, ref)
        tmp.sum().backward()
        assert torch.isfinite(logits.grad).all()
    # own draws: the chunked chain draws hand out fresh values every step
    b1 = step(table, ids, cache.buffer, off, pitch, hh, ww)
    b2 = step(table, ids, cache.buffer, off, pitch, hh, ww)
    assert not (np.array_equal(b1.autoaug[0], b2.autoaug[0]) and np.array_equal(b1.autoaug[1], b2.autoaug[1]))
