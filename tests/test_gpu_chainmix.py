"""GPU parity of the fused chain + mix kernels (lib/core/function.py:137-146,158-164 on chains that
are recomputed from the uint8 crop): against the materialised-chain path, against the oracle expression, and against
the REAL reference __getitem__ chains (tests/golden/replay.npz hashes / getitem.npz tensors)."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mix as OM       # noqa: E402


def dev():
    return torch.device("cuda:0")


def _pipeline_batch(golden, n=8, seed=21):
    from advmix_b200.dataset import AdvMixBatchPipeline
    g = golden("replay")
    recs = [{"image": g["images"][int(g["image_index"][i])], "center": g["centers"][i], "scale": g["scales"][i],
             "joints_3d": g["joints"][i], "joints_3d_vis": g["vis"][i]} for i in range(n)]
    pipe = AdvMixBatchPipeline(sample_times=3, is_train=True, draw_mode="reference", prob_half_body=0.3)
    np.random.seed(seed); random.seed(seed)
    inputs, tgts, tws, metas = pipe(recs)
    return pipe, inputs


def _force_all_ops(pipe, inputs):
    """Chain parameters that cover every op pair incl. both sharpness placements and gridmask on / off."""
    import advmix_b200 as A
    from advmix_b200 import chains as CH
    p = pipe.last_chain_params
    crop = p["crop_u8"]
    B = crop.shape[0]
    ops = torch.tensor([[1, 1], [2, 2], [1, 3], [3, 1], [5, 4], [4, 1], [0, 5], [0, 0]][:B], dtype=torch.int32)
    mags = torch.tensor([[0, 0], [5, 6], [0, 142.2], [170.7, 0], [1.63, 0], [0, 0], [0, 0.37], [0, 0]][:B], dtype=torch.float32)
    gm = torch.tensor([[1, 37, 5, 11], [0, 0, 0, 0], [1, 2, 1, 0], [1, 191, 100, 3], [1, 64, 0, 63], [0, 0, 0, 0], [1, 20, 19, 19],
                       [1, 100, 50, 50]][:B], dtype=torch.int32)
    clean = inputs[0]
    _, aug = CH.autoaug(crop, ops, mags, norm_dtype=torch.float32, want_u8=False)
    gimg, _ = CH.gridmask(clean, gm)
    return crop, ops, mags, gm, [clean, aug, gimg]


@pytest.mark.parametrize("forced", [False, True])
def test_chainmix_equals_materialised_mix(built_library, golden, forced):
    import advmix_b200 as A
    pipe, inputs = _pipeline_batch(golden)
    if forced:
        crop, ops, mags, gm, chains = _force_all_ops(pipe, inputs)
    else:
        p = pipe.last_chain_params
        crop, (ops, mags), gm, chains = p["crop_u8"], p["autoaug"], p["gridmask"], inputs
    plans = A.autoaug_plan(crop, ops, mags)
    B, _, H, W = chains[0].shape
    gen = torch.Generator(device="cuda").manual_seed(3)
    logits = torch.randn((B, 3, H, W), device="cuda", generator=gen)
    # G_input == cat(inputs, 1)
    assert torch.equal(A.chains_g_input(crop, plans, gm), torch.cat(chains, dim=1))
    # weights given: bit-exact vs the materialised mix and vs the oracle expression
    w = torch.softmax(logits, dim=1)
    fused = A.chain_mix(crop, plans, gm, w)
    assert torch.equal(fused, A.mix(chains, w))
    assert torch.equal(fused.cpu(), OM.mix_from_weights([c.cpu() for c in chains], w.cpu()))
    # fused softmax + backward: identical code path to mix.cu -> identical bits
    l1 = logits.clone().requires_grad_(True)
    l2 = logits.clone().requires_grad_(True)
    o1 = A.chain_mix_from_logits(crop, plans, gm, l1)
    o2 = A.mix_from_logits(chains, l2)
    assert torch.equal(o1, o2)
    go = torch.randn(o1.shape, device="cuda", generator=gen)
    o1.backward(go); o2.backward(go)
    assert torch.equal(l1.grad, l2.grad)
    # and against the oracle (torch autograd on the reference expression, CPU float32)
    exp, _ = OM.mix_from_logits([c.cpu() for c in chains], logits.cpu())
    assert torch.allclose(o1.detach().cpu(), exp, atol=2e-6, rtol=0)
    gl = OM.mix_backward([c.cpu() for c in chains], logits.cpu(), go.cpu())
    assert torch.allclose(l1.grad.cpu(), gl, atol=5e-6, rtol=0)
    # grad w.r.t. given weights
    w1 = w.clone().requires_grad_(True); w2 = w.clone().requires_grad_(True)
    A.chain_mix(crop, plans, gm, w1).backward(go); A.mix(chains, w2).backward(go)
    assert torch.equal(w1.grad, w2.grad)


def test_chainmix_on_real_reference_chains(built_library, golden):
    """The three chain tensors the REAL __getitem__ produced (getitem.npz) mixed by the oracle expression == the fused kernel
    fed only the uint8 crop and the replayed chain parameters."""
    import advmix_b200 as A
    from advmix_b200.dataset import AdvMixBatchPipeline
    g = golden("getitem")
    pipe = AdvMixBatchPipeline(sample_times=3, is_train=True, draw_mode="reference")
    for i in range(3):
        rec = {"image": g["images"][i], "center": g["centers"][i], "scale": g["scales"][i], "joints_3d": g["joints"][i],
               "joints_3d_vis": g["vis"][i]}
        np.random.seed(700 + i); random.seed(700 + i)
        pipe([rec])
        p = pipe.last_chain_params
        plans = A.autoaug_plan(p["crop_u8"], *p["autoaug"])
        ref_inputs = [torch.from_numpy(g["k3_in%d" % i][k])[None] for k in range(3)]
        logits = torch.randn((1, 3, 256, 192), generator=torch.Generator().manual_seed(i))
        w = torch.softmax(logits, dim=1)
        tmp = ref_inputs[0] * w[:, 0, ...].unsqueeze(dim=1)           # function.py:142-144 verbatim
        for k in range(1, 3):
            tmp += ref_inputs[k] * w[:, k].unsqueeze(dim=1)
        fused = A.chain_mix(p["crop_u8"], plans, p["gridmask"], w.cuda())
        assert torch.equal(fused.cpu(), tmp), i


@pytest.mark.parametrize("w_dtype,out_dtype", [(torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16), (torch.bfloat16, torch.float32)])
def test_chainmix_bf16_variants(built_library, golden, w_dtype, out_dtype):
    """bfloat16 logits in / mix out (737 280 B per sample): equals the float32 kernel on the up-cast logits, rounded."""
    import advmix_b200 as A
    pipe, inputs = _pipeline_batch(golden)
    crop, ops, mags, gm, chains = _force_all_ops(pipe, inputs)
    plans = A.autoaug_plan(crop, ops, mags)
    logits = torch.randn(chains[0].shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)).to(w_dtype)
    l1 = logits.clone().requires_grad_(True)
    out = A.chain_mix_from_logits(crop, plans, gm, l1, out_dtype=out_dtype)
    assert out.dtype == out_dtype
    l2 = logits.float().clone().requires_grad_(True)
    exp = A.chain_mix_from_logits(crop, plans, gm, l2)
    assert torch.equal(out, exp.to(out_dtype))
    go = torch.randn(out.shape, device="cuda").to(out_dtype)
    out.backward(go); exp.backward(go.float())
    assert l1.grad.dtype == w_dtype
    assert torch.equal(l1.grad, l2.grad.to(w_dtype))
    assert torch.equal(A.chains_g_input(crop, plans, gm, dtype=torch.bfloat16), torch.cat(chains, 1).to(torch.bfloat16))


@pytest.mark.parametrize("K", [1, 2, 3, 4])
def test_mix_u8_equals_normalised_mix(built_library, K):
    """General uint8-chain mix (corruption chains of the 15x5 set): equals mix() over the normalised tensors."""
    import advmix_b200 as A
    gen = torch.Generator(device="cuda").manual_seed(K)
    B, H, W = 5, 64, 48
    chains = [torch.randint(0, 256, (B, H, W, 3), device="cuda", dtype=torch.uint8, generator=gen) for _ in range(K)]
    norm = [A.to_tensor_normalize(c) for c in chains]
    logits = torch.randn((B, K, H, W), device="cuda", generator=gen)
    l1 = logits.clone().requires_grad_(True); l2 = logits.clone().requires_grad_(True)
    o1 = A.mix_u8_from_logits(chains, l1); o2 = A.mix_from_logits(norm, l2)
    assert torch.equal(o1, o2)
    go = torch.randn(o1.shape, device="cuda", generator=gen)
    o1.backward(go); o2.backward(go)
    assert torch.equal(l1.grad, l2.grad)
    w = torch.softmax(logits, 1)
    assert torch.equal(A.mix_u8(chains, w), A.mix(norm, w))
    assert torch.equal(A.mix_u8_from_logits(chains, logits.bfloat16(), out_dtype=torch.bfloat16),
                       A.mix_u8_from_logits(chains, logits.bfloat16().float()).to(torch.bfloat16))


def test_chainmix_full_size_property_and_errors(built_library):
    """B = 256 (BASELINE global batch): with one-hot weights the mix returns exactly the selected chain; shape errors raise."""
    import advmix_b200 as A
    from advmix_b200 import chains as CH
    gen = torch.Generator(device="cuda").manual_seed(9)
    B, H, W = 256, 256, 192
    crop = torch.randint(0, 256, (B, H, W, 3), device="cuda", dtype=torch.uint8, generator=gen)
    rng = np.random.default_rng(1)
    ops, mags = CH.sample_autoaug_batch(B, rng)
    gm = CH.sample_gridmask_batch(B, H, W, rng)
    plans = A.autoaug_plan(crop, ops, mags)
    gi = A.chains_g_input(crop, plans, gm)
    for k in range(3):
        w = torch.zeros((B, 3, H, W), device="cuda"); w[:, k] = 1
        assert torch.equal(A.chain_mix(crop, plans, gm, w), gi[:, 3 * k:3 * k + 3] + 0.0)
    with pytest.raises(ValueError):
        A.chain_mix(crop, plans, gm, torch.zeros((B, 2, H, W), device="cuda"))
    with pytest.raises(A.AdvmixError):
        A.chain_mix(crop[:, :, :190], plans, gm, torch.zeros((B, 3, H, 190), device="cuda"))
