"""the REAL lib/core/function.py:train_advmix (imported from the live reference tree, or from the
byte-compiled oracle/_ref on the GPU box) driven with a fake loader that yields AdvMixBatchPipeline output.
Proves the a8 contract - `inputs[k]`, `targets[0]`, `target_weights[0]`, `metas[0]` (function.py:129-133) - is
consumable unchanged, and that advmix_b200.mix reproduces the `tmp` the loop feeds to D (function.py:142-146)."""
import copy
import random
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _records(g, n):
    return [{"image": g["images"][int(g["image_index"][i])], "center": g["centers"][i], "scale": g["scales"][i],
             "joints_3d": g["joints"][i], "joints_3d_vis": g["vis"][i]} for i in range(n)]


def test_real_train_advmix_consumes_pipeline_output(built_library, golden):
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("reference modules not available (neither /root/reference nor oracle/_ref)")
    import advmix_b200 as A
    from advmix_b200.dataset import AdvMixBatchPipeline
    ns = ref_harness.load()
    g = golden("replay")
    B = 8
    pipe = AdvMixBatchPipeline(sample_times=3, is_train=True, draw_mode="reference", prob_half_body=0.3)
    np.random.seed(5); random.seed(5)
    batches = [pipe(_records(g, B)) for _ in range(2)]

    torch.manual_seed(0)
    D = torch.nn.Sequential(torch.nn.Conv2d(3, 17, 4, stride=4)).cuda()          # [B,3,256,192] -> [B,17,64,48]
    G = torch.nn.Sequential(torch.nn.Conv2d(9, 3, 3, padding=1)).cuda()          # gen_input_chn=9 -> K=3 logits
    T = copy.deepcopy(D)
    opt, opt_G = torch.optim.SGD(D.parameters(), lr=1e-2), torch.optim.SGD(G.parameters(), lr=1e-2)
    crit = ns.loss.JointsMSELoss(use_target_weight=True).cuda()
    cfg = types.SimpleNamespace(PRINT_FREQ=1, DEBUG=types.SimpleNamespace(DEBUG=False))
    args = types.SimpleNamespace(alpha=0.5, adv_loss_weight=1.0)
    scalars = []
    writer = types.SimpleNamespace(add_scalar=lambda *a: scalars.append(a))
    writer_dict = {"writer": writer, "train_global_steps": 0}

    # what D is fed: first tmp.detach(), then tmp (function.py:146,160)
    seen = []
    D.register_forward_pre_hook(lambda mod, inp: seen.append(inp[0].detach().clone()))
    g0 = copy.deepcopy(G.state_dict())
    d0 = [p.detach().clone() for p in D.parameters()]
    gp0 = [p.detach().clone() for p in G.parameters()]

    ns.function.train_advmix(cfg, args, batches, [D, G, T], crit, [opt, opt_G], 0, "/tmp", "/tmp", writer_dict)

    assert writer_dict["train_global_steps"] == 2 and len(scalars) == 4
    assert all(np.isfinite(s[1]) for s in scalars)
    assert any(not torch.equal(a, b) for a, b in zip(d0, D.parameters())), "D step did not update the pose net"
    assert any(not torch.equal(a, b) for a, b in zip(gp0, G.parameters())), "G step got no gradient through the mix"
    assert all(torch.equal(a, b) for a, b in zip(T.parameters(), d0)), "teacher must stay frozen"

    # advmix_b200.mix == the loop's tmp for step 0 (G weights as they were before the step), bit for bit
    G0 = copy.deepcopy(G)
    G0.load_state_dict(g0)
    inputs = batches[0][0]
    with torch.no_grad():
        w = torch.softmax(G0(torch.cat(inputs, dim=1)), dim=1)
        ours = A.mix(inputs, w)
        fused = A.mix_from_logits(inputs, G0(torch.cat(inputs, dim=1)))
    assert torch.equal(ours, seen[0]) and torch.equal(ours, seen[1])
    assert torch.allclose(fused, seen[0], atol=2e-6, rtol=0)
