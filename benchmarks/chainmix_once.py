"""A few launches of the fused chain + mix kernels (and the materialised mix) at B = 256 for ncu captures."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import advmix_b200 as A                    # noqa: E402
from advmix_b200 import chains as CH       # noqa: E402

dev = torch.device("cuda:0")
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 256, 192
dt = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float32
rng = np.random.default_rng(0)
g = torch.Generator(device=dev).manual_seed(0)
crop = torch.randint(0, 256, (B, H, W, 3), device=dev, dtype=torch.uint8, generator=g)
ops, mags = CH.sample_autoaug_batch(B, rng)
gm = torch.as_tensor(CH.sample_gridmask_batch(B, H, W, rng)).to(dev)
plans = A.autoaug_plan(crop, ops, mags)
logits = torch.randn((B, 3, H, W), device=dev, generator=g).to(dt).requires_grad_(True)
for _ in range(3):
    gi = A.chains_g_input(crop, plans, gm, dtype=dt)
    out = A.chain_mix_from_logits(crop, plans, gm, logits, out_dtype=dt)
    out.backward(torch.ones_like(out))
    xs = [gi[:, 3 * k:3 * k + 3].contiguous() for k in range(3)]
    lf = logits.detach().float().requires_grad_(True)
    o2 = A.mix_from_logits(xs, lf)
    o2.backward(torch.ones_like(o2))
torch.cuda.synchronize()
print("ok")
