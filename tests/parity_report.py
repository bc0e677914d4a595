"""Per-op mismatch statistics of the CUDA corruptions vs the oracle (injected draws), 256x192, all severities."""
import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
from oracle import corruptions as OK
from advmix_b200 import corruptions as K
from test_gpu_chains_corruptions import natural, pack_draws
dev = torch.device('cuda:0')
H, W = 256, 192
print("op severity n_values n_diff max_diff")
for name in OK.get_corruption_names(sys.argv[1] if len(sys.argv) > 1 else "all"):
    for sev in range(1, 6):
        rng = np.random.default_rng(1000 + sev)
        imgs = np.stack([natural(rng, H, W), rng.integers(0, 256, (H, W, 3), dtype=np.uint8)])
        imgs[0, :40, :40] = 255; imgs[0, -40:, -40:] = 0       # saturated regions
        bank = OK.synthetic_frost_bank(n=5, fh=H + 64, fw=W + 64)
        draws = [OK.make_draws(name, sev, H, W, rng, bank.shape) for _ in imgs]
        field, param = pack_draws(name, sev, H, W, draws)
        out = K.corrupt_batch(torch.from_numpy(imgs).to(dev), name, sev, rand_field=field, rand_param=param,
                              frost_bank=torch.from_numpy(bank).to(dev)).cpu().numpy()
        nd, md = 0, 0
        for i in range(len(imgs)):
            exp = OK.corrupt_with_draws(imgs[i], sev, name, draws[i], bank)
            d = np.abs(out[i].astype(int) - exp.astype(int)); nd += int((d > 0).sum()); md = max(md, int(d.max()))
        print(name, sev, out.size, nd, md, flush=True)
