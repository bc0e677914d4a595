import sys, json, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import advmix_b200 as A
from advmix_b200 import corruptions as K
dev = torch.device('cuda:0'); N, H, W = 512, 256, 192
g = torch.Generator(device=dev).manual_seed(1)
low = torch.rand((N, 3, H // 16 + 2, W // 16 + 2), device=dev, generator=g)
img = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False) * 255
img = (img + torch.randint(-8, 9, img.shape, device=dev, generator=g)).clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
out = torch.empty_like(img)
names = sys.argv[1].split(",") if len(sys.argv) > 1 else A.get_corruption_names("all")
tot = 0
for n in names:
    row = []
    for s in range(1, 6):
        for _ in range(2): K.corrupt_batch(img, n, s, seed=3, out=out, fast=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): K.corrupt_batch(img, n, s, seed=3, out=out, fast=True)
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 3 * 1e3 / N
        row.append(us); tot += us
    print('%-18s us/img %s   frac@s3 %.3f' % (n, ' '.join('%6.3f' % v for v in row), 294912 / (row[2] * 1e-6) / 6550.1e9))
print('sum us per image over all', tot)
