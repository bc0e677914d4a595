"""Each corruption once per severity in argv[2] (default 3,5) on N images: run under ncu to get the per-kernel time split."""
import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import advmix_b200 as A
from advmix_b200 import corruptions as K
dev = torch.device('cuda:0'); N, H, W = 256, 256, 192
g = torch.Generator(device=dev).manual_seed(1)
low = torch.rand((N, 3, H // 16 + 2, W // 16 + 2), device=dev, generator=g)
img = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False) * 255
img = (img + torch.randint(-8, 9, img.shape, device=dev, generator=g)).clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
out = torch.empty_like(img)
names = sys.argv[1].split(",") if len(sys.argv) > 1 and sys.argv[1] != "all" else A.get_corruption_names("all")
sevs = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "3,5").split(",")]
encode = "jpeg_encode" in names            # also run the device JPEG encoder (tools/make_datasets.py:45) once
names = [n for n in names if n != "jpeg_encode"]
for n in names:
    for s in sevs:
        K.corrupt_batch(img, n, s, seed=3, out=out, fast=True)   # warm (tables)
if encode:
    from advmix_b200 import jpeg as J
    J.encode_batch_device(img); J.encode_batch_device(img)
torch.cuda.synchronize()
for n in names:
    for s in sevs:
        torch.cuda.nvtx.range_push("%s_%d" % (n, s))
        K.corrupt_batch(img, n, s, seed=3, out=out, fast=True)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
