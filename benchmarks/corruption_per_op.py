"""Per-op timing of the corruption kernels on 512 resident 256x192 crops: us per image at severity 1..5 and the
fraction of the measured HBM peak at severity 3 on 294 912 algorithmic bytes per unit.
    python benchmarks/corruption_per_op.py [names|all|common] [fast|exact|both] [out.json]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import advmix_b200 as A                            # noqa: E402
from advmix_b200 import corruptions as K           # noqa: E402

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
dev = torch.device("cuda:0")
N, H, W = 512, 256, 192
g = torch.Generator(device=dev).manual_seed(1)
low = torch.rand((N, 3, H // 16 + 2, W // 16 + 2), device=dev, generator=g)
img = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False) * 255
img = (img + torch.randint(-8, 9, img.shape, device=dev, generator=g)).clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
out = torch.empty_like(img)
arg = sys.argv[1] if len(sys.argv) > 1 else "all"
names = A.get_corruption_names(arg) if arg in ("all", "common", "validation") else arg.split(",")
modes = {"fast": [True], "exact": [False], "both": [False, True]}[sys.argv[2] if len(sys.argv) > 2 else "fast"]
report = {"peak_gbs": PEAK, "n_images": N, "size": [H, W], "unit_bytes": 2 * H * W * 3, "ops": {}}
for fast in modes:
    tot = 0.0
    for n in names:
        row = []
        for s in range(1, 6):
            for _ in range(2):
                K.corrupt_batch(img, n, s, seed=3, out=out, fast=fast)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                K.corrupt_batch(img, n, s, seed=3, out=out, fast=fast)
            b.record()
            torch.cuda.synchronize()
            us = a.elapsed_time(b) / 3 * 1e3 / N
            row.append(us)
            tot += us
        fr = [2 * H * W * 3 / (v * 1e-6) / (PEAK * 1e9) for v in row]
        report["ops"].setdefault(n, {})["fast" if fast else "exact"] = {"us_per_image": row, "frac_of_measured_hbm": fr}
        print("%-5s %-18s us/img %s   frac %s" % ("fast" if fast else "exact", n, " ".join("%6.3f" % v for v in row), " ".join("%.3f" % v for v in fr)))
    report["sum_us_per_image_" + ("fast" if fast else "exact")] = tot
    # the five severities in one call (advmix_corrupt_sweep_u8c3)
    out5 = torch.empty((5,) + tuple(img.shape), dtype=torch.uint8, device=dev)
    tots = 0.0
    for n in names:
        for _ in range(2):
            K.corrupt_sweep(img, n, seed=3, out=out5, fast=fast)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            K.corrupt_sweep(img, n, seed=3, out=out5, fast=fast)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / 3 * 1e3 / N
        tots += us
        report["ops"][n]["fast_sweep5" if fast else "exact_sweep5"] = us
        print("%-5s %-18s five severities in one call: %6.3f us/img (separate calls: %6.3f)" % (
            "fast" if fast else "exact", n, us, sum(report["ops"][n]["fast" if fast else "exact"]["us_per_image"])))
    report["sum_us_per_image_sweep5_" + ("fast" if fast else "exact")] = tots
    print("%s sweep: %.2f us per image -> %.0f outputs/s" % ("fast" if fast else "exact", tots, 5 * len(names) / tots * 1e6))
    del out5
    print("%s: sum us per image over %d ops x 5 severities = %.2f  ->  %.0f outputs/s" % ("fast" if fast else "exact", len(names), tot, 5 * len(names) / tot * 1e6))
if len(sys.argv) > 3:
    json.dump(report, open(sys.argv[3], "w"), indent=1)
