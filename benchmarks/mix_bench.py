"""Mix kernels (row a3 / N1) timed alone: the materialised-chain mix, the fused chain + mix (chains recomputed from the
uint8 crop) and eager torch, forward and backward, at the reference's per-GPU batch (32) and the global batch (256).
Bytes are the ALGORITHMIC bytes of each variant; `frac` is against MEASURED_PEAKS.json hbm_gbs.
    python benchmarks/mix_bench.py [out.json]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import advmix_b200 as A                    # noqa: E402
from advmix_b200 import chains as CH       # noqa: E402

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
dev = torch.device("cuda:0")
H, W = 256, 192
P = H * W


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3       # us


rows = []
for B in (32, 256):
    rng = np.random.default_rng(B)
    g = torch.Generator(device=dev).manual_seed(B)
    crop = torch.randint(0, 256, (B, H, W, 3), device=dev, dtype=torch.uint8, generator=g)
    ops, mags = CH.sample_autoaug_batch(B, rng)
    gm = torch.as_tensor(CH.sample_gridmask_batch(B, H, W, rng)).to(dev)
    plans = A.autoaug_plan(crop, ops, mags)
    for dt in (torch.float32, torch.bfloat16):
        es = 2 if dt == torch.bfloat16 else 4
        gi = A.chains_g_input(crop, plans, gm, dtype=dt)
        xs = [gi[:, 3 * k:3 * k + 3].contiguous() for k in range(3)]
        logits = torch.randn((B, 3, H, W), device=dev, generator=g)
        lg = logits.to(dt)
        w = torch.softmax(logits, 1)
        go = torch.randn((B, 3, H, W), device=dev, generator=g).to(dt)
        cases = [
            # name, fn, algorithmic bytes per sample
            ("mix(weights) materialised", lambda: A.mix(xs, w), P * (9 * es + 12 + 3 * es)),
            ("mix_from_logits materialised", lambda: A.mix_from_logits(xs, logits), P * (9 * es + 12 + 3 * es + 12)),
            ("chain_mix_from_logits FUSED", lambda: A.chain_mix_from_logits(crop, plans, gm, lg, out_dtype=dt), P * (3 + 3 * es + 3 * es)),
            ("chains_g_input (emit)", lambda: A.chains_g_input(crop, plans, gm, dtype=dt), P * (3 + 9 * es)),
            ("eager torch (reference expr)", lambda: (xs[0] * w[:, 0:1].to(dt) + xs[1] * w[:, 1:2].to(dt) + xs[2] * w[:, 2:3].to(dt)), P * (9 * es + 12 + 3 * es)),
        ]
        lib = A.load_library()
        from advmix_b200 import _lib
        gw = torch.empty((B, 3, H, W), device=dev)
        lut = A.transforms.normalize_lut(device=dev) if hasattr(A, "transforms") else None
        from advmix_b200 import transforms as TF
        lut = TF.normalize_lut(device=dev)
        from advmix_b200.mix import _xptrs
        cases += [
            ("mix_bwd materialised (through softmax)",
             lambda: _lib.check(lib.advmix_mix_bwd(_xptrs(xs), _lib.ptr(w), _lib.ptr(go), _lib.ptr(gw), 1, B, 3, 3, H, W, _lib.dtype_code(dt), _lib.stream_ptr())),
             P * (9 * es + 12 + 3 * es + 12)),
            ("chainmix_bwd FUSED (softmax recomputed)",
             lambda: _lib.check(lib.advmix_chainmix_bwd(_lib.ptr(crop), _lib.ptr(plans), _lib.ptr(gm), _lib.ptr(lut), _lib.ptr(lg), _lib.dtype_code(dt), 1,
                                                        _lib.ptr(go), _lib.dtype_code(dt), _lib.ptr(gw), B, H, W, _lib.stream_ptr())),
             P * (3 + 3 * es + 3 * es + 12)),
        ]
        for name, fn, bps in cases:
            us = timeit(fn)
            gbs = B * bps / us / 1e3
            rows.append({"B": B, "dtype": str(dt).replace("torch.", ""), "kernel": name, "us": us, "bytes_per_sample": bps,
                         "gbs": gbs, "frac_of_measured_hbm": gbs / PEAK, "samples_per_s": B / us * 1e6})
            print("B=%-3d %-8s %-42s %8.1f us  %7.0f B/sample-px*P  %6.0f GB/s  frac %.2f  %9.0f samples/s" % (
                B, rows[-1]["dtype"], name, us, bps, gbs, gbs / PEAK, B / us * 1e6))
if len(sys.argv) > 1:
    json.dump({"peak_gbs": PEAK, "note": "B=32 working sets fit the 126 MB L2 (numbers above the HBM peak are L2-resident); B=256 does not", "rows": rows},
              open(sys.argv[1], "w"), indent=1)
