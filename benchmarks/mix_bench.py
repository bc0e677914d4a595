import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import advmix_b200 as A
dev = torch.device('cuda:0')
for B, dt in [(32, torch.float32), (256, torch.float32), (256, torch.bfloat16)]:
    xs = [torch.randn(B, 3, 256, 192, device=dev).to(dt) for _ in range(3)]
    logits = torch.randn(B, 3, 256, 192, device=dev)
    w = torch.softmax(logits, 1)
    for name, fn in [("mix(weights)", lambda: A.mix(xs, w)), ("mix_from_logits", lambda: A.mix_from_logits(xs, logits)),
                     ("eager torch", lambda: (xs[0] * w[:, 0:1] + xs[1] * w[:, 1:2] + xs[2] * w[:, 2:3]))]:
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): fn()
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 20 * 1e3
        es = 2 if dt == torch.bfloat16 else 4
        nbytes = B * 49152 * (3 * 3 * es + 3 * 4 + 3 * es + (12 if name == "mix_from_logits" else 0))
        print(B, dt, name, '%.1f us' % us, '%.0f GB/s' % (nbytes / us / 1e3), 'frac %.2f' % (nbytes / us / 1e3 / 6550.1))
