"""Timing of the device JPEG encoder (tools/make_datasets.py:45 replacement) next to PIL on one host core."""
import io, sys, time
import numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from advmix_b200 import jpeg as J
sys.path.insert(0, sys.path[0] + "/tests"); from test_gpu_jpeg import natural
rng = np.random.default_rng(0)
for (H, W, n) in ((256, 192, 512), (480, 640, 128)):
    imgs = np.stack([natural(rng, H, W) for _ in range(16)])
    x = torch.from_numpy(imgs).cuda().repeat(n // 16, 1, 1, 1).contiguous()
    for _ in range(3):
        f, l = J.encode_batch_device(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f, l = J.encode_batch_device(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    t0 = time.perf_counter()
    from PIL import Image
    for i in range(16):
        b = io.BytesIO(); Image.fromarray(imgs[i]).save(b, "JPEG")
    cpu = (time.perf_counter() - t0) / 16
    print("encode %dx%d n=%d: %.3f ms/batch = %.1f k images/s, mean file %d B; PIL one core %.2f k images/s" %
          (W, H, n, ms, n / ms, int(l.float().mean()), 1e-3 / cpu))
