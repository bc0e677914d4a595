"""Throughput of the K=3 `__getitem__` contract (chains clean / autoaug / gridmask, lib/dataset/JointsDataset.py:117-133)
through AdvMixBatchPipeline with the sources resident in HBM: crop (uint8 + normalised) -> autoaug chain -> gridmask chain ->
three heat-map targets.  Prints samples/s and the host time needed to issue one step."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import advmix_b200 as A
from advmix_b200.dataset import AdvMixBatchPipeline
dev = torch.device("cuda:0")
for B in (32, 256):
    rng = np.random.default_rng(bench.SEED)
    recs = bench.synth_records(B, rng); c, s, rot, flip = bench.synth_draws(recs, rng)
    for r in recs: r["width"], r["height"] = bench.SRC_W, bench.SRC_H
    sources = A.SourceBatch.from_tensor(bench.natural_images_torch(B, dev, bench.SEED))
    for K in (1, 3):
        pipe = AdvMixBatchPipeline(sample_times=K, is_train=True, device=dev)
        for _ in range(5): pipe(recs, sources=sources, draws=(c, s, rot, flip))
        torch.cuda.synchronize()
        n = 30
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        for _ in range(n): out = pipe(recs, sources=sources, draws=(c, s, rot, flip))
        b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        print("B=%3d K=%d: %.3f ms/step (host issue %.3f ms) = %.1f k samples/s" % (B, K, ms, (t1 - t0) / n * 1e3, B / ms))
