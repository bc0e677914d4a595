"""Secondary bench workloads (BASELINE.json configs[0] and configs[2]); same JSON contract as
bench.py's default line.  Used via `bench.py --workload coco_c|advmix_mix`."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W = 256, 192                     # COCO crops; `--workload mpii_c` switches to 256 x 256 (configs[4]) via set_size()
UNIT_BYTES = 2 * H * W * 3          # read u8 + write u8 per (image, corruption, severity)


def set_size(h, w):
    global H, W, UNIT_BYTES
    H, W, UNIT_BYTES = h, w, 2 * h * w * 3


def _peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def _time(fn, reps, torch):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run_jpeg_crop(args, rank, local_rank, world, dev, seed, metric, cfg_name):
    """the crop + targets step fed from ENCODED sources (what cv2.imread consumes)."""
    import time
    import cv2
    import torch
    import torch.distributed as dist
    import bench
    import advmix_b200 as A
    from advmix_b200 import jpeg as J
    from advmix_b200.dataset import AdvMixBatchPipeline
    peak, peak_src = _peak()
    B = args.batch
    rng = np.random.default_rng(seed + rank)
    recs = bench.synth_records(B, rng)
    c, s, rot, flip = bench.synth_draws(recs, rng)
    for r_ in recs:
        r_["width"], r_["height"] = bench.SRC_W, bench.SRC_H
    imgs = bench.natural_images_torch(B, dev, seed + rank).cpu().numpy()
    files = [cv2.imencode(".jpg", im, [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for im in imgs]
    enc = J.EncodedBatch(files)
    pipe = AdvMixBatchPipeline(sample_times=1, is_train=True, device=dev)
    out = torch.empty(B * bench.SRC_H * ((bench.SRC_W * 3 + 15) // 16 * 16) + 4096, dtype=torch.uint8, device=dev)
    pb = J.PlannedBatch(enc)
    files_d, plans_d = pb.to_device(dev)

    def dev_step():                                # encoded bytes resident in HBM
        sb = J.decode_planned(pb, files_d, plans_d, "bgr", out)
        return pipe(recs, sources=sb, draws=(c, s, rot, flip))
    tw_host = [torch.empty((B, 17, 1), dtype=torch.float32).pin_memory() for _ in range(2)]
    tw_done = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_run(n):                                # host parse + H2D of the files + decode + crop/targets + D2H (one step late)
        for i in range(n):
            p2 = J.PlannedBatch(enc)
            fd, pd = p2.to_device(dev)
            sb = J.decode_planned(p2, fd, pd, "bgr", out)
            _inp, _t, tw, _m = pipe(recs, sources=sb, draws=(c, s, rot, flip))
            tw_host[i & 1].copy_(tw, non_blocking=True)
            tw_done[i & 1].record()
            if i > 0:
                tw_done[(i - 1) & 1].synchronize()
        tw_done[(n - 1) & 1].synchronize()
    for _ in range(max(3, args.warmup)):
        dev_step()
    e2e_run(3)
    dec_ms = _time(lambda: J.decode_planned(pb, files_d, plans_d, "bgr", out), 10, torch)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = max(5, min(args.steps, 20))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        dev_step()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    a.record()
    e2e_run(steps)
    b.record()
    torch.cuda.synchronize()
    t2 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    ms, ms2 = float(t.item()), float(t2.item())
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:      # cv2.imdecode on one core, bounded sample (the crop path's CPU number is the default workload's)
        t0 = time.perf_counter()
        for f in files[:64]:
            cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION)
        cpu = {"value": 64 / (time.perf_counter() - t0), "unit": "images/s", "cores": 1, "kind": "reference",
               "sample": "cv2.imdecode (libjpeg-turbo) of 64 of the same files on one host core; decode only"}
    file_bytes = int(enc.nbytes)
    alg = file_bytes + pb.out_bytes                 # encoded in + decoded pixels out
    return {"metric": metric, "value": world * B * steps / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": cfg_name, "batch_per_gpu": B, "files": "640x480 4:2:0 baseline JPEG, quality 90, %.1f KB mean" % (file_bytes / B / 1e3),
                       "step": "device JPEG decode -> matrices -> crop || joints + heat maps"},
            "roofline": {"kernel": "advmix_jpeg_decode (Huffman + IDCT + colour)", "bound": "hbm", "achieved": alg / (dec_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": alg / (dec_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "us_per_launch": dec_ms * 1e3, "images_per_s_decode_only": B / (dec_ms * 1e-3),
                         "note": "entropy decoding is instruction-issue bound (bit-serial state machine per sub-sequence), not HBM bound"},
            "cpu_baseline": cpu,
            "e2e": {"value": world * B * steps / (ms2 * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": pb.h2d_bytes + B * 900,
                    "d2h_bytes_per_step": B * 17 * 4, "path": "pinned encoded files -> header parse on the host -> H2D of the files -> decode -> crop + targets -> D2H of target_weight one step late"},
            "gpu_launches": None, "impl": "advmix_b200"}


def _coco_c_e2e(args, rank, world, dev, seed, img, names, torch, dist):
    """tools/make_datasets.py process() end to end: uint8 images in pinned HOST memory -> H2D -> the 75 corruptions ->
    device JPEG encode (byte-identical to PIL's Image.save) -> the encoded files back in host memory.  Disk I/O excluded."""
    from advmix_b200 import corruptions as K, jpeg as J
    B = 256
    NV = 5 * len(names)
    CAP = (H * W * 7 // 8 + 1023) // 1024 * 1024      # bytes of every file copied back unconditionally (42 KB at 256x192; typical file: 11 KB)
    host = img[:B].cpu().pin_memory()
    x = torch.empty_like(img[:B])
    out = torch.empty_like(x)
    stride = (H * W * 3 // 2 + 4096 + 15) & ~15
    files_all = torch.empty((NV, B, stride), dtype=torch.uint8, device=dev)       # every file set stays on the device until
    lengths_all = torch.empty((NV, B), dtype=torch.int32, device=dev)             # the lengths have been checked
    files_h = torch.empty((NV, B, CAP), dtype=torch.uint8).pin_memory()
    lengths_h = torch.empty((NV, B), dtype=torch.int32).pin_memory()
    copy_stream = torch.cuda.Stream()
    d2h = [0]

    def one_pass():
        # compute stream: H2D, corrupt, encode.  copy stream: the first CAP bytes of every file, as soon as a set is encoded.
        # One synchronisation per pass; files longer than CAP (none at this size) are fetched afterwards.
        x.copy_(host, non_blocking=True)
        j = 0
        for n in names:
            for s in range(1, 6):
                K.corrupt_batch(x, n, s, seed=seed, sample_base=rank * B, out=out, fast=True)
                J.encode_batch_device(out, out=(files_all[j], lengths_all[j]))
                ev = torch.cuda.Event()
                ev.record()
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ev)
                    files_h[j].copy_(files_all[j, :, :CAP], non_blocking=True)
                j += 1
        lengths_h.copy_(lengths_all, non_blocking=True)
        torch.cuda.current_stream().wait_stream(copy_stream)
        torch.cuda.synchronize()
        ln = lengths_h.numpy()
        assert ln.min() > 0, "a file did not fit its device buffer"
        tails = 0
        for (jj, ii) in zip(*np.nonzero(ln > CAP)):
            tails += int(files_all[jj, ii, CAP:int(ln[jj, ii])].cpu().numel())
        d2h[0] += files_h.numel() + lengths_h.numel() * 4 + tails
        return ln
    one_pass()
    d2h[0] = 0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = max(1, min(args.steps, 3))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        ln = one_pass()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"value": world * B * NV * steps / (ms * 1e-3), "unit": "outputs/s", "h2d_bytes_per_step": int(host.numel()),
            "d2h_bytes_per_step": int(d2h[0] // steps), "images_per_step": B, "mean_file_bytes": float(ln.mean()),
            "max_file_bytes": int(ln.max()),
            "path": "pinned uint8 images -> corrupt_batch x75 -> jpeg.encode_batch_device -> first %d KB of every file copied to " % (CAP // 1024) + ""
                    "pinned host memory on a second stream (longer files fetched after the length check); one sync per pass"}


def _cpu_one_image(arg):
    """One image through the oracle port of make_datasets.py process(): 15 x 5 corrupt() calls + PIL save to memory."""
    import io
    import cv2
    from PIL import Image
    from oracle import corruptions as OK
    cv2.setNumThreads(1)
    img, names, seed = arg
    H, W = img.shape[0], img.shape[1]                    # (worker processes do not see set_size())
    rng = np.random.default_rng(seed)
    bank = OK.synthetic_frost_bank(n=5, fh=H + 64, fw=W + 64)
    nbytes = 0
    for name in names:
        for sev in range(1, 6):
            d = OK.make_draws(name, sev, H, W, rng, bank.shape)
            o = OK.corrupt_with_draws(img, sev, name, d, bank)
            b = io.BytesIO()
            Image.fromarray(o).save(b, "JPEG")
            nbytes += b.tell()
    return nbytes


def _coco_c_cpu_baseline(imgs, names, per_core=2):
    """Reported baseline: the oracle port (kind "port": imagecorruptions itself is not installable offline) on all host
    cores, one worker process per core like the DataLoader workers of make_datasets.py:53-56; bounded sample."""
    import multiprocessing as mp
    import time
    cores = os.cpu_count() or 1
    n = min(len(imgs), per_core * cores)
    work = [(imgs[i], list(names), 1000 + i) for i in range(n)]
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_one_image, work[:cores])                 # warm: imports, numba / table set-up
        t0 = time.perf_counter()
        pool.map(_cpu_one_image, work)
        dt = time.perf_counter() - t0
    calls = 5 * len(names)
    return {"value": n * calls / dt, "unit": "outputs/s", "cores": cores, "kind": "port",
            "sample": "%d images x %d (corruption, severity) calls + PIL JPEG save to memory, %d worker processes, %.1f s" % (n, calls, cores, dt)}


def run(args, rank, local_rank, world, dev, seed, metric, cfg_name):
    import torch
    import torch.distributed as dist
    import advmix_b200 as A
    from advmix_b200 import corruptions as K
    from advmix_b200.dataset import corruption_chains
    if args.workload == "jpeg_crop":
        return run_jpeg_crop(args, rank, local_rank, world, dev, seed, metric, cfg_name)
    peak, peak_src = _peak()
    g = torch.Generator(device=dev).manual_seed(seed + rank)
    names = A.get_corruption_names("common")
    if args.workload == "mpii_c":
        set_size(256, 256)
    if args.workload in ("coco_c", "mpii_c"):
        N = 1024                                          # 151 MB in + 151 MB out per op > L2
        low = torch.rand((N, 3, H // 16 + 2, W // 16 + 2), device=dev, generator=g)
        img = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False) * 255
        img = (img + torch.randint(-8, 9, img.shape, device=dev, generator=g)).clamp_(0, 255).to(torch.uint8)
        img = img.permute(0, 2, 3, 1).contiguous()
        out = torch.empty_like(img)
        K.set_frost_bank(K.default_frost_bank(384, 384), dev)

        def sweep():
            for n in names:
                for s in range(1, 6):
                    K.corrupt_batch(img, n, s, seed=seed, sample_base=rank * N, out=out, fast=True)
        for _ in range(max(1, min(args.warmup, 3))):
            sweep()
        per_op = {}
        for n in names:
            for s in range(1, 6):
                ms = _time(lambda: K.corrupt_batch(img, n, s, seed=seed, sample_base=rank * N, out=out, fast=True), 3, torch)
                per_op["%s/%d" % (n, s)] = {"us_per_image": ms * 1e3 / N, "gbs": N * UNIT_BYTES / (ms * 1e-3) / 1e9,
                                            "frac": N * UNIT_BYTES / (ms * 1e-3) / 1e9 / peak}
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        steps = max(1, min(args.steps, 5))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            sweep()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        units = world * N * 75 * steps
        e2e = _coco_c_e2e(args, rank, world, dev, seed, img, names, torch, dist)
        cpu = _coco_c_cpu_baseline(img[:128].cpu().numpy(), names, per_core=6) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
        slow = max(per_op, key=lambda k: per_op[k]["us_per_image"])
        by_op = {}
        for k, v in per_op.items():
            by_op.setdefault(k.split("/")[0], []).append(v["frac"])
        return {"metric": "%s corrupted %dx%d outputs/sec" % ("COCO-C" if args.workload == "coco_c" else "MPII-C", H, W), "value": units / (ms * 1e-3), "unit": "outputs/s",
                "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": cfg_name, "images_per_gpu": N, "units_per_step": N * 75,
                           "l2": "%d MB in+out per op > 126 MB L2" % (N * UNIT_BYTES // 1000000), "random_draws": "in-register Philox (perf mode)", "arithmetic": "ADVMIX_CORRUPT_FAST: float32 kernels for gaussian_noise / contrast (<=1 LSB), every other op in the reference's float64/float32 order"},
                "roofline": {"kernel": "whole sweep (75 op x severity calls)", "bound": "hbm",
                             "achieved": units / world * UNIT_BYTES / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": units / world * UNIT_BYTES / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                             "peak_source": peak_src, "slowest": slow,
                             "mean_frac_by_op": {k: float(np.mean(v)) for k, v in by_op.items()}},
                "per_op": per_op, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": None, "impl": "advmix_b200"}

    # advmix_mix: configs[2]
    B = 32 if args.batch == 256 else args.batch
    crop = torch.randint(0, 256, (B, H, W, 3), device=dev, dtype=torch.uint8, generator=g)
    logits = torch.randn((B, 3, H, W), device=dev, generator=g)
    nm1 = [names[(2 * b) % 15] for b in range(B)]; sv1 = [1 + b % 5 for b in range(B)]
    nm2 = [names[(2 * b + 1) % 15] for b in range(B)]; sv2 = [1 + (b + 2) % 5 for b in range(B)]

    def step():
        clean = A.to_tensor_normalize(crop)
        _, x1 = corruption_chains(crop, nm1, sv1, seed=seed, sample_base=rank * B)
        _, x2 = corruption_chains(crop, nm2, sv2, seed=seed + 1, sample_base=rank * B)
        return A.mix_from_logits([clean, x1, x2], logits)
    for _ in range(max(3, args.warmup)):
        step()
    xs = [A.to_tensor_normalize(crop) for _ in range(3)]
    mix_ms = _time(lambda: A.mix_from_logits(xs, logits), 50, torch)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    mix_bytes = B * (5 * 3 * H * W * 4)          # 3 chains + logits in, mix out (fp32): 2 949 120 B/sample
    return {"metric": metric, "value": world * B * args.steps / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg_name, "batch_per_gpu": B, "chains": "clean + 2 chains drawn round-robin from the 15x5 set"},
            "roofline": {"kernel": "mix_fwd_kernel<float,3>", "bound": "hbm", "achieved": mix_bytes / (mix_ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": mix_bytes / (mix_ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "us_per_launch": mix_ms * 1e3,
                         "note": "B=32 working set (94 MB) fits L2; see DESIGN.md"},
            "cpu_baseline": None, "e2e": None, "gpu_launches": None, "impl": "advmix_b200"}
