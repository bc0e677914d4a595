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


def _coco_c_e2e(args, rank, world, dev, seed, img, names, torch, dist, fast=True):
    """tools/make_datasets.py process() end to end: uint8 images in pinned HOST memory -> H2D -> the 75 corruptions ->
    device JPEG encode (byte-identical to PIL's Image.save) -> the encoded files back in host memory.  Disk I/O excluded."""
    from advmix_b200 import corruptions as K, jpeg as J
    B = 256
    NV = 5 * len(names)
    CAP = (H * W * 7 // 8 + 1023) // 1024 * 1024      # bytes of every file copied back unconditionally (42 KB at 256x192; typical file: 11 KB)
    host = img[:B].cpu().pin_memory()
    x = torch.empty_like(img[:B])
    out5 = torch.empty((5,) + tuple(x.shape), dtype=torch.uint8, device=dev)
    stride = (H * W * 3 // 2 + 4096 + 15) & ~15
    files = torch.empty((B, stride), dtype=torch.uint8, device=dev)
    lengths = torch.empty(B, dtype=torch.int32, device=dev)
    packed_all = torch.empty((NV, B * CAP), dtype=torch.uint8, device=dev)        # set j packed back to back (advmix_pack_files); stays on the
    offsets_all = torch.empty((NV, B + 1), dtype=torch.int64, device=dev)         # device until its byte count has reached the host
    files_h = torch.empty((NV, B * CAP), dtype=torch.uint8).pin_memory()
    offs_h = torch.empty((NV, B + 1), dtype=torch.int64).pin_memory()
    copy_stream = torch.cuda.Stream()
    ev_set = [torch.cuda.Event() for _ in range(NV)]
    ev_off = [torch.cuda.Event() for _ in range(NV)]
    d2h = [0]
    LAG = 3

    def fetch(k):
        # the offsets of set k reached the host LAG sets ago: copy exactly the encoded bytes
        ev_off[k].synchronize()
        total = int(offs_h[k, B])
        assert 0 < total <= B * CAP, "a file set did not fit its device buffer"
        with torch.cuda.stream(copy_stream):
            files_h[k, :total].copy_(packed_all[k, :total], non_blocking=True)
        d2h[0] += total + (B + 1) * 8

    def one_pass():
        # compute stream: H2D, corrupt (five severities per call), encode, pack.  copy stream: the offsets of every set, then - once they
        # are on the host - exactly the encoded bytes of the set.  One synchronisation per pass.
        x.copy_(host, non_blocking=True)
        j = 0
        for n in names:
            K.corrupt_sweep(x, n, seed=seed, sample_base=rank * B, out=out5, fast=fast)      # the five severities from one read
            for s in range(5):
                J.encode_batch_device(out5[s], out=(files, lengths))
                J.pack_files(files, lengths, out=(packed_all[j], offsets_all[j]))
                ev_set[j].record()
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ev_set[j])
                    offs_h[j].copy_(offsets_all[j], non_blocking=True)
                    ev_off[j].record(copy_stream)
                if j >= LAG:
                    fetch(j - LAG)
                j += 1
        for k in range(max(NV - LAG, 0), NV):
            fetch(k)
        torch.cuda.current_stream().wait_stream(copy_stream)
        torch.cuda.synchronize()
        ln = np.diff(offs_h.numpy(), axis=1)
        assert ln.min() > 0, "a file did not fit its device buffer"
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
            "path": "pinned uint8 images -> corrupt_sweep x15 (five severities per call) -> jpeg.encode_batch_device x75 -> jpeg.pack_files (files back to back on the device) -> "
                    "offsets, then exactly the encoded bytes of every set copied to pinned host memory on a second stream; one sync per pass"}


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


def _natural_crops(N, dev, g, torch):
    low = torch.rand((N, 3, H // 16 + 2, W // 16 + 2), device=dev, generator=g)
    img = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=False) * 255
    img = (img + torch.randint(-8, 9, img.shape, device=dev, generator=g)).clamp_(0, 255).to(torch.uint8)
    return img.permute(0, 2, 3, 1).contiguous()


def coco_c_record(args, rank, world, dev, seed, cfg_name, fast, with_e2e, with_cpu, sampler_cls=None, local_rank=0, N=1024):
    """configs[0] / configs[4]: the 15 x 5 sweep of tools/make_datasets.py:38-45 over N resident crops.  `fast` selects
    ADVMIX_CORRUPT_FAST (float32 / fixed-point kernels, <= 1 LSB) or the reference's float64 operation order (bit-exact)."""
    import torch
    import torch.distributed as dist
    import advmix_b200 as A
    from advmix_b200 import corruptions as K
    peak, peak_src = _peak()
    g = torch.Generator(device=dev).manual_seed(seed + rank)
    names = A.get_corruption_names("common")
    img = _natural_crops(N, dev, g, torch)                 # 151 MB in + 151 MB out per op > L2
    out = torch.empty_like(img)
    out5 = torch.empty((5,) + tuple(img.shape), dtype=torch.uint8, device=dev)
    K.set_frost_bank(K.default_frost_bank(384, 384), dev)

    def sweep():
        # make_datasets.py:38-45: severities innermost over the same image -> one advmix_corrupt_sweep_u8c3 call per corruption
        for n in names:
            K.corrupt_sweep(img, n, seed=seed, sample_base=rank * N, out=out5, fast=fast)
    for _ in range(max(1, min(args.warmup, 2))):
        sweep()
    per_op = {}
    for n in names:
        for s in range(1, 6):
            ms = _time(lambda: K.corrupt_batch(img, n, s, seed=seed, sample_base=rank * N, out=out, fast=fast), 2, torch)
            per_op["%s/%d" % (n, s)] = {"us_per_image": ms * 1e3 / N, "frac": N * UNIT_BYTES / (ms * 1e-3) / 1e9 / peak}
    sampler = sampler_cls(local_rank) if (sampler_cls and rank == 0) else None
    if sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = max(1, min(args.steps, 4))
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
    e2e = _coco_c_e2e(args, rank, world, dev, seed, img, names, torch, dist, fast) if with_e2e else None
    clocks = sampler.stop() if sampler else None
    cpu = _coco_c_cpu_baseline(img[:128].cpu().numpy(), names, per_core=4) if (with_cpu and rank == 0) else None
    per_sweep = {n: _time(lambda: K.corrupt_sweep(img, n, seed=seed, sample_base=rank * N, out=out5, fast=fast), 2, torch) * 1e3 / N for n in names}
    slow = max(per_op, key=lambda k: per_op[k]["us_per_image"])
    by_op = {}
    for k, v in per_op.items():
        by_op.setdefault(k.split("/")[0], []).append(v)
    arith = ("ADVMIX_CORRUPT_FAST: float32 / 24-bit fixed-point kernels for the noise, contrast and stencil ops (<= 1 LSB, < 0.2 % of values); "
             "shot / impulse / pixelate / jpeg / brightness / frost unchanged") if fast else "the reference's float64 / float32 operation order (bit-exact against the oracle)"
    arith += "; five severities per call (advmix_corrupt_sweep_u8c3: bit-identical to the per-severity calls)"
    return {"metric": "%s corrupted %dx%d outputs/sec" % ("MPII-C" if H == W else "COCO-C", H, W), "value": units / (ms * 1e-3), "unit": "outputs/s",
            "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": cfg_name, "images_per_gpu": N, "units_per_step": N * 75, "arithmetic": arith,
                       "l2": "%d MB in+out per op > 126 MB L2" % (N * UNIT_BYTES // 1000000), "random_draws": "in-register Philox (perf mode)"},
            "roofline": {"kernel": "whole sweep (15 advmix_corrupt_sweep_u8c3 calls = 75 op x severity outputs)", "bound": "hbm",
                         "achieved": units / world * UNIT_BYTES / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": units / world * UNIT_BYTES / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "slowest": slow, "algorithmic_bytes_per_unit": UNIT_BYTES,
                         "fused_accounting": {"bytes_per_image": (1 + 5) * 15 * (UNIT_BYTES // 2), "note": "crop read once per corruption + five writes (SURVEY 8d: one read per sweep would be 11.2 MB / image)",
                                              "frac": units / world / 75 * (1 + 5) * 15 * (UNIT_BYTES // 2) / (ms * 1e-3) / 1e9 / peak},
                         "us_per_image_five_severities_one_call_by_op": per_sweep,
                         "us_per_image_sum_over_severities_by_op": {k: float(np.sum([x["us_per_image"] for x in v])) for k, v in by_op.items()},
                         "mean_frac_by_op": {k: float(np.mean([x["frac"] for x in v])) for k, v in by_op.items()}},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": None, "clocks": clocks, "impl": "advmix_b200"}


def advmix_mix_record(args, rank, world, dev, seed, cfg_name, sampler_cls=None, local_rank=0, B=32, dtype_name="float32"):
    """configs[2], the AdvMix inner loop around the pose network (lib/core/function.py:137-146 forward, :158-164 backward) on the
    reference's per-GPU batch: uint8 crops -> per-image autoaug plans -> G_input = cat(clean, autoaug, gridmask) for the generator
    -> [generator: out of scope, its logits are a resident tensor] -> fused chain + mix forward (softmax inside) -> [pose network:
    out of scope, its input gradient is a resident tensor] -> fused backward to the logits.  Four launches + the plan's two,
    captured in one CUDA graph; no chain tensor is materialised for the mix."""
    import torch
    import torch.distributed as dist
    import advmix_b200 as A
    from advmix_b200 import _lib, chains as CH, transforms as TF
    peak, peak_src = _peak()
    dt = getattr(torch, dtype_name)
    es = 2 if dt == torch.bfloat16 else 4
    g = torch.Generator(device=dev).manual_seed(seed + 7 * rank)
    rng = np.random.default_rng(seed + rank)
    lib = A.load_library()
    P, S = _lib.ptr, _lib.stream_ptr
    NB = 8                                                  # batches cycled so that the working set (8 x 33 MB) exceeds L2
    crops = [_natural_crops(B, dev, g, torch) for _ in range(NB)]
    ops_mags = [CH.sample_autoaug_batch(B, rng) for _ in range(NB)]
    ops_d = [torch.as_tensor(o).to(dev, torch.int32).contiguous() for o, _ in ops_mags]
    mags_d = [torch.as_tensor(m).to(dev, torch.float32).contiguous() for _, m in ops_mags]
    gms = [torch.as_tensor(CH.sample_gridmask_batch(B, H, W, rng)).to(dev, torch.int32).contiguous() for _ in range(NB)]
    logits = [torch.randn((B, 3, H, W), device=dev, generator=g).to(dt) for _ in range(NB)]
    gouts = [torch.randn((B, 3, H, W), device=dev, generator=g).to(dt) for _ in range(NB)]
    lut = TF.normalize_lut(device=dev)
    plan_bytes = int(lib.advmix_autoaug_plan_bytes(1))
    plans = torch.empty((B, plan_bytes), dtype=torch.uint8, device=dev)
    ws = torch.empty(B * 768 * 4, dtype=torch.uint8, device=dev)
    g_in = torch.empty((B, 9, H, W), dtype=dt, device=dev)
    tmp = torch.empty((B, 3, H, W), dtype=dt, device=dev)
    gw = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
    dc = _lib.dtype_code(dt)

    def k_plan(j):
        _lib.check(lib.advmix_autoaug_plan_u8c3(P(crops[j]), P(ops_d[j]), P(mags_d[j]), P(plans), B, H, W, P(ws), ws.numel(), S()))

    def k_emit(j):
        _lib.check(lib.advmix_chains_emit_u8c3(P(crops[j]), P(plans), P(gms[j]), P(lut), P(g_in), B, H, W, dc, S()))

    def k_fwd(j):
        _lib.check(lib.advmix_chainmix_fwd(P(crops[j]), P(plans), P(gms[j]), P(lut), P(logits[j]), dc, 1, P(tmp), dc, None, B, H, W, S()))

    def k_bwd(j):
        _lib.check(lib.advmix_chainmix_bwd(P(crops[j]), P(plans), P(gms[j]), P(lut), P(logits[j]), dc, 1, P(gouts[j]), dc, P(gw), B, H, W, S()))

    def step(j):
        k_plan(j); k_emit(j); k_fwd(j); k_bwd(j)
    for j in range(NB):
        step(j)
    torch.cuda.synchronize()
    graphs = []
    for j in range(NB):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            step(j)
        graphs.append(gr)
    for gr in graphs:
        gr.replay()
    torch.cuda.synchronize()
    # per-kernel device time: NB launches of one kernel (one per batch) captured in a graph, so the ~15 us a ctypes call costs
    # on the host does not hide a 10 us kernel
    per_kernel = {}
    for name, fn in (("autoaug_plan", k_plan), ("chains_emit (G_input)", k_emit), ("chainmix_fwd", k_fwd), ("chainmix_bwd", k_bwd)):
        gk = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gk):
            for j in range(NB):
                fn(j)
        per_kernel[name] = _time(gk.replay, 10, torch) / NB * 1e3       # us per launch
    sampler = sampler_cls(local_rank) if (sampler_cls and rank == 0) else None
    if sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = max(NB, args.steps // NB * NB)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        graphs[i % NB].replay()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    # e2e: the whole K = 3 data path through the public API, from the rank's pinned host shard: HBM source cache lookup (misses cross
    # PCIe), host draws, ONE pinned parameter buffer + ONE library call (fastpath.AdvMixStep: crop, autoaug plans, targets of the
    # clean and gridmask chains), G_input for the generator, fused mix forward on the generator's logits (resident: the network is
    # out of scope), backward with the pose network's input gradient (resident), scalar read-back.  Steady state (epochs >= 2).
    import bench
    from advmix_b200 import fastpath as FP
    D = 8 * B
    host_all = torch.empty((D, bench.SRC_H, bench.SRC_W, 3), dtype=torch.uint8, pin_memory=True)
    for k in range(0, D, B):
        host_all[k:k + B].copy_(bench.natural_images_torch(B, dev, seed + 13 * rank + k))
    torch.cuda.synchronize()
    recs_all = bench.synth_records(D, rng)
    table = FP.RecordTable.from_records(recs_all, widths=np.full(D, bench.SRC_W), heights=np.full(D, bench.SRC_H))
    cache = FP.SourceCache(D * bench.SRC_H * bench.SRC_W * 3 + D * 256, D, dev)
    astep = FP.AdvMixStep(B, device=dev, seed=seed + rank, want_gridmask_targets=True, ring=4, out_ring=4, graph=True, prefetch_streams=2)
    res_h = torch.empty(1, dtype=torch.float32).pin_memory()
    perm_e = rng.permutation(D)

    def e2e_step(i):
        ids = perm_e[(i * B) % D:(i * B) % D + B]
        off, pitch, hh, ww = cache.ensure(ids, lambda k_: host_all[k_])
        batch = astep(table, ids, cache.buffer, off, pitch, hh, ww, after=cache.take_upload_event())
        gi = batch.g_input(dt)
        lg = logits[i % NB].detach().requires_grad_(True)
        o = batch.mix(lg, out_dtype=dt)
        o.backward(gouts[i % NB])
        res_h.copy_(o.detach()[0, 0, 0, :1].float(), non_blocking=True)
        return gi, lg.grad
    for j in range(D // B + 2):                           # first epoch fills the cache
        e2e_step(j)
    torch.cuda.synchronize()
    e2e_steps = 4 * (D // B)
    up0 = cache.uploaded_bytes
    a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a2.record()
    for i in range(e2e_steps):
        e2e_step(i)
    b2.record()
    torch.cuda.synchronize()
    _ = float(res_h[0])
    e2e_h2d = (cache.uploaded_bytes - up0) // e2e_steps + astep.nbytes
    t2 = torch.tensor([a2.elapsed_time(b2)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler else None
    ms, ms2 = float(t.item()), float(t2.item())
    px = H * W
    fwd_bytes = B * px * (3 + 3 * es + 3 * es)              # u8 crop + logits in, tmp out
    bwd_bytes = B * px * (3 + 3 * es + 3 * es + 12)         # u8 crop + logits + grad_out in, fp32 grad_logits out
    emit_bytes = B * px * (3 + 9 * es)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = _mix_cpu_baseline(crops[0].cpu().numpy(), B)
    return {"metric": "AdvMix inner-loop mix steps: samples/sec (chains + G_input + mix forward + backward)", "value": world * B * steps / (ms * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": NB, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if es == 4 else "bf16", "data": "synthetic",
            "config": {"workload": cfg_name, "batch_per_gpu": B, "chains": "clean + autoaug + gridmask, recomputed from the uint8 crop inside the kernels (JointsDataset.py:124-131)",
                       "io_dtype": dtype_name, "cuda_graph": True, "l2": "%d batches cycled: %d MB per cycle > 126 MB L2" % (NB, NB * (fwd_bytes + bwd_bytes + emit_bytes) // 1000000)},
            "roofline": {"kernel": "chain_kernel<FWD> (advmix_chainmix_fwd)", "bound": "hbm", "achieved": fwd_bytes / (per_kernel["chainmix_fwd"] * 1e-6) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": fwd_bytes / (per_kernel["chainmix_fwd"] * 1e-6) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_sample": {"fwd": fwd_bytes // B, "bwd": bwd_bytes // B, "g_input": emit_bytes // B, "reference_materialised_fwd": px * 5 * 3 * 4},
                         "per_kernel_us": per_kernel,
                         "frac_by_kernel": {"chainmix_fwd": fwd_bytes / (per_kernel["chainmix_fwd"] * 1e-6) / 1e9 / peak,
                                            "chainmix_bwd": bwd_bytes / (per_kernel["chainmix_bwd"] * 1e-6) / 1e9 / peak,
                                            "chains_emit": emit_bytes / (per_kernel["chains_emit (G_input)"] * 1e-6) / 1e9 / peak},
                         "step_frac": (fwd_bytes + bwd_bytes + emit_bytes) / (ms / steps * 1e-3) / 1e9 / peak},
            "cpu_baseline": cpu,
            "e2e": {"value": world * B * e2e_steps / (ms2 * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(e2e_h2d),
                    "d2h_bytes_per_step": 4, "steps": e2e_steps, "shard_images_per_rank": D,
                    "path": "pinned host shard -> HBM source cache -> fastpath.AdvMixStep (ring=4, CUDA-graph replay on two prefetch streams; one step call: 256x192 crop from 640x480 sources, autoaug plans, targets "
                            "of the clean + gridmask chains) -> batch.g_input() -> batch.mix(logits) (autograd) + backward -> scalar read-back; steady state"},
            "gpu_launches": 6 * steps, "clocks": clocks, "impl": "advmix_b200"}


def _mix_cpu_one(arg):
    """The reference's CPU share of one sample of the K=3 step (JointsDataset.get_var x3: PIL autoaug, grid_aug, ToTensor/Normalize),
    oracle/ port, followed by the function.py:138-144 expression in torch on the CPU."""
    import torch
    from oracle import affine as OA, chains as OC
    crop, seed = arg
    rng = np.random.default_rng(seed)
    lut = OA.normalize_lut()
    clean = OA.to_tensor_normalize(crop, lut)
    aa = OC.autoaug(crop, int(rng.integers(0, 12)), float(rng.random()), float(rng.random()), use_pil=True)
    x1 = OA.to_tensor_normalize(np.asarray(aa), lut)
    hh, ww = crop.shape[:2]
    d = int(rng.integers(2, min(hh, ww)))
    x2, _ = OC.gridmask(clean.copy(), np.zeros((17, 3)), np.ones((17, 3)), True, d, int(rng.integers(d)), int(rng.integers(d)))
    xs = [torch.from_numpy(np.ascontiguousarray(v)) for v in (clean, x1, x2)]
    w = torch.softmax(torch.randn(3, hh, ww), 0)
    tmp = xs[0] * w[0:1] + xs[1] * w[1:2] + xs[2] * w[2:3]
    return float(tmp[0, 0, 0])


def _mix_cpu_baseline(crops, B):
    import multiprocessing as mp
    import time
    cores = os.cpu_count() or 1
    n = 16 * cores
    work = [(crops[i % len(crops)], 100 + i) for i in range(n)]
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_mix_cpu_one, work[:cores])
        t0 = time.perf_counter()
        pool.map(_mix_cpu_one, work)
        dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": "%d samples: oracle/ port of the three chains (PIL autoaug, grid_aug, ToTensor/Normalize) + the function.py:138-144 "
                      "softmax / mix expression in torch on the CPU, %d worker processes, %.1f s (forward only)" % (n, cores, dt)}


def bottomup512_record(args, rank, world, dev, seed, cfg_name, sampler_cls=None, local_rank=0, B=32):
    """configs[3]: HrHRNet-W32 512x512 input, AdvMix mix + multi-resolution heat-map targets (128x128 and 256x256, 17 joints).
    One step on resident 640x480 sources: matrices -> 512x512 uint8 crop -> {joints -> heat maps at both resolutions} ||
    {autoaug plans -> G_input -> fused chain + mix forward -> fused backward}; captured in one CUDA graph per batch."""
    import torch
    import torch.distributed as dist
    import advmix_b200 as A
    import bench
    from advmix_b200 import _lib, chains as CH, transforms as TF, targets as TG
    peak, peak_src = _peak()
    S = 512
    lib = A.load_library()
    P, St = _lib.ptr, _lib.stream_ptr
    rng = np.random.default_rng(seed + rank)
    g = torch.Generator(device=dev).manual_seed(seed + 3 * rank)
    NB = 4
    Jn = 17
    lut = TF.normalize_lut(device=dev)
    gtab = TG.gaussian_table(2, dev)
    perm = TF.flip_perm(Jn, [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]], dev)
    sets = []
    for j in range(NB):
        recs = bench.synth_records(B, rng)
        c, s, rot, flip = bench.synth_draws(recs, rng)
        s = s * np.array([1.0, 0.75])[None, :]                       # square 512x512 output: scale w == h in pixels (bottom-up keeps the whole person box)
        images = bench.natural_images_torch(B, dev, seed + 31 * rank + j)
        src = A.SourceBatch.from_tensor(images)
        ops, mags = CH.sample_autoaug_batch(B, rng)
        sets.append(dict(src=src, c=torch.from_numpy(c).to(dev), s=torch.from_numpy(s).to(dev), r=torch.from_numpy(rot).to(dev),
                         f=torch.from_numpy(flip.astype(np.uint8)).to(dev),
                         jin=torch.from_numpy(np.stack([r_["joints_3d"] for r_ in recs])).to(dev),
                         vin=torch.from_numpy(np.stack([r_["joints_3d_vis"] for r_ in recs])).to(dev),
                         ops=torch.as_tensor(ops).to(dev, torch.int32), mags=torch.as_tensor(mags).to(dev, torch.float32),
                         gm=torch.as_tensor(CH.sample_gridmask_batch(B, S, S, rng)).to(dev, torch.int32),
                         logits=torch.randn((B, 3, S, S), device=dev, generator=g), gout=torch.randn((B, 3, S, S), device=dev, generator=g)))
    M = torch.empty((B, 2, 3), dtype=torch.float64, device=dev)
    crop = torch.empty((B, S, S, 3), dtype=torch.uint8, device=dev)
    jo = torch.empty((B, Jn, 3), dtype=torch.float64, device=dev); vo = torch.empty_like(jo)
    hm1 = torch.empty((B, Jn, 128, 128), dtype=torch.float32, device=dev)
    hm2 = torch.empty((B, Jn, 256, 256), dtype=torch.float32, device=dev)
    mu = torch.empty((B, Jn, 2), dtype=torch.float32, device=dev); tw = torch.empty((B, Jn, 1), dtype=torch.float32, device=dev)
    plans = torch.empty((B, int(lib.advmix_autoaug_plan_bytes(1))), dtype=torch.uint8, device=dev)
    ws = torch.empty(B * 768 * 4, dtype=torch.uint8, device=dev)
    g_in = torch.empty((B, 9, S, S), dtype=torch.float32, device=dev)
    tmp = torch.empty((B, 3, S, S), dtype=torch.float32, device=dev)
    gw = torch.empty((B, 3, S, S), dtype=torch.float32, device=dev)
    side = torch.cuda.Stream(device=dev)

    def kernels(d):
        sb = d["src"]
        return [
            ("affine_matrices", lambda: _lib.check(lib.advmix_affine_matrices(P(d["c"]), P(d["s"]), 0, P(d["r"]), P(M), B, S, S, St()))),
            ("warp_affine 512x512 u8", lambda: _lib.check(lib.advmix_warp_affine_u8c3(P(sb.buffer), P(sb.offsets), P(sb.heights), P(sb.widths), P(sb.pitches), P(d["f"]), P(M),
                                                                                   P(crop), None, P(lut), B, S, S, _lib.F32, St()))),
            ("joints_flip_affine", lambda: _lib.check(lib.advmix_joints_flip_affine(P(d["jin"]), P(d["vin"]), P(d["f"]), P(sb.widths), P(perm), P(M), P(jo), P(vo), B, Jn, St()))),
            ("heatmap_targets 128x128", lambda: _lib.check(lib.advmix_heatmap_targets(P(jo), P(vo), P(gtab), None, P(hm1), P(mu), P(tw), B, Jn, 128, 128, S, S, 2, St()))),
            ("heatmap_targets 256x256", lambda: _lib.check(lib.advmix_heatmap_targets(P(jo), P(vo), P(gtab), None, P(hm2), P(mu), P(tw), B, Jn, 256, 256, S, S, 2, St()))),
            ("autoaug_plan", lambda: _lib.check(lib.advmix_autoaug_plan_u8c3(P(crop), P(d["ops"]), P(d["mags"]), P(plans), B, S, S, P(ws), ws.numel(), St()))),
            ("chains_emit (G_input)", lambda: _lib.check(lib.advmix_chains_emit_u8c3(P(crop), P(plans), P(d["gm"]), P(lut), P(g_in), B, S, S, _lib.F32, St()))),
            ("chainmix_fwd", lambda: _lib.check(lib.advmix_chainmix_fwd(P(crop), P(plans), P(d["gm"]), P(lut), P(d["logits"]), _lib.F32, 1, P(tmp), _lib.F32, None, B, S, S, St()))),
            ("chainmix_bwd", lambda: _lib.check(lib.advmix_chainmix_bwd(P(crop), P(plans), P(d["gm"]), P(lut), P(d["logits"]), _lib.F32, 1, P(d["gout"]), _lib.F32, P(gw), B, S, S, St()))),
        ]

    def step(d):
        ks = dict(kernels(d))
        ks["affine_matrices"]()
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ks["joints_flip_affine"](); ks["heatmap_targets 128x128"](); ks["heatmap_targets 256x256"]()
        ks["warp_affine 512x512 u8"](); ks["autoaug_plan"](); ks["chains_emit (G_input)"](); ks["chainmix_fwd"](); ks["chainmix_bwd"]()
        main.wait_stream(side)
    for d in sets:
        step(d)
    torch.cuda.synchronize()
    graphs = []
    for d in sets:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            step(d)
        graphs.append(gr)
    for gr in graphs:
        gr.replay()
    torch.cuda.synchronize()
    per_kernel = {}
    for ki, (name, _f) in enumerate(kernels(sets[0])):
        gk = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gk):
            for d in sets:
                kernels(d)[ki][1]()
        per_kernel[name] = _time(gk.replay, 5, torch) / NB * 1e3
    sampler = sampler_cls(local_rank) if (sampler_cls and rank == 0) else None
    if sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = max(NB, min(args.steps, 40) // NB * NB)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        graphs[i % NB].replay()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler else None
    ms = float(t.item())
    px = S * S
    # SURVEY 8d config 4 accounting per sample: uint8 crop out (786 432) + heat maps 17*(128^2 + 256^2)*4 (5 570 560) + the fused mix
    # (crop + fp32 logits in, fp32 out: 7 077 888) + backward (10 223 616) + G_input (10 223 616); the source footprint is on top
    alg = {"crop_u8": 3 * px, "heatmaps": Jn * (128 * 128 + 256 * 256) * 4, "chainmix_fwd": px * (3 + 12 + 12), "chainmix_bwd": px * (3 + 12 + 12 + 12),
           "g_input": px * (3 + 36)}
    hm_us = per_kernel["heatmap_targets 128x128"] + per_kernel["heatmap_targets 256x256"]
    return {"metric": "bottom-up 512x512 AdvMix steps: samples/sec (crop + heat maps 128^2/256^2 + G_input + mix fwd/bwd)", "value": world * B * steps / (ms * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": NB, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
            "config": {"workload": cfg_name, "batch_per_gpu": B, "cuda_graph": True, "l2": "%d batches cycled" % NB},
            "roofline": {"kernel": "chain_kernel<FWD> (advmix_chainmix_fwd, 512x512)", "bound": "hbm",
                         "achieved": B * alg["chainmix_fwd"] / (per_kernel["chainmix_fwd"] * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": B * alg["chainmix_fwd"] / (per_kernel["chainmix_fwd"] * 1e-6) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_sample": alg, "per_kernel_us": per_kernel,
                         "frac_by_kernel": {"chainmix_fwd": B * alg["chainmix_fwd"] / (per_kernel["chainmix_fwd"] * 1e-6) / 1e9 / peak,
                                            "chainmix_bwd": B * alg["chainmix_bwd"] / (per_kernel["chainmix_bwd"] * 1e-6) / 1e9 / peak,
                                            "chains_emit": B * alg["g_input"] / (per_kernel["chains_emit (G_input)"] * 1e-6) / 1e9 / peak,
                                            "heatmaps_128_256": B * alg["heatmaps"] / (hm_us * 1e-6) / 1e9 / peak},
                         "step_frac": B * sum(alg.values()) / (ms / steps * 1e-3) / 1e9 / peak},
            "cpu_baseline": None, "e2e": None, "gpu_launches": 10 * steps, "clocks": clocks, "impl": "advmix_b200"}


def sub_records(args, rank, local_rank, world, dev, seed, sampler_cls):
    """The sub-records of bench.py's default line: COCO-C sweep (FAST and exact arithmetic) and the AdvMix mix step."""
    cfg0 = "configs[0]: COCO-C sweep, 15 corruptions x 5 severities on 256x192 crops"
    cfg2 = "configs[2]: AdvMix inner loop, K=3 chains + per-pixel mix (forward + backward), batch 32/GPU, 256x192"
    set_size(256, 192)
    cpu_ok = world == 1 and not args.no_cpu_baseline
    out = {}
    out["coco_c_fast"] = coco_c_record(args, rank, world, dev, seed, cfg0, True, True, cpu_ok, sampler_cls, local_rank)
    out["coco_c_exact"] = coco_c_record(args, rank, world, dev, seed, cfg0, False, False, False, sampler_cls, local_rank)
    out["advmix_mix"] = advmix_mix_record(args, rank, world, dev, seed, cfg2, sampler_cls, local_rank)
    return out


def run(args, rank, local_rank, world, dev, seed, metric, cfg_name):
    import bench
    if args.workload == "jpeg_crop":
        return run_jpeg_crop(args, rank, local_rank, world, dev, seed, metric, cfg_name)
    if args.workload in ("coco_c", "mpii_c"):
        set_size(256, 256 if args.workload == "mpii_c" else 192)
        cpu_ok = world == 1 and not args.no_cpu_baseline
        line = coco_c_record(args, rank, world, dev, seed, cfg_name, True, True, cpu_ok, bench.ClockSampler, local_rank)
        line["exact_arithmetic"] = coco_c_record(args, rank, world, dev, seed, cfg_name, False, False, False, bench.ClockSampler, local_rank)
        return line
    B = 32 if args.batch == 256 else args.batch
    if args.workload == "bottomup512":
        return bottomup512_record(args, rank, world, dev, seed, cfg_name, bench.ClockSampler, local_rank, B=B)
    set_size(256, 192)
    line = advmix_mix_record(args, rank, world, dev, seed, cfg_name, bench.ClockSampler, local_rank, B=B)
    line["bf16_io"] = advmix_mix_record(args, rank, world, dev, seed, cfg_name, bench.ClockSampler, local_rank, B=B, dtype_name="bfloat16")
    return line
