#!/usr/bin/env python
"""bench.py - headline benchmark of the AdvMix augmentation + target hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload crop_targets|coco_c|mpii_c|advmix_mix|jpeg_crop|bottomup512]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path on the host cores

Default workload = BASELINE.json configs[1]: JointsDataset affine crop + generate_target
heat maps 64x48x17, batch 256 per GPU, COCO top-down 256x192.  One "step" = one batch through
get_affine_transform -> warpAffine(+flip)+ToTensor/Normalize -> joints -> generate_target.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the byte accounting.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20261017
SRC_H, SRC_W = 480, 640
OUT_W, OUT_H = 192, 256
HM_W, HM_H, J = 48, 64, 17
BATCH = 256


# ------------------------------------------------------------------------------------ inputs
def synth_records(B, rng, uniform_images=False):
    """per-sample centre / bbox / joints; images made separately."""
    recs = []
    for _ in range(B):
        cx = SRC_W * (0.5 + rng.uniform(-0.25, 0.25))
        cy = SRC_H * (0.5 + rng.uniform(-0.25, 0.25))
        w, h = rng.uniform(60, 400), rng.uniform(80, 440)
        x, y = cx - w / 2, cy - h / 2
        center = np.array([x + w * 0.5, y + h * 0.5], np.float32)
        ar = OUT_W / OUT_H
        if w > ar * h:
            h = w / ar
        elif w < ar * h:
            w = h * ar
        scale = np.array([w / 200.0, h / 200.0], np.float32) * 1.25
        joints = np.zeros((J, 3))
        joints[:, 0] = rng.uniform(x, x + w, J)
        joints[:, 1] = rng.uniform(y, y + h, J)
        v = (rng.random(J) < 0.8).astype(np.float64)
        vis = np.stack([v, v, np.zeros(J)], 1)
        recs.append({"center": center, "scale": scale, "joints_3d": joints, "joints_3d_vis": vis})
    return recs


def synth_draws(recs, rng):
    """Augmentation draws of JointsDataset.py:177-188 (scale 0.3, rot 40, flip) - fixed per sample."""
    B = len(recs)
    c = np.stack([r["center"] for r in recs]).copy()
    s = np.stack([r["scale"] for r in recs]).astype(np.float64)
    s = s * np.clip(rng.standard_normal(B) * 0.3 + 1, 0.7, 1.3)[:, None]
    rot = np.where(rng.random(B) <= 0.6, np.clip(rng.standard_normal(B) * 40, -80, 80), 0.0)
    flip = rng.random(B) <= 0.5
    c[:, 0] = np.where(flip, SRC_W - c[:, 0] - 1, c[:, 0])
    return c.astype(np.float32), s, rot, flip


def natural_images_torch(B, device, seed):
    """(S) 'natural-like' images generated on the device: 4-octave bilinear noise + 8-LSB noise."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    acc = torch.zeros((B, 3, SRC_H, SRC_W), device=device)
    for o in range(4):
        sdiv = 2 ** (o + 3)
        low = torch.rand((B, 3, SRC_H // sdiv + 2, SRC_W // sdiv + 2), device=device, generator=g)
        acc += torch.nn.functional.interpolate(low, size=(SRC_H, SRC_W), mode="bilinear", align_corners=False) / (o + 1)
    acc = acc / acc.amax(dim=(1, 2, 3), keepdim=True) * 255
    acc += torch.randint(-8, 9, acc.shape, device=device, generator=g)
    return acc.clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def src_footprint_bytes(M_fwd, flip):
    """Algorithmic source bytes of one crop: area of the dst rectangle mapped back into the source,
    clipped to the image, x 3 B (Sutherland-Hodgman)."""
    A = np.vstack([M_fwd, [0, 0, 1]])
    inv = np.linalg.inv(A)
    quad = [(inv @ np.array([x, y, 1.0]))[:2] for x, y in ((0, 0), (OUT_W, 0), (OUT_W, OUT_H), (0, OUT_H))]

    def clip(poly, axis, bound, keep_less):
        out = []
        for i in range(len(poly)):
            p, q = poly[i], poly[(i + 1) % len(poly)]
            pin = p[axis] <= bound if keep_less else p[axis] >= bound
            qin = q[axis] <= bound if keep_less else q[axis] >= bound
            if pin:
                out.append(p)
            if pin != qin:
                t = (bound - p[axis]) / (q[axis] - p[axis])
                out.append(p + t * (q - p))
        return out
    poly = [np.array(p) for p in quad]
    for axis, bound, less in ((0, 0.0, False), (0, float(SRC_W), True), (1, 0.0, False), (1, float(SRC_H), True)):
        if not poly:
            break
        poly = clip(poly, axis, bound, less)
    if len(poly) < 3:
        return 0.0
    x = np.array([p[0] for p in poly]); y = np.array([p[1] for p in poly])
    return 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))) * 3.0


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU reference arm
def natural_images_numpy(n, rng):
    """Natural-like uint8 sources on the host (same recipe as the device generator: smooth colour field + noise)."""
    import cv2
    out = []
    for _ in range(n):
        acc = np.zeros((SRC_H, SRC_W, 3), np.float32)
        for o in range(4):
            sdiv = 2 ** (o + 3)
            low = rng.random((SRC_H // sdiv + 2, SRC_W // sdiv + 2, 3)).astype(np.float32)
            acc += cv2.resize(low, (SRC_W, SRC_H), interpolation=cv2.INTER_LINEAR) / (o + 1)
        acc = acc / acc.max() * 255 + rng.integers(-8, 9, acc.shape)
        out.append(np.clip(acc, 0, 255).astype(np.uint8))
    return out


class _PortCropTargets:
    """Fallback when neither /root/reference nor oracle/_ref is present: the oracle/ restatement of
    JointsDataset.get_clean (get_affine_transform + cv2.warpAffine + ToTensor/Normalize + generate_target)."""

    def __init__(self, images, recs, draws):
        self.images, self.recs, self.draws = images, recs, draws
        from torchvision import transforms as T
        self.tf = T.Compose([T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])

    def __len__(self):
        return len(self.recs)

    def __getitem__(self, i):
        import torch
        from oracle import affine as OA, targets as OT
        c, s, rot, flip = (d[i] for d in self.draws)
        rec = self.recs[i]
        img = self.images[i % len(self.images)]
        joints, vis = rec["joints_3d"].copy(), rec["joints_3d_vis"].copy()
        if flip:
            img = img[:, ::-1, :]
            joints, vis = OA.fliplr_joints(joints, vis, img.shape[1], [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]])
        trans = OA.get_affine_transform(c, s, rot, (OUT_W, OUT_H))
        crop = OA.warp_affine_cv2(img, trans, (OUT_W, OUT_H))
        inp = self.tf(crop)
        joints = OA.transform_joints(joints, vis, trans)
        target, tw = OT.generate_target(joints, vis, (OUT_W, OUT_H), (HM_W, HM_H), 2)
        return inp, torch.from_numpy(target[0]), torch.from_numpy(tw)


_MEM_IMAGES = None


def _mem_imread(path, flags=None):
    """The reference reads its sources with cv2.imread (JointsDataset.py:148); both arms of this bench start from DECODED
    sources in host RAM (the GPU arm's e2e leg reads pinned decoded pixels), so the harness serves "mem:<i>" paths from RAM."""
    return _MEM_IMAGES[int(path.split(":")[1]) % len(_MEM_IMAGES)]


def _ref_worker_init(_wid):
    import cv2
    cv2.setNumThreads(1)
    cv2.imread = _mem_imread


def run_cpu_reference(batches, warmup_batches, batch=BATCH, rng_seed=SEED, n_img=256):
    """`batches` timed 256-sample batches (after `warmup_batches` untimed ones) of the reference's CPU path for
    configs[1] through torch DataLoader(batch_size=256, num_workers=all cores), like tools/train.py:165-171.
    kind "reference": the REAL lib/dataset/JointsDataset.__getitem__ (sample_times=1 -> get_clean: imread ->
    get_affine_transform -> cv2.warpAffine -> ToTensor/Normalize -> generate_target), imported from /root/reference or from
    the byte-compiled oracle/_ref; kind "port": the oracle/ restatement when neither is present."""
    global _MEM_IMAGES
    import torch
    from torch.utils.data import DataLoader
    # A DataLoader worker builds WHOLE batches, so batches arrive in bursts of `cores`; the timed window must start and end on
    # a burst boundary or the batches already finished (prefetched) when the clock starts count as free: both counts are
    # rounded up to multiples of the worker count (the JSON line reports the counts actually used).
    ncores = os.cpu_count() or 1
    warmup_batches = max(1, -(-max(1, warmup_batches) // ncores)) * ncores
    batches = max(1, -(-batches // ncores)) * ncores
    from oracle import ref_harness
    rng = np.random.default_rng(rng_seed)
    cores = os.cpu_count() or 1
    images = natural_images_numpy(n_img, rng)
    n_samples = batch * (batches + warmup_batches)
    recs = synth_records(n_samples, rng)
    kind = "reference" if ref_harness.available() else "port"
    if kind == "reference":
        _MEM_IMAGES = images
        db = [{"image": "mem:%d" % i, "center": r["center"], "scale": r["scale"], "joints_3d": r["joints_3d"],
               "joints_3d_vis": r["joints_3d_vis"], "filename": "", "imgnum": 0} for i, r in enumerate(recs)]
        ds = ref_harness.make_dataset(db, is_train=True, sample_times=1)      # scale 0.3 / rot 40 / flip: the same draws' distributions
        what = "the reference's own JointsDataset.__getitem__ (%s)" % ("live tree" if ref_harness.kind() == "source" else "byte-compiled oracle/_ref")
    else:
        ds = _PortCropTargets(images, recs, synth_draws(recs, rng))
        what = "oracle/ port of get_clean"
    loader = DataLoader(ds, batch_size=batch, shuffle=False, num_workers=cores, worker_init_fn=_ref_worker_init,
                        prefetch_factor=2, persistent_workers=False)
    t0 = None
    seen = 0
    for bi, _b in enumerate(loader):
        if bi == warmup_batches - 1:
            t0 = time.perf_counter()              # the last warm-up batch has arrived: everything after it is timed
        elif bi >= warmup_batches:
            seen += 1
    dt = time.perf_counter() - t0
    value = seen * batch / dt
    return {"value": value, "unit": "samples/s", "cores": cores, "kind": kind, "batches": seen, "warmup_batches": warmup_batches,
            "sample": "%d timed batches of %d samples (after %d warm-up batches) through DataLoader(batch_size=%d, num_workers=%d): %s; "
                      "%d distinct natural-like %dx%d uint8 sources decoded in host RAM, %.1f s"
                      % (seen, batch, warmup_batches, batch, cores, what, n_img, SRC_W, SRC_H, dt), "seconds": dt}


def bench_config(cfg_name, batch, gpus):
    """The `config` object both arms print (identical keys and values for identical flags)."""
    return {"workload": cfg_name, "global_batch": batch * gpus, "batch_per_gpu": batch,
            "src": "%dx%d uint8 HWC, natural-like synthetic, decoded, resident in memory before the timed region" % (SRC_W, SRC_H),
            "out": "fp32 [3,256,192] normalised + fp32 heatmaps [17,64,48] + target_weight",
            "draws": "scale 0.3, rotation 40 (p 0.6), flip 0.5 per sample (JointsDataset.py:177-188)"}


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="advmix_b200", choices=["advmix_b200", "reference"])
    ap.add_argument("--workload", default="crop_targets", choices=["crop_targets", "coco_c", "mpii_c", "advmix_mix", "jpeg_crop", "bottomup512"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--shard-images", type=int, default=1024, help="decoded sources in a rank's pinned host shard (e2e leg)")
    ap.add_argument("--serial-steps", action="store_true", help="issue every step on one stream (default: two alternating streams)")
    ap.add_argument("--no-extras", action="store_true", help="skip the coco_c / advmix_mix sub-records of the default line")
    ap.add_argument("--sets", type=int, default=4, help="different 256-sample batches a rank cycles through")
    ap.add_argument("--shard", type=int, default=None, help="use this one synthetic draw-set for every batch (experiments)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3)
    metric = "augmented 256x192 samples/sec"
    cfg_name = {"crop_targets": "configs[1]: JointsDataset affine crop + generate_target heatmaps 64x48x17, batch %d/GPU, COCO top-down 256x192" % args.batch,
                "coco_c": "configs[0]: COCO-C sweep, 15 corruptions x 5 severities on 256x192 crops",
                "mpii_c": "configs[4]: MPII-C construction, 15 corruptions x 5 severities on 256x256 crops",
                "advmix_mix": "configs[2]: AdvMix inner loop, K=3 corruption chains + per-pixel mix, batch 32/GPU",
                "bottomup512": "configs[3]: HrHRNet-W32 512x512 bottom-up input, AdvMix mix + multi-resolution heatmap targets (128x128, 256x256), batch 32/GPU",
                "jpeg_crop": "configs[1] fed from encoded sources (SURVEY 8f rank 1): device JPEG decode + affine crop + heatmaps, batch %d/GPU" % args.batch}[args.workload]

    if args.impl == "reference":
        if rank != 0:
            return
        # one step = one 256-sample batch of the reference's CPU path on all host cores; K timed steps after W warm-up ones
        # (capped so that the run stays within minutes: ~75 ms per batch on 16 cores)
        steps = max(1, min(args.steps, 400))
        r = run_cpu_reference(batches=steps, warmup_batches=warmup, batch=args.batch)
        steps, warmup = r["batches"], r["warmup_batches"]
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": "samples/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * args.batch / r["value"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": bench_config(cfg_name, args.batch, args.gpus),
                "impl_notes": {"note": "CPU path on the host cores of the box; the GPU count does not apply (rank 0 runs it alone)"},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    numa_cpus = None
    if world > 1:
        # one process per GPU: keep the rank (and the pinned buffers it allocates) on the GPU's own NUMA node
        from advmix_b200.dist import bind_to_gpu_numa
        numa_cpus = bind_to_gpu_numa(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import advmix_b200 as A
    from advmix_b200 import _lib, transforms as TF, targets as TG
    _lib.check(A.load_library().advmix_device_check(local_rank), "device_check")

    # --- control block: rank 0 broadcasts {seed, epoch} once per epoch; per-step
    # counters are derived locally, so the timed region has no collective.
    ctrl = torch.tensor([SEED, 0], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(ctrl, src=0)
    seed = int(ctrl[0].item())

    if args.workload != "crop_targets":
        from benchmarks import extra_workloads
        line = extra_workloads.run(args, rank, local_rank, world, dev, seed, metric, cfg_name)
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    B = args.batch
    perm = TF.flip_perm(J, [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]], dev)
    lut = TF.normalize_lut(device=dev)
    gtab = TG.gaussian_table(2, dev)
    lib = A.load_library()
    P, S = _lib.ptr, _lib.stream_ptr

    # static output buffers (the step is allocation-free, so it can be captured in a CUDA graph)
    M = torch.empty((B, 2, 3), dtype=torch.float64, device=dev)
    inp = torch.empty((B, 3, OUT_H, OUT_W), dtype=torch.float32, device=dev)
    jo = torch.empty((B, J, 3), dtype=torch.float64, device=dev); vo = torch.empty_like(jo)
    hm = torch.empty((B, J, HM_H, HM_W), dtype=torch.float32, device=dev)
    mu = torch.empty((B, J, 2), dtype=torch.float32, device=dev)
    tw = torch.empty((B, J, 1), dtype=torch.float32, device=dev)
    side = torch.cuda.Stream(device=dev)

    # The step time depends on the draws of the batch (more rotated / down-scaled samples -> more pipeline items
    # in the crop kernel: 83..95 us for equal algorithmic bytes), so a rank does not time ONE batch: it cycles
    # through NSETS different 256-sample batches.  The draw-sets are the same on every rank (the pixels are
    # not), which makes the per-GPU work identical - the definition of weak scaling.
    NSETS = max(1, args.sets)

    def make_set(j):
        sid = j if args.shard is None else args.shard
        rng = np.random.default_rng(seed + 1000 * sid)
        recs = synth_records(B, rng)
        c, s, rot, flip = synth_draws(recs, rng)
        images = natural_images_torch(B, dev, seed + 97 * rank + sid)    # [B,480,640,3] uint8 resident in HBM (236 MB > L2)
        sources = A.SourceBatch.from_tensor(images)
        c_t = torch.from_numpy(c).to(dev); s_t = torch.from_numpy(s).to(dev)
        r_t = torch.from_numpy(rot).to(dev); f_t = torch.from_numpy(flip.astype(np.uint8)).to(dev)
        joints_in = torch.from_numpy(np.stack([r["joints_3d"] for r in recs])).to(dev)
        vis_in = torch.from_numpy(np.stack([r["joints_3d_vis"] for r in recs])).to(dev)

        def k_matrices():
            _lib.check(lib.advmix_affine_matrices(P(c_t), P(s_t), 0, P(r_t), P(M), B, OUT_W, OUT_H, S()))

        def k_warp():
            _lib.check(lib.advmix_warp_affine_u8c3(P(sources.buffer), P(sources.offsets), P(sources.heights), P(sources.widths),
                                                   P(sources.pitches), P(f_t), P(M), None, P(inp), P(lut), B, OUT_W, OUT_H,
                                                   _lib.F32, S()))

        def k_joints():
            _lib.check(lib.advmix_joints_flip_affine(P(joints_in), P(vis_in), P(f_t), P(sources.widths), P(perm), P(M),
                                                     P(jo), P(vo), B, J, S()))

        def k_heatmap():
            _lib.check(lib.advmix_heatmap_targets(P(jo), P(vo), P(gtab), None, P(hm), P(mu), P(tw), B, J, HM_H, HM_W,
                                                  OUT_W, OUT_H, 2, S()))

        def step():
            # matrices -> { warp (main stream) || joints + heat maps (side stream) }: the crop is issue-bound,
            # the targets are store-bound, so the two branches overlap on the SMs.
            diag = os.environ.get("ADVMIX_BENCH_DIAG", "")          # diagnosis only: which kernels make the step longer than the crop
            if "nomat" not in diag:
                k_matrices()
            main = torch.cuda.current_stream()
            if "noside" not in diag:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    k_joints()
                    k_heatmap()
            k_warp()
            if "noside" not in diag:
                main.wait_stream(side)

        k_matrices()
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        run = step
        graph = None
        if not args.no_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            run = graph.replay
        for _ in range(warmup):
            run()
        torch.cuda.synchronize()
        # byte accounting: per sample src footprint + fp32 normalised crop + heat maps + small
        Mh = M.cpu().numpy()
        foot = np.array([src_footprint_bytes(Mh[b], flip[b]) for b in range(B)])
        bytes_warp = float(foot.sum()) + B * (3 * OUT_H * OUT_W * 4)
        return dict(recs=recs, draws=(c, s, rot, flip), images=images, run=run, graph=graph, bytes_warp=bytes_warp,
                    kernels=[("affine_matrices", k_matrices), ("warp_affine", k_warp), ("joints_flip_affine", k_joints),
                             ("heatmap_targets", k_heatmap)], keep=(sources, c_t, s_t, r_t, f_t, joints_in, vis_in))

    sets = [make_set(j) for j in range(NSETS)]
    use_graph = not args.no_graph
    recs, (c, s, rot, flip), images = sets[0]["recs"], sets[0]["draws"], sets[0]["images"]     # the e2e leg below uses set 0
    bytes_warp = float(np.mean([st["bytes_warp"] for st in sets]))
    bytes_hm = B * (J * HM_H * HM_W * 4 + J * 4 + J * 8 + 2 * J * 24)
    bytes_step = bytes_warp + bytes_hm + B * (6 * 8 + 2 * 2 * J * 24)
    kernels = sets[0]["kernels"]

    # --- timed region: K steps, device events, barrier + sync on both sides, max over ranks
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Consecutive steps work on different batch sets, so they are issued on two alternating streams (set j always on stream
    # j % 2): the matrices / joints / heat-map kernels and the launch gaps of step i+1 run under the crop kernel of step i, as
    # in fastpath.CropTargetsStep(prefetch_streams=2).  --serial-steps puts every step on one stream.
    two = use_graph and not args.serial_steps and NSETS % 2 == 0
    main_s = torch.cuda.current_stream()
    lanes = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)] if two else None
    e0.record()
    if two:
        for ln in lanes:
            ln.wait_stream(main_s)
        for i in range(args.steps):
            torch.cuda.set_stream(lanes[i & 1])
            sets[i % NSETS]["run"]()
        torch.cuda.set_stream(main_s)
        for ln in lanes:
            main_s.wait_stream(ln)
    else:
        for i in range(args.steps):
            sets[i % NSETS]["run"]()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # --- per-kernel durations (events on the launching stream) for the roofline of the dominant kernel
    per_kernel = {}
    for ki, (name, _k) in enumerate(kernels):
        ks = [st["kernels"][ki][1] for st in sets]      # the same kernel on every batch set, round robin
        for i in range(3 * NSETS):
            ks[i % NSETS]()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(20, args.steps) // NSETS * NSETS
        a0.record()
        for i in range(reps):
            ks[i % NSETS]()
        a1.record()
        torch.cuda.synchronize()
        per_kernel[name] = a0.elapsed_time(a1) / reps * 1e3   # us

    # --- e2e: public API with HOST buffers.  The rank's shard of the dataset (`--shard-images` decoded uint8 sources +
    # their records) lives in pinned host memory.  Every step: the host draws the augmentation, gathers the records of the
    # batch, makes sure the batch's sources are in the HBM source cache (advmix_b200.fastpath.SourceCache: an image crosses
    # PCIe the first time an epoch touches it), packs ONE pinned parameter buffer, issues one H2D copy and one library call
    # (advmix_crop_targets_step), and reads target_weight back one step late.  `e2e.value` is the steady state (epochs >= 2 of a
    # 210-epoch schedule, tools/train.py); the first, cold epoch and the no-cache streaming path of round 1 are reported next to it.
    from advmix_b200 import fastpath as FP
    from advmix_b200.dataset import AdvMixBatchPipeline
    D = max(B, args.shard_images // B * B)
    host_all = torch.empty((D, SRC_H, SRC_W, 3), dtype=torch.uint8, pin_memory=True)
    for k in range(0, D, B):
        host_all[k:k + B].copy_(images if k == 0 else natural_images_torch(B, dev, seed + 97 * rank + 10 + k))
    torch.cuda.synchronize()
    rng_e = np.random.default_rng(seed + 5)
    recs_all = synth_records(D, rng_e)
    table = FP.RecordTable.from_records(recs_all, widths=np.full(D, SRC_W), heights=np.full(D, SRC_H))
    cache = FP.SourceCache(D * SRC_H * SRC_W * 3 + D * 256, D, dev)
    fstep = FP.CropTargetsStep(B, device=dev, seed=seed + rank, ring=4, out_ring=4, graph=True, prefetch_streams=2)     # record rows on the device, 4-deep ring of (pinned slot, output set, captured graph)
    tw_host = [torch.empty((B, J, 1), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    tw_done = [torch.cuda.Event(), torch.cuda.Event()]
    tw_np = [t.numpy() for t in tw_host]                 # views of the pinned read-back buffers (indexing a tensor costs ~5 us, a numpy view 0.2)
    fetch = lambda i: host_all[i]
    perm_e = rng_e.permutation(D)

    def e2e_run(first, n):
        checksum = 0.0
        for i in range(first, first + n):
            ids = perm_e[(i * B) % D:(i * B) % D + B]
            off, pitch, hh, ww = cache.ensure(ids, fetch)
            _inp, _target, _tw, _meta = fstep(table, ids, cache.buffer, off, pitch, hh, ww, after=cache.take_upload_event())
            k = i & 1
            tw_host[k].copy_(_tw, non_blocking=True)
            tw_done[k].record()
            if i > first:
                tw_done[k ^ 1].synchronize()
                checksum += float(tw_np[k ^ 1][0, 0, 0])
        tw_done[(first + n - 1) & 1].synchronize()
        return checksum + float(tw_np[(first + n - 1) & 1][0, 0, 0])

    def timed(fn, *a_):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        fn(*a_)
        b1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())
    epoch_steps = D // B
    e2e_run(0, len(fstep.slots))                         # allocator / first-call warm-up: every ring entry captures its graph once
    cache.off[:] = -1; cache.used = 0; cache.uploaded_bytes = 0
    ms_cold = timed(e2e_run, 0, epoch_steps)             # epoch 1: every source crosses PCIe once
    cold_h2d = cache.uploaded_bytes // epoch_steps + fstep.nbytes
    e2e_run(epoch_steps, 3)
    e2e_steps = max(2 * epoch_steps, min(args.steps, 64) // epoch_steps * epoch_steps)
    up0 = cache.uploaded_bytes
    ms_warm = timed(e2e_run, 2 * epoch_steps, e2e_steps)
    e2e_value = world * B * e2e_steps / (ms_warm * 1e-3)
    h2d_bytes = (cache.uploaded_bytes - up0) // e2e_steps + fstep.nbytes
    e2e_cold_value = world * B * epoch_steps / (ms_cold * 1e-3)
    # the round-1 path for comparison: no cache, AdvMixBatchPipeline, zero-copy gather of the boxes the crops read, every step
    hsb = TF.HostSourceBatch.from_tensor(host_all[:B], dev)
    pipe = AdvMixBatchPipeline(sample_times=1, is_train=True, device=dev)
    for r_ in recs:
        r_["width"], r_["height"] = SRC_W, SRC_H

    def stream_run(n):
        for i in range(n):
            _inp, _target, _tw, _meta = pipe(recs, draws=(c, s, rot, flip), host_sources=hsb)
            tw_host[i & 1].copy_(_tw, non_blocking=True)
            tw_done[i & 1].record()
            if i > 0:
                tw_done[(i - 1) & 1].synchronize()
        tw_done[(n - 1) & 1].synchronize()
    stream_run(3)
    if getattr(hsb, "bytes_sent", None) is not None:
        hsb.bytes_sent.zero_()
    ms_stream = timed(stream_run, 10)
    e2e_stream_value = world * B * 10 / (ms_stream * 1e-3)
    stream_h2d = (int(hsb.bytes_sent.item()) // 10) if getattr(hsb, "bytes_sent", None) is not None else int(pipe.last_h2d_bytes)
    clocks = sampler.stop() if sampler else None      # sampled over all timed regions (step loop, per-kernel, e2e)

    # --- the other north-star workloads as sub-records of the same line: COCO-C 15x5 sweep (exact and
    # FAST arithmetic) and the AdvMix inner-loop mix step (fused chain + mix, forward + backward); each with its own clocks
    extras = None
    if not args.no_extras:
        from benchmarks import extra_workloads
        extras = extra_workloads.sub_records(args, rank, local_rank, world, dev, seed, ClockSampler)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom = "warp_affine"
    achieved = bytes_warp / (per_kernel[dom] * 1e-6) / 1e9
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_warp, "us_per_launch": per_kernel[dom],
                "step_achieved_gbs": bytes_step / (ms_per_step * 1e-3) / 1e9,
                "step_frac": bytes_step / (ms_per_step * 1e-3) / 1e9 / peak,
                "per_kernel_us": per_kernel,
                "heatmap_gbs": bytes_hm / (per_kernel["heatmap_targets"] * 1e-6) / 1e9}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(dom)
        except Exception:
            pass
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = run_cpu_reference(batches=32, warmup_batches=1, batch=B)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": bench_config(cfg_name, B, world),
            "impl_notes": {"l2": "inputs+outputs per step = %.0f MB > 126 MB L2 (no flush needed)" % ((images.numel() + inp.numel() * 4 + hm.numel() * 4) / 1e6),
                           "batch_sets": "%d different %d-sample batches per rank, cycled step by step; the same draw-sets on every rank (identical per-GPU work), different pixels" % (NSETS, B),
                           "cuda_graph": use_graph, "streams": "warp || (joints, heat maps) after the matrix kernel; consecutive steps (different batch sets) on two alternating streams" if two else "warp || (joints, heat maps) after the matrix kernel; all steps on one stream", "parallelism": "sample-sharded, %d rank(s), no data-path collective" % world,
                           "rank0_numa_local_cpus": (len(numa_cpus) if numa_cpus else None)},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(tw_host[0].numel() * 4),
                    "steps": e2e_steps, "shard_images_per_rank": D,
                    "path": "advmix_b200.fastpath: pinned host shard (decoded uint8 sources + records; joints / visibility rows uploaded once) -> per step: host draws (pre-drawn 64 steps per numpy call), HBM source cache lookup (misses cross PCIe), one pinned parameter buffer (draws + source addresses + record indices), one H2D copy + one advmix_crop_targets_step_rec call into a 4-deep output ring (copy and call captured into one CUDA graph per ring entry, replayed on two alternating prefetch streams so that step i+1's copy / matrices / heat maps run under step i's crop), D2H of target_weight read one step late; steady state (epochs >= 2)",
                    "first_epoch": {"value": e2e_cold_value, "unit": "samples/s", "h2d_bytes_per_step": int(cold_h2d), "note": "cold cache: every decoded source of the shard crosses PCIe once"},
                    "streaming_no_cache": {"value": e2e_stream_value, "unit": "samples/s", "h2d_bytes_per_step": int(stream_h2d) + B * (8 + 16 + 8 + 1 + 2 * J * 24) + B * 64,
                                           "note": "round-1 path: AdvMixBatchPipeline(records, host_sources=...) gathers the source boxes of every crop out of pinned host memory every step"}},
            "value_api": e2e_value,        # throughput through the public API (fastpath.CropTargetsStep), not the bench's own graph - the e2e leg IS that call
            "gpu_launches": len(kernels) * args.steps, "clocks": clocks, "impl": "advmix_b200",
            "workloads": extras}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
