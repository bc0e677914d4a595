"""The K = 1 training step (JointsDataset.get_clean, lib/dataset/JointsDataset.py:258-364) with the host reduced to what it has
to do: draw, pack ONE pinned buffer, issue ONE copy and ONE library call (advmix_crop_targets_step).

`AdvMixBatchPipeline` mirrors the reference's `__getitem__` contract record by record (lists of dicts in, the reference's
structure out) and pays ~1 ms of Python per 256 samples for it - ten times the device time of the step.  A training loop that
owns its dataset does not need that generality per step:

  * `RecordTable`   the reference's `db` (list of dicts) as a structure of arrays, built once;
  * `SourceCache`   decoded source images resident in HBM, keyed by dataset index: an image crosses PCIe the first time an epoch
                    touches it and never again (a B200 holds ~190 k decoded 640x480 sources in its 180 GB; a rank's shard of
                    COCO train2017 is ~15 k).  The round-1 path shipped every crop's source rows every step and topped out at
                    the host's memory bandwidth with 8 ranks (0.48 scaling efficiency);
  * `CropTargetsStep`  draws for the whole batch (numpy Generator, the distributions of JointsDataset.py:177-188), packs the
                    per-step arrays into one pinned buffer, one H2D copy, one C call.

The outputs are the reference's: (input [B,3,H,W], [heatmap [B,J,Hh,Wh], mu [B,J,2]], target_weight [B,J,1], meta).
"""
import numpy as np
import torch

from . import _lib
from . import targets as TG
from . import transforms as TF


class RecordTable:
    """Structure-of-arrays form of the reference's db records ('center', 'scale', 'joints_3d', 'joints_3d_vis' + image size)."""

    def __init__(self, centers, scales, joints, vis, widths, heights):
        self.centers = np.ascontiguousarray(centers, np.float32)
        self.scales = np.ascontiguousarray(scales, np.float32)
        self.joints = np.ascontiguousarray(joints, np.float64)
        self.vis = np.ascontiguousarray(vis, np.float64)
        self.widths = np.ascontiguousarray(widths, np.int32)
        self.heights = np.ascontiguousarray(heights, np.int32)

    @classmethod
    def from_records(cls, records, widths=None, heights=None):
        n = len(records)
        w = widths if widths is not None else [r["width"] for r in records]
        h = heights if heights is not None else [r["height"] for r in records]
        return cls(np.array([r["center"] for r in records], np.float32).reshape(n, 2),
                   np.array([r["scale"] for r in records], np.float32).reshape(n, 2),
                   np.array([r["joints_3d"] for r in records], np.float64),
                   np.array([r["joints_3d_vis"] for r in records], np.float64), w, h)

    def __len__(self):
        return self.centers.shape[0]


class SourceCache:
    """Decoded uint8 HWC sources resident in one device buffer, keyed by dataset index.  `ensure(ids, fetch)` uploads the
    images that are not resident yet (`fetch(i)` -> pinned uint8 [H,W,3] host tensor) and returns the per-sample
    (offset, pitch, height, width) arrays of the batch.  Rows are stored with a 16-byte aligned pitch (the crop kernel's
    staged path).  When the buffer is full the cache starts over (a shard that fits never gets there)."""

    def __init__(self, capacity_bytes, n_items, device="cuda"):
        self.device = torch.device(device)
        self.buffer = torch.empty(int(capacity_bytes), dtype=torch.uint8, device=self.device)
        self.off = np.full(n_items, -1, np.int64)
        self.pitch = np.zeros(n_items, np.int64)
        self.h = np.zeros(n_items, np.int32)
        self.w = np.zeros(n_items, np.int32)
        self.used = 0
        self.uploaded_bytes = 0

    def ensure(self, ids, fetch):
        miss = ids[self.off[ids] < 0]
        if miss.size:
            for i in np.unique(miss):
                img = fetch(int(i))
                H, W = int(img.shape[0]), int(img.shape[1])
                pitch = (3 * W + 15) // 16 * 16
                need = (H * pitch + 255) // 256 * 256
                if self.used + need > self.buffer.numel():
                    self.off[:] = -1                                   # start over
                    self.used = 0
                    return self.ensure(ids, fetch)
                dst = self.buffer[self.used:self.used + H * pitch].view(H, pitch)
                if pitch == 3 * W:
                    dst.copy_(img.reshape(H, 3 * W), non_blocking=True)
                else:
                    dst[:, :3 * W].copy_(img.reshape(H, 3 * W), non_blocking=True)
                self.off[i], self.pitch[i], self.h[i], self.w[i] = self.used, pitch, H, W
                self.used += need
                self.uploaded_bytes += H * W * 3
        return self.off[ids], self.pitch[ids], self.h[ids], self.w[ids]


class CropTargetsStep:
    def __init__(self, batch, image_size=(192, 256), heatmap_size=(48, 64), sigma=2, num_joints=17, flip_pairs=None,
                 scale_factor=0.3, rot_factor=40, flip=True, is_train=True, joints_weight=None, norm_dtype=torch.float32,
                 device="cuda", seed=0, ring=4):
        from .dataset import COCO_FLIP_PAIRS
        self.B, self.J = int(batch), int(num_joints)
        self.image_size, self.heatmap_size, self.sigma = tuple(image_size), tuple(heatmap_size), int(sigma)
        self.scale_factor, self.rot_factor, self.flip, self.is_train = scale_factor, rot_factor, flip, is_train
        self.norm_dtype = norm_dtype
        self.device = torch.device(device)
        self.rng = np.random.default_rng(seed)
        self.lib = _lib.load()
        self.perm = TF.flip_perm(self.J, COCO_FLIP_PAIRS if flip_pairs is None else flip_pairs, self.device)
        self.lut = TF.normalize_lut(device=self.device)
        self.gtab = TG.gaussian_table(self.sigma, self.device)
        self.jw = None if joints_weight is None else torch.as_tensor(joints_weight, dtype=torch.float32).reshape(-1).to(self.device)
        self.nbytes = int(self.lib.advmix_step_params_bytes(self.B, self.J))
        # pinned staging ring + numpy views of its sections (layout documented at advmix_crop_targets_step)
        B, J = self.B, self.J
        al = lambda v: (v + 15) & ~15
        sections = [("src_off", np.int64, (B,)), ("src_pitch", np.int64, (B,)), ("src_h", np.int32, (B,)), ("src_w", np.int32, (B,)),
                    ("scale", np.float64, (B, 2)), ("rot", np.float64, (B,)), ("center", np.float32, (B, 2)), ("flip", np.uint8, (B,)),
                    ("joints", np.float64, (B, J, 3)), ("vis", np.float64, (B, J, 3))]
        self.slots = []
        for _ in range(ring):
            host = torch.empty(self.nbytes, dtype=torch.uint8).pin_memory()
            hv, views, o = host.numpy(), {}, 0
            for name, dt, shape in sections:
                n = int(np.prod(shape)) * np.dtype(dt).itemsize
                views[name] = hv[o:o + n].view(dt).reshape(shape)
                o += al(n)
            assert o == self.nbytes, (o, self.nbytes)
            self.slots.append({"host": host, "views": views, "dev": torch.empty(self.nbytes, dtype=torch.uint8, device=self.device), "event": None})
        self.next = 0

    def draw(self, centers, scales, widths):
        """The augmentation draws of JointsDataset.py:177-188 for the whole batch -> (center f32 [B,2] already mirrored where
        flipped, scale f64 [B,2], rot f64 [B], flip bool [B])."""
        B, g = len(centers), self.rng
        c = centers.copy()
        s = scales.astype(np.float64)
        rot = np.zeros(B)
        flip = np.zeros(B, bool)
        if self.is_train:
            sf, rf = self.scale_factor, self.rot_factor
            s = s * np.clip(g.standard_normal(B) * sf + 1, 1 - sf, 1 + sf)[:, None]
            rot = np.where(g.random(B) <= 0.6, np.clip(g.standard_normal(B) * rf, -rf * 2, rf * 2), 0.0)
            if self.flip:
                flip = g.random(B) <= 0.5
                c[:, 0] = np.where(flip, widths - c[:, 0] - 1, c[:, 0])
        return c, s, rot, flip

    def __call__(self, table, ids, src_base, src_off, src_pitch, src_h, src_w, draws=None):
        """table: RecordTable; ids: int array [B] of dataset indices; src_*: where the decoded sources of these samples live
        (SourceCache.ensure(...) output, or a SourceBatch's fields as numpy arrays) - src_base is the device buffer."""
        B, J, lib, P = self.B, self.J, self.lib, _lib.ptr
        assert len(ids) == B
        c, s, rot, flip = draws if draws is not None else self.draw(table.centers[ids], table.scales[ids], table.widths[ids])
        slot = self.slots[self.next]
        self.next = (self.next + 1) % len(self.slots)
        if slot["event"] is not None:
            slot["event"].synchronize()                  # the copy that last read this pinned buffer has finished
        v = slot["views"]
        v["src_off"][:] = src_off; v["src_pitch"][:] = src_pitch; v["src_h"][:] = src_h; v["src_w"][:] = src_w
        v["scale"][:] = s; v["rot"][:] = rot; v["center"][:] = c; v["flip"][:] = flip
        v["joints"][:] = table.joints[ids]; v["vis"][:] = table.vis[ids]
        dev = self.device
        slot["dev"].copy_(slot["host"], non_blocking=True)
        slot["event"] = torch.cuda.Event()
        slot["event"].record()
        W, H = self.image_size
        Wh, Hh = self.heatmap_size
        M = torch.empty((B, 2, 3), dtype=torch.float64, device=dev)
        inp = torch.empty((B, 3, H, W), dtype=self.norm_dtype, device=dev)
        jo = torch.empty((B, J, 3), dtype=torch.float64, device=dev)
        vo = torch.empty((B, J, 3), dtype=torch.float64, device=dev)
        hm = torch.empty((B, J, Hh, Wh), dtype=torch.float32, device=dev)
        mu = torch.empty((B, J, 2), dtype=torch.float32, device=dev)
        tw = torch.empty((B, J, 1), dtype=torch.float32, device=dev)
        _lib.check(lib.advmix_crop_targets_step(P(src_base), P(slot["dev"]), P(self.perm), P(self.lut), P(self.gtab), P(self.jw),
                                                P(M), P(inp), _lib.dtype_code(self.norm_dtype), P(jo), P(vo), P(hm), P(mu), P(tw),
                                                B, J, W, H, Hh, Wh, self.sigma, _lib.stream_ptr()), "advmix_crop_targets_step")
        meta = {"joints": jo, "joints_vis": vo, "center": c, "scale": s, "rotation": rot, "flip": flip, "trans": M, "index": ids}
        return inp, [hm, mu], tw, meta


class AdvMixStep(CropTargetsStep):
    """The K = 3 (sample_times = 3) step for the fused chain + mix path: ONE library call (advmix_crop_chains_step) produces the
    uint8 crop, the per-image autoaug plans, the gridmask parameters, the joints and the heat-map targets of the clean / autoaug
    chains and of the gridmask chain - the three normalised chain tensors of the reference's `inputs` list are never
    materialised.  The AdvMix inner loop (lib/core/function.py:134-146,158-164) then runs on the returned batch:

        batch = step(table, ids, cache.buffer, off, pitch, h, w)
        G_input = batch.g_input()                        # torch.cat(inputs, 1) for the generator
        tmp = batch.mix(model_G(G_input))                # softmax + mix, differentiable w.r.t. the logits
        loss(model_T(tmp), batch.target, batch.target_weight) ...
    """

    class Batch:
        def __init__(self, **kw):
            self.__dict__.update(kw)

        def g_input(self, dtype=torch.float32):
            from .mix import chains_g_input
            return chains_g_input(self.crop_u8, self.plans, self.gridmask, dtype=dtype)

        def mix(self, logits, out_dtype=torch.float32):
            from .mix import chain_mix_from_logits
            return chain_mix_from_logits(self.crop_u8, self.plans, self.gridmask, logits, out_dtype=out_dtype)

        def inputs(self, dtype=torch.float32):
            """The reference's `inputs` list [clean, autoaug, gridmask] (materialised; for code that still wants it)."""
            g = self.g_input(dtype)
            return [g[:, 0:3], g[:, 3:6], g[:, 6:9]]

    def __init__(self, batch, want_gridmask_targets=True, **kw):
        super().__init__(batch, **kw)
        from . import chains as CH
        self.CH = CH
        self.want_gm_targets = want_gridmask_targets
        B, J = self.B, self.J
        al = lambda v: (v + 15) & ~15
        base = self.nbytes
        self.nbytes = int(self.lib.advmix_chains_step_params_bytes(B, J))
        self.plan_bytes = int(self.lib.advmix_autoaug_plan_bytes(1))
        self.plan_ws = torch.empty(B * 768 * 4, dtype=torch.uint8, device=self.device)
        extra = [("aa_ops", np.int32, (B, 2)), ("aa_mags", np.float32, (B, 2)), ("gm", np.int32, (B, 4))]
        for slot in self.slots:
            host = torch.empty(self.nbytes, dtype=torch.uint8).pin_memory()
            hv, o = host.numpy(), 0
            views = {}
            # same K = 1 sections first (re-created on the larger buffer), then the chain parameters
            for name, v in slot["views"].items():
                n = v.nbytes
                views[name] = hv[o:o + n].view(v.dtype).reshape(v.shape)
                o += al(n)
            assert o == base
            self.gm_offset = None
            for name, dt, shape in extra:
                n = int(np.prod(shape)) * np.dtype(dt).itemsize
                if name == "gm":
                    self.gm_offset = o
                views[name] = hv[o:o + n].view(dt).reshape(shape)
                o += al(n)
            assert o == self.nbytes, (o, self.nbytes)
            slot.update(host=host, views=views, dev=torch.empty(self.nbytes, dtype=torch.uint8, device=self.device))

    def __call__(self, table, ids, src_base, src_off, src_pitch, src_h, src_w, draws=None, chain_draws=None):
        B, J, lib, P = self.B, self.J, self.lib, _lib.ptr
        assert len(ids) == B
        c, s, rot, flip = draws if draws is not None else self.draw(table.centers[ids], table.scales[ids], table.widths[ids])
        W, H = self.image_size
        if chain_draws is not None:
            (ops, mags), gm = chain_draws
        else:
            ops, mags = self.CH.sample_autoaug_batch(B, self.rng)
            gm = self.CH.sample_gridmask_batch(B, H, W, self.rng)
        slot = self.slots[self.next]
        self.next = (self.next + 1) % len(self.slots)
        if slot["event"] is not None:
            slot["event"].synchronize()
        v = slot["views"]
        v["src_off"][:] = src_off; v["src_pitch"][:] = src_pitch; v["src_h"][:] = src_h; v["src_w"][:] = src_w
        v["scale"][:] = s; v["rot"][:] = rot; v["center"][:] = c; v["flip"][:] = flip
        v["joints"][:] = table.joints[ids]; v["vis"][:] = table.vis[ids]
        v["aa_ops"][:] = ops; v["aa_mags"][:] = mags; v["gm"][:] = gm
        dev = self.device
        pbuf = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)         # a fresh buffer: the batch keeps a view of its gridmask section
        pbuf.copy_(slot["host"], non_blocking=True)
        slot["event"] = torch.cuda.Event()
        slot["event"].record()
        Wh, Hh = self.heatmap_size
        M = torch.empty((B, 2, 3), dtype=torch.float64, device=dev)
        crop = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
        plans = torch.empty((B, self.plan_bytes), dtype=torch.uint8, device=dev)
        jo = torch.empty((B, J, 3), dtype=torch.float64, device=dev)
        vo = torch.empty((B, J, 3), dtype=torch.float64, device=dev)
        hm = torch.empty((B, J, Hh, Wh), dtype=torch.float32, device=dev)
        mu = torch.empty((B, J, 2), dtype=torch.float32, device=dev)
        tw = torch.empty((B, J, 1), dtype=torch.float32, device=dev)
        vg = hg = tg = None
        if self.want_gm_targets:
            vg = torch.empty((B, J, 3), dtype=torch.float64, device=dev)
            hg = torch.empty((B, J, Hh, Wh), dtype=torch.float32, device=dev)
            tg = torch.empty((B, J, 1), dtype=torch.float32, device=dev)
        _lib.check(lib.advmix_crop_chains_step(P(src_base), P(pbuf), P(self.perm), P(self.gtab), P(self.jw), P(M), P(crop), P(plans),
                                               P(self.plan_ws), self.plan_ws.numel(), P(jo), P(vo), P(vg), P(hm), P(mu), P(tw), P(hg), P(tg),
                                               B, J, W, H, Hh, Wh, self.sigma, _lib.stream_ptr()), "advmix_crop_chains_step")
        gm_dev = pbuf[self.gm_offset:self.gm_offset + B * 16].view(torch.int32).view(B, 4)
        return AdvMixStep.Batch(crop_u8=crop, plans=plans, gridmask=gm_dev, target=hm, mu=mu, target_weight=tw, target_gridmask=hg,
                                target_weight_gridmask=tg, joints=jo, joints_vis=vo, joints_vis_gridmask=vg, trans=M,
                                center=c, scale=s, rotation=rot, flip=flip, index=ids, autoaug=(ops, mags))
