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

    def device_rows(self, device):
        """joints_3d / joints_3d_vis of the whole table as device tensors (uploaded once per device): the step then sends row
        indices instead of 2 x B x J x 3 doubles (advmix_crop_targets_step_rec)."""
        key = str(torch.device(device))
        cache = self.__dict__.setdefault("_dev_rows", {})
        if key not in cache:
            cache[key] = (torch.from_numpy(self.joints).to(device), torch.from_numpy(self.vis).to(device))
        return cache[key]


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
        self._upload_event = torch.cuda.Event()
        self._upload_pending = False

    def take_upload_event(self):
        """Event recorded after the uploads `ensure` issued since the last call (None if there were none): a step that runs on
        another stream than the one `ensure` was called on must wait for it (CropTargetsStep(..., after=...))."""
        if not self._upload_pending:
            return None
        self._upload_pending = False
        return self._upload_event

    def ensure(self, ids, fetch):
        miss = ids[self.off[ids] < 0]
        if miss.size:
            for i in np.unique(miss):
                img = fetch(int(i))
                H, W = int(img.shape[0]), int(img.shape[1])
                pitch = (3 * W + 15) // 16 * 16
                need = (H * pitch + 255) // 256 * 256
                if self.used + need > self.buffer.numel():
                    torch.cuda.synchronize(self.device)                # steps still reading the old contents (possibly on prefetch streams)
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
            self._upload_event.record()
            self._upload_pending = True
        return self.off[ids], self.pitch[ids], self.h[ids], self.w[ids]


class CropTargetsStep:
    """One K = 1 training step per call.  Host work per step (measured on the bench box: ~70 us for 256 samples, below the
    ~90 us the device needs, so the step is device-bound): slice pre-drawn augmentation draws, gather the batch's centre / scale
    rows, write ~14 KB into a pinned buffer, one H2D copy, one C call.

    record_rows="device" (default) keeps the table's joints / visibility rows on the device and sends row indices
    (advmix_crop_targets_step_rec); "host" gathers them on the host into the packed buffer (advmix_crop_targets_step).
    out_ring = 0 returns freshly allocated tensors every step; out_ring = n >= 2 cycles through n preallocated output sets
    (a batch stays valid until n - 1 further steps have been issued) and saves the seven allocations per step.
    graph=True (needs out_ring == ring >= 2): the parameter copy and the library call of every ring entry are
    captured into a CUDA graph the first time the entry is used and replayed afterwards - one launch call per step instead of a
    copy, four kernel launches and four event operations.
    prefetch_streams=2 (graph mode, even ring >= 4): consecutive steps are issued on two alternating streams owned by the step object, so
    the parameter copy, the matrices / joints / heat-map kernels and the launch gaps of step i+1 run under the crop kernel of step
    i (what a DataLoader's prefetching does for the reference).  The caller's current stream waits for the step it is handed;
    an output set is reused only after the work the caller queued on that batch has finished (events, no host synchronisation)."""

    DRAW_CHUNK = 64          # steps of augmentation draws generated per numpy call

    def __init__(self, batch, image_size=(192, 256), heatmap_size=(48, 64), sigma=2, num_joints=17, flip_pairs=None,
                 scale_factor=0.3, rot_factor=40, flip=True, is_train=True, joints_weight=None, norm_dtype=torch.float32,
                 device="cuda", seed=0, ring=4, record_rows="device", out_ring=0, graph=False, prefetch_streams=1):
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
        assert record_rows in ("device", "host")
        self.rec_rows = record_rows == "device"
        self._setup_slots(ring)
        self.out_ring = int(out_ring)
        self._outs, self._out_next = [], 0
        self._draws, self._draw_k = None, 0
        self.graph = bool(graph)
        if self.graph and not (self.out_ring == len(self.slots) and self.out_ring >= 2):
            raise ValueError("graph=True needs out_ring == ring >= 2")
        self._graphs = {}
        self.prefetch_streams = int(prefetch_streams)
        if self.prefetch_streams not in (1, 2) or (self.prefetch_streams == 2 and not (self.graph and len(self.slots) % 2 == 0 and len(self.slots) >= 4)):
            raise ValueError("prefetch_streams=2 needs graph=True and an even ring >= 4")
        if self.prefetch_streams == 2:
            n = len(self.slots)
            self._pstreams = [torch.cuda.Stream(self.device) for _ in range(2)]
            self._consumed = [torch.cuda.Event() for _ in range(n)]      # the caller's work on the batch of ring entry e has been queued before this
            self._consumed_set = [False] * n
            self._last = None

    def _sections(self):
        B, J = self.B, self.J
        head = [("src_off", np.int64, (B,)), ("src_pitch", np.int64, (B,)), ("src_h", np.int32, (B,)), ("src_w", np.int32, (B,)),
                ("scale", np.float64, (B, 2)), ("rot", np.float64, (B,)), ("center", np.float32, (B, 2)), ("flip", np.uint8, (B,))]
        if self.rec_rows:
            return head + [("rec_idx", np.int32, (B,))]
        return head + [("joints", np.float64, (B, J, 3)), ("vis", np.float64, (B, J, 3))]

    def _params_bytes(self):
        fn = self.lib.advmix_step_rec_params_bytes if self.rec_rows else self.lib.advmix_step_params_bytes
        return int(fn(self.B, self.J))

    def _setup_slots(self, ring):
        # pinned staging ring + numpy views of its sections (layouts documented at advmix_crop_targets_step[_rec])
        self.nbytes = self._params_bytes()
        al = lambda v: (v + 15) & ~15
        self.slots = []
        for _ in range(ring):
            host = torch.empty(self.nbytes, dtype=torch.uint8).pin_memory()
            hv, views, o = host.numpy(), {}, 0
            for name, dt, shape in self._sections():
                n = int(np.prod(shape)) * np.dtype(dt).itemsize
                views[name] = hv[o:o + n].view(dt).reshape(shape)
                o += al(n)
            assert o <= self.nbytes, (o, self.nbytes)
            self.slots.append({"host": host, "views": views, "dev": torch.empty(self.nbytes, dtype=torch.uint8, device=self.device),
                               "event": torch.cuda.Event(), "used": False})
        self.next = 0

    def _refill_draws(self):
        C, B, g = self.DRAW_CHUNK, self.B, self.rng
        sf, rf = self.scale_factor, self.rot_factor
        zs = np.clip(g.standard_normal((C, B)) * sf + 1, 1 - sf, 1 + sf)
        rot = np.where(g.random((C, B)) <= 0.6, np.clip(g.standard_normal((C, B)) * rf, -rf * 2, rf * 2), 0.0)
        flip = (g.random((C, B)) <= 0.5) if self.flip else np.zeros((C, B), bool)
        self._draws, self._draw_k = (zs, rot, flip), 0

    def draw(self, centers, scales, widths):
        """The augmentation draws of JointsDataset.py:177-188 for the whole batch -> (center f32 [B,2] already mirrored where
        flipped, scale f64 [B,2], rot f64 [B], flip bool [B]).  The random numbers come from a block drawn DRAW_CHUNK steps at
        a time (one numpy call per distribution per 64 steps instead of per step)."""
        B = len(centers)
        c = centers.copy()
        s = scales.astype(np.float64)
        if not self.is_train:
            return c, s, np.zeros(B), np.zeros(B, bool)
        assert B == self.B
        if self._draws is None or self._draw_k == self.DRAW_CHUNK:
            self._refill_draws()
        k = self._draw_k
        self._draw_k += 1
        zs, rot, flip = self._draws[0][k], self._draws[1][k], self._draws[2][k]
        s *= zs[:, None]
        if self.flip:
            c[:, 0] = np.where(flip, widths - c[:, 0] - 1, c[:, 0])
        return c, s, rot, flip

    def _alloc_outputs(self):
        B, J, dev = self.B, self.J, self.device
        W, H = self.image_size
        Wh, Hh = self.heatmap_size
        t = {"M": torch.empty((B, 2, 3), dtype=torch.float64, device=dev),
             "inp": torch.empty((B, 3, H, W), dtype=self.norm_dtype, device=dev),
             "jo": torch.empty((B, J, 3), dtype=torch.float64, device=dev),
             "vo": torch.empty((B, J, 3), dtype=torch.float64, device=dev),
             "hm": torch.empty((B, J, Hh, Wh), dtype=torch.float32, device=dev),
             "mu": torch.empty((B, J, 2), dtype=torch.float32, device=dev),
             "tw": torch.empty((B, J, 1), dtype=torch.float32, device=dev)}
        t["ptrs"] = {k: _lib.ptr(v) for k, v in t.items()}
        return t

    def _outputs(self):
        if self.out_ring < 2:
            return self._alloc_outputs()
        if len(self._outs) < self.out_ring:
            self._outs.append(self._alloc_outputs())
            return self._outs[-1]
        o = self._outs[self._out_next]
        self._out_next = (self._out_next + 1) % self.out_ring
        return o

    def _fill_extra(self, views, extra):
        pass

    def _fill(self, table, ids, src_off, src_pitch, src_h, src_w, draws, extra=None):
        """Pick the next pinned slot and fill it; returns (slot index, slot, c, s, rot, flip)."""
        c, s, rot, flip = draws if draws is not None else self.draw(table.centers[ids], table.scales[ids], table.widths[ids])
        i = self.next
        slot = self.slots[i]
        self.next = (i + 1) % len(self.slots)
        if slot["used"]:
            slot["event"].synchronize()                  # the copy that last read this pinned buffer has finished
        v = slot["views"]
        v["src_off"][:] = src_off; v["src_pitch"][:] = src_pitch; v["src_h"][:] = src_h; v["src_w"][:] = src_w
        v["scale"][:] = s; v["rot"][:] = rot; v["center"][:] = c; v["flip"][:] = flip
        if self.rec_rows:
            v["rec_idx"][:] = ids
        else:
            np.take(table.joints, ids, axis=0, out=v["joints"]); np.take(table.vis, ids, axis=0, out=v["vis"])
        self._fill_extra(v, extra)
        return i, slot, c, s, rot, flip

    def _launch(self, table, src_base, slot, o, stream):
        B, J, lib, P = self.B, self.J, self.lib, _lib.ptr
        q = o["ptrs"]
        W, H = self.image_size
        Wh, Hh = self.heatmap_size
        sp = _lib.C.c_void_p(stream.cuda_stream)
        slot["dev"].copy_(slot["host"], non_blocking=True)
        if self.rec_rows:
            rj, rv = table.device_rows(self.device)
            _lib.check(lib.advmix_crop_targets_step_rec(P(src_base), P(slot["dev"]), P(rj), P(rv), P(self.perm), P(self.lut), P(self.gtab),
                                                        P(self.jw), q["M"], q["inp"], _lib.dtype_code(self.norm_dtype), q["jo"], q["vo"],
                                                        q["hm"], q["mu"], q["tw"], B, J, W, H, Hh, Wh, self.sigma, sp),
                       "advmix_crop_targets_step_rec")
        else:
            _lib.check(lib.advmix_crop_targets_step(P(src_base), P(slot["dev"]), P(self.perm), P(self.lut), P(self.gtab), P(self.jw),
                                                    q["M"], q["inp"], _lib.dtype_code(self.norm_dtype), q["jo"], q["vo"], q["hm"], q["mu"],
                                                    q["tw"], B, J, W, H, Hh, Wh, self.sigma, sp), "advmix_crop_targets_step")

    def _run(self, table, ids, src_base, src_off, src_pitch, src_h, src_w, draws, after, extra=None):
        """Stage the parameters and issue the step (eagerly, or as the replay of the ring entry's CUDA graph, on the caller's
        stream or on the next prefetch stream).  Returns (slot, outputs, c, s, rot, flip)."""
        assert len(ids) == self.B
        i, slot, c, s, rot, flip = self._fill(table, ids, src_off, src_pitch, src_h, src_w, draws, extra)
        stream = torch.cuda.current_stream(self.device)
        if self.graph and not torch.cuda.is_current_stream_capturing():
            if len(self._outs) < self.out_ring:
                self._outs.append(self._alloc_outputs())
            o = self._outs[i]                               # ring entry i: pinned slot i, device buffer i, output set i
            key = (i, src_base.data_ptr(), id(table))
            g = self._graphs.get(key)
            issue = stream
            if self.prefetch_streams == 2:
                issue = self._pstreams[i & 1]
                if self._last is not None:                  # everything the caller queued on the previous batch is on `stream` by now
                    self._consumed[self._last].record(stream)
                    self._consumed_set[self._last] = True
                if self._consumed_set[i]:
                    issue.wait_event(self._consumed[i])     # output set i is free once the caller's work on its last batch is done
                else:
                    issue.wait_stream(stream)               # first round: order after whatever produced the inputs (cache uploads)
                if after is not None:
                    issue.wait_event(after)
                self._last = i
                torch.cuda.set_stream(issue)
            try:
                if g is None:
                    self._launch(table, src_base, slot, o, issue)           # first use: eager (one-time set-up inside the library), then capture
                    torch.cuda.synchronize(self.device)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._launch(table, src_base, slot, o, torch.cuda.current_stream(self.device))
                    self._graphs[key] = g
                else:
                    g.replay()
            finally:
                if issue is not stream:
                    torch.cuda.set_stream(stream)
            slot["event"].record(issue)       # also the "step of ring entry i has finished" event the caller's stream waits for
            slot["used"] = True
            if issue is not stream:
                stream.wait_event(slot["event"])
        else:
            o = self._outputs()
            self._launch(table, src_base, slot, o, stream)
            slot["event"].record(stream)
            slot["used"] = True
        return slot, o, c, s, rot, flip

    def __call__(self, table, ids, src_base, src_off, src_pitch, src_h, src_w, draws=None, after=None):
        """table: RecordTable; ids: int array [B] of dataset indices; src_*: where the decoded sources of these samples live
        (SourceCache.ensure(...) output, or a SourceBatch's fields as numpy arrays) - src_base is the device buffer.
        after: a CUDA event the step must wait for (SourceCache.take_upload_event()); only needed with prefetch_streams=2, where
        the step does not run on the caller's stream."""
        slot, o, c, s, rot, flip = self._run(table, ids, src_base, src_off, src_pitch, src_h, src_w, draws, after)
        meta = {"joints": o["jo"], "joints_vis": o["vo"], "center": c, "scale": s, "rotation": rot, "flip": flip, "trans": o["M"], "index": ids}
        return o["inp"], [o["hm"], o["mu"]], o["tw"], meta


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

    CHAIN_DRAW_CHUNK = 64

    def __init__(self, batch, want_gridmask_targets=True, **kw):
        kw["record_rows"] = "host"                        # advmix_crop_chains_step takes the rows in the packed buffer
        from . import chains as CH
        self.CH = CH
        self.want_gm_targets = want_gridmask_targets
        super().__init__(batch, **kw)
        self.plan_bytes = int(self.lib.advmix_autoaug_plan_bytes(1))
        self.plan_ws = torch.empty(self.B * 768 * 4, dtype=torch.uint8, device=self.device)
        self._chain_draws, self._chain_k = None, 0

    def _sections(self):
        B = self.B
        return super()._sections() + [("aa_ops", np.int32, (B, 2)), ("aa_mags", np.float32, (B, 2)), ("gm", np.int32, (B, 4))]

    def _params_bytes(self):
        return int(self.lib.advmix_chains_step_params_bytes(self.B, self.J))

    def _setup_slots(self, ring):
        super()._setup_slots(ring)
        v = self.slots[0]["views"]["gm"]
        self.gm_offset = v.__array_interface__["data"][0] - self.slots[0]["host"].numpy().__array_interface__["data"][0]

    def _fill_extra(self, v, extra):
        (ops, mags), gm = extra
        v["aa_ops"][:] = ops; v["aa_mags"][:] = mags; v["gm"][:] = gm

    def draw_chains(self):
        """Autoaug sub-policy / magnitude and gridmask parameters for one batch (advaug.py:53-60, 118-146), taken from a block drawn
        CHAIN_DRAW_CHUNK steps at a time."""
        B, C = self.B, self.CHAIN_DRAW_CHUNK
        if self._chain_draws is None or self._chain_k == C:
            W, H = self.image_size
            ops, mags = self.CH.sample_autoaug_batch(C * B, self.rng)
            gm = self.CH.sample_gridmask_batch(C * B, H, W, self.rng)
            self._chain_draws, self._chain_k = (ops.reshape(C, B, 2), mags.reshape(C, B, 2), gm.reshape(C, B, 4)), 0
        k = self._chain_k
        self._chain_k += 1
        d = self._chain_draws
        return (d[0][k], d[1][k]), d[2][k]

    def _alloc_outputs(self):
        B, J, dev = self.B, self.J, self.device
        W, H = self.image_size
        Wh, Hh = self.heatmap_size
        t = {"M": torch.empty((B, 2, 3), dtype=torch.float64, device=dev),
             "crop": torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev),
             "plans": torch.empty((B, self.plan_bytes), dtype=torch.uint8, device=dev),
             "jo": torch.empty((B, J, 3), dtype=torch.float64, device=dev),
             "vo": torch.empty((B, J, 3), dtype=torch.float64, device=dev),
             "hm": torch.empty((B, J, Hh, Wh), dtype=torch.float32, device=dev),
             "mu": torch.empty((B, J, 2), dtype=torch.float32, device=dev),
             "tw": torch.empty((B, J, 1), dtype=torch.float32, device=dev),
             "vg": None, "hg": None, "tg": None}
        if self.want_gm_targets:
            t["vg"] = torch.empty((B, J, 3), dtype=torch.float64, device=dev)
            t["hg"] = torch.empty((B, J, Hh, Wh), dtype=torch.float32, device=dev)
            t["tg"] = torch.empty((B, J, 1), dtype=torch.float32, device=dev)
        # the batch keeps a view of the gridmask section of ITS parameter buffer: with an output ring the device buffer of the ring
        # entry, otherwise a buffer of its own
        t["pbuf"] = None if self.out_ring >= 2 else torch.empty(self.nbytes, dtype=torch.uint8, device=dev)
        t["ptrs"] = {k: _lib.ptr(v) for k, v in t.items() if k != "ptrs"}
        return t

    def _launch(self, table, src_base, slot, o, stream):
        B, J, lib, P = self.B, self.J, self.lib, _lib.ptr
        q = o["ptrs"]
        W, H = self.image_size
        Wh, Hh = self.heatmap_size
        pbuf = slot["dev"] if o["pbuf"] is None else o["pbuf"]
        pbuf.copy_(slot["host"], non_blocking=True)
        _lib.check(lib.advmix_crop_chains_step(P(src_base), P(pbuf), P(self.perm), P(self.gtab), P(self.jw), q["M"], q["crop"], q["plans"],
                                               P(self.plan_ws), self.plan_ws.numel(), q["jo"], q["vo"], q["vg"], q["hm"], q["mu"], q["tw"],
                                               q["hg"], q["tg"], B, J, W, H, Hh, Wh, self.sigma, _lib.C.c_void_p(stream.cuda_stream)),
                   "advmix_crop_chains_step")

    def __call__(self, table, ids, src_base, src_off, src_pitch, src_h, src_w, draws=None, chain_draws=None, after=None):
        B = self.B
        extra = chain_draws if chain_draws is not None else self.draw_chains()
        slot, o, c, s, rot, flip = self._run(table, ids, src_base, src_off, src_pitch, src_h, src_w, draws, after, extra)
        pbuf = slot["dev"] if o["pbuf"] is None else o["pbuf"]
        gm_dev = pbuf[self.gm_offset:self.gm_offset + B * 16].view(torch.int32).view(B, 4)
        return AdvMixStep.Batch(crop_u8=o["crop"], plans=o["plans"], gridmask=gm_dev, target=o["hm"], mu=o["mu"], target_weight=o["tw"],
                                target_gridmask=o["hg"], target_weight_gridmask=o["tg"], joints=o["jo"], joints_vis=o["vo"],
                                joints_vis_gridmask=o["vg"], trans=o["M"], center=c, scale=s, rotation=rot, flip=flip, index=ids,
                                autoaug=extra[0])
