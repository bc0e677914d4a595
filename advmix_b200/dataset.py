"""Batched, device-resident mirror of lib/dataset/JointsDataset.py.

`AdvMixBatchPipeline(cfg...)(records)` does for a whole batch what the reference's
`JointsDataset.__getitem__` + `default_collate` do per sample in CPU workers
(JointsDataset.py:117-133; get_base :135-223; get_var :225-256; get_clean :258-364):
the random draws, the affine crop, the K=3 chains ['clean', 'autoaug', 'gridmask'] or the
K=1 single-sample path (incl. its optional --random_corruption branch, :284-286), ToTensor+Normalize
and generate_target - and returns exactly the reference's structure (row a8 of SURVEY.md): four lists
of K batched CUDA tensors / metas for sample_times=3, or `(input, [heatmap, mu], target_weight, meta)`
for sample_times=1.

The host only DRAWS: in "reference" mode it consumes `np.random` and `random` in the reference's own
order, so seeding both RNGs reproduces the reference's augmentation sample by sample; "batched" mode
draws the same distributions vectorised from a numpy Generator.  Everything computed from the draws -
half-body boxes, centre / scale bookkeeping (advmix_base_cs), get_affine_transform with cv2's LU,
the crop, the chains, the joints and the heat maps - runs in libadvmix_b200.so.
"""
import random as pyrandom

import numpy as np
import torch

from . import _lib
from . import chains as CH
from . import corruptions as CO
from . import targets as TG
from . import transforms as TF

COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]
COCO_UPPER_BODY = (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10)
COCO_JOINTS_WEIGHT = np.array([1., 1., 1., 1., 1., 1., 1., 1.2, 1.2, 1.5, 1.5, 1., 1., 1.2, 1.2, 1.5, 1.5],
                              dtype=np.float32).reshape((17, 1))
CHAINS = ("clean", "autoaug", "gridmask")         # JointsDataset.py:124
RANDOM_CORRUPTIONS = CO.CORRUPTIONS[:15]          # JointsDataset.py:259-264


class AdvMixBatchPipeline:
    def __init__(self, image_size=(192, 256), heatmap_size=(48, 64), sigma=2, num_joints=17,
                 flip_pairs=COCO_FLIP_PAIRS, upper_body_ids=COCO_UPPER_BODY, scale_factor=0.3, rot_factor=40,
                 flip=True, prob_half_body=0.0, num_joints_half_body=8, is_train=True, sample_times=3,
                 use_different_joints_weight=False, joints_weight=None, random_corruption=False, color_rgb=False,
                 norm_dtype=torch.float32, device="cuda", draw_mode="reference", seed=0):
        self.image_size = np.array(image_size)
        self.heatmap_size = np.array(heatmap_size)
        self.sigma = sigma
        self.num_joints = num_joints
        self.flip_pairs = flip_pairs
        self.upper_body_ids = upper_body_ids
        self.scale_factor, self.rotation_factor, self.flip = scale_factor, rot_factor, flip
        self.prob_half_body, self.num_joints_half_body = prob_half_body, num_joints_half_body
        self.is_train, self.sample_times = is_train, sample_times
        self.joints_weight = (joints_weight if joints_weight is not None else COCO_JOINTS_WEIGHT) \
            if use_different_joints_weight else None
        self.random_corruption = random_corruption
        self.color_rgb = color_rgb                 # cfg.DATASET.COLOR_RGB: cv2.cvtColor(BGR2RGB) after imread (:151-152)
        self.aspect_ratio = image_size[0] * 1.0 / image_size[1]
        self.pixel_std = 200
        self.norm_dtype = norm_dtype
        self.device = torch.device(device)
        self.draw_mode = draw_mode
        self.seed = seed
        self.rng = np.random.default_rng(seed)
        self.step = 0
        self._perm = None
        self._upper = None

    # ---- draws of get_base / get_clean (:167-188), reference RNG order ------------------------
    def _draw_base_reference(self, n_vis):
        """The draws one __getitem__ makes, in its order: (take_half_body, hb_randn, s_factor, rot, flip).
        `np.random.rand()` only when the sample has enough visible joints (short-circuit `and`, :168-169), the
        `np.random.randn()` of half_body_transform (:80) only when that is entered."""
        take, hb = False, 0.0
        sfac, r, flip = 1.0, 0.0, False
        if self.is_train:
            if n_vis > self.num_joints_half_body and np.random.rand() < self.prob_half_body:
                take, hb = True, np.random.randn()
            sf, rf = self.scale_factor, self.rotation_factor
            sfac = np.clip(np.random.randn() * sf + 1, 1 - sf, 1 + sf)
            r = np.clip(np.random.randn() * rf, -rf * 2, rf * 2) if pyrandom.random() <= 0.6 else 0
            if self.flip and pyrandom.random() <= 0.5:
                flip = True
        return take, hb, float(sfac), float(r), flip

    def _draw_base_batched(self, n_vis):
        B = len(n_vis)
        g = self.rng
        take = np.zeros(B, bool); hb = np.zeros(B)
        sfac = np.ones(B); rot = np.zeros(B); flip = np.zeros(B, bool)
        if self.is_train:
            if self.prob_half_body > 0:
                take = (n_vis > self.num_joints_half_body) & (g.random(B) < self.prob_half_body)
                hb = g.standard_normal(B)
            sf, rf = self.scale_factor, self.rotation_factor
            sfac = np.clip(g.standard_normal(B) * sf + 1, 1 - sf, 1 + sf)
            rot = np.where(g.random(B) <= 0.6, np.clip(g.standard_normal(B) * rf, -rf * 2, rf * 2), 0.0)
            if self.flip:
                flip = g.random(B) <= 0.5
        return take, hb, sfac, rot, flip

    def _stage_h2d(self, arrays):
        """The small per-step host arrays go to the device as ONE asynchronous copy out of a pinned ring buffer
        (a pageable `.to(device)` per array would block the host on the stream's previous work, i.e. on the
        previous step's source upload)."""
        offs, total = [], 0
        for a in arrays:
            offs.append(total)
            total += (a.nbytes + 255) & ~255
        ring = getattr(self, "_h2d_ring", None)
        if ring is None or ring[0][0].numel() < total:
            ring = self._h2d_ring = [[torch.empty(max(total, 1), dtype=torch.uint8).pin_memory(), None] for _ in range(4)]
            self._h2d_next = 0
        slot = ring[self._h2d_next]
        self._h2d_next = (self._h2d_next + 1) % len(ring)
        if slot[1] is not None:
            slot[1].synchronize()                      # the copy that last read this buffer has finished
        hbuf = slot[0].numpy()
        for a, o in zip(arrays, offs):
            hbuf[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
        dbuf = torch.empty(total, dtype=torch.uint8, device=self.device)
        dbuf.copy_(slot[0][:total], non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record()
        out = []
        for a, o in zip(arrays, offs):
            out.append(dbuf[o:o + a.nbytes].view(torch.from_numpy(a[:0]).dtype).view(a.shape))
        return out

    def _perm_tensor(self):
        if self._perm is None:
            self._perm = TF.flip_perm(self.num_joints, self.flip_pairs, self.device)
        return self._perm

    def _upper_mask(self):
        if self._upper is None:
            m = torch.zeros(self.num_joints, dtype=torch.uint8)
            m[list(self.upper_body_ids)] = 1
            self._upper = m.to(self.device)
        return self._upper

    def _base_cs(self, rec_c, rec_s, joints, vis, take, hb, sfac, flip, widths_t):
        """advmix_base_cs: half-body boxes, s * factor, c[0] mirror -> (center f32 [B,2], scale f64 [B,2])."""
        lib = _lib.load()
        B = rec_c.shape[0]
        c = torch.empty((B, 2), dtype=torch.float32, device=self.device)
        s = torch.empty((B, 2), dtype=torch.float64, device=self.device)
        _lib.check(lib.advmix_base_cs(_lib.ptr(rec_c), _lib.ptr(rec_s), _lib.ptr(joints), _lib.ptr(vis), _lib.ptr(self._upper_mask()),
                                      _lib.ptr(take), _lib.ptr(hb), _lib.ptr(sfac), _lib.ptr(flip), _lib.ptr(widths_t),
                                      _lib.ptr(c), _lib.ptr(s), B, self.num_joints, float(self.aspect_ratio),
                                      float(self.pixel_std), _lib.stream_ptr()), "advmix_base_cs")
        return c, s

    def _corrupt_sources(self, sources, names, sevs):
        """get_clean's --random_corruption branch (:284-286): corrupt(data_numpy, name, severity) on the FULL source
        image, before the crop.  Sources of one size and one (name, severity) go through one batched call; the
        result is a new dense SourceBatch."""
        B = len(sources)
        hs, ws = sources.heights.cpu().numpy(), sources.widths.cpu().numpy()
        offs, ps = sources.offsets.cpu().numpy(), sources.pitches.cpu().numpy()
        new_off, total = [], 0
        for b in range(B):
            new_off.append(total)
            total += (int(hs[b]) * int(ws[b]) * 3 + 255) // 256 * 256
        out = torch.empty(max(total, 16), dtype=torch.uint8, device=self.device)
        groups = {}
        for b in range(B):
            groups.setdefault((int(hs[b]), int(ws[b]), names[b], int(sevs[b])), []).append(b)
        for (H, W, name, sev), members in groups.items():
            dense = torch.empty((len(members), H, W, 3), dtype=torch.uint8, device=self.device)
            for i, b in enumerate(members):
                p = int(ps[b])
                dense[i] = sources.buffer[int(offs[b]):int(offs[b]) + H * p].view(H, p)[:, :3 * W].reshape(H, W, 3)
            # draws keyed by (pipeline seed, step, sample): np.random is left to the reference-order draws
            res = CO.corrupt_batch(dense, name, sev, seed=(self.seed << 20) ^ self.step, sample_base=members[0])
            for i, b in enumerate(members):
                out[new_off[b]:new_off[b] + H * W * 3] = res[i].reshape(-1)
        dev = self.device
        return TF.SourceBatch(out, torch.tensor(new_off, dtype=torch.int64, device=dev), sources.heights, sources.widths,
                              torch.tensor([int(w) * 3 for w in ws], dtype=torch.int64, device=dev))

    # ---- the batch -----------------------------------------------------------------------------
    def __call__(self, records, sources=None, draws=None, host_sources=None):
        """records: list of db dicts {'image': uint8 HWC ndarray as cv2.imread returns it (BGR; anything if `sources`
        is given), 'center' f32[2], 'scale' f32[2], 'joints_3d' f64[J,3], 'joints_3d_vis' f64[J,3], ...}.
        sources: optional transforms.SourceBatch already resident on the device (already RGB if color_rgb).
        draws: optional explicit (center [B,2], scale [B,2], rot [B], flip [B]) replacing the
        random draws of get_base / get_clean (the centre already mirrored where flip is set).
        host_sources: optional transforms.HostSourceBatch (pinned host images): only the source rows the
        crops of this step read are sent to the device before the crop kernel runs."""
        B = len(records)
        dev = self.device
        if host_sources is not None:
            sources = host_sources.dev
        if sources is None:
            for r in records:
                if r.get("image") is None:          # JointsDataset.py:155-157
                    raise ValueError("Fail to read {}".format(r.get("image_file", "")))
            sources = TF.SourceBatch.from_numpy([r["image"] for r in records], dev, bgr_to_rgb=self.color_rgb)
        widths_np = sources.widths.cpu().numpy() if "width" not in records[0] else np.array([r["width"] for r in records])
        k3 = self.is_train and self.sample_times != 1
        H, W = int(self.image_size[1]), int(self.image_size[0])
        joints_np = np.array([r["joints_3d"] for r in records], dtype=np.float64)        # one C-level conversion (np.stack
        vis_np = np.array([r["joints_3d_vis"] for r in records], dtype=np.float64)       # costs ~0.1 ms more per 256 records)
        rec_c = np.array([r["center"] for r in records], dtype=np.float32).reshape(B, 2)
        rec_s = np.array([r["scale"] for r in records], dtype=np.float32).reshape(B, 2)

        aa = gm = None
        rc_names = rc_sevs = None
        explicit = draws is not None
        if explicit:
            c, s, rot, flip = (np.asarray(d) for d in draws)
            flip = flip.astype(bool)
            if k3:
                aa = CH.sample_autoaug_batch(B, self.rng)
                gm = CH.sample_gridmask_batch(B, H, W, self.rng)
        else:
            n_vis = vis_np[:, :, 0].sum(1)
            if self.draw_mode == "reference":
                take = np.zeros(B, bool); hb = np.zeros(B); sfac = np.ones(B); rot = np.zeros(B); flip = np.zeros(B, bool)
                aa_ops, aa_mags = np.zeros((B, 2), np.int32), np.zeros((B, 2), np.float32)
                gm_params = np.zeros((B, 4), np.int32)
                if self.random_corruption and not k3:
                    rc_names, rc_sevs = [], []
                for b in range(B):                  # per-sample order == one __getitem__ after another
                    if rc_names is not None:        # get_clean :284-286, drawn before everything else
                        rc_names.append(pyrandom.choice(RANDOM_CORRUPTIONS))
                        rc_sevs.append(pyrandom.randint(1, 5))
                    take[b], hb[b], sfac[b], rot[b], flip[b] = self._draw_base_reference(n_vis[b])
                    if k3:
                        o, m = CH.sample_autoaug(1)
                        aa_ops[b], aa_mags[b] = o[0], m[0]
                        gm_params[b] = CH.sample_gridmask(1, H, W)[0]
                if k3:
                    aa, gm = (aa_ops, aa_mags), gm_params
            else:
                if self.random_corruption and not k3:
                    rc_names = [RANDOM_CORRUPTIONS[i] for i in self.rng.integers(0, 15, B)]
                    rc_sevs = list(self.rng.integers(1, 6, B))
                take, hb, sfac, rot, flip = self._draw_base_batched(n_vis)
                if k3:
                    aa = CH.sample_autoaug_batch(B, self.rng)
                    gm = CH.sample_gridmask_batch(B, H, W, self.rng)

        if rc_names is not None:
            sources = self._corrupt_sources(sources, rc_names, rc_sevs)

        self.last_h2d_bytes = 0
        if host_sources is not None:
            heights = np.array([r["height"] for r in records]) if "height" in records[0] else sources.heights.cpu().numpy()
            if explicit:
                bc, bs = c, s
                whole = np.zeros(B, bool)
            else:
                # conservative boxes from the host's view of the draws; half-body samples (their box is computed on
                # the device) ship the whole image
                bs = rec_s.astype(np.float64) * np.asarray(sfac)[:, None]
                bc = rec_c.copy()
                bc[:, 0] = np.where(flip, widths_np - bc[:, 0] - 1, bc[:, 0])
                whole = np.asarray(take, bool)
            lo, hi, blo, bhi, quad = TF.source_boxes(bc, bs, rot, flip, heights, widths_np, host_sources.pitches_h, self.image_size)
            if whole.any():
                lo[whole] = 0; hi[whole] = heights[whole]; blo[whole] = 0
                bhi[whole] = (host_sources.pitches_h[whole] & ~15)
                wq, hq = widths_np[whole].astype(np.float32), heights[whole].astype(np.float32)
                quad[whole, :, 0] = np.stack([0 * wq, wq, wq, 0 * wq], 1)
                quad[whole, :, 1] = np.stack([0 * hq, 0 * hq, hq, hq], 1)
            self.last_h2d_bytes = host_sources.upload_boxes(lo, hi, blo, bhi, quad)

        if explicit:
            staged = [np.ascontiguousarray(c, np.float32), np.ascontiguousarray(s)]      # scale keeps numpy's dtype (f32 or f64)
        else:
            staged = [rec_c, rec_s, np.ascontiguousarray(take, np.uint8), np.ascontiguousarray(hb, np.float64),
                      np.ascontiguousarray(sfac, np.float64), np.ascontiguousarray(widths_np, np.int32)]
        n_head = len(staged)
        staged += [np.ascontiguousarray(rot, np.float64), np.ascontiguousarray(flip, np.uint8), joints_np, vis_np]
        if k3:                                              # chain parameters ride in the same pinned copy
            staged += [np.ascontiguousarray(aa[0], np.int32), np.ascontiguousarray(aa[1], np.float32),
                       np.ascontiguousarray(gm, np.int32)]
        staged = self._stage_h2d(staged)
        r_t, f_t, joints, vis = staged[n_head:n_head + 4]
        if explicit:
            c_t, s_t = staged[0], staged[1]
        else:
            rc_t, rs_t, take_t, hb_t, sfac_t, w_t = staged[:6]
            any_hb = bool(np.any(take))
            c_t, s_t = self._base_cs(rc_t, rs_t, joints, vis, take_t if any_hb else None, hb_t if any_hb else None,
                                     sfac_t if self.is_train else None, f_t, w_t)
            if not self.is_train:
                s_t = s_t.to(torch.float32)          # evaluation keeps the record's float32 scale (selects numpy's float32 products)
        if k3:
            aa, gm = (staged[n_head + 4], staged[n_head + 5]), staged[n_head + 6]

        trans = TF.get_affine_transform(c_t, s_t, r_t, self.image_size)
        crop_u8, clean = TF.warp_affine(sources, trans, self.image_size, flip=f_t, want_u8=k3,
                                        norm_dtype=self.norm_dtype)
        joints, vis = TF.fliplr_affine_joints(joints, vis, trans, flip=f_t, widths=sources.widths,
                                              perm=self._perm_tensor())
        gt = dict(image_size=self.image_size, heatmap_size=self.heatmap_size, sigma=self.sigma,
                  joints_weight=self.joints_weight)
        meta = {"joints": joints, "joints_vis": vis, "center": c_t, "scale": s_t, "rotation": r_t,
                "flip": f_t, "trans": trans,
                "image": [r.get("image_file", "") for r in records],
                "filename": [r.get("filename", "") for r in records],
                "imgnum": [r.get("imgnum", 0) for r in records],
                "score": [r.get("score", 1) for r in records]}
        if rc_names is not None:
            meta["random_corruption"] = list(zip(rc_names, [int(v) for v in rc_sevs]))
        self.step += 1
        if not k3:
            target, tw = TG.generate_target(joints, vis, **gt)
            return clean, target, tw, meta

        inputs, tgts, tws, metas = [], [], [], []
        # chain 'clean'
        t0, w0 = TG.generate_target(joints, vis, **gt)
        inputs.append(clean); tgts.append(t0[0]); tws.append(w0); metas.append(meta)
        # chain 'autoaug' (PIL sub-policy on the uint8 crop, then transform)
        _, aug = CH.autoaug(crop_u8, aa[0], aa[1], norm_dtype=self.norm_dtype, want_u8=False)
        inputs.append(aug); tgts.append(t0[0]); tws.append(w0); metas.append(meta)
        # chain 'gridmask' (on the normalised tensor; drops visibility of masked joints)
        gimg, gvis = CH.gridmask(clean, gm, joints, vis)
        t2, w2 = TG.generate_target(joints, gvis, **gt)
        # the reference shares ONE meta dict between the chains and gridmask mutates its
        # joints_vis in place: after __getitem__ every meta shows the masked vis.
        meta["joints_vis"] = gvis
        inputs.append(gimg); tgts.append(t2[0]); tws.append(w2); metas.append(meta)
        self.last_chain_params = {"crop_u8": crop_u8, "autoaug": aa, "gridmask": gm}     # for the fused chain+mix (advmix_b200.mix)
        return inputs, tgts, tws, metas


def corruption_chains(crop_u8, names, severities, seed, sample_base=0, norm_dtype=torch.float32):
    """Target-workload chains (BASELINE config 3): per-sample (corruption, severity) drawn from the
    15x5 set.  names/severities: length-B sequences.  Samples are grouped by (op, severity) and each
    group is one batched call through the `idx` indirection (no gather copies).
    Returns (uint8 [B,H,W,3], normalised [B,3,H,W])."""
    B = crop_u8.shape[0]
    out = torch.empty_like(crop_u8)
    groups = {}
    for b, (n, s) in enumerate(zip(names, severities)):
        groups.setdefault((n, int(s)), []).append(b)
    for (n, s), members in groups.items():
        idx = torch.tensor(members, dtype=torch.int32, device=crop_u8.device)
        CO.corrupt_batch(crop_u8, n, s, seed=seed, sample_base=sample_base, idx=idx, out=out)
    return out, TF.to_tensor_normalize(out, dtype=norm_dtype)
