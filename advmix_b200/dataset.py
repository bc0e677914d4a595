"""Batched, device-resident mirror of lib/dataset/JointsDataset.py.

`AdvMixBatchPipeline(cfg...)(records)` does for a whole batch what the reference's
`JointsDataset.__getitem__` + `default_collate` do per sample in CPU workers
(JointsDataset.py:117-133; get_base :135-223; get_var :225-256; get_clean :258-364):
the random draws, the affine crop, the K=3 chains ['clean', 'autoaug', 'gridmask'] or the
K=1 single-sample path, ToTensor+Normalize and generate_target - and returns exactly the
reference's structure (row a8 of SURVEY.md): four lists of K batched CUDA tensors / metas for
sample_times=3, or `(input, [heatmap, mu], target_weight, meta)` for sample_times=1.

The draws consume `np.random` and `random` in the reference's own order ("reference" mode),
so seeding both RNGs reproduces the reference's augmentation sample by sample; "batched"
mode draws the same distributions vectorised from a numpy Generator.
All pixel / heat-map work runs in libadvmix_b200.so; only the O(J) per-sample bookkeeping
(half-body box, draw bookkeeping) is host numpy, as in the reference.
"""
import random as pyrandom

import numpy as np
import torch

from . import chains as CH
from . import corruptions as CO
from . import targets as TG
from . import transforms as TF

COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]
COCO_UPPER_BODY = (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10)
COCO_JOINTS_WEIGHT = np.array([1., 1., 1., 1., 1., 1., 1., 1.2, 1.2, 1.5, 1.5, 1., 1., 1.2, 1.2, 1.5, 1.5],
                              dtype=np.float32).reshape((17, 1))
CHAINS = ("clean", "autoaug", "gridmask")         # JointsDataset.py:124


def xywh2cs(x, y, w, h, aspect_ratio, pixel_std=200):
    """lib/dataset/coco.py:205-220."""
    center = np.zeros((2), dtype=np.float32)
    center[0] = x + w * 0.5
    center[1] = y + h * 0.5
    if w > aspect_ratio * h:
        h = w * 1.0 / aspect_ratio
    elif w < aspect_ratio * h:
        w = h * aspect_ratio
    scale = np.array([w * 1.0 / pixel_std, h * 1.0 / pixel_std], dtype=np.float32)
    if center[0] != -1:
        scale = scale * 1.25
    return center, scale


class AdvMixBatchPipeline:
    def __init__(self, image_size=(192, 256), heatmap_size=(48, 64), sigma=2, num_joints=17,
                 flip_pairs=COCO_FLIP_PAIRS, upper_body_ids=COCO_UPPER_BODY, scale_factor=0.3, rot_factor=40,
                 flip=True, prob_half_body=0.0, num_joints_half_body=8, is_train=True, sample_times=3,
                 use_different_joints_weight=False, joints_weight=None, random_corruption=False,
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
        self.aspect_ratio = image_size[0] * 1.0 / image_size[1]
        self.pixel_std = 200
        self.norm_dtype = norm_dtype
        self.device = torch.device(device)
        self.draw_mode = draw_mode
        self.rng = np.random.default_rng(seed)
        self.step = 0
        self._perm = None

    # ---- JointsDataset.half_body_transform (:69-111), host numpy like the reference ----------
    def half_body_transform(self, joints, joints_vis):
        upper_joints, lower_joints = [], []
        for joint_id in range(self.num_joints):
            if joints_vis[joint_id][0] > 0:
                (upper_joints if joint_id in self.upper_body_ids else lower_joints).append(joints[joint_id])
        if np.random.randn() < 0.5 and len(upper_joints) > 2:
            selected_joints = upper_joints
        else:
            selected_joints = lower_joints if len(lower_joints) > 2 else upper_joints
        if len(selected_joints) < 2:
            return None, None
        selected_joints = np.array(selected_joints, dtype=np.float32)
        center = selected_joints.mean(axis=0)[:2]
        left_top = np.amin(selected_joints, axis=0)
        right_bottom = np.amax(selected_joints, axis=0)
        w = right_bottom[0] - left_top[0]
        h = right_bottom[1] - left_top[1]
        if w > self.aspect_ratio * h:
            h = w * 1.0 / self.aspect_ratio
        elif w < self.aspect_ratio * h:
            w = h * self.aspect_ratio
        scale = np.array([w * 1.0 / self.pixel_std, h * 1.0 / self.pixel_std], dtype=np.float32)
        return center, scale * 1.5

    # ---- draws of get_base / get_clean (:167-188), reference RNG order ------------------------
    def _draw_base(self, rec, width):
        c = np.array(rec["center"], dtype=np.float32).copy()
        s = np.array(rec["scale"], dtype=np.float32).copy()
        r, flip = 0, False
        if self.is_train:
            jv = rec["joints_3d_vis"]
            if np.sum(jv[:, 0]) > self.num_joints_half_body and np.random.rand() < self.prob_half_body:
                c_hb, s_hb = self.half_body_transform(rec["joints_3d"], jv)
                if c_hb is not None and s_hb is not None:
                    c, s = c_hb, s_hb
            sf, rf = self.scale_factor, self.rotation_factor
            s = s * np.clip(np.random.randn() * sf + 1, 1 - sf, 1 + sf)
            r = np.clip(np.random.randn() * rf, -rf * 2, rf * 2) if pyrandom.random() <= 0.6 else 0
            if self.flip and pyrandom.random() <= 0.5:
                flip = True
                c[0] = width - c[0] - 1
        return c, s, float(r), flip

    def _draw_batch_vectorised(self, records, widths):
        B = len(records)
        g = self.rng
        c = np.stack([np.asarray(r["center"], np.float32) for r in records]).copy()
        s = np.stack([np.asarray(r["scale"], np.float32) for r in records]).copy()
        rot = np.zeros(B)
        flip = np.zeros(B, bool)
        if self.is_train:
            sf, rf = self.scale_factor, self.rotation_factor
            s = (s * np.clip(g.standard_normal(B) * sf + 1, 1 - sf, 1 + sf)[:, None]).astype(np.float32)
            rot = np.where(g.random(B) <= 0.6, np.clip(g.standard_normal(B) * rf, -rf * 2, rf * 2), 0.0)
            if self.flip:
                flip = g.random(B) <= 0.5
                c[:, 0] = np.where(flip, widths - c[:, 0] - 1, c[:, 0])
        return c, s, rot, flip

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

    # ---- the batch -----------------------------------------------------------------------------
    def __call__(self, records, sources=None, draws=None, host_sources=None):
        """records: list of db dicts {'image': uint8 HWC ndarray (or anything if `sources` given),
        'center' f32[2], 'scale' f32[2], 'joints_3d' f64[J,3], 'joints_3d_vis' f64[J,3], ...}.
        sources: optional transforms.SourceBatch already resident on the device.
        draws: optional explicit (center [B,2], scale [B,2], rot [B], flip [B]) replacing the
        random draws of get_base / get_clean (the centre already mirrored where flip is set).
        host_sources: optional transforms.HostSourceBatch (pinned host images): only the source rows the
        crops of this step read are sent to the device before the crop kernel runs."""
        B = len(records)
        dev = self.device
        if host_sources is not None:
            sources = host_sources.dev
        if sources is None:
            sources = TF.SourceBatch.from_numpy([r["image"] for r in records], dev)
        widths_np = sources.widths.cpu().numpy() if "width" not in records[0] else np.array([r["width"] for r in records])
        k3 = self.is_train and self.sample_times != 1

        aa = gm = None
        if draws is not None:
            c, s, rot, flip = (np.asarray(d) for d in draws)
            flip = flip.astype(bool)
            if k3:
                H, W = int(self.image_size[1]), int(self.image_size[0])
                aa = CH.sample_autoaug_batch(B, self.rng)
                gm = CH.sample_gridmask_batch(B, H, W, self.rng)
        elif self.draw_mode == "reference":
            cs, ss, rs, fs = [], [], [], []
            aa_ops, aa_mags = np.zeros((B, 2), np.int32), np.zeros((B, 2), np.float32)
            gm_params = np.zeros((B, 4), np.int32)
            H, W = int(self.image_size[1]), int(self.image_size[0])
            for b, rec in enumerate(records):       # per-sample order == one __getitem__ after another
                c, s, r, f = self._draw_base(rec, int(widths_np[b]))
                cs.append(c); ss.append(s); rs.append(r); fs.append(f)
                if k3:
                    o, m = CH.sample_autoaug(1)
                    aa_ops[b], aa_mags[b] = o[0], m[0]
                    gm_params[b] = CH.sample_gridmask(1, H, W)[0]
            c, s, rot, flip = np.stack(cs), np.stack(ss), np.array(rs, np.float64), np.array(fs, bool)
            if k3:
                aa, gm = (aa_ops, aa_mags), gm_params
        else:
            c, s, rot, flip = self._draw_batch_vectorised(records, widths_np)
            if k3:
                H, W = int(self.image_size[1]), int(self.image_size[0])
                aa = CH.sample_autoaug_batch(B, self.rng)
                gm = CH.sample_gridmask_batch(B, H, W, self.rng)

        self.last_h2d_bytes = 0
        if host_sources is not None:
            heights = np.array([r["height"] for r in records]) if "height" in records[0] else sources.heights.cpu().numpy()
            widths = np.array([r["width"] for r in records]) if "width" in records[0] else sources.widths.cpu().numpy()
            lo, hi, blo, bhi, quad = TF.source_boxes(c, s, rot, flip, heights, widths, host_sources.pitches_h, self.image_size)
            self.last_h2d_bytes = host_sources.upload_boxes(lo, hi, blo, bhi, quad)
        staged = [
            np.ascontiguousarray(c, np.float32), np.ascontiguousarray(s),       # scale keeps numpy's dtype (f32 or f64)
            np.ascontiguousarray(rot, np.float64), flip.astype(np.uint8),
            np.array([r["joints_3d"] for r in records], dtype=np.float64),          # one C-level conversion (np.stack costs
            np.array([r["joints_3d_vis"] for r in records], dtype=np.float64)]      # ~0.1 ms of Python per 256 records)
        if k3:                                              # chain parameters ride in the same pinned copy
            staged += [np.ascontiguousarray(aa[0], np.int32), np.ascontiguousarray(aa[1], np.float32),
                       np.ascontiguousarray(gm, np.int32)]
        staged = self._stage_h2d(staged)
        c_t, s_t, r_t, f_t, joints, vis = staged[:6]
        if k3:
            aa, gm = (staged[6], staged[7]), staged[8]

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
