"""Device form of `make_data.process` (tools/make_datasets.py:36-45), the loop that builds COCO-C / MPII-C:

    image = np.asarray(Image.open(img))
    for corruption in get_corruption_names('all'):
        for severity in range(5):
            np.random.seed(1)
            corrupted = corrupt(image, corruption_name=corruption, severity=severity+1)
            Image.fromarray(corrupted).save(corrupted_path)

Here a batch of encoded source files goes in and the encoded corrupted files come out; decode
(`jpeg.decode_batch`), the corruptions (`corrupt_batch`) and the encode (`jpeg.encode_batch`, byte-identical
to PIL's `Image.save`) all run on the device, so only encoded bytes cross PCIe.  Directory layout and file
I/O stay with the caller (`<root>/<dataset>-C/<corruption>/<severity>/<basename>`, make_datasets.py:42).
"""
import numpy as np
import torch

from . import jpeg
from .corruptions import corrupt_batch, corrupt_sweep, get_corruption_names


def _dense_group(sb, pb, members):
    """uint8 [n, H, W, 3] tensor of the decoded images `members` (all the same size) of a SourceBatch."""
    H, W = int(pb.heights[members[0]]), int(pb.widths[members[0]])
    out = torch.empty((len(members), H, W, 3), dtype=torch.uint8, device=sb.buffer.device)
    for k, i in enumerate(members):
        off, pitch = int(pb.out_off[i]), int(pb.out_pitch[i])
        out[k] = sb.buffer[off:off + H * pitch].view(H, pitch)[:, :3 * W].reshape(H, W, 3)
    return out


def process_files(files, corruption_names=None, severities=(1, 2, 3, 4, 5), seed=1, quality=75, fast=False, device="cuda",
                  frost_bank=None):
    """files: list of `bytes` (baseline JPEG files, any sizes >= 32x32).  Returns {(corruption_name, severity): [bytes]}
    with one encoded file per input file, in input order.

    corruption_names: default `get_corruption_names('all')` like make_datasets.py:38.  seed: the reference re-seeds
    np.random with 1 before every call; here the draws are Philox streams keyed by (seed, position in the batch,
    corruption), so a run is reproducible for a given batch composition.  quality: PIL's default 75.
    frost_bank: uint8 [N,fh,fw,3] RGB frost textures (the package's frost1-6 images, loaded by the caller); without it
    (and without a prior `set_frost_bank`) the 'frost' split is built from synthetic stand-ins and a warning is issued.
    Images are grouped by size; per group and corruption one `corrupt_sweep` call produces the five severities (the loop nest of
    make_datasets.py:38-45; `corrupt_batch` per severity when only some are asked for), then one `encode_batch` per output."""
    if frost_bank is not None:
        from .corruptions import set_frost_bank
        set_frost_bank(frost_bank, device)
    names = list(corruption_names) if corruption_names is not None else get_corruption_names("all")
    enc = jpeg.EncodedBatch(files)
    pb = jpeg.PlannedBatch(enc)
    files_d, plans_d = pb.to_device(device)
    sb = jpeg.decode_planned(pb, files_d, plans_d, color="rgb")
    groups = {}
    for i in range(len(files)):
        H, W = int(pb.heights[i]), int(pb.widths[i])
        if H < 32 or W < 32:
            raise AttributeError("Image width and height must be at least 32 pixels")      # imagecorruptions.corrupt
        groups.setdefault((H, W), []).append(i)
    result = {(n, int(s)): [None] * len(files) for n in names for s in severities}
    for (H, W), members in groups.items():
        x = _dense_group(sb, pb, members)
        all_five = sorted(int(s) for s in severities) == [1, 2, 3, 4, 5]
        out5 = torch.empty((5,) + tuple(x.shape), dtype=torch.uint8, device=x.device) if all_five else None
        out = None if all_five else torch.empty_like(x)
        for n in names:
            if all_five:
                corrupt_sweep(x, n, seed=seed, sample_base=members[0], out=out5, fast=fast)     # bit-identical to the per-severity calls
            for s in severities:
                if not all_five:
                    corrupt_batch(x, n, int(s), seed=seed, sample_base=members[0], out=out, fast=fast)
                for i, f in zip(members, jpeg.encode_batch(out5[int(s) - 1] if all_five else out, quality=quality)):
                    result[(n, int(s))][i] = f
    return result
