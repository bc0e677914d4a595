"""GPU parity for SURVEY row f1: device JPEG decode vs cv2.imdecode / PIL (libjpeg-turbo), bit for bit."""
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def natural(rng, H, W):
    import cv2
    acc = np.zeros((H, W, 3), np.float32)
    for o in range(4):
        s = 2 ** (o + 2)
        acc += cv2.resize(rng.random((H // s + 2, W // s + 2, 3)).astype(np.float32), (W, H)) / (o + 1)
    acc = acc / acc.max() * 255 + rng.integers(-8, 9, acc.shape)
    return np.clip(acc, 0, 255).astype(np.uint8)


def pil_jpeg(img, **kw):
    from PIL import Image
    b = io.BytesIO()
    Image.fromarray(img).save(b, "JPEG", **kw)
    return b.getvalue()


def image_of(sb, i):
    off, H, W, pitch = int(sb.offsets[i]), int(sb.heights[i]), int(sb.widths[i]), int(sb.pitches[i])
    return sb.buffer[off:off + H * pitch].view(H, pitch)[:, :3 * W].reshape(H, W, 3).cpu().numpy()


def corpus():
    import cv2
    rng = np.random.default_rng(5)
    files = []
    for (H, W) in ((480, 640), (250, 333), (37, 53), (16, 16), (8, 8), (1, 1), (97, 16), (427, 640)):
        img = natural(rng, max(H, 8), max(W, 8))[:H, :W]
        for sub in (0, 1, 2):
            files.append(pil_jpeg(img, quality=int(rng.integers(20, 96)), subsampling=sub))
        files.append(pil_jpeg(img, quality=85, optimize=True))                       # custom Huffman tables
        files.append(pil_jpeg(img[..., 0], quality=70))                              # grayscale
        files.append(pil_jpeg(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), quality=100, subsampling=0))   # dense noise
        ok, buf = cv2.imencode(".jpg", img[..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, 60, cv2.IMWRITE_JPEG_RST_INTERVAL, 3])
        files.append(buf.tobytes())                                                  # restart markers
    return files


def test_decode_matches_cv2_and_pil(built_library):
    import cv2
    from PIL import Image
    from advmix_b200 import jpeg as J
    files = corpus()
    bgr = J.decode_batch(files, color="bgr")
    rgb = J.decode_batch(files, color="rgb")
    assert len(bgr) == len(files)
    for i, f in enumerate(files):
        exp = cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR | cv2.IMREAD_IGNORE_ORIENTATION)
        got = image_of(bgr, i)
        assert got.shape == exp.shape, (i, got.shape, exp.shape)
        assert np.array_equal(got, exp), "file %d (%dx%d): %d px differ" % (i, exp.shape[1], exp.shape[0], (got != exp).any(-1).sum())
        pil = np.array(Image.open(io.BytesIO(f)).convert("RGB"))
        assert np.array_equal(image_of(rgb, i), pil), "file %d vs PIL" % i


def test_decode_feeds_the_crop_kernel(built_library):
    """The decoded batch is a SourceBatch: crops taken from it equal crops taken from cv2-decoded images."""
    import cv2
    import advmix_b200 as A
    from advmix_b200 import jpeg as J
    rng = np.random.default_rng(9)
    imgs = [natural(rng, 480, 640), natural(rng, 300, 420)]
    files = [pil_jpeg(im, quality=80) for im in imgs]
    sb = J.decode_batch(files, color="bgr")
    ref = A.SourceBatch.from_numpy([cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR) for f in files])
    dev = torch.device("cuda:0")
    c = torch.tensor([[320., 240.], [200., 150.]], dtype=torch.float32, device=dev)
    s = torch.tensor([[1.5, 2.0], [1.0, 1.3]], dtype=torch.float32, device=dev)
    r = torch.tensor([25.0, -40.0], dtype=torch.float64, device=dev)
    M = A.get_affine_transform(c, s, r, (192, 256))
    a, _ = A.warp_affine(sb, M, (192, 256))
    b, _ = A.warp_affine(ref, M, (192, 256))
    assert torch.equal(a, b)


def test_unsupported_files_raise(built_library):
    import cv2
    import advmix_b200 as A
    from advmix_b200 import jpeg as J
    rng = np.random.default_rng(1)
    img = natural(rng, 64, 64)
    prog = pil_jpeg(img, quality=80, progressive=True)
    with pytest.raises(A.AdvmixError, match="progressive"):
        J.decode_batch([pil_jpeg(img), prog])
    with pytest.raises(A.AdvmixError):
        J.decode_batch([b"\xff\xd8\xff\xd9"])
    with pytest.raises(A.AdvmixError):
        J.decode_batch([pil_jpeg(img)[:200]])                     # cut inside the headers
    # a file cut inside the scan still decodes (zeros for the missing bits, like libjpeg's warning path)
    cut = pil_jpeg(img, quality=90)
    sb = J.decode_batch([cut[:len(cut) // 2]])
    assert int(sb.heights[0]) == 64 and int(sb.widths[0]) == 64


# ---- encode: the files tools/make_datasets.py:45 writes with PIL, produced on the device ----------------------------
ENC_SIZES = ((256, 192), (256, 256), (480, 640), (375, 500), (427, 640), (37, 53), (100, 41), (16, 16), (8, 8), (1, 1),
             (97, 16), (9, 9), (17, 33), (24, 40))


def first_diff(a, b):
    n = min(len(a), len(b))
    d = np.nonzero(np.frombuffer(a[:n], np.uint8) != np.frombuffer(b[:n], np.uint8))[0]
    return (int(d[0]) if len(d) else n), len(a), len(b)


@pytest.mark.parametrize("quality", [75, 25, 7, 95])
def test_encode_is_byte_identical_to_pil(built_library, quality):
    from advmix_b200 import jpeg as J
    rng = np.random.default_rng(quality)
    for (H, W) in ENC_SIZES:
        imgs = np.stack([natural(rng, max(H, 8), max(W, 8))[:H, :W] for _ in range(3)] +
                        [rng.integers(0, 256, (H, W, 3), dtype=np.uint8), np.zeros((H, W, 3), np.uint8),
                         np.full((H, W, 3), 255, np.uint8)])
        got = J.encode_batch(torch.from_numpy(imgs).cuda(), quality=quality)
        for i, g in enumerate(got):
            exp = pil_jpeg(imgs[i], quality=quality)
            assert g == exp, "%dx%d image %d q%d: first difference (byte, len got, len exp) = %s" % (W, H, i, quality, first_diff(g, exp))


def test_encode_default_quality_and_round_trip(built_library):
    """PIL's default save (what make_datasets.py calls) and decode(encode(x)) through both device paths."""
    from PIL import Image
    from advmix_b200 import jpeg as J
    rng = np.random.default_rng(3)
    imgs = np.stack([natural(rng, 256, 192) for _ in range(8)])
    got = J.encode_batch(torch.from_numpy(imgs).cuda())
    for i, g in enumerate(got):
        b = io.BytesIO()
        Image.fromarray(imgs[i]).save(b, "JPEG")                  # no quality argument, like the reference
        assert g == b.getvalue()
    sb = J.decode_batch(got, color="rgb")
    for i, g in enumerate(got):
        assert np.array_equal(image_of(sb, i), np.array(Image.open(io.BytesIO(g)).convert("RGB")))


def test_encode_dense_noise_quality_100_and_overflow(built_library):
    """Worst-case entropy: every coefficient non-zero, many FF bytes to stuff; and the -1 length on a buffer that is too small."""
    import advmix_b200 as A
    from advmix_b200 import _lib, jpeg as J
    rng = np.random.default_rng(11)
    imgs = rng.integers(0, 256, (4, 64, 80, 3), dtype=np.uint8)
    got = J.encode_batch(torch.from_numpy(imgs).cuda(), quality=100)
    for i, g in enumerate(got):
        exp = pil_jpeg(imgs[i], quality=100)
        assert g == exp, first_diff(g, exp)
    lib = _lib.load()
    x = torch.from_numpy(imgs).cuda()
    out = torch.zeros((4, 1024), dtype=torch.uint8, device="cuda")
    ln = torch.zeros(4, dtype=torch.int32, device="cuda")
    wsb = int(lib.advmix_jpeg_encode_workspace_bytes(4, 64, 80))
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    _lib.check(lib.advmix_jpeg_encode_u8c3(_lib.ptr(x), 4, 64, 80, 100, _lib.ptr(out), 1024, _lib.ptr(ln), _lib.ptr(ws), wsb,
                                           _lib.stream_ptr()))
    assert (ln.cpu().numpy() == -1).all()
    with pytest.raises(A.AdvmixError):
        _lib.check(lib.advmix_jpeg_encode_u8c3(_lib.ptr(x), 4, 64, 80, 0, _lib.ptr(out), 1024, _lib.ptr(ln), _lib.ptr(ws), wsb,
                                               _lib.stream_ptr()))
    assert J.encode_batch(torch.zeros((0, 16, 16, 3), dtype=torch.uint8, device="cuda")) == []


def test_process_files_like_make_datasets(built_library):
    """tools/make_datasets.py process(): encoded sources in, encoded corrupted files out; the integer-exact corruptions
    give the very bytes PIL would have written for the oracle's output."""
    import io
    from PIL import Image
    from advmix_b200 import datasets_c
    from oracle import corruptions as OK
    rng = np.random.default_rng(21)
    imgs = [natural(rng, 96, 128), natural(rng, 64, 80), natural(rng, 96, 128), natural(rng, 50, 70)]
    files = [pil_jpeg(im, quality=88) for im in imgs]
    names = ["pixelate", "jpeg_compression", "gaussian_noise", "zoom_blur"]
    res = datasets_c.process_files(files, names, severities=(2, 5))
    assert set(res) == {(n, s) for n in names for s in (2, 5)}
    for i, f in enumerate(files):
        src = np.array(Image.open(io.BytesIO(f)).convert("RGB"))                 # what np.asarray(Image.open(img)) yields
        for name in ("pixelate", "jpeg_compression"):
            for sev in (2, 5):
                exp = OK.corrupt_with_draws(src, sev, name, {})
                assert res[(name, sev)][i] == pil_jpeg(exp), (i, name, sev)
        for name in ("gaussian_noise", "zoom_blur"):
            got = np.array(Image.open(io.BytesIO(res[(name, 5)][i])))
            assert got.shape == src.shape and got.dtype == np.uint8
    z = np.array(Image.open(io.BytesIO(res[("zoom_blur", 2)][1]))).astype(int)
    e = OK.corrupt_with_draws(np.array(Image.open(io.BytesIO(files[1])).convert("RGB")), 2, "zoom_blur", {})
    assert np.abs(z - np.array(Image.open(io.BytesIO(pil_jpeg(e)))).astype(int)).max() <= 16    # same picture through the same codec
    # all five severities: one corrupt_sweep call per corruption and size group - the same files as the per-severity path
    res5 = datasets_c.process_files(files, names, fast=True)
    res2 = datasets_c.process_files(files, names, severities=(2, 5), fast=True)
    assert set(res5) == {(n, s) for n in names for s in (1, 2, 3, 4, 5)}
    for n in names:
        for sev in (2, 5):
            assert res5[(n, sev)] == res2[(n, sev)], (n, sev)
    with pytest.raises(AttributeError):
        datasets_c.process_files([pil_jpeg(natural(rng, 24, 40))], ["pixelate"], severities=(1,))


def test_many_plans_alive_before_upload(built_library):
    """More PlannedBatch objects alive than the pinned plan ring is deep, all parsed before any to_device(): the ring
    must not hand a buffer out twice (plans overwrote each other)."""
    import cv2
    from advmix_b200 import jpeg as J
    rng = np.random.default_rng(31)
    imgs = [natural(rng, 40 + 8 * k, 56 + 8 * k) for k in range(7)]
    encs = [J.EncodedBatch([pil_jpeg(im, quality=80)]) for im in imgs]
    plans = [J.PlannedBatch(e) for e in encs]                 # 7 pending plans, ring depth 4
    for k, pb in enumerate(plans):
        fd, pd = pb.to_device("cuda")
        sb = J.decode_planned(pb, fd, pd, "rgb")
        exp = cv2.imdecode(np.frombuffer(bytes(encs[k].host[:encs[k].lengths[0]].numpy()), np.uint8), cv2.IMREAD_COLOR)[..., ::-1]
        assert np.array_equal(image_of(sb, 0), exp), k
