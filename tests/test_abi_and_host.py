"""not-gpu: the C-ABI library builds, loads and exports exactly what include/advmix_b200.h declares;
host-side logic (draw order, sharding, argument validation) without touching a GPU."""
import os
import random
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "advmix_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(advmix_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_header_library_binding_agree(built_library):
    from advmix_b200 import _lib
    hdr = header_functions()
    assert set(hdr) == set(_lib.SIGNATURES), set(hdr) ^ set(_lib.SIGNATURES)
    for name, nargs in hdr.items():
        assert len(_lib.SIGNATURES[name][1]) == nargs, name
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (advmix_\w+)", nm))
    assert exported == set(hdr), exported ^ set(hdr)
    assert built_library.advmix_abi_version() == _lib.ABI_VERSION
    # pure size queries need no device
    assert built_library.advmix_corrupt_rand_field_bytes(0, 1, 256, 192) == 256 * 192 * 3 * 4
    assert built_library.advmix_corrupt_rand_field_bytes(4, 3, 256, 192) == 3 * 256 * 192 * 2
    assert built_library.advmix_corrupt_rand_field_bytes(9, 1, 256, 192) == 256 * 256 * 4
    assert built_library.advmix_corrupt_workspace_bytes(11, 1, 4, 256, 192) >= 4 * 3 * 8
    assert built_library.advmix_corrupt_workspace_bytes(99, 1, 4, 256, 192) == 0


def test_library_is_sm100a_only(built_library):
    from advmix_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_in_product():
    """The product never imports the oracle and fails loudly without the shared library."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "advmix_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    code = ("import sys; sys.path.insert(0, %r); from advmix_b200 import _lib; _lib.LIB_PATH = '/nonexistent/lib.so'\n"
            "try:\n    _lib.load()\nexcept ImportError as e:\n    print('LOUD', e)\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "LOUD" in out and "no CPU or PyTorch fallback" in out


def test_corrupt_validation_matches_package_before_touching_cuda():
    import advmix_b200 as A
    img = np.zeros((64, 48, 3), np.uint8)
    with pytest.raises(AttributeError):
        A.corrupt([1, 2, 3], 1, "contrast")
    with pytest.raises(AttributeError):
        A.corrupt(img.astype(np.float32), 1, "contrast")
    with pytest.raises(AttributeError):
        A.corrupt(img[:31], 1, "contrast")
    with pytest.raises(AttributeError):
        A.corrupt(np.zeros((64, 48, 2), np.uint8), 1, "contrast")
    with pytest.raises(AttributeError):
        A.corrupt(img, 0, "contrast")
    with pytest.raises(ValueError):
        A.corrupt(img, 1)
    assert A.get_corruption_names() == A.get_corruption_names("common") and len(A.get_corruption_names("all")) == 19
    from oracle import corruptions as OK
    for subset in ("common", "validation", "all", "noise", "blur", "weather", "digital"):
        assert A.get_corruption_names(subset) == OK.get_corruption_names(subset)
    with pytest.raises(ValueError):
        A.get_corruption_names("nope")


def test_draw_order_matches_reference(golden):
    """sample_gridmask / sample_autoaug consume np.random / random exactly like advaug.py does."""
    from advmix_b200 import chains as C
    from oracle import chains as OC
    g = golden("chains")
    for i in range(len(g["gm_params"])):
        np.random.seed(500 + i)                     # seeds used by oracle/make_golden.py around the real grid_aug
        assert np.array_equal(C.sample_gridmask(1, 64, 48)[0], g["gm_params"][i])
    for i, (pidx, c1, c2, s1, s2) in enumerate(g["aa_plan"]):
        # replay SubPolicy.__call__'s stream for this fixture: coin1 [, sign1], coin2 [, sign2]
        random.seed(1000 + i)
        p1, op1, m1, p2, op2, m2 = C.POLICIES[int(pidx)]
        ops = np.zeros(2, np.int32); mags = np.zeros(2, np.float32)
        if random.random() < p1:
            s = random.choice([-1, 1]) if op1 == "sharpness" else 1
            ops[0], mags[0] = C._stage(op1, m1, s)
        if random.random() < p2:
            s = random.choice([-1, 1]) if op2 == "sharpness" else 1
            ops[1], mags[1] = C._stage(op2, m2, s)
        eo, em = C.plan_autoaug(int(pidx), c1, c2, int(s1), int(s2))
        assert tuple(ops) == eo and np.allclose(mags, em)
    assert C.POLICIES == OC.POLICIES
    random.seed(3)
    ops, mags = C.sample_autoaug(64)
    assert ops.shape == (64, 2) and set(np.unique(ops)) <= {0, 1, 2, 3, 4, 5}


def test_shards():
    from advmix_b200.dist import shard_range, sample_base
    for n, w in [(256, 8), (257, 8), (5, 8), (0, 2), (32, 1)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert sample_base(0, 3, 256, 32) == 3 * 256 + 32


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from advmix_b200.dist import broadcast_control, shard_range
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    seed, epoch = broadcast_control(1234 if rank == 0 else -1, 7 if rank == 0 else -1)
    lo, hi = shard_range(257, rank, world)
    # the "G on rank 0" arrangement: rank 0's [B_global, K, H, W] logits reach every rank, each keeps its shard
    import torch
    from advmix_b200.dist import broadcast_mix_weights
    full = torch.arange(5 * 3 * 4 * 2, dtype=torch.float32).reshape(5, 3, 4, 2)
    mine = broadcast_mix_weights(full if rank == 0 else ((5, 3, 4, 2), torch.float32, "cpu"), 5)
    l5, h5 = shard_range(5, rank, world)
    assert torch.equal(mine, full[l5:h5]), "mix-weight broadcast shard mismatch on rank %d" % rank
    q.put((rank, seed, epoch, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_control_broadcast_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(1234, 7), (1234, 7)]
    assert res[0][3:] == (0, 129) and res[1][3:] == (129, 257)


def test_jpeg_plan_is_host_code(built_library):
    """advmix_jpeg_plan_h runs on the host (header parse only): geometry, sampling and the status field."""
    import ctypes as C
    import io
    from PIL import Image
    from advmix_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (45, 70, 3), dtype=np.uint8)
    files = []
    for kw in (dict(subsampling=2), dict(subsampling=0), dict(progressive=True)):
        b = io.BytesIO(); Image.fromarray(img).save(b, "JPEG", quality=75, **kw); files.append(b.getvalue())
    buf = np.concatenate([np.frombuffer(f, np.uint8) for f in files])
    lens = np.array([len(f) for f in files], np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    stride = int(lib.advmix_jpeg_plan_stride())
    plans = np.zeros((3, stride), np.uint8)
    totals = np.zeros(3, np.int64)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.advmix_jpeg_plan_h(vp(buf), vp(offs), vp(lens), 3, vp(plans), vp(totals[0:]), vp(totals[1:]), vp(totals[2:]))
    assert rc == -3 and b"1 of 3" in lib.advmix_last_error()
    i32 = lambda b, o: int(plans[b, o:o + 4].view(np.int32)[0])
    assert [i32(b, 156) for b in range(3)] == [0, 0, 2]                  # ok, ok, progressive
    assert (i32(0, 32), i32(0, 36), i32(0, 40)) == (70, 45, 3)
    assert i32(0, 224) == 80 and i32(0, 240) == 48                       # 4:2:0 luma plane padded to whole 16x16 MCUs
    assert i32(1, 224) == 72 and i32(1, 240) == 48                       # 4:4:4: 8x8 MCUs
    assert int(plans[0, 24:32].view(np.int64)[0]) == 224                 # out_pitch = 3*70 rounded up to 16


def test_batch_samplers_match_the_reference_order_samplers():
    """sample_*_batch draw the same distributions as the per-sample samplers that follow advaug.py's RNG order."""
    import random
    from advmix_b200 import chains as C
    g = np.random.default_rng(5)
    ops, mags = C.sample_autoaug_batch(60000, g)
    random.seed(5)
    ops_r, mags_r = C.sample_autoaug(30000)
    fa, fr = np.bincount(ops.ravel(), minlength=8) / ops.size, np.bincount(ops_r.ravel(), minlength=8) / ops_r.size
    assert np.abs(fa - fr).max() < 0.01, (fa, fr)
    for code in np.unique(ops_r[ops_r > 0]):                    # every op draws from the same set of magnitudes
        assert set(np.unique(mags[ops == code]).tolist()) == set(np.unique(mags_r[ops_r == code]).tolist()), code
    assert (mags[ops == 0] == 0).all()
    p = C.sample_gridmask_batch(60000, 256, 192, g)
    np.random.seed(5)
    pr = C.sample_gridmask(30000, 256, 192)
    assert abs(p[:, 0].mean() - pr[:, 0].mean()) < 0.01
    on = p[:, 0] > 0
    assert p[on, 1].min() >= 2 and p[on, 1].max() <= 191 and (p[on, 2] < p[on, 1]).all() and (p[on, 3] < p[on, 1]).all()
    assert (p[~on] == 0).all()
    assert abs(p[on, 1].mean() - pr[pr[:, 0] > 0, 1].mean()) < 1.5 and abs(p[on, 2].mean() - pr[pr[:, 0] > 0, 2].mean()) < 1.5
