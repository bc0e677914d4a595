"""ctypes binding of libadvmix_b200.so (the C ABI in include/advmix_b200.h).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ADVMIX_B200_LIB") or os.path.join(_HERE, "libadvmix_b200.so")   # env override: kernel-variant builds

ABI_VERSION = 1
F32, BF16 = 0, 1
CORRUPT_FAST = 0x100

OPS = ("gaussian_noise", "shot_noise", "impulse_noise", "defocus_blur", "glass_blur", "motion_blur",
       "zoom_blur", "snow", "frost", "fog", "brightness", "contrast", "elastic_transform", "pixelate",
       "jpeg_compression", "speckle_noise", "gaussian_blur", "spatter", "saturate")

_p = C.c_void_p
_i = C.c_int
_sz = C.c_size_t

# name -> (restype, argtypes); kept in one table so tests can check every header symbol
SIGNATURES = {
    "advmix_abi_version": (_i, []),
    "advmix_last_error": (C.c_char_p, []),
    "advmix_device_check": (_i, [_i]),
    "advmix_warp_affine_u8c3": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "advmix_h2d_source_rows": (_i, [_p, _p, _p, _p, _p, _p, _i, _p]),
    "advmix_h2d_source_boxes": (_i, [_p, _p, _p, _i, _i, _p, _p]),
    "advmix_affine_matrices": (_i, [_p, _p, _i, _p, _p, _i, _i, _i, _p]),
    "advmix_step_params_bytes": (C.c_size_t, [_i, _i]),
    "advmix_chains_step_params_bytes": (C.c_size_t, [_i, _i]),
    "advmix_crop_chains_step": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, C.c_size_t, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "advmix_crop_targets_step": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "advmix_joints_flip_affine": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "advmix_joints_flip_affine_rec": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "advmix_step_rec_params_bytes": (C.c_size_t, [_i, _i]),
    "advmix_crop_targets_step_rec": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "advmix_xywh2cs": (_i, [_p, _p, _p, _i, C.c_double, C.c_double, _p]),
    "advmix_half_body_cs": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, C.c_double, C.c_double, _p]),
    "advmix_select_data": (_i, [_p, _p, _p, _p, _p, _i, _i, C.c_double, _p]),
    "advmix_base_cs": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, C.c_double, C.c_double, _p]),
    "advmix_jpeg_plan_stride": (_sz, []),
    "advmix_jpeg_plan_h": (_i, [_p, _p, _p, _i, _p, _p, _p, _p]),
    "advmix_jpeg_decode": (_i, [_p, _p, _i, _i, _i, _p, _p, _sz, C.c_int64, C.c_int64, C.c_int64, _i, _i, _p]),
    "advmix_jpeg_encode_workspace_bytes": (_sz, [_i, _i, _i]),
    "advmix_jpeg_encode_u8c3": (_i, [_p, _i, _i, _i, _i, _p, _sz, _p, _p, _sz, _p]),
    "advmix_heatmap_decode": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _i, _i, _i, _i, _p]),
    "advmix_flip_merge": (_i, [_p, _p, _p, _i, _p, _i, _i, _i, _i, _p]),
    "advmix_crop_csr_u8c3": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "advmix_joints_csr": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "advmix_normalize_u8c3": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "advmix_heatmap_targets": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "advmix_mix_fwd": (_i, [C.POINTER(_p), _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "advmix_mix_bwd": (_i, [C.POINTER(_p), _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "advmix_autoaug_workspace_bytes": (_sz, [_i, _i, _i]),
    "advmix_autoaug_u8c3": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _sz, _p]),
    "advmix_autoaug_plan_bytes": (_sz, [_i]),
    "advmix_autoaug_plan_u8c3": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _sz, _p]),
    "advmix_chains_emit_u8c3": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "advmix_chainmix_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _p, _i, _p, _i, _i, _i, _p]),
    "advmix_chainmix_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _p, _i, _p, _i, _i, _i, _p]),
    "advmix_mix_u8_fwd": (_i, [C.POINTER(_p), _p, _p, _i, _i, _p, _i, _p, _i, _i, _i, _i, _p]),
    "advmix_mix_u8_bwd": (_i, [C.POINTER(_p), _p, _p, _i, _i, _p, _i, _p, _i, _i, _i, _i, _p]),
    "advmix_gridmask": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "advmix_corrupt_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "advmix_corrupt_rand_field_bytes": (_sz, [_i, _i, _i, _i]),
    "advmix_corrupt_fill_rand": (_i, [_i, _i, _i, _i, _i, C.c_uint64, C.c_int64, _p, _p, _p, _i, _i, _i, _p]),
    "advmix_corrupt_u8c3": (_i, [_i, _i, _p, _p, _i, _p, _i, _i, _p, _p, C.c_uint64, C.c_int64, _p, _i, _i, _i,
                                 _p, _sz, _p]),
    "advmix_pack_files": (_i, [_p, _sz, _p, _i, _p, _p, _p]),
    "advmix_corrupt_sweep_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "advmix_corrupt_sweep_u8c3": (_i, [_i, _p, _p, _i, _p, _i, _i, C.c_uint64, C.c_int64, _p, _i, _i, _i, _p, _sz, _p]),
}

_lib = None


class AdvmixError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libadvmix_b200.so not found at %s - build it with `python -m advmix_b200.build` "
            "(there is no CPU or PyTorch fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.advmix_abi_version() != ABI_VERSION:
        raise ImportError("libadvmix_b200.so ABI %d != binding ABI %d" % (lib.advmix_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().advmix_last_error()
        raise AdvmixError("%s failed (%d): %s" % (what or "advmix call", rc, msg.decode() if msg else ""))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(dt):
    import torch
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TypeError("advmix_b200 supports float32 and bfloat16 outputs, got %r" % (dt,))
