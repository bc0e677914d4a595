"""advmix_b200 - B200-native (sm_100a) implementation of AdvMix's augmentation + target hot path.

Public surface (mirrors the reference's call boundaries):
  corrupt, get_corruption_names, corrupt_batch      <- imagecorruptions API
  generate_target                                   <- JointsDataset.generate_target
  get_affine_transform, warp_affine, ...            <- lib/utils/transforms.py + cv2.warpAffine
  mix, mix_from_logits                              <- lib/core/function.py:137-146
  chain_mix(_from_logits), chains_g_input, mix_u8   <- the same, fused with the chain ops (reads the uint8 crop)
  AdvMixBatchPipeline                               <- JointsDataset.__getitem__ + collate
  get_max_preds, get_final_preds, flip_merge        <- lib/core/inference.py, function.py:241-261
  jpeg.decode_batch / encode_batch, datasets_c      <- cv2.imread / PIL Image.save, tools/make_datasets.py process()

Everything runs through libadvmix_b200.so (hand-written CUDA, C ABI in include/advmix_b200.h).
Importing the package does not need a GPU; calling any op without the built library or
without a CUDA device raises.
"""
from ._lib import AdvmixError, load as load_library  # noqa: F401
from .corruptions import corrupt, corrupt_batch, corrupt_sweep, get_corruption_names  # noqa: F401
from . import datasets_c, jpeg, records  # noqa: F401
from .inference import flip_back, flip_merge, get_final_preds, get_max_preds  # noqa: F401
from .mix import (autoaug_plan, chain_mix, chain_mix_from_logits, chains_g_input, mix, mix_from_logits,  # noqa: F401
                  mix_u8, mix_u8_from_logits)
from .targets import generate_target  # noqa: F401
from .transforms import (SourceBatch, crop_csr, fliplr_affine_joints, get_affine_transform,  # noqa: F401
                         joints_csr, to_tensor_normalize, warp_affine)

__all__ = ["corrupt", "corrupt_batch", "corrupt_sweep", "get_corruption_names", "mix", "mix_from_logits", "chain_mix", "chain_mix_from_logits",
           "chains_g_input", "autoaug_plan", "mix_u8", "mix_u8_from_logits", "generate_target",
           "SourceBatch", "get_affine_transform", "warp_affine", "crop_csr", "joints_csr", "fliplr_affine_joints", "to_tensor_normalize",
           "get_max_preds", "get_final_preds", "flip_merge", "flip_back", "load_library", "AdvmixError"]
