"""ctypes binding of the C-ABI library ``libdeepatlas_b200.so`` (declared in include/deepatlas_b200.h).

There is no CPU fallback: if the library is missing or a call fails, a RuntimeError is raised with
``da_last_error()``.  The library never allocates device memory; callers pass PyTorch-allocated
buffers (outputs, saved tensors, workspaces) as raw pointers plus the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DA_LIB_PATH") or os.path.join(_HERE, "libdeepatlas_b200.so")  # override: kernel experiments

_T = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_int64, "f": ctypes.c_float,
      "d": ctypes.c_double, "s": ctypes.c_void_p}

# name -> (argument codes, returns int64 size instead of status)
SIGNATURES = {
    "da_version": ("", "int"),
    "da_last_error": ("", "str"),
    "da_memset_zero": ("pls", "rc"),
    "da_launch_count": ("", "size"),
    # warp3d
    "da_warp3d_fwd": ("ppippiiiiiiiis", "rc"),
    "da_warp3d_bwd": ("pppippiiiiiiiis", "rc"),
    "da_warp3d_bwd_cl": ("pppippiiiiiiiis", "rc"),
    # dice
    "da_dice_workspace_bytes": ("iil", "size"),
    "da_dice_sums_fwd": ("ppiiiilppls", "rc"),
    "da_dice_sums_bwd": ("ppiiiilppppps", "rc"),
    "da_warp_dice_fwd_workspace_bytes": ("iill", "size"),
    "da_warp_dice_bwd_workspace_bytes": ("il", "size"),
    "da_warp_dice_sums_fwd": ("ppipiiiiiiiiipppls", "rc"),
    "da_warp_dice_sums_bwd": ("ppipipppiiiiiiiipppls", "rc"),
    "da_softmax_fwd": ("ppiils", "rc"),
    "da_softmax_dice_fwd": ("ppiiilpppls", "rc"),
    "da_softmax_dice_bwd": ("ppiiilppppps", "rc"),
    "da_softmax_bwd": ("pppiils", "rc"),
    "da_head_dice_supported": ("iil", "size"),
    "da_head_dice_workspace_bytes": ("iil", "size"),
    "da_head_dice_fwd": ("ppppiiiilpppls", "rc"),
    "da_head_dice_bwd": ("ppppiiiilppppppipls", "rc"),
    "da_argmax_counts": ("ppiiilpps", "rc"),
    # lncc
    "da_lncc_coef_bytes": ("iiiiii", "size"),
    "da_lncc_fwd_workspace_bytes": ("iiiii", "size"),
    "da_lncc_bwd_workspace_bytes": ("iiiii", "size"),
    "da_lncc_fwd": ("ppiiiiidipppls", "rc"),
    "da_lncc_bwd": ("ppppiiiiiiippls", "rc"),
    # bending
    "da_bending_fwd_workspace_bytes": ("i", "size"),
    "da_bending_bwd_workspace_bytes": ("iiii", "size"),
    "da_bending_fwd": ("piiiippls", "rc"),
    "da_bending_bwd": ("ppiiiippls", "rc"),
    "da_bending_fwd_ex": ("piiiiippls", "rc"),
    "da_bending_bwd_ex": ("ppiiiiippls", "rc"),
    # conv
    "da_umma_debug_read": ("p", "rc"),
    "da_set_conv_impl": ("i", "rc"),
    "da_set_conv_split": ("i", "rc"),
    "da_conv3d_pack_bytes": ("iii", "size"),
    "da_conv3d_dgrad_workspace_bytes": ("iiiiiiii", "size"),
    "da_conv3d_wgrad_workspace_bytes": ("iii", "size"),
    "da_conv3d_fwd": ("pipipippiiiiiiiiifpls", "rc"),
    "da_conv3d_dgrad": ("ppipiiiiiiiiiiipls", "rc"),
    "da_conv3d_wgrad": ("pipipippiiiiiiiipls", "rc"),
    "da_absmax": ("plplps", "rc"),
    "da_conv3d_fwd_ex": ("pipipippiiiiiiiiifplspi", "rc"),
    "da_conv3d_dgrad_ex": ("ppipiiiiiiiiiiiplspi", "rc"),
    "da_conv3d_wgrad_ex": ("pipipippiiiiiiiiplspipii", "rc"),
    "da_channel_sum_workspace_bytes": ("i", "size"),
    "da_channel_sum": ("piilppls", "rc"),
    # bn / act / pool / upsample
    "da_bn_workspace_bytes": ("i", "size"),
    "da_bn_stats": ("piilffpppppls", "rc"),
    "da_bn_stats_ex": ("piilffppppppifppls", "rc"),
    "da_bn_act_fwd": ("pppppiilifps", "rc"),
    "da_bn_act_bwd": ("ppppppiiliifppppls", "rc"),
    "da_bn_act_bwd_ex": ("ppppppiiliifppppipls", "rc"),
    "da_act_bwd": ("ppflps", "rc"),
    "da_maxpool2_fwd": ("ppliiis", "rc"),
    "da_maxpool2_bwd": ("pppliiis", "rc"),
    "da_upsample_nearest_fwd": ("ppliiiiiis", "rc"),
    "da_upsample_nearest_bwd": ("ppliiiiiis", "rc"),
    # deconv k2 s2
    "da_deconv_k2s2_wgrad_workspace_bytes": ("ii", "size"),
    "da_deconv_k2s2_fwd": ("ppppiiiiiis", "rc"),
    "da_deconv_k2s2_dgrad": ("pppiiiiiis", "rc"),
    "da_deconv_k2s2_wgrad": ("ppppiiiiiipls", "rc"),
    "da_deconv_k2s2_wgrad_ex": ("ppppiiiiiiipls", "rc"),
    # remaining registry losses
    "da_pair_moments_workspace_bytes": ("i", "size"),
    "da_pair_moments_fwd": ("ppilppls", "rc"),
    "da_affine2": ("pppilps", "rc"),
    "da_gradient_loss_workspace_bytes": ("ii", "size"),
    "da_gradient_loss_fwd": ("piiiiiippls", "rc"),
    "da_gradient_loss_bwd": ("ppiiiiiips", "rc"),
    "da_xent_workspace_bytes": ("i", "size"),
    "da_xent_fwd": ("ppiiiilpfilppls", "rc"),
    "da_xent_bwd": ("ppiiiilpfilppps", "rc"),
    # UNet_generator variants, input stage
    "da_upsample_trilinear2_fwd": ("ppliiis", "rc"),
    "da_upsample_trilinear2_bwd": ("ppliiis", "rc"),
    "da_add_bcast": ("ppiiilps", "rc"),
    "da_channel_reduce": ("piilps", "rc"),
    "da_crop_clip_f32": ("ppliiiiiiiiiffs", "rc"),
    "da_crop_u8": ("ppliiiiiiiiis", "rc"),
    "da_label_overlap_counts": ("pipiiilps", "rc"),
    "da_lncc_ms_workspace_bytes": ("", "size"),
    "da_lncc_ms_coef_bytes": ("iiiiiii", "size"),
    "da_lncc_ms_fwd": ("ppiiiiiiipppls", "rc"),
    "da_lncc_ms_bwd": ("pppipfiiiiiiiips", "rc"),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library once; raise loudly if it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"deepatlas_b200: CUDA library not built ({LIB_PATH} missing). "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` from the repo root.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (codes, ret) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.argtypes = [_T[c] for c in codes]
        fn.restype = {"rc": ctypes.c_int, "int": ctypes.c_int, "size": ctypes.c_int64,
                      "str": ctypes.c_char_p}[ret]
    _lib = lib
    return lib


def last_error() -> str:
    return load().da_last_error().decode("utf-8", "replace")


def call(name: str, *args):
    """Invoke a status-returning entry point; raise RuntimeError(da_last_error()) on failure."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {last_error()}")


def size(name: str, *args) -> int:
    return int(getattr(load(), name)(*args))
