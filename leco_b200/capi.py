"""ctypes binding of the C ABI in include/leco_b200.h.

The product path has NO fallback: if the shared library is missing or a call fails,
this module raises.  (`python -m leco_b200.build` or `__graft_entry__.build()` builds it.)
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_void_p

P, I, L, F = c_void_p, c_int32, c_int64, c_float

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libleco_b200.so")
_lib = None


class LecoError(RuntimeError):
    pass


class GemmArgs(Structure):
    _fields_ = [
        ("a", c_void_p), ("b", c_void_p), ("d", c_void_p),
        ("mode", c_int32), ("M", c_int32), ("N", c_int32), ("K", c_int32),
        ("lda", c_int64), ("ldb", c_int64), ("ldd", c_int64),
        ("batch0", c_int32), ("batch1", c_int32),
        ("a_bs0", c_int64), ("a_bs1", c_int64), ("b_bs0", c_int64), ("b_bs1", c_int64),
        ("d_bs0", c_int64), ("d_bs1", c_int64),
        ("cn", c_int32), ("ch", c_int32), ("cw", c_int32), ("cc", c_int32),
        ("a2", c_void_p), ("b2", c_void_p), ("K2", c_int32),
        ("lda2", c_int64), ("ldb2", c_int64),
        ("bias", c_void_p), ("rowbias", c_void_p), ("rows_per_group", c_int32),
        ("ld_rowbias", c_int64),
        ("residual", c_void_p), ("ldr", c_int64),
        ("epilogue", c_int32), ("alpha", c_float), ("out_fp32", c_int32), ("block_n", c_int32),
        ("b_rows", c_int32), ("cta_pair", c_int32),
        ("splitk_ws", c_void_p), ("splitk_ws_bytes", c_int64),
        ("fl_ad", c_void_p), ("fl_bup", c_void_p), ("fl_kl", c_int32), ("fl_rank", c_int32),
        ("fl_ld_ad", c_int64), ("fl_ld_bup", c_int64), ("fl_scale", c_float), ("fl_t_out", c_void_p),
        ("fl_ld_t", c_int64), ("debug_mode", c_int32),
    ]


def lib_path() -> str:
    return _LIB_PATH


def load():
    """dlopen the library (works without a GPU: no CUDA call happens at load time)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise LecoError(
            f"{_LIB_PATH} not built. Run `python -m leco_b200.build` (needs nvcc). "
            "leco_b200 has no CPU / PyTorch fallback path.")
    lib = ctypes.CDLL(_LIB_PATH)
    lib.leco_last_error.restype = c_char_p
    lib.leco_abi_version.restype = c_int32
    lib.leco_launch_count.restype = c_int64
    lib.leco_device_info.argtypes = [POINTER(c_int32)] * 3
    lib.leco_gemm_bf16.argtypes = [POINTER(GemmArgs), c_void_p]
    lib.leco_group_norm_workspace_bytes.restype = c_int64
    lib.leco_group_norm_workspace_bytes.argtypes = [c_int32, c_int32]
    lib.leco_group_norm_barrier_bytes.restype = c_int64
    lib.leco_group_norm_barrier_bytes.argtypes = [c_int32]
    _declare_ops(lib)
    _lib = lib
    return lib


# (name, argtypes) of every other entry point; filled in as kernels are added
_OPS: list[tuple[str, list]] = [
    ("leco_conv_in", [P, I, P, P, P, I, I, I, I, P]),
    ("leco_conv_out", [P, P, P, P, I, I, I, I, I, P]),
    ("leco_conv_out_bwd", [P, P, P, I, I, I, I, I, P]),
    ("leco_cols_to_nchw", [P, I, P, P, I, I, I, P]),
    ("leco_timestep_embedding", [P, P, I, I, P]),
    ("leco_silu", [P, P, L, P]),
    ("leco_add_inplace", [P, P, L, P]),
    ("leco_geglu_fwd", [P, P, L, I, P]),
    ("leco_geglu_bwd", [P, P, P, L, I, P]),
    ("leco_copy_cols", [P, L, I, P, L, I, L, I, P]),
    ("leco_upsample2x", [P, P, I, I, I, I, P]),
    ("leco_upsample2x_bwd", [P, P, I, I, I, I, P]),
    ("leco_im2col_s2", [P, P, I, I, I, I, P]),
    ("leco_col2im_s2", [P, P, I, I, I, I, P]),
    ("leco_im2col_s1", [P, P, I, I, I, I, P]),
    ("leco_rowgroup_sum", [P, P, I, I, I, P]),
    ("leco_transpose", [P, P, I, I, I, L, L, L, L, L, L, I, I, P]),
    ("leco_softmax_rows", [P, P, L, I, I, L, L, P]),
    ("leco_softmax_rows_causal", [P, P, L, I, I, L, L, I, P]),
    ("leco_embed_tokens", [P, P, P, P, L, I, I, I, P]),
    ("leco_activation", [P, P, L, I, P]),
    ("leco_softmax_bwd_rows", [P, P, P, L, I, I, L, L, F, P]),
    ("leco_flash_attn_fwd", [P, L, P, L, P, L, P, L, P, L, I, I, I, I, I, F, P]),
    ("leco_flash_attn_fwd_lse", [P, L, P, L, P, L, P, L, P, I, I, I, I, I, F, P]),
    ("leco_attn_bwd_prep", [P, L, P, L, P, I, I, I, I, P]),
    ("leco_flash_attn_bwd", [P, L, P, L, P, L, P, L, P, P, P, P, L, P, L, I, I, I, I, I, F, P]),
    ("leco_attn_dq_cast", [P, P, L, I, I, I, I, P]),
    ("leco_group_norm", [P, P, P, P, P, I, I, I, I, F, I, P, P]),
    ("leco_group_norm_bwd", [P, P, P, P, P, P, I, I, I, I, I, P, P]),
    ("leco_group_norm_fused", [P, P, P, P, P, I, I, I, I, F, I, P, P, P]),
    ("leco_group_norm_v2", [P, P, P, P, P, I, I, I, I, F, I, P, P, P]),
    ("leco_group_norm_v3", [P, P, P, P, P, I, I, I, I, F, I, P, P, P]),
    ("leco_layer_norm", [P, P, P, P, P, L, I, F, P]),
    ("leco_layer_norm_bwd", [P, P, P, P, P, L, I, P]),
    ("leco_tn_reduce", [P, L, P, L, P, L, L, I, I, F, I, P]),
    ("leco_adamw_flat", [P, P, P, P, I, P, P, L, I, P]),
    ("leco_optim_flat", [P, P, P, P, I, P, P, L, I, P]),
    ("leco_optim_flat_master", [P, P, P, P, P, P, P, L, I, P]),
    ("leco_transpose_tiles", [P, P, P, I, P]),
    ("leco_set_deterministic", [I]),
    ("leco_guided_step", [P, P, P, P, P, L, P]),
    ("leco_sched_step", [P, P, P, P, P, P, L, P]),
    ("leco_scale_by_dev", [P, P, P, I, L, P]),
    ("leco_loss", [P, P, P, P, F, P, P, L, P]),
    ("leco_axpby", [P, P, P, F, F, L, I, P]),
    ("leco_cast_f32_to_bf16", [P, P, L, P]),
    ("leco_cast_bf16_to_f32", [P, P, L, P]),
]

# every symbol include/leco_b200.h declares (tests check the .so exports all of them)
EXPORTED = ["leco_last_error", "leco_abi_version", "leco_launch_count", "leco_device_info",
            "leco_gemm_bf16", "leco_group_norm_workspace_bytes", "leco_group_norm_barrier_bytes"] + [n for n, _ in _OPS]


def _declare_ops(lib):
    for name, argtypes in _OPS:
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int32


def check(rc: int, what: str):
    if rc != 0:
        msg = load().leco_last_error().decode(errors="replace")
        raise LecoError(f"{what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(load().leco_launch_count())
