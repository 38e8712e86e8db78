"""ctypes binding of libunigen_b200.so (C ABI: include/unigen_b200.h).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  A missing library
or a missing symbol is a hard error -- the product has no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UNIGEN_B200_LIB") or os.path.join(_HERE, "csrc", "libunigen_b200.so")   # env override: kernel experiments only

UG_ABI_VERSION = 2
SYMBOLS = (
    "ug_abi_version", "ug_last_error", "ug_decode_step_workspace_bytes", "ug_decode_step_bf16",
    "ug_mask_by_random_topk", "ug_export_noise", "ug_omni_attention_bf16", "ug_omni_attention_rows_bf16", "ug_add_rmsnorm", "ug_rope_qk", "ug_swiglu", "ug_embed_rows", "ug_gemm_bf16", "ug_gemm_residual_bf16", "ug_qkv_rope_bf16", "ug_lfq_unpack_bf16", "ug_lfq_postquant_bf16", "ug_conv_nhwc_bf16", "ug_softmax_rows_bf16", "ug_groupnorm_workspace_floats", "ug_groupnorm_nhwc_bf16", "ug_bias_add_nhwc_bf16", "ug_bias_add_stats_nhwc_bf16", "ug_cfg_mix_bf16",
    "ug_upsample2x_nhwc_bf16",
)

# include/unigen_b200_test.h: the same entry points with explicit, per-call test instrumentation (stateless)
TEST_SYMBOLS = ("ugt_decode_step_bf16", "ugt_omni_attention_rows_bf16")

_lib = None


class DecodeTestOptions(ctypes.Structure):
    """ugt_decode_options of include/unigen_b200_test.h."""
    _fields_ = [("rowsum_out", ctypes.c_void_p), ("force_exact_bits", ctypes.c_int), ("exact_kernel", ctypes.c_int),
                ("stats", ctypes.c_void_p)]


class UnigenB200Error(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UnigenB200Error(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no fallback path.")
    L = ctypes.CDLL(LIB_PATH)
    for s in SYMBOLS + TEST_SYMBOLS:
        if not hasattr(L, s):
            raise UnigenB200Error(f"{LIB_PATH} does not export {s}")
    c = ctypes
    L.ug_abi_version.restype = c.c_int
    L.ug_last_error.restype = c.c_char_p
    L.ug_decode_step_workspace_bytes.restype = c.c_size_t
    L.ug_decode_step_workspace_bytes.argtypes = [c.c_int, c.c_int, c.c_int]
    L.ug_decode_step_bf16.restype = c.c_int
    L.ug_decode_step_bf16.argtypes = [
        c.c_void_p, c.c_void_p, c.c_float,
        c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
        c.c_void_p, c.c_void_p,
        c.c_uint64, c.c_uint64, c.c_int64,
        c.c_int, c.c_int, c.c_int, c.c_int64, c.c_int, c.c_float,
        c.c_void_p, c.c_size_t, c.c_void_p]
    L.ug_mask_by_random_topk.restype = c.c_int
    L.ug_mask_by_random_topk.argtypes = [
        c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint64, c.c_uint64, c.c_int64,
        c.c_float, c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_void_p]
    L.ug_export_noise.restype = c.c_int
    L.ug_export_noise.argtypes = [c.c_uint64, c.c_uint64, c.c_int64, c.c_int, c.c_int, c.c_int,
                                  c.c_void_p, c.c_void_p, c.c_void_p]
    L.ug_omni_attention_bf16.restype = c.c_int
    L.ug_omni_attention_bf16.argtypes = [
        c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
        c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_float,
        c.c_int64, c.c_int64, c.c_int64, c.c_int64, c.c_int64, c.c_int64, c.c_int64, c.c_int64, c.c_int64,
        c.c_void_p]
    L.ug_omni_attention_rows_bf16.restype = c.c_int
    L.ug_omni_attention_rows_bf16.argtypes = [
        c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
        c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_float,
        c.c_int64, c.c_int64, c.c_int64, c.c_int64, c.c_int64, c.c_int64, c.c_int64, c.c_int64, c.c_int64,
        c.c_void_p]
    L.ugt_decode_step_bf16.restype = c.c_int
    L.ugt_decode_step_bf16.argtypes = L.ug_decode_step_bf16.argtypes + [c.POINTER(DecodeTestOptions)]
    L.ugt_omni_attention_rows_bf16.restype = c.c_int
    L.ugt_omni_attention_rows_bf16.argtypes = L.ug_omni_attention_rows_bf16.argtypes + [c.c_void_p]
    L.ug_add_rmsnorm.restype = c.c_int
    L.ug_add_rmsnorm.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_float, c.c_void_p]
    L.ug_rope_qk.restype = c.c_int
    L.ug_rope_qk.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_int, c.c_int, c.c_int,
                             c.c_void_p, c.c_int, c.c_int, c.c_void_p]
    L.ug_swiglu.restype = c.c_int
    L.ug_swiglu.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_void_p]
    L.ug_embed_rows.restype = c.c_int
    L.ug_embed_rows.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_int64, c.c_void_p]
    L.ug_gemm_bf16.restype = c.c_int
    L.ug_gemm_bf16.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_int,
                               c.c_int64, c.c_int64, c.c_int64, c.c_int, c.c_void_p, c.c_int, c.c_float, c.c_void_p]
    L.ug_gemm_residual_bf16.restype = c.c_int
    L.ug_gemm_residual_bf16.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_int,
                                        c.c_int64, c.c_int64, c.c_void_p]
    L.ug_qkv_rope_bf16.restype = c.c_int
    L.ug_qkv_rope_bf16.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int,
                                   c.c_int64, c.c_int64, c.c_int64, c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_int,
                                   c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_int, c.c_float, c.c_void_p]
    L.ug_lfq_unpack_bf16.restype = c.c_int
    L.ug_lfq_unpack_bf16.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_void_p]
    L.ug_lfq_postquant_bf16.restype = c.c_int
    L.ug_lfq_postquant_bf16.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_int, c.c_void_p]
    L.ug_conv_nhwc_bf16.restype = c.c_int
    L.ug_conv_nhwc_bf16.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p]
    L.ug_softmax_rows_bf16.restype = c.c_int
    L.ug_softmax_rows_bf16.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_int, c.c_float, c.c_void_p]
    L.ug_groupnorm_workspace_floats.restype = c.c_size_t
    L.ug_groupnorm_workspace_floats.argtypes = [c.c_int, c.c_int64, c.c_int]
    L.ug_groupnorm_nhwc_bf16.restype = c.c_int
    L.ug_groupnorm_nhwc_bf16.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int64,
                                         c.c_int, c.c_float, c.c_int, c.c_int, c.c_void_p]
    L.ug_bias_add_stats_nhwc_bf16.restype = c.c_int
    L.ug_bias_add_stats_nhwc_bf16.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int64,
                                              c.c_int, c.c_void_p]
    L.ug_bias_add_nhwc_bf16.restype = c.c_int
    L.ug_bias_add_nhwc_bf16.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_void_p]
    L.ug_upsample2x_nhwc_bf16.restype = c.c_int
    L.ug_upsample2x_nhwc_bf16.argtypes = [c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p]
    L.ug_cfg_mix_bf16.restype = c.c_int
    L.ug_cfg_mix_bf16.argtypes = [c.c_void_p, c.c_void_p, c.c_float, c.c_int64, c.c_void_p]
    if L.ug_abi_version() != UG_ABI_VERSION:
        raise UnigenB200Error(f"ABI version {L.ug_abi_version()} != {UG_ABI_VERSION}")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise UnigenB200Error(f"libunigen_b200 rc={rc}: {lib().ug_last_error().decode()}")
