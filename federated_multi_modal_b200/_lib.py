"""ctypes binding of libmfk.so (the C ABI declared in include/mfk.h).

There is NO fallback: if the shared library is missing it is built with nvcc; if that fails, or a
kernel returns a non-zero status, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MFK_LIB_PATH") or os.path.join(_HERE, "libmfk.so")  # env: A/B experiment builds

P, I, L, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float

# name -> argument ctypes (return type is int unless listed in _RET)
SIGNATURES = {
    "mfk_version": [],
    "mfk_error_string": [I],
    "mfk_debug_set_attn_trace": [P],
    "mfk_debug_set_gemm_trace": [P],
    "mfk_gemm_bf16": [P, L, P, L, I, I, I, P, I, P, L, P, L, P, L, P, L, P, L, I, P, L, P],
    "mfk_gemm_bf16_at_b": [P, L, P, L, I, I, I, P, L, P],
    "mfk_attn_fwd": [P, P, P, I, I, I, I, P],
    "mfk_attn_fwd_tc": [P, P, P, I, I, I, I, P],
    "mfk_attn_bwd": [P, P, P, P, P, P, I, I, I, I, P],
    "mfk_attn_bwd_tc": [P, P, P, P, P, P, I, I, I, I, P],
    "mfk_attn_bwd_fused": [P, P, P, P, P, P, I, I, I, P],
    "mfk_layernorm_fwd": [P, P, P, P, P, P, P, P, P, I, I, F, P],
    "mfk_layernorm_fwd_splice": [P, P, P, P, P, P, P, P, P, I, I, F, P, I, I, I, P],
    "mfk_layernorm_bwd_splice": [P, I, P, P, P, P, P, P, P, P, P, P, I, I, I, P, I, I, I, P],
    "mfk_ln_bwd_ctas": [I],
    "mfk_layernorm_bwd": [P, I, P, P, P, P, P, P, P, P, P, P, I, I, I, P],
    "mfk_colsum": [P, I, L, I, I, P, P, I, P],
    "mfk_patch_im2col": [P, P, I, I, P],
    "mfk_vis_assemble_lnpre": [P, P, P, P, P, P, P, P, P, P, I, I, I, I, F, P],
    "mfk_text_assemble": [P, P, P, P, P, I, I, I, I, I, P],
    "mfk_prompt_splice_fwd": [P, P, I, I, I, I, I, P],
    "mfk_prompt_splice_bwd": [P, P, P, I, I, I, I, I, I, I, P],
    "mfk_prompt_splice_bwd_batched": [P, L, P, L, I, I, I, I, I, I, I, P],
    "mfk_scatter_rows": [P, P, P, P, I, I, P],
    "mfk_scatter_rows_dense": [P, P, P, I, I, I, P],
    "mfk_gather_rows": [P, P, P, I, L, I, P],
    "mfk_transpose_bf16": [P, I, L, P, L, P, L, I, I, P],
    "mfk_cast_f32_bf16": [P, P, L, P],
    "mfk_split_bf16x3": [P, P, I, I, P],
    "mfk_quickgelu_split_bf16x3": [P, P, I, I, P],
    "mfk_patch_im2col_f32": [P, P, I, I, P],
    "mfk_attn_fwd_f32": [P, P, I, I, I, I, P],
    "mfk_attn_bwd_f32": [P, P, P, P, I, I, I, I, P],
    "mfk_dquickgelu_mul_f32": [P, P, P, L, P],
    "mfk_split_bf16x3_rhs": [P, P, I, I, P],
    "mfk_attn_rows_fwd": [P, P, P, P, I, I, I, I, P],
    "mfk_attn_rows_bwd": [P, P, P, P, P, I, I, I, I, P],
    "mfk_linear_small_fwd": [P, P, P, P, I, I, I, P],
    "mfk_linear_small_bwd": [P, P, P, P, P, P, P, I, I, I, P],
    "mfk_linear_small_fwd_grouped": [P, I, I, P],
    "mfk_linear_small_bwd_grouped": [P, I, I, I, I, P],
    "mfk_repack_grouped": [P, I, I, P],
    "mfk_partial_reduce_grouped": [P, I, I, P],
    "mfk_rrc_flip_normalize": [P, I, I, I, P, P, P, P, P, I, I, P],
    "mfk_head_workspace_floats": [I, I, I],
    "mfk_head_forward_backward": [P, P, P, P, P, P, P, P, P, I, I, I, P],
    "mfk_fedavg_reduce": [P, P, F, I, L, I, P, P, P, P],
    "mfk_fedavg_reduce_scatter": [P, P, F, I, L, L, L, P, P, I, P],
    "mfk_check_finite": [P, L, I, P, P],
    "mfk_grad_norm": [P, L, P, P, P],
    "mfk_sgd_step": [P, P, P, L, P, P, P, P, P],
}
_RET = {"mfk_error_string": ctypes.c_char_p, "mfk_head_workspace_floats": L}
_NO_STATUS = {"mfk_version", "mfk_error_string", "mfk_head_workspace_floats", "mfk_ln_bwd_ctas",
              "mfk_debug_set_attn_trace", "mfk_debug_set_gemm_trace"}

_lock = threading.Lock()
_lib = None
launch_count = 0  # C-ABI calls that launched work
kernel_count = 0  # CUDA kernels those calls launched (bench.py reports it as gpu_launches)


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise RuntimeError(f"{LIB_PATH} is missing: run `python -m federated_multi_modal_b200.csrc.build`")
            # one builder at a time: N ranks of a torchrun job may all find the library missing
            import fcntl
            from .csrc.build import build
            with open(LIB_PATH + ".lock", "w") as lk:
                fcntl.flock(lk, fcntl.LOCK_EX)
                if not os.path.exists(LIB_PATH):
                    build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.argtypes = args
            fn.restype = _RET.get(name, I)
        _lib = lib
        return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args, kernels: int = 1):
    """Invoke a C-ABI entry point; tensors are passed as raw pointers. Raises on non-zero status.
    ``kernels`` = number of CUDA kernels this call launches (for the launch accounting only)."""
    global launch_count, kernel_count
    lib = load()
    conv = [a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args]
    rc = getattr(lib, name)(*conv)
    if name in _NO_STATUS:
        return rc
    launch_count += 1
    kernel_count += kernels
    if rc != 0:
        msg = lib.mfk_error_string(rc)
        raise RuntimeError(f"{name} failed with status {rc}: {msg.decode() if msg else '?'}")
    return 0
