"""Thin tensor-level wrappers over the C ABI (one function per entry point of include/mfk.h).

All tensors must live on the current CUDA device; work is enqueued on torch's current stream.
No op here has a CPU or PyTorch fallback.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import call, stream_ptr

BF16, F32 = torch.bfloat16, torch.float32


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: libmfk kernels need CUDA tensors (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if t.dim() >= 2 and t.stride(-1) != 1:
        raise ValueError(f"{name}: innermost dimension must be contiguous")


def gemm(a: torch.Tensor, b: torch.Tensor, *, bias: Optional[torch.Tensor] = None, act: int = 0,
         aux: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         out_f32: Optional[torch.Tensor] = None, out_bf16: Optional[torch.Tensor] = None,
         out_pre: Optional[torch.Tensor] = None, k: Optional[int] = None, tile_n: int = 0,
         ws: Optional[torch.Tensor] = None):
    """out[M,N] = epi(a[M,K] @ b[N,K]^T); see mfk_gemm_bf16. `k` overrides K (padded operands). `ws`: zero-filled
    split-K workspace (splitk_workspace()) owned by the GEMMs of one stream; None disables the split-K tail."""
    _chk(a, BF16, "a"); _chk(b, BF16, "b")
    M, N = a.shape[0], b.shape[0]
    K = a.shape[1] if k is None else k
    ld = lambda t: t.stride(0) if t is not None else 0
    call("mfk_gemm_bf16", a, a.stride(0), b, b.stride(0), M, N, K, bias, act, aux, ld(aux), residual, ld(residual),
         out_f32, ld(out_f32), out_bf16, ld(out_bf16), out_pre, ld(out_pre), tile_n, ws,
         ws.numel() * ws.element_size() if ws is not None else 0, stream_ptr())


SPLITK_WS_BYTES = 4096 + 148 * 128 * 256 * 4


def splitk_workspace(device) -> torch.Tensor:
    """Zero-filled workspace for the split-K tail of ops.gemm (one per stream that runs large GEMMs)."""
    return torch.zeros(SPLITK_WS_BYTES // 4, device=device, dtype=F32)


def gemm_at_b(at: torch.Tensor, bt: torch.Tensor, out_f32: torch.Tensor):
    """out[M,N] = at[K,M]^T @ bt[K,N] (bf16 in, fp32 out): wgrad straight from row-major dY and X."""
    _chk(at, BF16, "at"); _chk(bt, BF16, "bt")
    call("mfk_gemm_bf16_at_b", at, at.stride(0), bt, bt.stride(0), at.shape[1], bt.shape[1], at.shape[0], out_f32,
         out_f32.stride(0), stream_ptr())


TC_ATTN_MIN_T = 65  # sequences longer than this use the tcgen05 attention kernels


def attn_fwd(qkv, out, lse, N, T, heads, causal, impl: Optional[str] = None):
    """impl: None = by sequence length, "tc" = tcgen05/TMEM kernel, "mma" = warp-level mma.sync kernel."""
    _chk(qkv, BF16, "qkv"); _chk(out, BF16, "out")
    use_tc = (T >= TC_ATTN_MIN_T) if impl is None else impl == "tc"
    call("mfk_attn_fwd_tc" if use_tc else "mfk_attn_fwd", qkv, out, lse, N, T, heads, int(causal), stream_ptr())


def attn_bwd(qkv, out, d_out, lse, delta_ws, dqkv, N, T, heads, causal, impl: Optional[str] = None):
    if impl == "fused" or (impl is None and not causal and T >= TC_ATTN_MIN_T):
        call("mfk_attn_bwd_fused", qkv, out, d_out, lse, delta_ws, dqkv, N, T, heads, stream_ptr(), kernels=1)
        return
    use_tc = (TC_ATTN_MIN_T <= T <= 240) if impl is None else impl == "tc"
    call("mfk_attn_bwd_tc" if use_tc else "mfk_attn_bwd", qkv, out, d_out, lse, delta_ws, dqkv, N, T, heads,
         int(causal), stream_ptr(), kernels=1 if (not use_tc and T <= 32) else 3)  # T <= 32: one fused small-T launch


def layernorm_fwd(x, gamma, beta, *, rowidx=None, y_bf16=None, y_f32=None, x_save=None, mean=None, rstd=None,
                  M=None, eps=1e-5, splice=None):
    """splice = (prompt [n_ctx, D] fp32, T, row0, n_ctx): fused deep-prompt splice (x is updated in place)."""
    _chk(x, F32, "x")
    D = x.shape[-1]
    M = (rowidx.numel() if rowidx is not None else x.numel() // D) if M is None else M
    if splice is None:
        call("mfk_layernorm_fwd", x, rowidx, gamma, beta, y_bf16, y_f32, x_save, mean, rstd, M, D, eps, stream_ptr())
    else:
        prompt, T, row0, n_ctx = splice
        call("mfk_layernorm_fwd_splice", x, rowidx, gamma, beta, y_bf16, y_f32, x_save, mean, rstd, M, D, eps, prompt,
             T, row0, n_ctx, stream_ptr())


def ln_bwd_ctas(M: int) -> int:
    return call("mfk_ln_bwd_ctas", M)


def layernorm_bwd(dy, x, mean, rstd, gamma, *, g_in=None, g_out, g_out_bf16=None, dgamma=None, dbeta=None,
                  partial_ws=None, accumulate=False, M=None, defer=False, splice_grad=None):
    """defer: leave the per-CTA dgamma/dbeta partials in partial_ws (reduced later by partial_reduce_grouped).
    splice_grad = (gprompt [N, n_ctx, D] fp32, T, row0, n_ctx): fused backward of the deep-prompt splice."""
    D = x.shape[-1]
    M = x.numel() // D if M is None else M
    flags = int(accumulate) | (2 if defer else 0)
    k = 1 + ((dgamma is not None or dbeta is not None) and not defer)
    if splice_grad is None:
        call("mfk_layernorm_bwd", dy, int(dy.dtype == BF16), x, mean, rstd, gamma, g_in, g_out, g_out_bf16, dgamma,
             dbeta, partial_ws, flags, M, D, stream_ptr(), kernels=k)
    else:
        gp, T, row0, n_ctx = splice_grad
        call("mfk_layernorm_bwd_splice", dy, int(dy.dtype == BF16), x, mean, rstd, gamma, g_in, g_out, g_out_bf16,
             dgamma, dbeta, partial_ws, flags, M, D, gp, T, row0, n_ctx, stream_ptr(), kernels=k)


def partial_reduce_table(problems, device) -> torch.Tensor:
    """Device table of mfk_partial_reduce_problem from (partial, P, N, out0, out1, accumulate) tuples."""
    rows = [[pt.data_ptr(), P | (N << 32), o0.data_ptr() if o0 is not None else 0,
             o1.data_ptr() if o1 is not None else 0, int(acc)] for pt, P, N, o0, o1, acc in problems]
    return torch.tensor(rows, dtype=torch.int64, device=device).contiguous()


def partial_reduce_grouped(table, max_N):
    call("mfk_partial_reduce_grouped", table, table.shape[0], max_N, stream_ptr())


def colsum(x, out, partial_ws, accumulate=False):
    call("mfk_colsum", x, int(x.dtype == BF16), x.stride(0), x.shape[0], x.shape[1], out, partial_ws,
         int(accumulate), stream_ptr(), kernels=2)


def patch_im2col(img, out):
    _chk(img, F32, "img")
    call("mfk_patch_im2col", img, out, img.shape[0], img.shape[-1], stream_ptr())


def vis_assemble_lnpre(tok, cls, pos, shared_ctx, gamma, beta, x0_save, x, mean, rstd, B, T, n_ctx, eps=1e-5):
    call("mfk_vis_assemble_lnpre", tok, cls, pos, shared_ctx, gamma, beta, x0_save, x, mean, rstd, B, T, n_ctx,
         x.shape[-1], eps, stream_ptr())


def text_assemble(prefix, ctx, suffix, pos, x, C, Te, n_ctx, Tfull):
    call("mfk_text_assemble", prefix, ctx, suffix, pos, x, C, Te, n_ctx, Tfull, x.shape[-1], stream_ptr())


def prompt_splice_fwd(x, prompt, N, T, row0, n_ctx):
    call("mfk_prompt_splice_fwd", x, prompt, N, T, row0, n_ctx, x.shape[-1], stream_ptr())


def prompt_splice_bwd(g, g_bf16, dprompt, N, T, row0, n_ctx, round_fp16=True, zero_rows=True):
    call("mfk_prompt_splice_bwd", g, g_bf16, dprompt, N, T, row0, n_ctx, g.shape[-1], int(round_fp16),
         int(zero_rows), stream_ptr())


def prompt_splice_bwd_batched(g_all, dprompt_all, N, T, row0, n_ctx, round_fp16=True):
    """g_all [layers, N*T, D] (or any per-layer tensor with row = sequence*T + t), dprompt_all [layers, n_ctx, D]:
    dprompt[l, j] = sum over sequences of (fp16-rounded) g[l, b*T + row0 + j] in batch order, one launch."""
    Lr, D = g_all.shape[0], g_all.shape[-1]
    call("mfk_prompt_splice_bwd_batched", g_all, g_all.stride(0), dprompt_all, dprompt_all.stride(0), Lr, N, T, row0,
         n_ctx, D, int(round_fp16), stream_ptr())


def scatter_rows(dx, rowidx, g, g_bf16):
    call("mfk_scatter_rows", dx, rowidx, g, g_bf16, dx.shape[0], dx.shape[1], stream_ptr())


def scatter_rows_dense(dx, rowidx, g, N, T):
    """g[N*T, D] = zeros except g[rowidx[n]] = dx[n] (one consumed row per sequence); no prior zero fill needed."""
    _chk(dx, F32, "dx"); _chk(g, F32, "g")
    assert g.shape[0] == N * T and dx.shape[0] == N and rowidx.dtype == torch.int32
    call("mfk_scatter_rows_dense", dx, rowidx, g, N, T, dx.shape[1], stream_ptr())


def gather_rows(src, rowidx, dst, scatter=False):
    """dst[r] = src[rowidx[r]] (or the inverse scatter into a pre-zeroed dst); rows of equal byte length."""
    call("mfk_gather_rows", src, rowidx, dst, rowidx.numel(), src.shape[-1] * src.element_size(), int(scatter),
         stream_ptr())


def transpose_bf16(inp, out, copy=None):
    """out[N, ldo] = inp[M, N]^T as bf16 (inp fp32 or bf16)."""
    call("mfk_transpose_bf16", inp, int(inp.dtype == F32), inp.stride(0), out, out.stride(0), copy,
         copy.stride(0) if copy is not None else 0, inp.shape[0], inp.shape[1], stream_ptr())


def cast_bf16(inp, out):
    call("mfk_cast_f32_bf16", inp, out, inp.numel(), stream_ptr())


def split_bf16x3(x, out):
    call("mfk_split_bf16x3", x, out, x.shape[0], x.shape[1], stream_ptr())


def quickgelu_split_bf16x3(u, out):
    _chk(u, F32, "u")
    call("mfk_quickgelu_split_bf16x3", u, out, u.shape[0], u.shape[1], stream_ptr())


def patch_im2col_f32(img, out):
    _chk(img, F32, "img")
    call("mfk_patch_im2col_f32", img, out, img.shape[0], img.shape[-1], stream_ptr())


def attn_rows_fwd(qkv, rows, out_rows, lse_rows, N, T, heads, causal):
    """Attention core for ONE query row per sequence (rows int32 [N], global row indices): the last block's CLS/EOT."""
    _chk(qkv, BF16, "qkv"); _chk(out_rows, BF16, "out_rows")
    call("mfk_attn_rows_fwd", qkv, rows, out_rows, lse_rows, N, T, heads, int(causal), stream_ptr())


def attn_rows_bwd(qkv, rows, d_out_rows, lse_rows, dqkv, N, T, heads, causal):
    _chk(qkv, BF16, "qkv"); _chk(d_out_rows, BF16, "d_out_rows"); _chk(dqkv, BF16, "dqkv")
    call("mfk_attn_rows_bwd", qkv, rows, d_out_rows, lse_rows, dqkv, N, T, heads, int(causal), stream_ptr())


def attn_fwd_f32(qkv, out, N, T, heads, causal):
    _chk(qkv, F32, "qkv"); _chk(out, F32, "out")
    call("mfk_attn_fwd_f32", qkv, out, N, T, heads, int(causal), stream_ptr())


def attn_bwd_f32(qkv, d_out, dqkv, stat_ws, N, T, heads, causal):
    """fp32 training mode: dqkv fp32 [N*T, 3D] from d_out fp32 [N*T, D]; stat_ws: 2*N*heads*T floats."""
    _chk(qkv, F32, "qkv"); _chk(d_out, F32, "d_out"); _chk(dqkv, F32, "dqkv"); _chk(stat_ws, F32, "stat_ws")
    assert stat_ws.numel() >= 2 * N * heads * T
    call("mfk_attn_bwd_f32", qkv, d_out, dqkv, stat_ws, N, T, heads, int(causal), stream_ptr(), kernels=2)


def dquickgelu_mul_f32(dact, u, du):
    _chk(dact, F32, "dact"); _chk(u, F32, "u"); _chk(du, F32, "du")
    call("mfk_dquickgelu_mul_f32", dact, u, du, u.numel(), stream_ptr())


def split_bf16x3_rhs(x, out):
    """out[rows, 3D] = [hi | hi | lo] (B-side packing; split_bf16x3 gives the A-side [hi | lo | hi])."""
    _chk(x, F32, "x"); _chk(out, BF16, "out")
    call("mfk_split_bf16x3_rhs", x, out, x.shape[0], x.shape[1], stream_ptr())


def linear_small_fwd(x, W, b, y):
    call("mfk_linear_small_fwd", x, W, b, y, x.shape[0], W.shape[0], W.shape[1], stream_ptr())


def linear_small_bwd(x, W, dy, *, dW=None, db=None, dx_add=None, dx=None):
    call("mfk_linear_small_bwd", x, W, dy, dW, db, dx_add, dx, x.shape[0], W.shape[0], W.shape[1], stream_ptr(),
         kernels=(dW is not None) + (dx is not None))


def small_linear_table(problems, device) -> torch.Tensor:
    """Device table of mfk_small_linear_problem (see mfk.h) from dicts with tensors x, W, b, y, dy, dW, db, dx_add,
    dx (missing / None -> NULL). Pointers are captured: the tensors must stay alive and in place."""
    rows = []
    for pr in problems:
        ptr = lambda k: (pr[k].data_ptr() if pr.get(k) is not None else 0)
        m, (N, K) = pr["x"].shape[0], pr["W"].shape
        rows.append([ptr(k) for k in ("x", "W", "b", "y", "dy", "dW", "db", "dx_add", "dx")] + [m | (N << 32), K])
    return torch.tensor(rows, dtype=torch.int64, device=device).contiguous()


def linear_small_fwd_grouped(table, max_N):
    call("mfk_linear_small_fwd_grouped", table, table.shape[0], max_N, stream_ptr())


def linear_small_bwd_grouped(table, max_m, max_N, max_K):
    call("mfk_linear_small_bwd_grouped", table, table.shape[0], max_m, max_N, max_K, stream_ptr(), kernels=2)


def repack_table(problems, device):
    """(device table of mfk_repack_problem, total 64x64 tiles) from (master fp32 [M,N], out_t bf16 [N,M],
    copy bf16 [M,N]) triples."""
    rows, tile0 = [], 0
    for w, t, c in problems:
        M, N = w.shape
        tm, tn = (M + 63) // 64, (N + 63) // 64
        rows.append([w.data_ptr(), t.data_ptr(), c.data_ptr(), M | (N << 32), tile0 | (tn << 32)])
        tile0 += tm * tn
    return torch.tensor(rows, dtype=torch.int64, device=device).contiguous(), tile0


def repack_grouped(table, total_tiles):
    call("mfk_repack_grouped", table, table.shape[0], total_tiles, stream_ptr())


def rrc_flip_normalize(src_u8, boxes, flip, mean, std, out, round_u8=True):
    """RandomResizedCrop (antialiased bicubic) + flip + ToTensor + Normalize of a uint8 [B,3,H,W] batch on the GPU;
    boxes int32 [B,4] = (top, left, height, width), flip uint8 [B] (drawn on the host), out fp32 [B,3,S,S]."""
    assert src_u8.is_cuda and src_u8.dtype == torch.uint8 and src_u8.is_contiguous() and src_u8.dim() == 4
    assert boxes.dtype == torch.int32 and flip.dtype == torch.uint8 and out.dtype == F32 and out.is_contiguous()
    B, _, H, W = src_u8.shape
    call("mfk_rrc_flip_normalize", src_u8, B, H, W, boxes, flip, mean, std, out, out.shape[-1], int(round_u8),
         stream_ptr())


def head_workspace_floats(B, C, E) -> int:
    return call("mfk_head_workspace_floats", B, C, E)


def head_forward_backward(img_feat, txt_feat, logit_scale, label, logits, loss, d_img, d_txt, ws):
    B, E = img_feat.shape
    C = txt_feat.shape[0]
    # training: ONE fused launch (8-CTA cluster) when the features fit its shared memory (mfk_head.cu), else six kernels
    nb = (B + 7) // 8
    fused = label is not None and E == 512 and 4 * (2 * C * E + nb * E + 2 * nb * C + 4 * nb + C + B) <= 227 * 1024
    call("mfk_head_forward_backward", img_feat, txt_feat, logit_scale, label, logits, loss, d_img, d_txt, ws, B, C, E,
         stream_ptr(), kernels=2 if label is None else (1 if fused else 6))


def fedavg_reduce(ptrs_dev, weights_dev, divisor, K, n, in_is_fp16, out_f32, out_f16, flags_dev):
    call("mfk_fedavg_reduce", ptrs_dev, weights_dev, float(divisor), K, n, int(in_is_fp16), out_f32, out_f16,
         flags_dev, stream_ptr())


def fedavg_reduce_scatter(ptrs_dev, weights_dev, divisor, K, n, lo, hi, out32_ptrs_dev, out16_ptrs_dev, W):
    """Sharded fixed-order FedAvg: this rank reduces [lo, hi) and pushes the result into all W ranks' buffers."""
    call("mfk_fedavg_reduce_scatter", ptrs_dev, weights_dev, float(divisor), K, n, lo, hi, out32_ptrs_dev,
         out16_ptrs_dev, W, stream_ptr())


def check_finite(t, flag_dev):
    code = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[t.dtype]
    call("mfk_check_finite", t, t.numel(), code, flag_dev, stream_ptr())


def grad_norm(g, partial_ws, norm_out):
    call("mfk_grad_norm", g, g.numel(), partial_ws, norm_out, stream_ptr(), kernels=2)


def sgd_step(p, g, mom, hyper_dev, total_norm_dev, n=None, loss_dev=None, flag_dev=None):
    """loss_dev (fp32 [1]) / flag_dev (int32 [1]): optional device-side guard — a non-finite loss or a raised input
    flag skips the update (the reference raises before optim.step(), trainers/maple.py:375-376, 556-557)."""
    call("mfk_sgd_step", p, g, mom, p.numel() if n is None else n, hyper_dev, total_norm_dev, loss_dev, flag_dev,
         stream_ptr())
