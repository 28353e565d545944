"""GPU: every C-ABI kernel against a plain torch fp32 reference of the same op."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from federated_multi_modal_b200 import ops

DEV = "cuda"
BF16, F32 = torch.bfloat16, torch.float32


def rnd(*shape, std=1.0, dtype=F32, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * std).to(DEV).to(dtype)


def qgelu(u):
    return u * torch.sigmoid(1.702 * u)


def dqgelu(u):
    s = torch.sigmoid(1.702 * u)
    return s * (1 + 1.702 * u * (1 - s))


GEMM_SHAPES = [(128, 128, 64), (256, 256, 128), (100, 64, 72), (6368, 2304, 768), (6368, 768, 768),
               (6368, 3072, 768), (6368, 768, 3072), (770, 1536, 512), (770, 512, 2048), (32, 512, 768),
               (10, 512, 512), (12736, 768, 768), (300, 32, 40)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("tile_n", [0, 128, 2])  # auto | 128-wide tiles | CTA-pair (cta_group::2) kernel
def test_gemm_plain(M, N, K, tile_n):
    a, b = rnd(M, K, dtype=BF16, seed=1), rnd(N, K, std=K ** -0.5, dtype=BF16, seed=2)
    out32 = torch.full((M, N), float("nan"), device=DEV, dtype=F32)
    out16 = torch.empty(M, N, device=DEV, dtype=BF16)
    ops.gemm(a, b, out_f32=out32, out_bf16=out16, tile_n=tile_n)
    ref = a.float() @ b.float().t()
    torch.cuda.synchronize()
    assert torch.isfinite(out32).all()
    err = (out32 - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err
    assert (out16.float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K", [(6368, 768, 768), (770, 512, 512), (200, 96, 128)])
@pytest.mark.parametrize("tile_n", [0, 2])
def test_gemm_epilogues(M, N, K, tile_n):
    a, b = rnd(M, K, dtype=BF16, seed=3), rnd(N, K, std=K ** -0.5, dtype=BF16, seed=4)
    bias = rnd(N, std=0.1, seed=5)
    res = rnd(M, N, seed=6)
    ref = a.float() @ b.float().t() + bias
    # bias + residual -> fp32
    out32 = torch.empty(M, N, device=DEV, dtype=F32)
    ops.gemm(a, b, bias=bias, residual=res, out_f32=out32, tile_n=tile_n)
    assert (out32 - (ref + res)).abs().max().item() < 5e-3
    # in-place residual (out aliases residual), as the towers use it
    x = res.clone()
    ops.gemm(a, b, bias=bias, residual=x, out_f32=x, tile_n=tile_n)
    assert (x - (ref + res)).abs().max().item() < 5e-3
    # QuickGELU with pre-activation saved
    act = torch.empty(M, N, device=DEV, dtype=BF16)
    pre = torch.empty(M, N, device=DEV, dtype=BF16)
    ops.gemm(a, b, bias=bias, act=1, out_bf16=act, out_pre=pre, tile_n=tile_n)
    assert (pre.float() - ref).abs().max().item() < 3e-2
    assert (act.float() - qgelu(pre.float())).abs().max().item() < 2e-2
    # backward of QuickGELU: multiply by gelu'(aux)
    dg = torch.empty(M, N, device=DEV, dtype=BF16)
    ops.gemm(a, b, act=2, aux=pre, out_bf16=dg, tile_n=tile_n)
    want = (a.float() @ b.float().t()) * dqgelu(pre.float())
    assert (dg.float() - want).abs().max().item() < 3e-2


@pytest.mark.parametrize("M,N,K", [(6368, 768, 3072), (6368, 768, 2304), (12736, 768, 3072), (1000, 800, 3072),
                                   (3184, 768, 1536)])
def test_gemm_splitk_tail(M, N, K):
    """Split-K tail (workspace given): same results as the plain path for every epilogue, bit-reproducible run to
    run, and the device-side counters re-arm themselves (the workspace is reused without a memset)."""
    a, b = rnd(M, K, dtype=BF16, seed=7), rnd(N, K, std=K ** -0.5, dtype=BF16, seed=8)
    bias, res = rnd(N, std=0.1, seed=9), rnd(M, N, seed=10)
    ws = ops.splitk_workspace(DEV)
    ref = a.float() @ b.float().t()
    outs = []
    for rep in range(3):
        o32 = torch.full((M, N), float("nan"), device=DEV, dtype=F32)
        ops.gemm(a, b, bias=bias, residual=res, out_f32=o32, ws=ws)
        outs.append(o32)
    torch.cuda.synchronize()
    assert (ws.view(torch.int32)[:1024] == 0).all()          # counters re-armed
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert (outs[0] - (ref + bias + res)).abs().max().item() < 5e-3
    plain = torch.empty(M, N, device=DEV, dtype=F32)
    ops.gemm(a, b, bias=bias, residual=res, out_f32=plain)
    assert (plain - outs[0]).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())
    # bf16 output, QuickGELU with saved pre-activation, and its backward, all through the split-K fix-up
    o16 = torch.empty(M, N, device=DEV, dtype=BF16)
    ops.gemm(a, b, out_bf16=o16, ws=ws)
    assert (o16.float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())
    act = torch.empty(M, N, device=DEV, dtype=BF16)
    pre = torch.empty(M, N, device=DEV, dtype=BF16)
    ops.gemm(a, b, bias=bias, act=1, out_bf16=act, out_pre=pre, ws=ws)
    assert (pre.float() - (ref + bias)).abs().max().item() < 3e-2
    assert (act.float() - qgelu(pre.float())).abs().max().item() < 2e-2
    dg = torch.empty(M, N, device=DEV, dtype=BF16)
    ops.gemm(a, b, act=2, aux=pre, out_bf16=dg, ws=ws)
    assert (dg.float() - ref * dqgelu(pre.float())).abs().max().item() < 3e-2
    torch.cuda.synchronize()
    assert (ws.view(torch.int32)[:1024] == 0).all()


def test_gemm_padded_k_operands():
    # wgrad form: operands are transposed copies whose leading dimension is padded to a multiple of 8
    M, N, K, ld = 768, 512, 770, 776
    a = torch.zeros(M, ld, device=DEV, dtype=BF16); a[:, :K] = rnd(M, K, dtype=BF16, seed=7)
    b = torch.zeros(N, ld, device=DEV, dtype=BF16); b[:, :K] = rnd(N, K, dtype=BF16, seed=8)
    a[:, K:] = float("nan"); b[:, K:] = float("nan")  # padding must never be read (TMA bounds = K)
    out = torch.empty(M, N, device=DEV, dtype=F32)
    ops.gemm(a, b, out_f32=out, k=K)
    ref = a[:, :K].float() @ b[:, :K].float().t()
    assert (out - ref).abs().max().item() < 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(768, 3072, 6368), (3072, 768, 6368), (512, 512, 100), (1536, 512, 770),
                                   (128, 64, 64), (200, 96, 1000)])
def test_gemm_at_b_wgrad_form(M, N, K):
    at, bt = rnd(K, M, dtype=BF16, seed=60), rnd(K, N, dtype=BF16, seed=61)
    out = torch.full((M, N), float("nan"), device=DEV, dtype=F32)
    ops.gemm_at_b(at, bt, out)
    ref = at.float().t() @ bt.float()
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


def _attn_ref(qkv, N, T, heads, causal):
    D = heads * 64
    q, k, v = qkv.float().reshape(N, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    if causal:
        s = s + torch.full((T, T), float("-inf"), device=qkv.device).triu_(1)
    p = torch.softmax(s, -1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(N * T, D)
    return o, p, (q, k, v)


@pytest.mark.parametrize("N,T,heads,causal", [(2, 199, 12, False), (3, 77, 8, True), (4, 16, 8, True),
                                              (2, 64, 2, False), (1, 256, 1, True), (5, 11, 8, True),
                                              (32, 199, 12, False), (3, 10, 8, False), (2, 32, 8, True),
                                              (2, 33, 8, True)])  # T <= 32: one-launch small-T backward (text tower)
@pytest.mark.parametrize("impl", ["mma", "tc", "fused"])
def test_attention_fwd_bwd(N, T, heads, causal, impl):
    if impl == "fused" and causal:
        pytest.skip("the fused backward is the non-causal (vision) path")
    D = heads * 64
    qkv = rnd(N * T, 3 * D, dtype=BF16, seed=9)
    out = torch.empty(N * T, D, device=DEV, dtype=BF16)
    lse = torch.empty(N, heads, T, device=DEV, dtype=F32)
    ops.attn_fwd(qkv, out, lse, N, T, heads, causal, impl="tc" if impl == "fused" else impl)
    x = qkv.float().requires_grad_(True)
    o_ref, p, _ = _attn_ref(x, N, T, heads, causal)
    assert (out.float() - o_ref).abs().max().item() < 2e-2
    d_out = rnd(N * T, D, dtype=BF16, seed=10)
    o_ref.backward(d_out.float())
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(N * heads * T, device=DEV, dtype=F32)
    if impl == "tc" and T > 240:
        with pytest.raises(RuntimeError, match="unsupported shape"):
            ops.attn_bwd(qkv, out, d_out, lse, delta, dqkv, N, T, heads, causal, impl="tc")
        impl = None  # dispatch by length: falls to the warp-level kernel beyond 240 keys
    ops.attn_bwd(qkv, out, d_out, lse, delta, dqkv, N, T, heads, causal, impl=impl)
    err = (dqkv.float() - x.grad).abs().max().item()
    assert err < 3e-2 * max(1.0, x.grad.abs().max().item()), err


@pytest.mark.parametrize("N,T,heads", [(3, 129, 2), (2, 144, 3), (2, 256, 2), (2, 128, 4), (20, 100, 8), (40, 199, 4),
                                       (1, 200, 1), (2, 240, 2), (3, 33, 2)])
def test_attention_fused_backward_edge_lengths(N, T, heads):
    """Fused tcgen05 backward at the lengths where its operand-load schedule changes: one query / key tile (T <= 128) vs
    two, second tiles of 16 .. 128 rows (loaded as short TMA boxes; the rows behind them alias the next buffer), several
    units per CTA in both regimes (160 units on 148 SMs), and the 3-D TMA stores of dQ / dK / dV, which must clip at the
    end of every sequence (the output buffer is NaN-filled first, and every row of it is compared)."""
    D = heads * 64
    qkv = rnd(N * T, 3 * D, dtype=BF16, seed=T)
    out = torch.empty(N * T, D, device=DEV, dtype=BF16)
    lse = torch.empty(N, heads, T, device=DEV, dtype=F32)
    ops.attn_fwd(qkv, out, lse, N, T, heads, False, impl="tc")
    x = qkv.float().requires_grad_(True)
    o_ref, _, _ = _attn_ref(x, N, T, heads, False)
    assert (out.float() - o_ref).abs().max().item() < 2e-2
    d_out = rnd(N * T, D, dtype=BF16, seed=T + 1)
    o_ref.backward(d_out.float())
    delta = torch.empty(N * heads * T, device=DEV, dtype=F32)
    dqkv = torch.full_like(qkv, float("nan"))
    ops.attn_bwd(qkv, out, d_out, lse, delta, dqkv, N, T, heads, False, impl="fused")
    assert torch.isfinite(dqkv.float()).all()
    err = (dqkv.float() - x.grad).abs().max().item()
    assert err < 3e-2 * max(1.0, x.grad.abs().max().item()), err
    d2 = torch.full_like(qkv, float("nan"))
    ops.attn_bwd(qkv, out, d_out, lse, delta, d2, N, T, heads, False, impl="fused")
    assert torch.equal(dqkv, d2)  # deterministic


@pytest.mark.parametrize("N,T,heads,causal", [(4, 199, 12, False), (6, 10, 8, True), (3, 77, 8, True)])
def test_attention_single_query_rows(N, T, heads, causal):
    """Last-block attention for one consumed row per sequence == the dense attention restricted to that row, forward
    and backward (dQ only on the row, rank-1 dK / dV)."""
    D = heads * 64
    qkv = rnd(N * T, 3 * D, dtype=BF16, seed=12)
    pos = torch.tensor([(3 * i + 1) % T for i in range(N)]) if causal else torch.zeros(N, dtype=torch.long)
    rows = (torch.arange(N) * T + pos).to(DEV, torch.int32)
    out_r = torch.empty(N, D, device=DEV, dtype=BF16)
    lse_r = torch.empty(N * heads, device=DEV, dtype=F32)
    ops.attn_rows_fwd(qkv, rows, out_r, lse_r, N, T, heads, causal)
    x = qkv.float().requires_grad_(True)
    o_ref, _, _ = _attn_ref(x, N, T, heads, causal)
    o_rows = o_ref[rows.long()]
    assert (out_r.float() - o_rows).abs().max().item() < 2e-2
    d_out_r = rnd(N, D, dtype=BF16, seed=13)
    o_rows.backward(d_out_r.float())
    dqkv = torch.full_like(qkv, float("nan"))
    ops.attn_rows_bwd(qkv, rows, d_out_r, lse_r, dqkv, N, T, heads, causal)
    assert torch.isfinite(dqkv.float()).all()
    err = (dqkv.float() - x.grad).abs().max().item()
    assert err < 3e-2 * max(1.0, x.grad.abs().max().item()), err


@pytest.mark.parametrize("N,T,heads,causal", [(3, 199, 12, False), (5, 10, 8, True), (2, 77, 8, True)])
def test_attention_fp32_mode(N, T, heads, causal):
    D = heads * 64
    qkv = rnd(N * T, 3 * D, seed=11)
    out = torch.empty(N * T, D, device=DEV)
    ops.attn_fwd_f32(qkv, out, N, T, heads, causal)
    o_ref, _, _ = _attn_ref(qkv, N, T, heads, causal)
    assert (out - o_ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("M,D", [(6368, 768), (770, 512), (37, 768), (5, 128)])
def test_layernorm_fwd_bwd(M, D):
    x = rnd(M, D, std=2.0, seed=11) + 0.5
    g, b = 1 + rnd(D, std=0.1, seed=12), rnd(D, std=0.1, seed=13)
    y16 = torch.empty(M, D, device=DEV, dtype=BF16)
    y32 = torch.empty(M, D, device=DEV, dtype=F32)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    ops.layernorm_fwd(x, g, b, y_bf16=y16, y_f32=y32, mean=mean, rstd=rstd)
    xr = x.clone().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5)
    assert (y32 - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())
    assert (y16.float() - ref).abs().max().item() < 2e-2
    for dy_dtype in (F32, BF16):
        dy = rnd(M, D, seed=14).to(dy_dtype)
        gin = rnd(M, D, seed=15)
        xr.grad = gr.grad = br.grad = None
        ref.backward(dy.float(), retain_graph=True)
        gout = torch.empty(M, D, device=DEV)
        g16 = torch.empty(M, D, device=DEV, dtype=BF16)
        dgam, dbet = torch.empty(D, device=DEV), torch.empty(D, device=DEV)
        ws = torch.empty(2 * D * ops.ln_bwd_ctas(M), device=DEV)
        ops.layernorm_bwd(dy, x, mean, rstd, g, g_in=gin, g_out=gout, g_out_bf16=g16, dgamma=dgam, dbeta=dbet,
                          partial_ws=ws)
        assert (gout - (gin + xr.grad)).abs().max().item() < 1e-4
        assert (g16.float() - gout).abs().max().item() < 3e-2
        assert (dgam - gr.grad).abs().max().item() < 1e-3 * max(1.0, gr.grad.abs().max().item())
        assert (dbet - br.grad).abs().max().item() < 1e-3 * max(1.0, br.grad.abs().max().item())


def test_layernorm_gather_rows():
    x = rnd(100, 768, seed=16)
    idx = torch.tensor([0, 7, 99, 42], device=DEV, dtype=torch.int32)
    g, b = 1 + rnd(768, std=0.1, seed=17), rnd(768, std=0.1, seed=18)
    y = torch.empty(4, 768, device=DEV, dtype=F32)
    xs = torch.empty(4, 768, device=DEV, dtype=F32)
    ops.layernorm_fwd(x, g, b, rowidx=idx, y_f32=y, x_save=xs)
    assert torch.equal(xs, x[idx.long()])
    assert (y - torch.nn.functional.layer_norm(x[idx.long()], (768,), g, b)).abs().max().item() < 1e-5


def test_colsum_transpose_cast():
    x = rnd(777, 192, seed=19)
    ws = torch.empty(32 * 192, device=DEV)
    out = torch.empty(192, device=DEV)
    ops.colsum(x, out, ws)
    assert (out - x.sum(0)).abs().max().item() < 1e-3
    xb = x.to(BF16)
    ops.colsum(xb, out, ws, accumulate=True)
    assert (out - (x.sum(0) + xb.float().sum(0))).abs().max().item() < 2e-3
    t = torch.zeros(192, 784, device=DEV, dtype=BF16)
    cp = torch.empty(777, 192, device=DEV, dtype=BF16)
    ops.transpose_bf16(x, t, cp)
    assert torch.equal(t[:, :777], x.to(BF16).t())
    assert torch.equal(cp, x.to(BF16))
    t2 = torch.zeros(192, 784, device=DEV, dtype=BF16)
    ops.transpose_bf16(xb, t2)
    assert torch.equal(t2[:, :777], xb.t())
    c = torch.empty(777 * 192 - 1, device=DEV, dtype=BF16)
    ops.cast_bf16(x.reshape(-1)[:-1], c)
    assert torch.equal(c, x.reshape(-1)[:-1].to(BF16))


def test_im2col_matches_conv():
    img = rnd(3, 3, 224, 224, seed=20)
    w = rnd(768, 3, 16, 16, std=0.05, seed=21)
    col = torch.empty(3 * 196, 768, device=DEV, dtype=BF16)
    ops.patch_im2col(img, col)
    ref = torch.nn.functional.conv2d(img.to(BF16).float(), w.to(BF16).float(), stride=16)
    ref = ref.reshape(3, 768, 196).permute(0, 2, 1).reshape(3 * 196, 768)
    got = col.float() @ w.to(BF16).float().reshape(768, -1).t()
    assert (got - ref).abs().max().item() < 1e-3


def test_assemble_and_splice():
    B, T, n, D = 3, 199, 2, 768
    tok, cls, pos = rnd(B * 196, D, seed=22), rnd(D, seed=23), rnd(197, D, seed=24)
    sc = rnd(n, D, std=0.02, seed=25)
    g, b = 1 + rnd(D, std=0.1, seed=26), rnd(D, std=0.1, seed=27)
    x0 = torch.empty(B * T, D, device=DEV); x = torch.empty(B * T, D, device=DEV)
    mean, rstd = torch.empty(B * T, device=DEV), torch.empty(B * T, device=DEV)
    ops.vis_assemble_lnpre(tok, cls, pos, sc, g, b, x0, x, mean, rstd, B, T, n)
    ref0 = torch.cat([torch.cat([cls.expand(B, 1, D), tok.reshape(B, 196, D)], 1) + pos,
                      sc.half().float().expand(B, n, D)], 1).reshape(B * T, D)
    assert torch.equal(x0, ref0)
    assert (x - torch.nn.functional.layer_norm(ref0, (D,), g, b)).abs().max().item() < 2e-5
    # text assembly, truncated to Te positions
    C, Te, Dt = 4, 16, 512
    pre, ctx, suf, tpos = rnd(C, 1, Dt, seed=28), rnd(n, Dt, seed=29), rnd(C, 74, Dt, seed=30), rnd(77, Dt, seed=31)
    xt = torch.empty(C * Te, Dt, device=DEV)
    ops.text_assemble(pre, ctx, suf, tpos, xt, C, Te, n, 77)
    reft = (torch.cat([pre, ctx.expand(C, n, Dt), suf], 1) + tpos)[:, :Te].reshape(C * Te, Dt)
    assert torch.equal(xt, reft)
    # splice fwd / bwd
    p = rnd(n, D, std=0.02, seed=32)
    xs = x.clone()
    ops.prompt_splice_fwd(xs, p, B, T, T - n, n)
    ref = x.clone().reshape(B, T, D); ref[:, T - n:, :] = p.half().float()
    assert torch.equal(xs.reshape(B, T, D), ref)
    gg = rnd(B * T, D, std=1e-3, seed=33)
    g16 = gg.to(BF16)
    dp = torch.empty(n, D, device=DEV)
    want = gg.reshape(B, T, D)[:, 1:1 + n, :].half().float()
    want = want[0] + want[1] + want[2]
    ops.prompt_splice_bwd(gg, g16, dp, B, T, 1, n)
    assert torch.equal(dp, want)
    assert gg.reshape(B, T, D)[:, 1:1 + n].abs().max().item() == 0
    assert g16.reshape(B, T, D)[:, 1:1 + n].float().abs().max().item() == 0
    idx = torch.tensor([5, 1, 300], device=DEV, dtype=torch.int32)
    dx = rnd(3, D, seed=34)
    gz = torch.zeros(B * T, D, device=DEV); gz16 = torch.zeros(B * T, D, device=DEV, dtype=BF16)
    ops.scatter_rows(dx, idx, gz, gz16)
    assert torch.equal(gz[idx.long()], dx) and gz.abs().sum().item() == pytest.approx(dx.abs().sum().item(), rel=1e-5)


def test_small_linear():
    m, N, K = 2, 768, 512
    x, W, b = rnd(m, K, seed=35), rnd(N, K, std=0.05, seed=36), rnd(N, seed=37)
    y = torch.empty(m, N, device=DEV)
    ops.linear_small_fwd(x, W, b, y)
    assert (y - (x @ W.t() + b)).abs().max().item() < 1e-4
    dy = rnd(m, N, seed=38)
    dW, db, dx = torch.empty(N, K, device=DEV), torch.empty(N, device=DEV), torch.empty(m, K, device=DEV)
    add = rnd(m, K, seed=39)
    ops.linear_small_bwd(x, W, dy, dW=dW, db=db, dx_add=add, dx=dx)
    assert (dW - dy.t() @ x).abs().max().item() < 1e-4
    assert (db - dy.sum(0)).abs().max().item() < 1e-5
    assert (dx - (dy @ W + add)).abs().max().item() < 1e-3


def test_grouped_small_linear_and_repack_match_single_calls():
    probs, ref = [], []
    for i, (N, K) in enumerate([(768, 512), (512, 768), (768, 512)]):
        x, W, b = rnd(2, K, seed=20 + i), rnd(N, K, std=0.02, seed=30 + i), rnd(N, std=0.1, seed=40 + i)
        dy, add = rnd(2, N, seed=50 + i), rnd(2, K, seed=60 + i)
        y, dW, db, dx = (torch.empty(2, N, device=DEV), torch.empty(N, K, device=DEV), torch.empty(N, device=DEV),
                         torch.empty(2, K, device=DEV))
        probs.append(dict(x=x, W=W, b=b, y=y, dy=dy, dW=dW, db=db, dx_add=add, dx=dx))
        y1, dW1, db1, dx1 = torch.empty_like(y), torch.empty_like(dW), torch.empty_like(db), torch.empty_like(dx)
        ops.linear_small_fwd(x, W, b, y1)
        ops.linear_small_bwd(x, W, dy, dW=dW1, db=db1, dx_add=add, dx=dx1)
        ref.append((y1, dW1, db1, dx1))
    tab = ops.small_linear_table(probs, DEV)
    ops.linear_small_fwd_grouped(tab, 768)
    ops.linear_small_bwd_grouped(tab, 2, 768, 768)
    torch.cuda.synchronize()
    for pr, (y1, dW1, db1, dx1) in zip(probs, ref):   # same arithmetic and reduction order: bit-identical
        assert torch.equal(pr["y"], y1) and torch.equal(pr["dW"], dW1)
        assert torch.equal(pr["db"], db1) and torch.equal(pr["dx"], dx1)
    # grouped repack: bf16 copy + transposed copy of the fp32 masters with the fixed-threshold dithered rounding
    # (see repack_grouped_kernel): copy and transpose hold the same values, every value is one of the two bf16
    # neighbours of the master, bf16-representable masters pass unchanged, the rounding is unbiased on the fp16 grid
    # (the reference's weights) where round-to-nearest-even is not, and it is deterministic
    ws = [rnd(96, 160, seed=70), rnd(512, 2048, seed=71), rnd(37, 51, seed=72),
          (rnd(512, 2048, seed=73) * 0.03).half().float(), rnd(64, 128, seed=74).to(BF16).float()]
    trip = [(w, torch.empty(w.shape[1], w.shape[0], device=DEV, dtype=BF16),
             torch.empty(w.shape, device=DEV, dtype=BF16)) for w in ws]
    ops.repack_grouped(*ops.repack_table(trip, DEV))
    for w, t, c in trip:
        assert torch.equal(t, c.t().contiguous())
        lo = (w.view(torch.int32) & -65536).view(torch.float32)             # truncation toward zero
        hi = ((w.view(torch.int32) & -65536) + 65536).view(torch.float32)   # next bf16 away from zero
        cf = c.float()
        assert bool(((cf == lo) | (cf == hi)).all())
        exact = lo == w
        assert torch.equal(cf[exact], w[exact])
    w, _, c = trip[3]
    ulp = (w.abs().clamp_min(1e-30).log2().floor() - 7).exp2()
    bias_dither = ((c.float() - w) / ulp).mean().item()
    bias_rne = ((w.to(BF16).float() - w) / ulp).mean().item()
    assert abs(bias_dither) < 2e-3, (bias_dither, bias_rne)
    # a coherent update far below one ulp is carried, on average, at its true size (RNE from the fp16 grid gives ~1/16 ulp)
    w2 = w + 0.01 * ulp
    base = [(w, torch.empty(w.shape[1], w.shape[0], device=DEV, dtype=BF16), torch.empty(w.shape, device=DEV, dtype=BF16))]
    ops.repack_grouped(*ops.repack_table(base, DEV))      # same problem slot (= same thresholds) as the updated copy
    c = base[0][2]
    trip2 = [(w2, torch.empty(w2.shape[1], w2.shape[0], device=DEV, dtype=BF16), torch.empty(w2.shape, device=DEV, dtype=BF16))]
    ops.repack_grouped(*ops.repack_table(trip2, DEV))
    moved = ((trip2[0][2].float() - c.float()) / ulp).mean().item()
    moved_rne = ((w2.to(BF16).float() - w.to(BF16).float()) / ulp).mean().item()
    print(f"repack of fp16-grid weights after a +0.01 ulp update: dithered copy moves {moved:.4f} ulp on average, RNE {moved_rne:.4f}")
    assert abs(moved - 0.01) < 4e-3 and moved_rne > 0.03
    again = [(w2, torch.empty_like(trip2[0][1]), torch.empty_like(trip2[0][2]))]
    ops.repack_grouped(*ops.repack_table(again, DEV))
    assert torch.equal(again[0][2], trip2[0][2])


@pytest.mark.parametrize("B,C", [(4, 10), (32, 10), (64, 21), (32, 38), (32, 1000)])  # last one: multi-kernel path
def test_head(B, C):
    E = 512
    fi = rnd(B, E, seed=40).requires_grad_(True)
    ft = rnd(C, E, seed=41).requires_grad_(True)
    ls = torch.tensor([math.log(1 / 0.07)], device=DEV)
    lab = torch.randint(0, C, (B,), generator=torch.Generator().manual_seed(1)).to(DEV)
    a, t = torch.nn.functional.normalize(fi, dim=-1, eps=1e-8), torch.nn.functional.normalize(ft, dim=-1, eps=1e-8)
    logits_ref = ls.exp().clamp(max=100) * a @ t.t()
    loss_ref = torch.nn.functional.cross_entropy(logits_ref, lab) + 0.5 * (
        1 - torch.nn.functional.cosine_similarity(a, t[lab]).mean())
    loss_ref.backward()
    logits = torch.empty(B, C, device=DEV); loss = torch.empty(1, device=DEV)
    dfi, dft = torch.empty(B, E, device=DEV), torch.empty(C, E, device=DEV)
    ws = torch.empty(ops.head_workspace_floats(B, C, E), device=DEV)
    ops.head_forward_backward(fi.detach(), ft.detach(), ls, lab, logits, loss, dfi, dft, ws)
    assert (logits - logits_ref).abs().max().item() < 2e-5
    assert abs(loss.item() - loss_ref.item()) < 1e-5
    assert (dfi - fi.grad).abs().max().item() < 1e-6 + 1e-4 * fi.grad.abs().max().item()
    assert (dft - ft.grad).abs().max().item() < 1e-6 + 1e-4 * ft.grad.abs().max().item()
    lg2 = torch.empty(B, C, device=DEV)
    ops.head_forward_backward(fi.detach(), ft.detach(), ls, None, lg2, None, None, None, ws)
    assert torch.equal(lg2, logits)


@pytest.mark.parametrize("K", [1, 2, 3, 8, 16, 17, 32, 33])
def test_fedavg_bit_exact_vs_oracle(K):
    from oracle.maple_cpu import fedavg_oracle
    n = 100003
    g = torch.Generator().manual_seed(K)
    xs = [torch.randn(n, generator=g) for _ in range(K)]
    if K > 1:
        xs[1][5] = float("nan"); xs[1][6] = float("inf"); xs[K - 1][7] = float("-inf")
    dev = [x.to(DEV) for x in xs]
    ptrs = torch.tensor([d.data_ptr() for d in dev], dtype=torch.int64, device=DEV)
    out32 = torch.empty(n, device=DEV); out16 = torch.empty(n, device=DEV, dtype=torch.float16)
    flags = torch.zeros(K, device=DEV, dtype=torch.int32)
    ops.fedavg_reduce(ptrs, None, float(K), K, n, False, out32, out16, flags)
    m32, m16 = fedavg_oracle(xs)
    assert torch.equal(out32.cpu(), m32)
    assert torch.equal(out16.cpu(), m16)
    if K > 1:
        want = [0] * K; want[1] |= 3; want[K - 1] |= 2
        assert flags.cpu().tolist() == want
    # weighted mode
    w = [float(10 + 3 * k) for k in range(K)]
    wd = torch.tensor(w, device=DEV)
    ops.fedavg_reduce(ptrs, wd, float(sum(w)), K, n, False, out32, None, None)
    m32w, _ = fedavg_oracle(xs, w)
    assert torch.equal(out32.cpu(), m32w)
    # fp16 inputs (state_dict tensors of the reference are mostly fp16)
    h = [x.half() for x in xs]
    hd = [x.to(DEV) for x in h]
    ptrs = torch.tensor([d.data_ptr() for d in hd], dtype=torch.int64, device=DEV)
    ops.fedavg_reduce(ptrs, None, float(K), K, n, True, out32, out16, None)
    m32h, m16h = fedavg_oracle(h)
    assert torch.equal(out32.cpu(), m32h) and torch.equal(out16.cpu(), m16h)


def test_check_finite_and_sgd():
    x = rnd(10001, seed=50)
    flag = torch.zeros(1, device=DEV, dtype=torch.int32)
    ops.check_finite(x, flag); assert flag.item() == 0
    x[17] = float("inf"); ops.check_finite(x, flag); assert flag.item() == 2
    x[18] = float("nan"); flag.zero_(); ops.check_finite(x.half(), flag); assert flag.item() == 3
    # clip_grad_norm_ + SGD(momentum, wd) two steps vs torch
    n = 50000
    p0, g1, g2 = rnd(n, seed=51), rnd(n, std=0.1, seed=52), rnd(n, std=0.01, seed=53)
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([pt], lr=0.0026, momentum=0.9, weight_decay=5e-4)
    p, mom = p0.clone(), torch.zeros(n, device=DEV)
    ws, norm = torch.empty(296, device=DEV), torch.empty(1, device=DEV)
    for step, gsrc in enumerate((g1, g2)):
        pt.grad = gsrc.clone()
        tn = torch.nn.utils.clip_grad_norm_([pt], 1.0)
        opt.step()
        g = gsrc.clone()
        hp = torch.tensor([0.0026, 0.9, 0.0, 5e-4, 1.0, 0.0, 1.0 if step == 0 else 0.0], device=DEV)
        ops.grad_norm(g, ws, norm)
        assert abs(norm.item() - tn.item()) < 1e-4 * tn.item()
        ops.sgd_step(p, g, mom, hp, norm)
        assert (g - pt.grad).abs().max().item() < 1e-6
        assert (p - pt.detach()).abs().max().item() < 1e-6


def test_gpu_augmentation_matches_torchvision():
    """mfk_rrc_flip_normalize == torchvision resized_crop(bicubic, antialias) -> hflip -> /255 -> normalize on uint8
    tensors (the PIL-compatible pipeline Dassl's build_transform produces for the reference's yaml), to the uint8
    grid: at most 1 LSB (1/255/std) on a vanishing fraction of pixels, exact elsewhere."""
    import torchvision.transforms.functional as TF
    from federated_multi_modal_b200.trainers.client_datamanager import GpuAugment
    g = torch.Generator().manual_seed(3)
    B, H, W, S = 6, 256, 256, 224
    raw = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8)
    # smooth content as well as noise: low-pass half of the batch
    raw[: B // 2] = torch.nn.functional.avg_pool2d(raw[: B // 2].float(), 9, 1, 4).round().to(torch.uint8)
    aug = GpuAugment(size=S, seed=11)
    boxes, flip = aug.draw(B, H, W)
    flip[0], flip[1] = 1, 0
    out = aug.apply(raw.cuda(), boxes, flip).cpu()
    mean, std = torch.tensor(GpuAugment.MEAN)[:, None, None], torch.tensor(GpuAugment.STD)[:, None, None]
    worst, n_off = 0.0, 0
    for b in range(B):
        t, l, h, w = [int(v) for v in boxes[b]]
        r = TF.resized_crop(raw[b], t, l, h, w, [S, S], interpolation=TF.InterpolationMode.BICUBIC, antialias=True)
        if int(flip[b]):
            r = TF.hflip(r)
        want = (r.float() / 255.0 - mean) / std
        lsb = ((out[b] - want).abs() * std * 255.0)
        worst = max(worst, lsb.max().item())
        n_off += int((lsb > 0.5).sum())
    assert worst <= 1.01, worst
    assert n_off <= 2e-3 * B * 3 * S * S, n_off


def test_fedavg_unaligned_rows_and_guarded_sgd():
    """(a) client rows whose addresses are NOT 16-byte aligned (rows of a packed [K, n] buffer with
    n % 4 != 0, fp16 views at odd offsets) must reduce bit-exactly through the element-load path instead of faulting;
    (b) the fused clip+SGD step leaves parameters and momentum untouched when the loss or the input flag is bad, and
    a NaN gradient norm propagates (torch.clamp semantics), it is not replaced by 'no clipping'."""
    from oracle.maple_cpu import fedavg_oracle
    K, n = 5, 1001
    g = torch.Generator().manual_seed(3)
    packed = torch.randn(K * n + 3, generator=g)
    base = packed.to(DEV)
    rows = [base[1 + k * n: 1 + (k + 1) * n] for k in range(K)]        # 4-byte aligned only
    assert any(r.data_ptr() % 16 for r in rows)
    ptrs = torch.tensor([r.data_ptr() for r in rows], dtype=torch.int64, device=DEV)
    out = torch.empty(n + 1, device=DEV)[1:]                            # misaligned output too
    ops.fedavg_reduce(ptrs, None, float(K), K, n, False, out, None, None)
    want = fedavg_oracle([packed[1 + k * n: 1 + (k + 1) * n] for k in range(K)])[0]
    assert torch.equal(out.cpu(), want)
    h = packed.half().to(DEV)
    hrows = [h[1 + k * n: 1 + (k + 1) * n] for k in range(K)]
    ptrs = torch.tensor([r.data_ptr() for r in hrows], dtype=torch.int64, device=DEV)
    out16 = torch.empty(n + 1, device=DEV, dtype=torch.float16)[1:]
    ops.fedavg_reduce(ptrs, None, float(K), K, n, True, None, out16, None)
    assert torch.equal(out16.cpu(), fedavg_oracle([packed.half()[1 + k * n: 1 + (k + 1) * n] for k in range(K)])[1])
    # ---- guarded update
    m = 4096
    p0, g0 = rnd(m, seed=60), rnd(m, std=0.1, seed=61)
    hp = torch.tensor([0.01, 0.9, 0.0, 5e-4, 1.0, 0.0, 1.0], device=DEV)
    ws, norm = torch.empty(296, device=DEV), torch.empty(1, device=DEV)
    for loss_v, flag_v in ((float("nan"), 0), (float("inf"), 0), (1.0, 1), (1.0, 2)):
        p, gg, mom = p0.clone(), g0.clone(), torch.zeros(m, device=DEV)
        ops.grad_norm(gg, ws, norm)
        ops.sgd_step(p, gg, mom, hp, norm, m, torch.tensor([loss_v], device=DEV),
                     torch.tensor([flag_v], device=DEV, dtype=torch.int32))
        assert torch.equal(p, p0) and float(mom.abs().max()) == 0.0 and torch.equal(gg, g0), (loss_v, flag_v)
    p, gg, mom = p0.clone(), g0.clone(), torch.zeros(m, device=DEV)
    ops.grad_norm(gg, ws, norm)
    ops.sgd_step(p, gg, mom, hp, norm, m, torch.tensor([1.0], device=DEV), torch.zeros(1, device=DEV, dtype=torch.int32))
    assert not torch.equal(p, p0)
    p, gg, mom = p0.clone(), g0.clone(), torch.zeros(m, device=DEV)
    gg[7] = float("nan")
    ops.grad_norm(gg, ws, norm)
    assert torch.isnan(norm).item()
    ops.sgd_step(p, gg, mom, hp, norm, m, torch.tensor([1.0], device=DEV), torch.zeros(1, device=DEV, dtype=torch.int32))
    assert torch.isnan(p).all().item()     # clip coefficient NaN * every gradient, as torch.clamp(max_norm / (norm + eps)) does
