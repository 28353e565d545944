"""TEST INFRASTRUCTURE — CPU oracle for the MaPLe hot path. NOT a product path.

A plain fp32 torch-CPU restatement, stage by stage and with a hand-derived
backward pass, of the reference's ``CustomCLIP`` forward/backward step. Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it; the product package never does
(the CUDA path fails loudly when its extension is missing).

Parity pin: ``tests/test_oracle_golden.py`` checks this file against fixtures in
``tests/golden/`` produced by running the *unmodified reference* in the build
container (``tests/golden/make_golden.py`` -> ``oracle/ref_harness.py``): logits,
loss, features and all 145 parameter gradients of the reference's own autograd.

Reference lines restated (all under /root/reference):
  prompt learner ........ trainers/maple.py:177-218
  text encoder .......... trainers/maple.py:52-79
  vision tower .......... clip/model.py:509-572
  residual block ........ clip/model.py:307-352 (+ nn.MultiheadAttention, 303-305)
  LayerNorm / QuickGELU . clip/model.py:153-164
  logits + loss ......... trainers/maple.py:325-372
  trainable set ......... trainers/maple.py:447-479

The "fp32-ref" oracle definition is used: all tensors fp32, the
three hard-coded ``.half()`` casts of the prompt splices (clip/model.py:327,344,537)
kept as an fp16 round-trip. ``gemm_round`` optionally emulates the CUDA path's bf16
rounding of GEMM operands (used only to tighten CUDA-vs-oracle tolerances).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch

F32 = torch.float32


def q16(x: torch.Tensor) -> torch.Tensor:
    """fp16 round trip == the ``.half()`` in the prompt splices (value only)."""
    return x.to(torch.float16).to(F32)


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(F32)


def _ident(x):
    return x


# ----------------------------------------------------------------------------- LN
def ln_fwd(x, g, b, eps=1e-5):
    """clip/model.py:153-159 (fp32 LayerNorm, eps 1e-5, biased variance)."""
    mean = x.mean(-1, keepdim=True)
    xc = x - mean
    var = (xc * xc).mean(-1, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    xhat = xc * rstd
    return xhat * g + b, (xhat, rstd)


def ln_bwd(dy, cache, g):
    xhat, rstd = cache
    dg = (dy * xhat).reshape(-1, xhat.shape[-1]).sum(0)
    db = dy.reshape(-1, xhat.shape[-1]).sum(0)
    dxh = dy * g
    dx = (dxh - dxh.mean(-1, keepdim=True) - xhat * (dxh * xhat).mean(-1, keepdim=True)) * rstd
    return dx, dg, db


# -------------------------------------------------------------------------- block
class _BlockCache:
    __slots__ = ("ln1", "h", "q", "k", "v", "p", "a", "ln2", "h2", "u", "gact")


def block_fwd(x, W, heads, causal, rnd):
    """One ResidualAttentionBlock_MaPLe *after* the splice (clip/model.py:350-351).
    x: [N, T, D] (batch-major; the reference's LND permute is layout only)."""
    c = _BlockCache()
    N, T, D = x.shape
    dh = D // heads
    h, c.ln1 = ln_fwd(x, W["ln_1.weight"], W["ln_1.bias"])
    c.h = rnd(h)
    qkv = c.h @ rnd(W["attn.in_proj_weight"]).t() + W["attn.in_proj_bias"]
    qkv = rnd(qkv)
    q, k, v = qkv.split(D, dim=-1)
    sh = lambda t: t.reshape(N, T, heads, dh).permute(0, 2, 1, 3)
    c.q, c.k, c.v = sh(q), sh(k), sh(v)
    s = (c.q @ c.k.transpose(-1, -2)) * (dh ** -0.5)
    if causal:
        mask = torch.full((T, T), float("-inf")).triu_(1)  # clip/model.py:679-685
        s = s + mask
    c.p = torch.softmax(s, dim=-1)
    o = c.p @ c.v
    c.a = rnd(o.permute(0, 2, 1, 3).reshape(N, T, D))
    x = x + c.a @ rnd(W["attn.out_proj.weight"]).t() + W["attn.out_proj.bias"]
    h2, c.ln2 = ln_fwd(x, W["ln_2.weight"], W["ln_2.bias"])
    c.h2 = rnd(h2)
    c.u = c.h2 @ rnd(W["mlp.c_fc.weight"]).t() + W["mlp.c_fc.bias"]
    c.u = rnd(c.u)
    c.gact = rnd(c.u * torch.sigmoid(1.702 * c.u))  # QuickGELU, clip/model.py:162-164
    x = x + c.gact @ rnd(W["mlp.c_proj.weight"]).t() + W["mlp.c_proj.bias"]
    return x, c


def block_bwd(dx, c: _BlockCache, W, heads, rnd, want_wgrad: bool):
    """Returns dx_in and a dict of parameter grads (LN always; weights if want_wgrad)."""
    g: Dict[str, torch.Tensor] = {}
    N, T, D = dx.shape
    dh = D // heads
    dxr = rnd(dx)
    # --- MLP branch
    dg = dxr @ rnd(W["mlp.c_proj.weight"])
    if want_wgrad:
        g["mlp.c_proj.weight"] = dxr.reshape(-1, D).t() @ c.gact.reshape(-1, 4 * D)
        g["mlp.c_proj.bias"] = dx.reshape(-1, D).sum(0)
    sig = torch.sigmoid(1.702 * c.u)
    du = rnd(dg * (sig * (1.0 + 1.702 * c.u * (1.0 - sig))))
    dh2 = du @ rnd(W["mlp.c_fc.weight"])
    if want_wgrad:
        g["mlp.c_fc.weight"] = du.reshape(-1, 4 * D).t() @ c.h2.reshape(-1, D)
        g["mlp.c_fc.bias"] = du.reshape(-1, 4 * D).sum(0)
    dln, g["ln_2.weight"], g["ln_2.bias"] = ln_bwd(dh2, c.ln2, W["ln_2.weight"])
    dx = dx + dln
    # --- attention branch
    dxr = rnd(dx)
    da = rnd(dxr @ rnd(W["attn.out_proj.weight"]))
    if want_wgrad:
        g["attn.out_proj.weight"] = dxr.reshape(-1, D).t() @ c.a.reshape(-1, D)
        g["attn.out_proj.bias"] = dx.reshape(-1, D).sum(0)
    do = da.reshape(N, T, heads, dh).permute(0, 2, 1, 3)
    dv = c.p.transpose(-1, -2) @ do
    dp = do @ c.v.transpose(-1, -2)
    ds = c.p * (dp - (dp * c.p).sum(-1, keepdim=True))
    scale = dh ** -0.5
    dq = (ds @ c.k) * scale
    dk = (ds.transpose(-1, -2) @ c.q) * scale
    un = lambda t: t.permute(0, 2, 1, 3).reshape(N, T, D)
    dqkv = rnd(torch.cat([un(dq), un(dk), un(dv)], dim=-1))
    dh1 = dqkv @ rnd(W["attn.in_proj_weight"])
    if want_wgrad:
        g["attn.in_proj_weight"] = dqkv.reshape(-1, 3 * D).t() @ c.h.reshape(-1, D)
        g["attn.in_proj_bias"] = dqkv.reshape(-1, 3 * D).sum(0)
    dln, g["ln_1.weight"], g["ln_1.bias"] = ln_bwd(dh1, c.ln1, W["ln_1.weight"])
    dx = dx + dln
    return dx, g


# ------------------------------------------------------------------------- oracle
class MapleOracle:
    """fp32 CPU restatement of CustomCLIP fwd/bwd on a CustomCLIP-layout state_dict.

    state_dict keys are the reference's: ``prompt_learner.*``,
    ``image_encoder.*``, ``text_encoder.*``, ``logit_scale``. ``tokenized_prompts`` is
    the int64 [C,77] attribute (trainers/maple.py:149).
    """

    def __init__(self, state_dict: Dict[str, torch.Tensor], tokenized_prompts: torch.Tensor,
                 n_ctx: int = 2, depth: int = 9, v_heads: int = 12, t_heads: int = 8,
                 patch: int = 16, gemm_round: Optional[str] = None, trainable: str = "reference",
                 grad_q16: bool = True):
        self.P = {k: v.detach().to(F32).clone() for k, v in state_dict.items()
                  if not k.startswith("clip_model2.")}
        self.tok = tokenized_prompts.clone()
        self.eot = self.tok.argmax(-1)  # trainers/maple.py:76
        self.n, self.J = n_ctx, depth
        self.vh, self.th, self.patch = v_heads, t_heads, patch
        self.rnd: Callable = bf16_round if gemm_round == "bf16" else _ident
        self.trainable = trainable
        # autograd of the splice ``.half()`` (clip/model.py:327,344,537) also rounds the incoming
        # gradient to fp16, per element, BEFORE the expand-backward sum over the batch.
        self.gq: Callable = q16 if grad_q16 else _ident
        self.vL = len({k.split(".")[3] for k in self.P if k.startswith("image_encoder.transformer.resblocks.")})
        self.tL = len({k.split(".")[3] for k in self.P if k.startswith("text_encoder.transformer.resblocks.")})

    # -- helpers
    def _blk(self, tower: str, i: int):
        pre = f"{tower}.transformer.resblocks.{i}."
        return {k[len(pre):]: v for k, v in self.P.items() if k.startswith(pre)}

    def _wants_wgrad(self, layer: int, nlayers: int) -> bool:
        # trainers/maple.py:472-474: name contains "transformer.resblocks.11"
        return self.trainable == "reference" and layer == 11

    # -- prompt learner (trainers/maple.py:177-218)
    def prompt_learner(self):
        P, n = self.P, self.n
        ctx = P["prompt_learner.ctx"]
        C = P["prompt_learner.token_prefix"].shape[0]
        prompts = torch.cat([P["prompt_learner.token_prefix"], ctx.unsqueeze(0).expand(C, -1, -1),
                             P["prompt_learner.token_suffix"]], dim=1)
        deep_text: List[torch.Tensor] = []
        deep_vis: List[torch.Tensor] = []
        for i in range(self.J - 1):
            Wi = P[f"prompt_learner.compound_prompt_projections.{i}.weight"]
            bi = P[f"prompt_learner.compound_prompt_projections.{i}.bias"]
            if i % 2 == 0:
                t = P[f"prompt_learner.compound_prompts_text_parameters.{i // 2}"]
                deep_vis.append(t @ Wi.t() + bi)
                deep_text.append(t)
            else:
                v = P[f"prompt_learner.visual_deep_prompts_parameters.{(i - 1) // 2}"]
                deep_text.append(v @ Wi.t() + bi)
                deep_vis.append(v)
        shared = ctx @ P["prompt_learner.proj_lang_to_vis.weight"].t() + P["prompt_learner.proj_lang_to_vis.bias"]
        return prompts, shared, deep_text, deep_vis

    # -- towers
    def text_forward(self, prompts, deep_text, keep=True):
        P, n = self.P, self.n
        x = prompts + P["text_encoder.positional_embedding"]
        caches, acts = [], []
        for l in range(self.tL):
            if l >= 1 and (l - 1) < len(deep_text):  # clip/model.py:340-347
                x = x.clone()
                x[:, 1:1 + n, :] = q16(deep_text[l - 1])
            x, c = block_fwd(x, self._blk("text_encoder", l), self.th, True, self.rnd)
            caches.append(c if keep else None)
            acts.append(x)
        C = x.shape[0]
        xe = x[torch.arange(C), self.eot]  # LN is row-wise: gather first == LN-all-then-gather
        y, lnc = ln_fwd(xe, P["text_encoder.ln_final.weight"], P["text_encoder.ln_final.bias"])
        y = self.rnd(y)
        feat = y @ self.rnd(P["text_encoder.text_projection"])
        return feat, dict(blocks=caches, lnf=lnc, y=y, acts=acts)

    def im2col(self, img):
        B, ps = img.shape[0], self.patch
        g = img.shape[-1] // ps
        # [B,3,g,ps,g,ps] -> [B,g,g,3,ps,ps] -> [B,g*g,3*ps*ps]; K index = c*256+ky*16+kx
        return img.reshape(B, 3, g, ps, g, ps).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, 3 * ps * ps)

    def vision_forward(self, img, shared, deep_vis, keep=True):
        P, n = self.P, self.n
        B = img.shape[0]
        Wc = P["image_encoder.conv1.weight"].reshape(P["image_encoder.conv1.weight"].shape[0], -1)
        tok = self.rnd(self.im2col(img.to(F32))) @ self.rnd(Wc).t()  # conv 16x16/s16 no bias == GEMM
        D = tok.shape[-1]
        cls = P["image_encoder.class_embedding"].reshape(1, 1, D).expand(B, 1, D)
        x = torch.cat([cls, tok], dim=1) + P["image_encoder.positional_embedding"]
        x = torch.cat([x, q16(shared).unsqueeze(0).expand(B, -1, -1)], dim=1)  # clip/model.py:536-538
        x, lnpre = ln_fwd(x, P["image_encoder.ln_pre.weight"], P["image_encoder.ln_pre.bias"])
        T = x.shape[1]
        caches, acts = [], []
        for l in range(self.vL):
            if l >= 1 and (l - 1) < len(deep_vis):  # clip/model.py:324-330
                x = x.clone()
                x[:, T - n:, :] = q16(deep_vis[l - 1])
            x, c = block_fwd(x, self._blk("image_encoder", l), self.vh, False, self.rnd)
            caches.append(c if keep else None)
            acts.append(x)
        y, lnpost = ln_fwd(x[:, 0, :], P["image_encoder.ln_post.weight"], P["image_encoder.ln_post.bias"])
        y = self.rnd(y)
        feat = y @ self.rnd(P["image_encoder.proj"])
        return feat, dict(blocks=caches, lnpre=lnpre, lnpost=lnpost, y=y, acts=acts, T=T)

    # -- head (trainers/maple.py:325-372)
    @staticmethod
    def head(fi, ft, logit_scale, label=None):
        s = min(math.exp(float(logit_scale)), 100.0)
        ni = fi.norm(dim=-1, keepdim=True).clamp_min(1e-8)
        nt = ft.norm(dim=-1, keepdim=True).clamp_min(1e-8)
        a, t = fi / ni, ft / nt
        logits = s * (a @ t.t())
        out = dict(logits=logits, a=a, t=t, ni=ni, nt=nt, s=s)
        if label is None:
            return out
        B = fi.shape[0]
        lse = torch.logsumexp(logits, dim=1)
        ce = (lse - logits[torch.arange(B), label]).mean()
        ty = t[label]
        na2 = a.norm(dim=-1).clamp_min(1e-8)
        nt2 = ty.norm(dim=-1).clamp_min(1e-8)
        cos = (a * ty).sum(-1) / (na2 * nt2)
        loss = ce + 0.5 * (1.0 - cos.mean())
        out.update(loss=loss, ce=ce, cos=cos, na2=na2, nt2=nt2, label=label)
        return out

    @staticmethod
    def head_bwd(h):
        """d loss / d image_features, d text_features (pre-normalisation)."""
        a, t, s, label = h["a"], h["t"], h["s"], h["label"]
        B = a.shape[0]
        dlog = torch.softmax(h["logits"], dim=1)
        dlog[torch.arange(B), label] -= 1.0
        dlog /= B
        da = s * (dlog @ t)
        dt = s * (dlog.t() @ a)
        # alignment: loss += 0.5*(1 - mean cos)
        ty, cos = t[label], h["cos"]
        na2, nt2 = h["na2"].unsqueeze(-1), h["nt2"].unsqueeze(-1)
        dcos = -0.5 / B
        da = da + dcos * (ty / (na2 * nt2) - cos.unsqueeze(-1) * a / (na2 * na2))
        dty = dcos * (a / (na2 * nt2) - cos.unsqueeze(-1) * ty / (nt2 * nt2))
        dt = dt.index_add(0, label, dty)
        # F.normalize backward
        dfi = (da - a * (a * da).sum(-1, keepdim=True)) / h["ni"]
        dft = (dt - t * (t * dt).sum(-1, keepdim=True)) / h["nt"]
        return dfi, dft

    # -- public API
    @torch.no_grad()
    def logits(self, img):
        prompts, shared, dt, dv = self.prompt_learner()
        ft, _ = self.text_forward(prompts, dt, keep=False)
        fi, _ = self.vision_forward(img, shared, dv, keep=False)
        return self.head(fi, ft, self.P["logit_scale"])["logits"]

    @torch.no_grad()
    def forward_backward(self, img, label):
        """Returns dict(loss, logits, image_features, text_features, grads{name: tensor}, acts)."""
        P, n, rnd = self.P, self.n, self.rnd
        prompts, shared, deep_text, deep_vis = self.prompt_learner()
        ft, tc = self.text_forward(prompts, deep_text)
        fi, vc = self.vision_forward(img, shared, deep_vis)
        h = self.head(fi, ft, P["logit_scale"], label)
        dfi, dft = self.head_bwd(h)
        G: Dict[str, torch.Tensor] = {}
        B, C = fi.shape[0], ft.shape[0]
        nd = self.J - 1
        d_deep_text = [None] * nd
        d_deep_vis = [None] * nd

        # ---------------- vision backward
        dy = rnd(dfi) @ rnd(P["image_encoder.proj"]).t()
        dcls, G["image_encoder.ln_post.weight"], G["image_encoder.ln_post.bias"] = \
            ln_bwd(dy, vc["lnpost"], P["image_encoder.ln_post.weight"])
        T = vc["T"]
        dx = torch.zeros(B, T, dcls.shape[-1])
        dx[:, 0, :] = dcls
        for l in reversed(range(self.vL)):
            W = self._blk("image_encoder", l)
            dx, g = block_bwd(dx, vc["blocks"][l], W, self.vh, rnd, self._wants_wgrad(l, self.vL))
            for k, v in g.items():
                G[f"image_encoder.transformer.resblocks.{l}.{k}"] = v
            if l >= 1 and (l - 1) < nd:
                d_deep_vis[l - 1] = self.gq(dx[:, T - n:, :]).sum(0)
                dx[:, T - n:, :] = 0
        dxpre, G["image_encoder.ln_pre.weight"], G["image_encoder.ln_pre.bias"] = \
            ln_bwd(dx, vc["lnpre"], P["image_encoder.ln_pre.weight"])
        d_shared = self.gq(dxpre[:, T - n:, :]).sum(0)

        # ---------------- text backward
        dy = rnd(dft) @ rnd(P["text_encoder.text_projection"]).t()
        deot, G["text_encoder.ln_final.weight"], G["text_encoder.ln_final.bias"] = \
            ln_bwd(dy, tc["lnf"], P["text_encoder.ln_final.weight"])
        dx = torch.zeros(C, self.tok.shape[1], deot.shape[-1])
        dx[torch.arange(C), self.eot] = deot
        for l in reversed(range(self.tL)):
            W = self._blk("text_encoder", l)
            dx, g = block_bwd(dx, tc["blocks"][l], W, self.th, rnd, self._wants_wgrad(l, self.tL))
            for k, v in g.items():
                G[f"text_encoder.transformer.resblocks.{l}.{k}"] = v
            if l >= 1 and (l - 1) < nd:
                d_deep_text[l - 1] = self.gq(dx[:, 1:1 + n, :]).sum(0)
                dx[:, 1:1 + n, :] = 0
        d_ctx = dx[:, 1:1 + n, :].sum(0)

        # deep prompts beyond the tower depth are never spliced: zero gradient
        d_deep_vis = [torch.zeros_like(deep_vis[i]) if d is None else d for i, d in enumerate(d_deep_vis)]
        d_deep_text = [torch.zeros_like(deep_text[i]) if d is None else d for i, d in enumerate(d_deep_text)]
        # ---------------- prompt learner backward
        pl = "prompt_learner."
        for i in range(nd):
            Wi = P[f"{pl}compound_prompt_projections.{i}.weight"]
            if i % 2 == 0:
                t = P[f"{pl}compound_prompts_text_parameters.{i // 2}"]
                dyv = d_deep_vis[i]
                G[f"{pl}compound_prompts_text_parameters.{i // 2}"] = d_deep_text[i] + dyv @ Wi
                G[f"{pl}compound_prompt_projections.{i}.weight"] = dyv.t() @ t
                G[f"{pl}compound_prompt_projections.{i}.bias"] = dyv.sum(0)
            else:
                v = P[f"{pl}visual_deep_prompts_parameters.{(i - 1) // 2}"]
                dyt = d_deep_text[i]
                G[f"{pl}visual_deep_prompts_parameters.{(i - 1) // 2}"] = d_deep_vis[i] + dyt @ Wi
                G[f"{pl}compound_prompt_projections.{i}.weight"] = dyt.t() @ v
                G[f"{pl}compound_prompt_projections.{i}.bias"] = dyt.sum(0)
        ctx = P[f"{pl}ctx"]
        G[f"{pl}ctx"] = d_ctx + d_shared @ P[f"{pl}proj_lang_to_vis.weight"]
        G[f"{pl}proj_lang_to_vis.weight"] = d_shared.t() @ ctx
        G[f"{pl}proj_lang_to_vis.bias"] = d_shared.sum(0)
        if self.trainable != "reference":
            G = {k: v for k, v in G.items() if k.startswith(pl)}
        return dict(loss=h["loss"], logits=h["logits"], image_features=fi, text_features=ft, grads=G,
                    vis_acts=vc["acts"], txt_acts=tc["acts"], dfi=dfi, dft=dft)


# ------------------------------------------------------------------------- FedAvg
def fedavg_oracle(tensors: List[torch.Tensor], weights: Optional[List[float]] = None):
    """trainers/maple_fed.py:309-315 restated with an explicit summation order.

    uniform (weights=None): fp32 cast -> nan_to_num(nan=0, posinf=1e4, neginf=-1e4) ->
    rows added sequentially in client order inside chunks of 16, chunk sums added
    sequentially, remainder last -> TRUE division by K (torch CPU mean == sum / K).
    This is the order torch's CPU ``sum(dim=0)`` (cascade sum, SumKernel.cpp) uses for the
    vectorised body of a [K, n] stack, i.e. every column below the last multiple of
    4 x SIMD-width (64 fp32 on AVX-512, 32 on AVX2) [probed in this container, torch 2.11].
    Tail columns (n mod 64) and 0-d tensors go through other torch code paths whose order
    depends on the host ISA; there the reference itself is not reproducible across hosts,
    so this oracle's order is normative and tests compare those columns after the
    reference's own ``.half()`` rounding. Every trainable MaPLe tensor has n % 64 == 0.
    weighted (north_star extension, absent from the reference): acc = sum_k float(n_k)*w_k
    in the same order, / float(sum n_k).
    Returns (fp32 mean, fp16-rounded mean == the reference's ``.half()`` output).
    """
    K = len(tensors)
    xs = [torch.nan_to_num(t.detach().to(F32), nan=0.0, posinf=1e4, neginf=-1e4) for t in tensors]
    if weights is not None:
        xs = [x * torch.tensor(float(w), dtype=F32) for x, w in zip(xs, weights)]
    def seq(rows):
        acc = rows[0].clone()
        for r in rows[1:]:
            acc = acc + r
        return acc
    full = K // 16
    if full == 0:
        total = seq(xs)
    else:
        total = seq([seq(xs[i * 16:(i + 1) * 16]) for i in range(full)])
        rem = xs[full * 16:]
        if rem:
            total = total + seq(rem)
    den = float(K) if weights is None else float(sum(float(w) for w in weights))
    mean = total / torch.tensor(den, dtype=F32)
    return mean, mean.to(torch.float16)
