"""Drop-in mirror of the reference's ``clip/model.py`` MaPLe branch on libmfk kernels.

Same class names, constructor signatures, parameter names/shapes/dtypes and forward
signatures as the reference (so ``state_dict`` round-trips in both directions):
  LayerNorm, QuickGELU ............ clip/model.py:153-164
  ResidualAttentionBlock_MaPLe .... clip/model.py:269-352   forward([x(L,N,D), deep, counter])
  Transformer (MaPLe branch) ...... clip/model.py:355-380
  VisionTransformer_MaPLe ......... clip/model.py:478-572   forward(x, shared_ctx, deep, clip_embeddings=None)
  CLIP, convert_weights, build_model  clip/model.py:575-793

The module-level ``forward``s here are the *standalone hooks*: they run the same CUDA kernels as
the fused engine (``engine.MapleEngine``, which is what ``CustomCLIP`` uses for the hot path), in
inference mode, converting from the reference's sequence-first (L,N,D) layout to the kernels'
token-major [N*T, D] layout. They need a CUDA device (no CPU fallback) and do not record autograd
history — gradients of the hot path come from the engine's explicit backward.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Tuple, Union

import numpy as np
import torch
from torch import nn

from .. import ops

BF16, F32 = torch.bfloat16, torch.float32


def _need_cuda(t: torch.Tensor, who: str):
    if not t.is_cuda:
        raise RuntimeError(f"{who}: libmfk kernels need CUDA tensors (there is no CPU fallback)")


class _Packed:
    """bf16 K-major copy of a Linear weight + fp32 bias, refreshed when the parameter changes."""

    def __init__(self):
        self.ver = None
        self.w = self.b = None

    def get(self, weight: torch.Tensor, bias):
        ver = (weight._version, weight.data_ptr(), None if bias is None else bias._version)
        if ver != self.ver:
            self.w = weight.detach().to(BF16).contiguous()
            self.b = None if bias is None else bias.detach().to(F32).contiguous()
            self.ver = ver
        return self.w, self.b


class LayerNorm(nn.LayerNorm):
    """fp32 LayerNorm on any input dtype (clip/model.py:153-159)."""

    def forward(self, x: torch.Tensor):
        _need_cuda(x, "LayerNorm")
        D = x.shape[-1]
        xf = x.detach().to(F32).reshape(-1, D).contiguous()
        y = torch.empty_like(xf)
        ops.layernorm_fwd(xf, self.weight.detach().to(F32), self.bias.detach().to(F32), y_f32=y, eps=self.eps)
        return y.reshape(x.shape).to(x.dtype)


class QuickGELU(nn.Module):
    def forward(self, x: torch.Tensor):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock_MaPLe(nn.Module):
    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None, design_details=None,
                 text_layer=False, i=0):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.text_layer = text_layer
        self.attn_mask = attn_mask
        self.compound_prompt_nctx = design_details["maple_length"]
        self.first_layer = i == 0
        self.n_head = n_head
        self._pk = [_Packed() for _ in range(4)]
        for name, p in self.named_parameters():  # reference: only the norms start trainable (291-297)
            p.requires_grad = ("ln_1" in name) or ("ln_2" in name)

    def _run(self, x: torch.Tensor) -> torch.Tensor:
        """x: [N, T, D] fp32 contiguous -> x + attn(ln_1 x) -> + mlp(ln_2 .) on the CUDA kernels."""
        N, T, D = x.shape
        M = N * T
        dev = x.device
        w_in, b_in = self._pk[0].get(self.attn.in_proj_weight, self.attn.in_proj_bias)
        w_out, b_out = self._pk[1].get(self.attn.out_proj.weight, self.attn.out_proj.bias)
        w_fc, b_fc = self._pk[2].get(self.mlp.c_fc.weight, self.mlp.c_fc.bias)
        w_pj, b_pj = self._pk[3].get(self.mlp.c_proj.weight, self.mlp.c_proj.bias)
        x = x.reshape(M, D)
        h = torch.empty(M, D, device=dev, dtype=BF16)
        ops.layernorm_fwd(x, self.ln_1.weight.detach().float(), self.ln_1.bias.detach().float(), y_bf16=h)
        qkv = torch.empty(M, 3 * D, device=dev, dtype=BF16)
        ops.gemm(h, w_in, bias=b_in, out_bf16=qkv)
        att = torch.empty(M, D, device=dev, dtype=BF16)
        ops.attn_fwd(qkv, att, None, N, T, self.n_head, self.attn_mask is not None)
        x2 = torch.empty(M, D, device=dev, dtype=F32)
        ops.gemm(att, w_out, bias=b_out, residual=x, out_f32=x2)
        ops.layernorm_fwd(x2, self.ln_2.weight.detach().float(), self.ln_2.bias.detach().float(), y_bf16=h)
        act = torch.empty(M, 4 * D, device=dev, dtype=BF16)
        ops.gemm(h, w_fc, bias=b_fc, act=1, out_bf16=act)
        out = torch.empty(M, D, device=dev, dtype=F32)
        ops.gemm(act, w_pj, bias=b_pj, residual=x2, out_f32=out)
        return out.reshape(N, T, D)

    @torch.no_grad()
    def forward(self, inputs):
        x, deep, counter = inputs[0], inputs[1], inputs[2]
        _need_cuda(x, "ResidualAttentionBlock_MaPLe")
        dtype = x.dtype
        xb = x.detach().permute(1, 0, 2).to(F32).contiguous()  # (L,N,D) -> [N,T,D]
        N, T, D = xb.shape
        n = self.compound_prompt_nctx
        if not self.first_layer and len(deep) > 0 and not (counter > len(deep) - 1):
            prompt = deep[counter].detach().to(F32).contiguous()
            row0 = 1 if self.text_layer else T - n
            ops.prompt_splice_fwd(xb.reshape(N * T, D), prompt, N, T, row0, n)
            counter += 1
        out = self._run(xb)
        return [out.permute(1, 0, 2).to(dtype), deep, counter]


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None, prompts_needed=0,
                 text_layer=False, design_details=None):
        super().__init__()
        self.width, self.layers = width, layers
        if design_details["trainer"] != "MaPLe":
            raise NotImplementedError("only the MaPLe branch of clip/model.py:370-373 is on the hot path")
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock_MaPLe(width, heads, attn_mask, design_details,
                                                                      text_layer, i) for i in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


class VisionTransformer_MaPLe(nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int,
                 design_details):
        super().__init__()
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        self.VPT_shallow = True
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.prompt_till_layer_visual = 0
        self.transformer = Transformer(width, layers, heads, design_details=design_details)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self._pk_conv, self._pk_proj = _Packed(), _Packed()
        for name, p in self.named_parameters():  # clip/model.py:501-507
            p.requires_grad = ("ln_pre" in name) or ("ln_post" in name)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, shared_ctx, compound_deeper_prompts, clip_embeddings=None):
        if clip_embeddings is not None:
            raise NotImplementedError("caption branch (clip/model.py:550-561) is out of scope: it draws fresh "
                                      "random weights on every call, so no parity target exists (SURVEY.md §2 #11)")
        _need_cuda(x, "VisionTransformer_MaPLe")
        dtype = x.dtype
        B, dev = x.shape[0], x.device
        D = self.conv1.weight.shape[0]
        P = (self.input_resolution // self.conv1.kernel_size[0]) ** 2
        n = shared_ctx.shape[0]
        T = P + 1 + n
        wc, _ = self._pk_conv.get(self.conv1.weight.reshape(D, -1), None)
        col = torch.empty(B * P, wc.shape[1], device=dev, dtype=BF16)
        ops.patch_im2col(x.detach().to(F32).contiguous(), col)
        tok = torch.empty(B * P, D, device=dev, dtype=F32)
        ops.gemm(col, wc, out_f32=tok)
        xt = torch.empty(B * T, D, device=dev, dtype=F32)
        f = lambda t: t.detach().to(F32).contiguous()
        ops.vis_assemble_lnpre(tok, f(self.class_embedding), f(self.positional_embedding), f(shared_ctx),
                               f(self.ln_pre.weight), f(self.ln_pre.bias), None, xt, None, None, B, T, n)
        seq = xt.reshape(B, T, D).permute(1, 0, 2)  # NLD -> LND, as the reference passes it on
        out = self.transformer([seq, compound_deeper_prompts, 0])[0]
        xo = out.permute(1, 0, 2).to(F32).contiguous().reshape(B * T, D)
        rows = (torch.arange(B, device=dev, dtype=torch.int32) * T).contiguous()
        y = torch.empty(B, D, device=dev, dtype=BF16)
        ops.layernorm_fwd(xo, f(self.ln_post.weight), f(self.ln_post.bias), rowidx=rows, y_bf16=y, M=B)
        if self.proj is None:
            return y.to(dtype)
        pT, _ = self._pk_proj.get(self.proj.detach().t(), None)
        feat = torch.empty(B, self.proj.shape[1], device=dev, dtype=F32)
        ops.gemm(y, pT, out_f32=feat)
        return feat.to(dtype)


class CLIP(nn.Module):
    def __init__(self, embed_dim: int, image_resolution: int, vision_layers: Union[Tuple[int, int, int, int], int],
                 vision_width: int, vision_patch_size: int, context_length: int, vocab_size: int,
                 transformer_width: int, transformer_heads: int, transformer_layers: int, design_details):
        super().__init__()
        self.context_length = context_length
        if isinstance(vision_layers, (tuple, list)) or design_details["trainer"] != "MaPLe":
            raise NotImplementedError("only the ViT + MaPLe configuration is on the hot path")
        self.visual = VisionTransformer_MaPLe(image_resolution, vision_patch_size, vision_width, vision_layers,
                                              vision_width // 64, embed_dim, design_details)
        self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads,
                                       attn_mask=self.build_attention_mask(),
                                       prompts_needed=design_details["language_depth"], text_layer=True,
                                       design_details=design_details)
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.initialize_parameters()

    def initialize_parameters(self):
        # same distributions as clip/model.py:650-677 (values are overwritten by load_state_dict)
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        w, L = self.transformer.width, self.transformer.layers
        for blk in self.transformer.resblocks:
            nn.init.normal_(blk.attn.in_proj_weight, std=w ** -0.5)
            nn.init.normal_(blk.attn.out_proj.weight, std=(w ** -0.5) * ((2 * L) ** -0.5))
            nn.init.normal_(blk.mlp.c_fc.weight, std=(2 * w) ** -0.5)
            nn.init.normal_(blk.mlp.c_proj.weight, std=(w ** -0.5) * ((2 * L) ** -0.5))
        nn.init.normal_(self.text_projection, std=w ** -0.5)

    def build_attention_mask(self):
        # additive causal mask (clip/model.py:679-685); the kernels implement it as a causal flag
        return torch.full((self.context_length, self.context_length), float("-inf")).triu_(1)

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype


def convert_weights(model: nn.Module):
    """fp16 storage for Conv/Linear/MHA weights, text_projection and proj (clip/model.py:726-747);
    LayerNorm parameters and the embeddings stay fp32."""
    def _to_half(m):
        if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Linear)):
            m.weight.data = m.weight.data.half()
            if m.bias is not None:
                m.bias.data = m.bias.data.half()
        if isinstance(m, nn.MultiheadAttention):
            for attr in ("in_proj_weight", "q_proj_weight", "k_proj_weight", "v_proj_weight", "in_proj_bias",
                         "bias_k", "bias_v"):
                t = getattr(m, attr)
                if t is not None:
                    t.data = t.data.half()
        for name in ("text_projection", "proj"):
            t = getattr(m, name, None)
            if isinstance(t, torch.Tensor):
                t.data = t.data.half()
    model.apply(_to_half)


def build_model(state_dict: dict, design_details):
    """CLIP from a checkpoint state_dict (clip/model.py:750-793), ViT only."""
    if "visual.proj" not in state_dict:
        raise NotImplementedError("ResNet CLIP backbones are outside the MaPLe ViT-B/16 hot path")
    vw = state_dict["visual.conv1.weight"].shape[0]
    vl = len([k for k in state_dict if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    ps = state_dict["visual.conv1.weight"].shape[-1]
    grid = round((state_dict["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    tw = state_dict["ln_final.weight"].shape[0]
    tl = len({k.split(".")[2] for k in state_dict if k.startswith("transformer.resblocks")})
    model = CLIP(state_dict["text_projection"].shape[1], ps * grid, vl, vw, ps,
                 state_dict["positional_embedding"].shape[0], state_dict["token_embedding.weight"].shape[0], tw,
                 tw // 64, tl, design_details)
    sd = {k: v for k, v in state_dict.items() if k not in ("input_resolution", "context_length", "vocab_size")}
    convert_weights(model)
    model.load_state_dict(sd)
    return model.eval()
