"""MapleEngine — explicit forward/backward schedule of the MaPLe CustomCLIP step on libmfk kernels.

This is the host side of the hot path: it owns the packed frozen CLIP
weights (bf16, plus K-major transposes for dgrad), one flat fp32 arena with every trainable tensor
(+ gradient and momentum arenas of the same layout), the saved activations, and issues the kernel
sequence through the C ABI. It mirrors ``oracle/maple_cpu.py`` stage by stage; there is no autograd
graph and no host synchronisation inside a step, so a step can be captured in a CUDA graph.

Reference call sites replaced: CustomCLIP.forward (trainers/maple.py:304-381),
MultiModalPromptLearner.forward (177-218), TextEncoder.forward (52-79),
VisionTransformer_MaPLe.forward (clip/model.py:509-572), ResidualAttentionBlock_MaPLe.forward
(clip/model.py:307-352) and loss.backward() (trainers/maple.py:590) for the reference's trainable
set (trainers/maple.py:447-479) or the prompt-only set.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

from . import ops

BF16, F32 = torch.bfloat16, torch.float32
_LIN = ("attn.in_proj", "attn.out_proj", "mlp.c_fc", "mlp.c_proj")


def _lin_keys(name):
    # nn.MultiheadAttention stores in_proj_weight / in_proj_bias; Linear stores .weight / .bias
    if name == "attn.in_proj":
        return "attn.in_proj_weight", "attn.in_proj_bias"
    return name + ".weight", name + ".bias"


class _Tower:
    """Static description + packed weights + workspaces of one transformer tower."""

    def __init__(self, name: str, D: int, heads: int, L: int, causal: bool):
        self.name, self.D, self.heads, self.L, self.causal = name, D, heads, L, causal
        self.w: List[Dict[str, torch.Tensor]] = []  # per layer packed tensors
        self.N = self.T = self.M = 0
        self.ws: Dict[str, torch.Tensor] = {}
        self.gemm_ws: Optional[torch.Tensor] = None  # split-K workspace (vision tower only: one user stream)
        self.ln_slot, self.ln_pending, self.ln_table, self.ln_key = 0, [], None, None  # deferred LN dgamma/dbeta


class MapleEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], tokenized_prompts: torch.Tensor, *, n_ctx: int = 2,
                 depth: int = 9, device: str = "cuda", trainable: str = "reference", text_truncate: bool = True,
                 patch: int = 16, share_from: Optional["MapleEngine"] = None, share_workspace: bool = True):
        """``share_from``: another engine on the same GPU (a co-located federated client). The frozen packed
        CLIP weights and the activation workspaces are shared with it; only the trainable arena, its
        gradient/momentum arenas and the bf16 copies of trainable block weights are per client."""
        if not torch.cuda.is_available():
            raise RuntimeError("MapleEngine needs a CUDA device: the libmfk kernels have no CPU fallback")
        assert trainable in ("reference", "prompt_only")
        self.dev = torch.device(device)
        self.n, self.J, self.patch = n_ctx, depth, patch
        self.trainable = trainable
        sd = {k: v for k, v in state_dict.items() if not k.startswith("clip_model2.")}
        self._sd_src = sd if share_from is None else share_from._sd_src  # fp32 mode re-reads frozen weights exactly
        self.tok = tokenized_prompts.clone().cpu()
        self.eot = self.tok.argmax(-1)
        self.C = self.tok.shape[0]
        self.Tfull = self.tok.shape[1]
        self.Te = int(self.eot.max().item()) + 1 if text_truncate else self.Tfull
        vD = sd["image_encoder.ln_pre.weight"].shape[0]
        tD = sd["text_encoder.ln_final.weight"].shape[0]
        vL = len({k.split(".")[3] for k in sd if k.startswith("image_encoder.transformer.resblocks.")})
        tL = len({k.split(".")[3] for k in sd if k.startswith("text_encoder.transformer.resblocks.")})
        self.vis = _Tower("image_encoder", vD, vD // 64, vL, False)
        self.txt = _Tower("text_encoder", tD, tD // 64, tL, True)
        self.E = sd["image_encoder.proj"].shape[1]
        self.P = (sd["image_encoder.positional_embedding"].shape[0] - 1)  # patches per image
        self.Tv = self.P + 1 + n_ctx
        self._build_arena(sd)
        self._pack_frozen(sd, share_from)
        self._bufs: Dict[str, torch.Tensor] = share_from._bufs if (share_from is not None and share_workspace) else {}
        # split-K tail workspace of the vision-tower GEMMs (main stream only; the text tower runs concurrently on a
        # side stream and a workspace must never be shared by two streams)
        if "__gemm_ws__" not in self._bufs:
            self._bufs["__gemm_ws__"] = ops.splitk_workspace(self.dev)
        self.vis.gemm_ws = self._bufs["__gemm_ws__"]
        self._Bmax = 0
        self._text_cache_valid = False
        self._eval_graphs = {}
        self.eval_graph = os.environ.get("MFK_EVAL_GRAPH", "1") != "0"
        self.eval_text_f32 = os.environ.get("MFK_EVAL_TEXT", "fp32") != "bf16"
        self.mom_initialized = False
        self.train_precision = "bf16"  # "fp32": forward_backward runs the split-operand fp32 training mode
        self.repack_trainable()

    # ------------------------------------------------------------------ parameters
    def _trainable_names(self, sd) -> List[str]:
        """Order of the arena: prompt learner | LayerNorms | resblocks.11 | never-updated."""
        pl = [k for k in sd if k.startswith("prompt_learner.") and "token_" not in k and "proj_vis_to_lang" not in k]
        ln = [k for k in sd if (".ln_" in k or "ln_pre." in k or "ln_post." in k or "ln_final." in k)]
        last = []
        for tw in (self.vis, self.txt):
            pre = f"{tw.name}.transformer.resblocks.{tw.L - 1}."
            if tw.L == 12:  # the reference unfreezes names containing "transformer.resblocks.11"
                last += [k for k in sd if k.startswith(pre) and ".ln_" not in k]
        tail = [k for k in sd if "proj_vis_to_lang" in k]  # trainable flag, never receives a gradient
        self._n_pl = sum(sd[k].numel() for k in pl)
        self._n_ln = sum(sd[k].numel() for k in ln)
        self._n_last = sum(sd[k].numel() for k in last)
        return pl + ln + last + tail

    def _build_arena(self, sd):
        names = self._trainable_names(sd)
        total = sum(sd[k].numel() for k in names)
        pad = (-total) % 64
        self.params = torch.zeros(total + pad, device=self.dev, dtype=F32)
        self.grads = torch.zeros_like(self.params)
        self.momentum = torch.zeros_like(self.params)
        self.p: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self.g: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self.offsets: "OrderedDict[str, tuple]" = OrderedDict()
        off = 0
        for k in names:
            n = sd[k].numel()
            # every tensor starts 16-byte aligned (all sizes are multiples of 4 elements)
            assert off % 4 == 0, k
            self.p[k] = self.params[off:off + n].view(sd[k].shape)
            self.g[k] = self.grads[off:off + n].view(sd[k].shape)
            self.p[k].copy_(sd[k].to(self.dev, F32))
            self.offsets[k] = (off, n)
            off += n
        self.n_params_total = off
        if self.trainable == "reference":
            self.n_update = self._n_pl + self._n_ln + self._n_last
        else:
            self.n_update = self._n_pl
        self.logit_scale = sd["logit_scale"].to(self.dev, F32).reshape(1).clone()
        self.prefix = sd["prompt_learner.token_prefix"].to(self.dev, F32).contiguous()
        self.suffix = sd["prompt_learner.token_suffix"].to(self.dev, F32).contiguous()

    @property
    def wgrad_last(self) -> bool:
        return self.trainable == "reference" and self.vis.L == 12

    def _pack_frozen(self, sd, share_from=None):
        dev = self.dev
        f32 = lambda k: sd[k].to(dev, F32).contiguous()
        for tw in (self.vis, self.txt):
            src_tw = None if share_from is None else (share_from.vis if tw is self.vis else share_from.txt)
            for l in range(tw.L):
                pre = f"{tw.name}.transformer.resblocks.{l}."
                w: Dict[str, torch.Tensor] = {}
                for lin in _LIN:
                    wk, bk = _lin_keys(lin)
                    full_w, full_b = pre + wk, pre + bk
                    if full_w in self.p:  # trainable master lives in the arena; bf16 copies made by repack
                        W = self.p[full_w]
                        w[lin + ".b"] = self.p[full_b]
                        w[lin + ".master"] = W
                        w[lin + ".w"] = torch.empty(W.shape, device=dev, dtype=BF16)
                        w[lin + ".wT"] = torch.empty(W.shape[1], W.shape[0], device=dev, dtype=BF16)
                    elif src_tw is not None and lin + ".master" not in src_tw.w[l]:
                        for sfx in (".w", ".wT", ".b"):
                            w[lin + sfx] = src_tw.w[l][lin + sfx]
                    else:
                        W = sd[full_w].to(dev, F32)
                        w[lin + ".w"] = W.to(BF16).contiguous()
                        w[lin + ".wT"] = W.t().to(BF16).contiguous()
                        w[lin + ".b"] = f32(full_b)
                for ln in ("ln_1", "ln_2"):
                    w[ln + ".g"], w[ln + ".b"] = self.p[pre + ln + ".weight"], self.p[pre + ln + ".bias"]
                tw.w.append(w)
        t = "text_encoder."
        self.eot_rows = (torch.arange(self.C) * self.Te + self.eot).to(dev, torch.int32)
        if share_from is not None:
            for a in ("conv_w", "cls", "vpos", "vproj", "vproj_T", "tpos", "tproj", "tproj_T"):
                setattr(self, a, getattr(share_from, a))
            return
        v = "image_encoder."
        conv = sd[v + "conv1.weight"].to(dev, F32)
        self.conv_w = conv.reshape(conv.shape[0], -1).to(BF16).contiguous()
        self.cls = f32(v + "class_embedding")
        self.vpos = f32(v + "positional_embedding")
        proj = sd[v + "proj"].to(dev, F32)
        self.vproj = proj.to(BF16).contiguous()          # [768,512]  (B operand of the dgrad GEMM)
        self.vproj_T = self._split_b(proj.t())           # [512,3*768] hi|hi|lo (B operand of the forward GEMM)
        self.tpos = f32(t + "positional_embedding")
        tp = sd[t + "text_projection"].to(dev, F32)
        self.tproj = tp.to(BF16).contiguous()
        self.tproj_T = self._split_b(tp.t())

    @staticmethod
    def _split_b(w: torch.Tensor) -> torch.Tensor:
        """[N,K] fp32 -> bf16 [N,3K] = hi | hi | lo (pairs with ops.split_bf16x3's hi | lo | hi A operand)."""
        w = w.contiguous().float()
        hi = w.to(BF16)
        lo = (w - hi.float()).to(BF16)
        return torch.cat([hi, hi, lo], dim=1).contiguous()

    def repack_trainable(self):
        """Refresh the bf16 (and transposed) copies of the trainable resblock weights from the fp32 arena: one
        grouped launch over all of them."""
        if getattr(self, "_repack_table", None) is None:
            probs = [(w[lin + ".master"], w[lin + ".wT"], w[lin + ".w"]) for tw in (self.vis, self.txt) for w in tw.w
                     for lin in _LIN if lin + ".master" in w]
            self._repack_table = ops.repack_table(probs, self.dev) if probs else False
        if self._repack_table is not False:
            ops.repack_grouped(*self._repack_table)
        self._text_cache_valid = False

    # ------------------------------------------------------------------ workspaces
    def _buf(self, name, shape, dtype):
        t = self._bufs.get(name)
        n = 1
        for s in shape:
            n *= s
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(n, device=self.dev, dtype=dtype)
            self._bufs[name] = t
            # CUDA graphs captured earlier hold raw pointers into the old buffers: bump the generation so
            # their owners re-capture before the next replay (see MaPLe.forward_backward)
            self._bufs["__generation__"] = self._bufs.get("__generation__", 0) + 1
        return t[:n].view(shape)

    @property
    def buffer_generation(self) -> int:
        return self._bufs.get("__generation__", 0)

    def _tower_bufs(self, tw: _Tower, N: int, T: int, train: bool):
        tw.N, tw.T, tw.M = N, T, N * T
        M, D, L = tw.M, tw.D, tw.L
        mode = "train" if train else "eval"  # separate pools: an eval batch never resizes training buffers
        b = lambda n, s, d: self._buf(f"{tw.name}.{mode}.{n}", s, d)
        ws = tw.ws = {}
        nl = L if train else 1
        ws["x1"] = b("x1", (nl + 1, M, D), F32)   # x1[l] = input of layer l (after splice); x1[L] = output
        ws["x2"] = b("x2", (nl, M, D), F32)       # x2[l] = after the attention residual
        ws["h"] = b("h", (M, D), BF16)
        ws["h2"] = b("h2", (M, D), BF16)
        ws["qkv"] = b("qkv", (nl, M, 3 * D), BF16)
        ws["att"] = b("att", (nl, M, D), BF16)
        ws["act"] = b("act", (M, 4 * D), BF16)
        # last block: only one row per sequence (CLS / EOT) is consumed -> out-proj + MLP on N gathered rows
        R = N
        ws["att_r"] = b("att_r", (R, D), BF16)
        ws["x1_r"] = b("x1_r", (R, D), F32)
        ws["x2_r"] = b("x2_r", (R, D), F32)
        ws["h2_r"] = b("h2_r", (R, D), BF16)
        ws["act_r"] = b("act_r", (R, 4 * D), BF16)
        ws["xout_r"] = b("xout_r", (R, D), F32)
        if train:
            ws["u_r"] = b("u_r", (R, 4 * D), BF16)
            ws["stat_r"] = b("stat_r", (2, R), F32)
            ws["lse_r"] = b("lse_r", (R * tw.heads,), F32)
            ws["g_r"] = b("g_r", (R, D), F32)
            ws["g16_r"] = b("g16_r", (R, D), BF16)
            ws["du_r"] = b("du_r", (R, 4 * D), BF16)
            ws["dh_r"] = b("dh_r", (R, D), BF16)
        if train:
            ws["u"] = b("u", (L, M, 4 * D), BF16)
            ws["lse"] = b("lse", (L, N * tw.heads * T), F32)
            ws["stat"] = b("stat", (L, 4, M), F32)  # mean1, rstd1, mean2, rstd2
            ws["g"] = b("g", (M, D), F32)
            ws["g16"] = b("g16", (M, D), BF16)
            ws["du"] = b("du", (M, 4 * D), BF16)
            ws["dh"] = b("dh", (M, D), BF16)
            ws["dqkv"] = b("dqkv", (M, 3 * D), BF16)
            ws["delta"] = b("delta", (N * tw.heads * T,), F32)
            # one dgamma/dbeta partial slot per LayerNorm backward of a step (2 per layer + final LN + ln_pre)
            ws["lnp_all"] = b("lnp_all", (2 * L + 3, 2 * D * ops.ln_bwd_ctas(M)), F32)
            ws["csum"] = b("csum", (32 * 4 * D,), F32)

    # ------------------------------------------------------------------ prompt learner
    def _prompt_learner_problems(self):
        """The J-1 compound projections + proj_lang_to_vis as one problem table (forward AND backward pointers).
        Everything it points to (arena slices, persistent workspaces) keeps its address for the engine's life."""
        p, G, nd, n = self.p, self.g, self.J - 1, self.n
        pl = "prompt_learner."
        self.deep_text: List[torch.Tensor] = []
        self.deep_vis: List[torch.Tensor] = []
        probs = []
        dpr = lambda tw, i: self._dprompt_all(tw)[i]
        for i in range(nd):
            Wn = f"{pl}compound_prompt_projections.{i}"
            W, b = p[Wn + ".weight"], p[Wn + ".bias"]
            y = self._buf(f"pl.y{i}", (n, W.shape[0]), F32)
            if i % 2 == 0:   # text parameter -> vision prompt; dx = dy W + d(text prompt i)
                pn = f"{pl}compound_prompts_text_parameters.{i // 2}"
                self.deep_vis.append(y)
                self.deep_text.append(p[pn])
                dy, dx_add = dpr(self.vis, i), dpr(self.txt, i)
            else:            # vision parameter -> text prompt
                pn = f"{pl}visual_deep_prompts_parameters.{(i - 1) // 2}"
                self.deep_text.append(y)
                self.deep_vis.append(p[pn])
                dy, dx_add = dpr(self.txt, i), dpr(self.vis, i)
            probs.append(dict(x=p[pn], W=W, b=b, y=y, dy=dy, dW=G[Wn + ".weight"], db=G[Wn + ".bias"], dx_add=dx_add,
                              dx=G[pn]))
        self.shared = self._buf("pl.shared", (n, self.vis.D), F32)
        self.d_shared = self._buf("pl.dshared", (n, self.vis.D), F32)
        self.d_ctx_t = self._buf("pl.dctx_t", (n, self.txt.D), F32)
        probs.append(dict(x=p[pl + "ctx"], W=p[pl + "proj_lang_to_vis.weight"], b=p[pl + "proj_lang_to_vis.bias"],
                          y=self.shared, dy=self.d_shared, dW=G[pl + "proj_lang_to_vis.weight"],
                          db=G[pl + "proj_lang_to_vis.bias"], dx_add=self.d_ctx_t, dx=G[pl + "ctx"]))
        self._pl_table = ops.small_linear_table(probs, self.dev)
        self._pl_keep = probs  # fixed-size buffers: _buf never reallocates them, so the pointers stay valid

    def _dprompt_all(self, tw: _Tower) -> torch.Tensor:
        """[J-1, n_ctx, D] gradients of the deep prompts of one tower (fixed address: the problem table points in)."""
        return self._buf(f"{tw.name}.dprompt_all", (max(self.J - 1, 1), self.n, tw.D), F32)

    def _prompt_learner_fwd(self):
        """MultiModalPromptLearner.forward (trainers/maple.py:194-215): all projections in one grouped launch."""
        if getattr(self, "_pl_table", None) is None:
            self._prompt_learner_problems()
        ops.linear_small_fwd_grouped(self._pl_table, max(self.vis.D, self.txt.D))

    # ------------------------------------------------------------------ one residual block
    def _slots(self, l: int, train: bool):
        """(x1 in, x1 out, per-layer slot) — training keeps every layer, inference ping-pongs two buffers."""
        return (l, l + 1, l) if train else (l % 2, (l + 1) % 2, 0)

    def _block_fwd(self, tw: _Tower, l: int, train: bool, rows=None, splice=None):
        """rows (int32 [N]) is given for the LAST block only: everything after the attention core then runs on
        those gathered rows and the compact [N, D] result is returned."""
        ws, w = tw.ws, tw.w[l]
        si, so, s = self._slots(l, train)
        x1, x1n, x2 = ws["x1"][si], ws["x1"][so], ws["x2"][s]
        st = ws["stat"][l] if train else (None, None, None, None)
        # splice = (prompt, T, row0, n_ctx): the deep-prompt splice of this layer is fused into ln_1 (x1 updated in place)
        ops.layernorm_fwd(x1, w["ln_1.g"], w["ln_1.b"], y_bf16=ws["h"], mean=st[0], rstd=st[1], splice=splice)
        qkv, att = ws["qkv"][s], ws["att"][s]
        ops.gemm(ws["h"], w["attn.in_proj.w"], bias=w["attn.in_proj.b"], out_bf16=qkv, ws=tw.gemm_ws)
        if rows is not None:
            # last block: the attention core is needed for the consumed (CLS / EOT) query row of each sequence only
            sr = ws["stat_r"] if train else (None, None)
            ops.attn_rows_fwd(qkv, rows, ws["att_r"], ws["lse_r"] if train else None, tw.N, tw.T, tw.heads, tw.causal)
            ops.gather_rows(x1, rows, ws["x1_r"])
            ops.gemm(ws["att_r"], w["attn.out_proj.w"], bias=w["attn.out_proj.b"], residual=ws["x1_r"],
                     out_f32=ws["x2_r"])
            ops.layernorm_fwd(ws["x2_r"], w["ln_2.g"], w["ln_2.b"], y_bf16=ws["h2_r"], mean=sr[0], rstd=sr[1])
            ops.gemm(ws["h2_r"], w["mlp.c_fc.w"], bias=w["mlp.c_fc.b"], act=1, out_bf16=ws["act_r"],
                     out_pre=ws["u_r"] if train else None)
            ops.gemm(ws["act_r"], w["mlp.c_proj.w"], bias=w["mlp.c_proj.b"], residual=ws["x2_r"],
                     out_f32=ws["xout_r"])
            return ws["xout_r"]
        ops.attn_fwd(qkv, att, ws["lse"][l], tw.N, tw.T, tw.heads, tw.causal) if train else \
            ops.attn_fwd(qkv, att, None, tw.N, tw.T, tw.heads, tw.causal)
        ops.gemm(att, w["attn.out_proj.w"], bias=w["attn.out_proj.b"], residual=x1, out_f32=x2, ws=tw.gemm_ws)
        ops.layernorm_fwd(x2, w["ln_2.g"], w["ln_2.b"], y_bf16=ws["h2"], mean=st[2], rstd=st[3])
        ops.gemm(ws["h2"], w["mlp.c_fc.w"], bias=w["mlp.c_fc.b"], act=1, out_bf16=ws["act"],
                 out_pre=ws["u"][l] if train else None, ws=tw.gemm_ws)
        ops.gemm(ws["act"], w["mlp.c_proj.w"], bias=w["mlp.c_proj.b"], residual=x2, out_f32=x1n, ws=tw.gemm_ws)
        return x1n

    def _tower_fwd(self, tw: _Tower, deep: List[torch.Tensor], row0: int, train: bool, rows=None):
        """Returns the compact [N, D] output rows (CLS / EOT) of the last block."""
        out = None
        for l in range(tw.L):
            splice = (deep[l - 1], tw.T, row0, self.n) if (l >= 1 and (l - 1) < len(deep)) else None
            out = self._block_fwd(tw, l, train, rows if l == tw.L - 1 else None, splice)
        return out

    def _wgrad(self, tw: _Tower, dy16, x16, dW, Nout, Kin):
        """dW[Nout,Kin] = dy^T x straight from the row-major bf16 activations (MN-major UMMA operands)."""
        ops.gemm_at_b(dy16, x16, dW)

    # ------------------------------------------------------------------ LayerNorm backward with deferred dgamma/dbeta
    def _ln_bwd(self, tw: _Tower, dy, x, mean, rstd, gamma, *, g_in=None, g_out, g_out_bf16=None, dgamma=None,
                dbeta=None, M=None, splice_grad=None):
        """LayerNorm backward whose per-CTA dgamma/dbeta partials go to this call's own slot of the tower's partial
        buffer; _ln_reduce() finishes every LayerNorm of the tower in one grouped launch (fixed order: the step's
        launch sequence is static, so slot k is always the same LayerNorm)."""
        if dgamma is None and dbeta is None:
            ops.layernorm_bwd(dy, x, mean, rstd, gamma, g_in=g_in, g_out=g_out, g_out_bf16=g_out_bf16, M=M,
                              splice_grad=splice_grad)
            return
        D = x.shape[-1]
        rows = x.numel() // D if M is None else M
        slot = tw.ln_slot
        tw.ln_slot += 1
        part = tw.ws["lnp_all"][slot]
        ops.layernorm_bwd(dy, x, mean, rstd, gamma, g_in=g_in, g_out=g_out, g_out_bf16=g_out_bf16, dgamma=dgamma,
                          dbeta=dbeta, partial_ws=part, M=M, defer=True, splice_grad=splice_grad)
        tw.ln_pending.append((part, ops.ln_bwd_ctas(rows), D, dgamma, dbeta, False))

    def _ln_reduce(self, tw: _Tower):
        if tw.ln_pending:
            key = (tw.ws["lnp_all"].data_ptr(), tw.M, len(tw.ln_pending))
            if tw.ln_table is None or tw.ln_key != key:
                tw.ln_table, tw.ln_key = ops.partial_reduce_table(tw.ln_pending, self.dev), key
            ops.partial_reduce_grouped(tw.ln_table, tw.D)
        tw.ln_pending, tw.ln_slot = [], 0

    def _last_block_bwd_rows(self, tw: _Tower, rows):
        """Backward of the last block's out-proj + MLP on the gathered rows. In: ws["g_r"] / ws["g16_r"] = gradient
        at the block output rows. Out: ws["g"] / ws["g16"] (full, zero except `rows`) = gradient after the attention
        residual, ws["dh_r"] (compact bf16 [R, D]) = gradient wrt the attention output rows."""
        l = tw.L - 1
        ws, w, D = tw.ws, tw.w[l], tw.D
        pre = f"{tw.name}.transformer.resblocks.{l}."
        wg, ln_grads, G = self.wgrad_last, self.trainable == "reference", self.g
        g, g16 = ws["g_r"], ws["g16_r"]
        ops.gemm(g16, w["mlp.c_proj.wT"], act=2, aux=ws["u_r"], out_bf16=ws["du_r"])
        if wg:
            ops.gemm_at_b(g16, ws["act_r"], G[pre + "mlp.c_proj.weight"])
            ops.colsum(g, G[pre + "mlp.c_proj.bias"], ws["csum"])
        ops.gemm(ws["du_r"], w["mlp.c_fc.wT"], out_bf16=ws["dh_r"])
        if wg:
            ops.gemm_at_b(ws["du_r"], ws["h2_r"], G[pre + "mlp.c_fc.weight"])
            ops.colsum(ws["du_r"], G[pre + "mlp.c_fc.bias"], ws["csum"])
        sr = ws["stat_r"]
        self._ln_bwd(tw, ws["dh_r"], ws["x2_r"], sr[0], sr[1], w["ln_2.g"], g_in=g, g_out=g, g_out_bf16=g16,
                          dgamma=G[pre + "ln_2.weight"] if ln_grads else None,
                          dbeta=G[pre + "ln_2.bias"] if ln_grads else None, M=g.shape[0])
        ops.gemm(g16, w["attn.out_proj.wT"], out_bf16=ws["dh_r"])
        if wg:
            ops.gemm_at_b(g16, ws["att_r"], G[pre + "attn.out_proj.weight"])
            ops.colsum(g, G[pre + "attn.out_proj.bias"], ws["csum"])
        # dense scatter: every row of ws["g"] is written (zeros except the consumed rows) in one pass; ws["g16"] is an
        # OUTPUT of the following ln_1 backward, so it needs neither the fill nor the scatter
        ops.scatter_rows_dense(g, rows, ws["g"], tw.N, tw.T)
        tw.last_rows = rows  # ws["dh_r"] (gradient wrt the attention output rows) feeds the single-query backward

    def _block_bwd(self, tw: _Tower, l: int, attn_only: bool = False, splice_grad=None):
        """attn_only: the MLP / out-proj part was already done on gathered rows (_last_block_bwd_rows)."""
        ws, w, D, M = tw.ws, tw.w[l], tw.D, tw.M
        g, g16 = ws["g"], ws["g16"]
        st = ws["stat"][l]
        pre = f"{tw.name}.transformer.resblocks.{l}."
        wg = self.wgrad_last and l == tw.L - 1
        ln_grads = self.trainable == "reference"
        G = self.g
        if attn_only:
            ops.attn_rows_bwd(ws["qkv"][l], tw.last_rows, ws["dh_r"], ws["lse_r"], ws["dqkv"], tw.N, tw.T, tw.heads,
                              tw.causal)
            ops.gemm(ws["dqkv"], w["attn.in_proj.wT"], out_bf16=ws["dh"], ws=tw.gemm_ws)
            if wg:
                self._wgrad(tw, ws["dqkv"], ws["h"], G[pre + "attn.in_proj_weight"], 3 * D, D)
                ops.colsum(ws["dqkv"], G[pre + "attn.in_proj_bias"], ws["csum"])
            self._ln_bwd(tw, ws["dh"], ws["x1"][l], st[0], st[1], w["ln_1.g"], g_in=g, g_out=g, g_out_bf16=g16,
                              dgamma=G[pre + "ln_1.weight"] if ln_grads else None,
                              dbeta=G[pre + "ln_1.bias"] if ln_grads else None, splice_grad=splice_grad)
            return
        # ---- MLP branch
        ops.gemm(g16, w["mlp.c_proj.wT"], act=2, aux=ws["u"][l], out_bf16=ws["du"], ws=tw.gemm_ws)
        if wg:
            self._wgrad(tw, g16, ws["act"], G[pre + "mlp.c_proj.weight"], D, 4 * D)
            ops.colsum(g, G[pre + "mlp.c_proj.bias"], ws["csum"])
        ops.gemm(ws["du"], w["mlp.c_fc.wT"], out_bf16=ws["dh"], ws=tw.gemm_ws)
        if wg:
            self._wgrad(tw, ws["du"], ws["h2"], G[pre + "mlp.c_fc.weight"], 4 * D, D)
            ops.colsum(ws["du"], G[pre + "mlp.c_fc.bias"], ws["csum"])
        self._ln_bwd(tw, ws["dh"], ws["x2"][l], st[2], st[3], w["ln_2.g"], g_in=g, g_out=g, g_out_bf16=g16,
                          dgamma=G[pre + "ln_2.weight"] if ln_grads else None,
                          dbeta=G[pre + "ln_2.bias"] if ln_grads else None)
        # ---- attention branch
        da = ws["dh"]
        ops.gemm(g16, w["attn.out_proj.wT"], out_bf16=da, ws=tw.gemm_ws)
        if wg:
            self._wgrad(tw, g16, ws["att"][l], G[pre + "attn.out_proj.weight"], D, D)
            ops.colsum(g, G[pre + "attn.out_proj.bias"], ws["csum"])
        ops.attn_bwd(ws["qkv"][l], ws["att"][l], da, ws["lse"][l], ws["delta"], ws["dqkv"], tw.N, tw.T, tw.heads,
                     tw.causal)
        ops.gemm(ws["dqkv"], w["attn.in_proj.wT"], out_bf16=ws["dh"], ws=tw.gemm_ws)
        if wg:
            self._wgrad(tw, ws["dqkv"], ws["h"], G[pre + "attn.in_proj_weight"], 3 * D, D)
            ops.colsum(ws["dqkv"], G[pre + "attn.in_proj_bias"], ws["csum"])
        self._ln_bwd(tw, ws["dh"], ws["x1"][l], st[0], st[1], w["ln_1.g"], g_in=g, g_out=g, g_out_bf16=g16,
                          dgamma=G[pre + "ln_1.weight"] if ln_grads else None,
                          dbeta=G[pre + "ln_1.bias"] if ln_grads else None, splice_grad=splice_grad)

    # ------------------------------------------------------------------ towers: embed + head rows
    def _vision_embed(self, img: torch.Tensor, train: bool, wait_before_assemble=None):
        B = img.shape[0]
        tw = self.vis
        self._tower_bufs(tw, B, self.Tv, train)
        col = self._buf("vis.col", (B * self.P, 3 * self.patch * self.patch), BF16)
        tok = self._buf("vis.tok", (B * self.P, tw.D), F32)
        ops.patch_im2col(img, col)
        ops.gemm(col, self.conv_w, out_f32=tok)
        p = self.p
        x0 = self._buf("vis.x0", (tw.M, tw.D), F32) if train else None
        self.vstat0 = self._buf("vis.stat0", (2, tw.M), F32)
        if wait_before_assemble is not None:
            torch.cuda.current_stream().wait_event(wait_before_assemble)
        ops.vis_assemble_lnpre(tok, self.cls, self.vpos, self.shared, p["image_encoder.ln_pre.weight"],
                               p["image_encoder.ln_pre.bias"], x0, tw.ws["x1"][0], self.vstat0[0], self.vstat0[1], B,
                               self.Tv, self.n)
        self.vx0 = x0

    def _features(self, tw: _Tower, xout, rows, ln_g, ln_b, projT, name, R, train, rowidx=None):
        """ln_post / ln_final on the R consumed rows (xout compact [R, D], or full with ``rowidx`` gathering them)
        + split-precision projection to the joint embedding."""
        y = self._buf(name + ".y", (R, tw.D), F32)
        y3 = self._buf(name + ".y3", (R, 3 * tw.D), BF16)
        xs = self._buf(name + ".xs", (R, tw.D), F32)
        stat = self._buf(name + ".st", (2, R), F32)
        ops.layernorm_fwd(xout, ln_g, ln_b, rowidx=rowidx, y_f32=y, x_save=xs, mean=stat[0], rstd=stat[1], M=R)
        # split-precision head (hi/lo bf16, one GEMM of depth 3D): keeps ~16 mantissa bits in the features
        ops.split_bf16x3(y, y3)
        feat = self._buf(name + ".feat", (R, self.E), F32)
        ops.gemm(y3, projT, out_f32=feat)
        return feat, xs, stat

    def _text_features(self, train: bool, class_range=None):
        """Text tower over all classes, or over the contiguous class shard [c0, c1) (inference, config 5)."""
        tw, p = self.txt, self.p
        c0, c1 = (0, self.C) if class_range is None else class_range
        Cn = c1 - c0
        self._tower_bufs(tw, Cn, self.Te, train)
        ops.text_assemble(self.prefix[c0:c1], p["prompt_learner.ctx"], self.suffix[c0:c1], self.tpos,
                          tw.ws["x1"][0], Cn, self.Te, self.n, self.Tfull)
        rows = self.eot_rows if class_range is None else \
            (self.eot_rows[c0:c1] - c0 * self.Te).contiguous()
        self.txt_rows = rows
        xout = self._tower_fwd(tw, self.deep_text, 1, train, rows)
        return self._features(tw, xout, rows, p["text_encoder.ln_final.weight"],
                              p["text_encoder.ln_final.bias"], self.tproj_T, "txt", Cn, train)

    def _image_features(self, img, train: bool, wait_before_assemble=None):
        tw, p = self.vis, self.p
        self._vision_embed(img, train, wait_before_assemble)
        B = img.shape[0]
        key = f"cls_rows{B}"
        if key not in self._bufs:
            self._bufs[key] = (torch.arange(B, device=self.dev, dtype=torch.int32) * self.Tv).contiguous()
        self.cls_rows = self._bufs[key]
        xout = self._tower_fwd(tw, self.deep_vis, self.Tv - self.n, train, self.cls_rows)
        return self._features(tw, xout, self.cls_rows, p["image_encoder.ln_post.weight"],
                              p["image_encoder.ln_post.bias"], self.vproj_T, "vis", B, train)

    # ------------------------------------------------------------------ public: inference
    @torch.no_grad()
    def logits(self, img: torch.Tensor, cache_text: bool = True, shard_classes: bool = False,
               precision: str = "bf16") -> torch.Tensor:
        """Eval path of CustomCLIP.forward (trainers/maple.py:381): returns logits [B, C] (fp32).
        ``precision="fp32"``: the parity mode of the contract (logits within 1e-3 of the reference's fp32 path) —
        every GEMM runs with bf16x3 split operands on the same tcgen05 kernel, LayerNorm / softmax / QuickGELU / the
        residual stream in fp32 (see _block_fwd_f32); slower, inference only.
        Text features are input independent and cached across eval batches until parameters change.
        ``shard_classes``: under torch.distributed each rank runs the text tower on C/world classes and the
        [C, E] feature matrix is all-gathered once; images stay data-parallel."""
        assert precision in ("bf16", "fp32")
        img = img.to(self.dev, F32).contiguous()
        B = img.shape[0]
        self._prompt_learner_fwd()
        if precision == "fp32":
            return self._logits_f32(img)
        if not (cache_text and self._text_cache_valid):
            import torch.distributed as dist
            if shard_classes and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                world, rank = dist.get_world_size(), dist.get_rank()
                if self.C % world:
                    raise ValueError(f"shard_classes needs C={self.C} divisible by world size {world}")
                per = self.C // world
                ft = self._eval_text_features((rank * per, (rank + 1) * per))
                full = torch.empty(self.C, self.E, device=self.dev, dtype=F32)
                dist.all_gather_into_tensor(full, ft.contiguous())
            else:
                full = self._eval_text_features()
            # persistent per-engine cache (never in the workspace shared by co-located clients: their prompts
            # differ); eval graphs hold its address across refreshes
            if getattr(self, "_ft_cache", None) is None:
                self._ft_cache = torch.empty(self.C, self.E, device=self.dev, dtype=F32)
            self._ft_cache.copy_(full)
            self._text_cache_valid = True
        if cache_text and self.eval_graph and not torch.cuda.is_current_stream_capturing():
            return self._logits_graphed(img)
        return self._logits_image_part(img, torch.empty(B, self.C, device=self.dev, dtype=F32))

    def _eval_text_features(self, class_range=None) -> torch.Tensor:
        """Text features of the evaluation path. They are input independent and cached across batches, so they are
        computed in the split-operand fp32 mode at no per-batch cost: with bf16 operands the 10-token text tower
        contributes ~3x the logit error of the 199-token vision tower (relative feature error 9e-3 against 3e-3,
        measured on the CPU with bf16-rounded operands), which is most of the distance to the reference's logits.
        MFK_EVAL_TEXT=bf16 keeps the bf16 tensor-core tower (the training step always uses it)."""
        if self.eval_text_f32:
            return self._text_features_f32(class_range)
        return self._text_features(False, class_range)[0]

    def _logits_image_part(self, img, out):
        """Vision tower + logits head against the cached text features (the per-batch part of evaluation)."""
        B = img.shape[0]
        fi, _, _ = self._image_features(img, False)
        self._last_fi = fi
        ws = self._buf("head.ws", (ops.head_workspace_floats(B, self.C, self.E),), F32)
        ops.head_forward_backward(fi, self._ft_cache, self.logit_scale, None, out, None, None, None, ws)
        return out

    def _logits_graphed(self, img):
        """The ~250 launches of one evaluation batch replayed from a CUDA graph (one per batch size): eager
        evaluation is bound by the host's launch rate, not by the GPU. The first batch of a size runs eagerly (it
        also allocates the workspaces), the graph is captured right after and replayed from then on; prompts and
        text features live in persistent buffers, so parameter updates do not invalidate it."""
        B = img.shape[0]
        ent = self._eval_graphs.get(B)
        if ent is not None and ent["gen"] == self.buffer_generation:
            ent["img"].copy_(img, non_blocking=True)
            ent["graph"].replay()
            return ent["out"].clone()
        out = self._logits_image_part(img, torch.empty(B, self.C, device=self.dev, dtype=F32))
        ent = {"img": img.clone(), "out": torch.empty(B, self.C, device=self.dev, dtype=F32)}
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._logits_image_part(ent["img"], ent["out"])
        ent["graph"], ent["gen"] = g, self.buffer_generation
        self._eval_graphs[B] = ent
        return out

    # ------------------------------------------------------------------ fp32 mode (parity contract, inference only)
    def _w3(self, tw: _Tower, l: int, lin: str) -> torch.Tensor:
        """[N, 3K] hi|hi|lo split of a block weight (cached for frozen weights, rebuilt for trainable ones)."""
        w = tw.w[l]
        key = lin + ".w3"
        if key not in w or lin + ".master" in w:
            src = w[lin + ".master"] if lin + ".master" in w else self._frozen_f32(tw, l, lin)
            w[key] = self._split_b(src)
        return w[key]

    def _frozen_f32(self, tw: _Tower, l: int, lin: str) -> torch.Tensor:
        # frozen weights are fp16 values in the reference: their bf16 copy is NOT exact, so the fp32 originals are kept
        return self._sd_src[f"{tw.name}.transformer.resblocks.{l}.{_lin_keys(lin)[0]}"].to(self.dev, F32)

    def _split_a(self, x: torch.Tensor, name: str) -> torch.Tensor:
        out = self._buf(name, (x.shape[0], 3 * x.shape[1]), BF16)
        ops.split_bf16x3(x, out)
        return out

    def _block_fwd_f32(self, tw: _Tower, l: int, x1: torch.Tensor, x1n: torch.Tensor):
        """ResidualAttentionBlock_MaPLe.forward (clip/model.py:350-351) in fp32 mode; x1 -> x1n (both fp32 [M, D])."""
        w, M, D = tw.w[l], tw.M, tw.D
        pre = f"{tw.name}.f32."
        hf = self._buf(pre + "hf", (M, D), F32)
        x2 = self._buf(pre + "x2", (M, D), F32)
        qkv = self._buf(pre + "qkv", (M, 3 * D), F32)
        att = self._buf(pre + "att", (M, D), F32)
        u = self._buf(pre + "u", (M, 4 * D), F32)
        ops.layernorm_fwd(x1, w["ln_1.g"], w["ln_1.b"], y_f32=hf)
        ops.gemm(self._split_a(hf, pre + "a3"), self._w3(tw, l, "attn.in_proj"), bias=w["attn.in_proj.b"], out_f32=qkv)
        ops.attn_fwd_f32(qkv, att, tw.N, tw.T, tw.heads, tw.causal)
        ops.gemm(self._split_a(att, pre + "a3"), self._w3(tw, l, "attn.out_proj"), bias=w["attn.out_proj.b"],
                 residual=x1, out_f32=x2)
        ops.layernorm_fwd(x2, w["ln_2.g"], w["ln_2.b"], y_f32=hf)
        ops.gemm(self._split_a(hf, pre + "a3"), self._w3(tw, l, "mlp.c_fc"), bias=w["mlp.c_fc.b"], out_f32=u)
        act3 = self._buf(pre + "act3", (M, 12 * D), BF16)
        ops.quickgelu_split_bf16x3(u, act3)
        ops.gemm(act3, self._w3(tw, l, "mlp.c_proj"), bias=w["mlp.c_proj.b"], residual=x2, out_f32=x1n)

    def _tower_fwd_f32(self, tw: _Tower, deep: List[torch.Tensor], row0: int, x: torch.Tensor) -> torch.Tensor:
        bufs = [x, self._buf(f"{tw.name}.f32.xb", tuple(x.shape), F32)]
        for l in range(tw.L):
            xin, xout = bufs[l % 2], bufs[(l + 1) % 2]
            if l >= 1 and (l - 1) < len(deep):
                ops.prompt_splice_fwd(xin, deep[l - 1], tw.N, tw.T, row0, self.n)
            self._block_fwd_f32(tw, l, xin, xout)
        return bufs[tw.L % 2]

    def _text_features_f32(self, class_range=None) -> torch.Tensor:
        """Text tower in the split-operand fp32 mode over all classes or the shard [c0, c1): [Cn, E] fp32 features.
        Rows after the last EOT are dead under the causal mask, so only T_eff positions run."""
        p = self.p
        tw = self.txt
        c0, c1 = (0, self.C) if class_range is None else class_range
        Cn = c1 - c0
        tw.N, tw.T, tw.M = Cn, self.Te, Cn * self.Te
        xt = self._buf("txt.f32.x", (tw.M, tw.D), F32)
        ops.text_assemble(self.prefix[c0:c1], p["prompt_learner.ctx"], self.suffix[c0:c1], self.tpos, xt, Cn, self.Te,
                          self.n, self.Tfull)
        xt = self._tower_fwd_f32(tw, self.deep_text, 1, xt)
        rows = self.eot_rows if class_range is None else (self.eot_rows[c0:c1] - c0 * self.Te).contiguous()
        ft, _, _ = self._features(tw, xt, None, p["text_encoder.ln_final.weight"], p["text_encoder.ln_final.bias"],
                                  self.tproj_T, "txt32", Cn, False, rowidx=rows)
        return ft

    def _logits_f32(self, img: torch.Tensor) -> torch.Tensor:
        p, B = self.p, img.shape[0]
        ft = self._text_features_f32()
        # ---- vision tower
        tw = self.vis
        tw.N, tw.T, tw.M = B, self.Tv, B * self.Tv
        colf = self._buf("vis.f32.col", (B * self.P, 3 * self.patch * self.patch), F32)
        tok = self._buf("vis.f32.tok", (B * self.P, tw.D), F32)
        ops.patch_im2col_f32(img, colf)
        if getattr(self, "_conv_w3", None) is None:
            self._conv_w3 = self._split_b(self._sd_src["image_encoder.conv1.weight"].to(self.dev, F32).reshape(tw.D, -1))
        ops.gemm(self._split_a(colf, "vis.f32.col3"), self._conv_w3, out_f32=tok)
        xv = self._buf("vis.f32.x", (tw.M, tw.D), F32)
        st = self._buf("vis.f32.st", (2, tw.M), F32)
        ops.vis_assemble_lnpre(tok, self.cls, self.vpos, self.shared, p["image_encoder.ln_pre.weight"],
                               p["image_encoder.ln_pre.bias"], None, xv, st[0], st[1], B, self.Tv, self.n)
        xv = self._tower_fwd_f32(tw, self.deep_vis, self.Tv - self.n, xv)
        key = f"cls_rows{B}"
        if key not in self._bufs:
            self._bufs[key] = (torch.arange(B, device=self.dev, dtype=torch.int32) * self.Tv).contiguous()
        fi, _, _ = self._features(tw, xv, None, p["image_encoder.ln_post.weight"], p["image_encoder.ln_post.bias"],
                                  self.vproj_T, "vis32", B, False, rowidx=self._bufs[key])
        self._last_fi = fi
        out = torch.empty(B, self.C, device=self.dev, dtype=F32)
        ws = self._buf("head.ws", (ops.head_workspace_floats(B, self.C, self.E),), F32)
        ops.head_forward_backward(fi, ft, self.logit_scale, None, out, None, None, None, ws)
        return out

    # ------------------------------------------------------------------ fp32 mode: training step
    # cfg.TRAINER.MAPLE.PREC = "fp32": the reference calls clip_model.float() and trains the fp32 model
    # (trainers/maple.py:438-439, 590). Same schedule as the bf16 step, with every contraction as a split-operand
    # (bf16x3) GEMM on the tcgen05 kernel — forward, dgrad and the wgrads of resblocks.11 — and fp32 LayerNorm,
    # attention, QuickGELU and residual stream in between. The last block runs in full (no consumed-row shortcut):
    # this is the parity mode (gradients within 1e-3 of the reference's fp32 autograd), not the fast path.
    def _w3T(self, tw: _Tower, l: int, lin: str) -> torch.Tensor:
        """[K, 3N] hi|hi|lo split of W^T: the B operand of the fp32-mode dgrad GEMM dX = dY W."""
        w = tw.w[l]
        key = lin + ".w3T"
        if key not in w or lin + ".master" in w:
            src = w[lin + ".master"] if lin + ".master" in w else self._frozen_f32(tw, l, lin)
            w[key] = self._split_b(src.t())
        return w[key]

    def _f32_train_bufs(self, tw: _Tower, N: int, T: int):
        tw.N, tw.T, tw.M = N, T, N * T
        M, D, L = tw.M, tw.D, tw.L
        b = lambda n, s, d=F32: self._buf(f"{tw.name}.f32t.{n}", s, d)
        S = dict(x1=b("x1", (L + 1, M, D)), x2=b("x2", (L, M, D)), qkv=b("qkv", (L, M, 3 * D)),
                 att=b("att", (L, M, D)), u=b("u", (L, M, 4 * D)), stat=b("stat", (L, 4, M)),
                 h1=b("h1", (M, D)), h2=b("h2", (M, D)), g=b("g", (M, D)), g_r=b("g_r", (N, D)),
                 dact=b("dact", (M, 4 * D)), du=b("du", (M, 4 * D)), dh=b("dh", (M, D)), dqkv=b("dqkv", (M, 3 * D)),
                 attn_ws=b("attn_ws", (2 * N * tw.heads * T,)), rhs=b("rhs", (M, 3 * D), BF16),
                 lnp_all=b("lnp_all", (2 * L + 3, 2 * D * ops.ln_bwd_ctas(M))), csum=b("csum", (32 * 4 * D,)))
        tw.ws = S  # _ln_bwd / _ln_reduce take their partial slots from tw.ws["lnp_all"]
        return S

    def _block_fwd_f32_train(self, tw: _Tower, l: int, splice=None):
        S, w, M, D = tw.ws, tw.w[l], tw.M, tw.D
        nm = f"{tw.name}.f32t."
        x1, x1n, x2, st = S["x1"][l], S["x1"][l + 1], S["x2"][l], S["stat"][l]
        ops.layernorm_fwd(x1, w["ln_1.g"], w["ln_1.b"], y_f32=S["h1"], mean=st[0], rstd=st[1], splice=splice)
        ops.gemm(self._split_a(S["h1"], nm + "a3"), self._w3(tw, l, "attn.in_proj"), bias=w["attn.in_proj.b"],
                 out_f32=S["qkv"][l])
        ops.attn_fwd_f32(S["qkv"][l], S["att"][l], tw.N, tw.T, tw.heads, tw.causal)
        ops.gemm(self._split_a(S["att"][l], nm + "a3"), self._w3(tw, l, "attn.out_proj"), bias=w["attn.out_proj.b"],
                 residual=x1, out_f32=x2)
        ops.layernorm_fwd(x2, w["ln_2.g"], w["ln_2.b"], y_f32=S["h2"], mean=st[2], rstd=st[3])
        ops.gemm(self._split_a(S["h2"], nm + "a3"), self._w3(tw, l, "mlp.c_fc"), bias=w["mlp.c_fc.b"],
                 out_f32=S["u"][l])
        act3 = self._buf(nm + "act3", (M, 12 * D), BF16)
        ops.quickgelu_split_bf16x3(S["u"][l], act3)
        ops.gemm(act3, self._w3(tw, l, "mlp.c_proj"), bias=w["mlp.c_proj.b"], residual=x2, out_f32=x1n)

    def _block_bwd_f32(self, tw: _Tower, l: int, splice_grad=None):
        """In / out: tw.ws["g"] = gradient at the block output / input (fp32). Mirrors _block_bwd."""
        S, w, M, D = tw.ws, tw.w[l], tw.M, tw.D
        nm = f"{tw.name}.f32t."
        g, st, G = S["g"], S["stat"][l], self.g
        pre = f"{tw.name}.transformer.resblocks.{l}."
        wg = self.wgrad_last and l == tw.L - 1
        ln_grads = self.trainable == "reference"
        rows3 = lambda t: t.view(3 * M, t.shape[1] // 3)  # [M, 3K] split buffer read as [3M, K]
        # ---- MLP branch
        g3 = self._split_a(g, nm + "g3")
        ops.gemm(g3, self._w3T(tw, l, "mlp.c_proj"), out_f32=S["dact"])
        if wg:
            act3 = self._buf(nm + "act3", (M, 12 * D), BF16)
            ops.quickgelu_split_bf16x3(S["u"][l], act3)
            ops.split_bf16x3_rhs(g, S["rhs"])
            ops.gemm_at_b(rows3(S["rhs"]), rows3(act3), G[pre + "mlp.c_proj.weight"])
            ops.colsum(g, G[pre + "mlp.c_proj.bias"], S["csum"])
        ops.dquickgelu_mul_f32(S["dact"], S["u"][l], S["du"])
        du3 = self._split_a(S["du"], nm + "du3")
        ops.gemm(du3, self._w3T(tw, l, "mlp.c_fc"), out_f32=S["dh"])
        if wg:
            ops.split_bf16x3_rhs(S["h2"], S["rhs"])
            ops.gemm_at_b(rows3(du3), rows3(S["rhs"]), G[pre + "mlp.c_fc.weight"])
            ops.colsum(S["du"], G[pre + "mlp.c_fc.bias"], S["csum"])
        self._ln_bwd(tw, S["dh"], S["x2"][l], st[2], st[3], w["ln_2.g"], g_in=g, g_out=g,
                     dgamma=G[pre + "ln_2.weight"] if ln_grads else None,
                     dbeta=G[pre + "ln_2.bias"] if ln_grads else None)
        # ---- attention branch
        g3 = self._split_a(g, nm + "g3")
        da = S["dh"]
        ops.gemm(g3, self._w3T(tw, l, "attn.out_proj"), out_f32=da)
        if wg:
            ops.split_bf16x3_rhs(g, S["rhs"])
            ops.gemm_at_b(rows3(S["rhs"]), rows3(self._split_a(S["att"][l], nm + "a3")),
                          G[pre + "attn.out_proj.weight"])
            ops.colsum(g, G[pre + "attn.out_proj.bias"], S["csum"])
        ops.attn_bwd_f32(S["qkv"][l], da, S["dqkv"], S["attn_ws"], tw.N, tw.T, tw.heads, tw.causal)
        dq3 = self._split_a(S["dqkv"], nm + "dq3")
        ops.gemm(dq3, self._w3T(tw, l, "attn.in_proj"), out_f32=S["dh"])
        if wg:
            ops.split_bf16x3_rhs(S["h1"], S["rhs"])
            ops.gemm_at_b(rows3(dq3), rows3(S["rhs"]), G[pre + "attn.in_proj_weight"])
            ops.colsum(S["dqkv"], G[pre + "attn.in_proj_bias"], S["csum"])
        self._ln_bwd(tw, S["dh"], S["x1"][l], st[0], st[1], w["ln_1.g"], g_in=g, g_out=g,
                     dgamma=G[pre + "ln_1.weight"] if ln_grads else None,
                     dbeta=G[pre + "ln_1.bias"] if ln_grads else None, splice_grad=splice_grad)

    def _tower_fwd_f32_train(self, tw: _Tower, deep, row0: int, rows, ln_g, ln_b, projT, name: str):
        for l in range(tw.L):
            splice = (deep[l - 1], tw.T, row0, self.n) if (l >= 1 and (l - 1) < len(deep)) else None
            self._block_fwd_f32_train(tw, l, splice)
        return self._features(tw, tw.ws["x1"][tw.L], None, ln_g, ln_b, projT, name, tw.N, True, rowidx=rows)

    def _tower_bwd_f32(self, tw: _Tower, dfeat, proj3, xs, stat, rows, lnname: str, deep_row0: int, name: str):
        """d(features) -> gradient at the tower's (post-embedding) input, left in tw.ws["g"]; fills the deep-prompt
        gradients of this tower (_dprompt_all). Mirrors _tower_bwd."""
        p, G, n, nd = self.p, self.g, self.n, self.J - 1
        ln_grads = self.trainable == "reference"
        S, D, R = tw.ws, tw.D, tw.N
        dy = self._buf(name + ".dy", (R, D), F32)
        ops.gemm(self._split_a(dfeat, name + ".df3"), proj3, out_f32=dy)
        self._ln_bwd(tw, dy, xs, stat[0], stat[1], p[lnname + ".weight"], g_out=S["g_r"],
                     dgamma=G[lnname + ".weight"] if ln_grads else None,
                     dbeta=G[lnname + ".bias"] if ln_grads else None, M=R)
        ops.scatter_rows_dense(S["g_r"], rows, S["g"], tw.N, tw.T)
        ns = max(0, min(nd, tw.L - 1))
        gp = self._buf(f"{tw.name}.gprompt", (max(ns, 1), tw.N * n, D), F32)
        dp_all = self._dprompt_all(tw)
        for l in reversed(range(tw.L)):
            sg = (gp[l - 1], tw.T, deep_row0, n) if 1 <= l <= ns else None
            self._block_bwd_f32(tw, l, splice_grad=sg)
        if ns > 0:
            ops.prompt_splice_bwd_batched(gp[:ns], dp_all[:ns], tw.N, n, 0, n, True)
        if ns < nd:
            dp_all[ns:].zero_()

    @torch.no_grad()
    def forward_backward_f32(self, img: torch.Tensor, label: torch.Tensor, loss_out: Optional[torch.Tensor] = None):
        """forward_backward in the fp32 mode (same contract: fills ``self.g``, returns (loss[1], logits[B, C]))."""
        assert img.is_cuda and img.dtype == F32 and label.is_cuda and label.dtype == torch.int64
        B, C, n = img.shape[0], self.C, self.n
        p, G = self.p, self.g
        ln_grads = self.trainable == "reference"
        self._text_cache_valid = False
        for tw in (self.vis, self.txt):
            tw.ln_slot, tw.ln_pending = 0, []
        self._prompt_learner_fwd()
        if getattr(self, "_proj3", None) is None:  # [D, 3E] splits of the two projections (B operands of their dgrads)
            self._proj3 = (self._split_b(self._sd_src["image_encoder.proj"].to(self.dev, F32)),
                           self._split_b(self._sd_src["text_encoder.text_projection"].to(self.dev, F32)))
        # ---- text tower forward
        tt = self.txt
        St = self._f32_train_bufs(tt, C, self.Te)
        ops.text_assemble(self.prefix, p["prompt_learner.ctx"], self.suffix, self.tpos, St["x1"][0], C, self.Te, n,
                          self.Tfull)
        self.txt_rows = self.eot_rows
        ft, txs, tstat = self._tower_fwd_f32_train(tt, self.deep_text, 1, self.eot_rows,
                                                   p["text_encoder.ln_final.weight"], p["text_encoder.ln_final.bias"],
                                                   self.tproj_T, "txt32t")
        # ---- vision tower forward
        tv = self.vis
        Sv = self._f32_train_bufs(tv, B, self.Tv)
        colf = self._buf("vis.f32.col", (B * self.P, 3 * self.patch * self.patch), F32)
        tok = self._buf("vis.f32.tok", (B * self.P, tv.D), F32)
        ops.patch_im2col_f32(img, colf)
        if getattr(self, "_conv_w3", None) is None:
            self._conv_w3 = self._split_b(self._sd_src["image_encoder.conv1.weight"].to(self.dev, F32).reshape(tv.D, -1))
        ops.gemm(self._split_a(colf, "vis.f32.col3"), self._conv_w3, out_f32=tok)
        x0 = self._buf("vis.f32t.x0", (tv.M, tv.D), F32)
        st0 = self._buf("vis.f32t.st0", (2, tv.M), F32)
        ops.vis_assemble_lnpre(tok, self.cls, self.vpos, self.shared, p["image_encoder.ln_pre.weight"],
                               p["image_encoder.ln_pre.bias"], x0, Sv["x1"][0], st0[0], st0[1], B, self.Tv, n)
        key = f"cls_rows{B}"
        if key not in self._bufs:
            self._bufs[key] = (torch.arange(B, device=self.dev, dtype=torch.int32) * self.Tv).contiguous()
        self.cls_rows = self._bufs[key]
        fi, vxs, vstat = self._tower_fwd_f32_train(tv, self.deep_vis, self.Tv - n, self.cls_rows,
                                                   p["image_encoder.ln_post.weight"], p["image_encoder.ln_post.bias"],
                                                   self.vproj_T, "vis32t")
        # ---- head
        logits = self._buf("head.logits", (B, C), F32)
        loss = loss_out if loss_out is not None else self._buf("head.loss", (1,), F32)
        dfi, dft = self._buf("head.dfi", (B, self.E), F32), self._buf("head.dft", (C, self.E), F32)
        hws = self._buf("head.ws", (ops.head_workspace_floats(B, C, self.E),), F32)
        ops.head_forward_backward(fi, ft, self.logit_scale, label, logits, loss, dfi, dft, hws)
        # ---- backward: text, then vision (one stream; the towers' workspaces are separate)
        self._tower_bwd_f32(tt, dft, self._proj3[1], txs, tstat, self.eot_rows, "text_encoder.ln_final", 1, "txt32t")
        ops.prompt_splice_bwd(St["g"], None, self.d_ctx_t, C, self.Te, 1, n, False, False)
        self._ln_reduce(tt)
        self._tower_bwd_f32(tv, dfi, self._proj3[0], vxs, vstat, self.cls_rows, "image_encoder.ln_post", self.Tv - n,
                            "vis32t")
        self._ln_bwd(tv, Sv["g"], x0, st0[0], st0[1], p["image_encoder.ln_pre.weight"], g_out=Sv["g"],
                     dgamma=G["image_encoder.ln_pre.weight"] if ln_grads else None,
                     dbeta=G["image_encoder.ln_pre.bias"] if ln_grads else None)
        ops.prompt_splice_bwd(Sv["g"], None, self.d_shared, B, self.Tv, self.Tv - n, n, True, False)
        self._ln_reduce(tv)
        ops.linear_small_bwd_grouped(self._pl_table, n, max(self.vis.D, self.txt.D), max(self.vis.D, self.txt.D))
        self.last = dict(image_features=fi, text_features=ft, dfi=dfi, dft=dft)
        return loss, logits

    def last_image_features(self) -> torch.Tensor:
        return self._last_fi.clone()

    def _side_stream(self):
        st = self._bufs.get("__side_stream__")
        if st is None:
            st = torch.cuda.Stream(device=self.dev)
            self._bufs["__side_stream__"] = st
        return st

    def _tower_bwd(self, tw: _Tower, dfeat, proj, xs, stat, rows, lnname, R, deep_row0):
        """Backward of one tower from d(features) down to the gradient at its (post-embedding) input, which is
        left in tw.ws["g"]. Returns {deep prompt index: gradient [n_ctx, D]}."""
        p, G, n, nd = self.p, self.g, self.n, self.J - 1
        ln_grads = self.trainable == "reference"
        ws, D = tw.ws, tw.D
        d16 = self._buf(tw.name + ".dfeat16", (R, self.E), BF16)
        ops.cast_bf16(dfeat, d16)
        dy = self._buf(tw.name + ".dy", (R, D), F32)
        ops.gemm(d16, proj, out_f32=dy)
        self._ln_bwd(tw, dy, xs, stat[0], stat[1], p[lnname + ".weight"], g_out=ws["g_r"], g_out_bf16=ws["g16_r"],
                          dgamma=G[lnname + ".weight"] if ln_grads else None,
                          dbeta=G[lnname + ".bias"] if ln_grads else None, M=R)
        self._last_block_bwd_rows(tw, rows)
        # Deep-prompt splices: ln_1 backward of a spliced layer diverts the gradient of the prompt rows to
        # gprompt[l - 1] ([N, n, D]) and zeroes them in the stream; one batched launch then sums all layers over the
        # batch (batch order, fp16-rounded like autograd through `.half()`).
        ns = max(0, min(nd, tw.L - 1))                       # spliced layers 1 .. ns
        gp = self._buf(f"{tw.name}.gprompt", (max(ns, 1), tw.N * n, D), F32)
        dp_all = self._dprompt_all(tw)
        for l in reversed(range(tw.L)):
            sg = (gp[l - 1], tw.T, deep_row0, n) if 1 <= l <= ns else None
            self._block_bwd(tw, l, attn_only=(l == tw.L - 1), splice_grad=sg)
        if ns > 0:
            ops.prompt_splice_bwd_batched(gp[:ns], dp_all[:ns], tw.N, n, 0, n, True)
        if ns < nd:  # deep prompts beyond the tower depth are never spliced: zero gradient
            dp_all[ns:].zero_()
        got = {i: dp_all[i] for i in range(nd)}
        return got

    # ------------------------------------------------------------------ public: training step
    @torch.no_grad()
    def forward_backward(self, img: torch.Tensor, label: torch.Tensor, loss_out: Optional[torch.Tensor] = None,
                         precision: Optional[str] = None):
        """Forward + backward of one batch. Fills ``self.g[name]`` (fp32 arena) for every trainable tensor and
        returns (loss[1], logits[B,C]) device tensors. No host synchronisation.
        ``precision`` (default: ``self.train_precision``, "bf16"): "fp32" runs the split-operand fp32 training mode
        (forward_backward_f32; cfg PREC = "fp32" of the reference, trainers/maple.py:438-439)."""
        assert img.is_cuda and img.dtype == F32 and label.is_cuda and label.dtype == torch.int64
        if (precision or self.train_precision) == "fp32":
            return self.forward_backward_f32(img, label, loss_out)
        B, C, n, nd = img.shape[0], self.C, self.n, self.J - 1
        p, G = self.p, self.g
        self._text_cache_valid = False
        for tw in (self.vis, self.txt):
            tw.ln_slot, tw.ln_pending = 0, []
        # The two towers are independent until the logits head: the (small) text tower runs on a side stream so
        # its latency-bound kernels fill the gaps of the vision tower; inside a CUDA graph this becomes two
        # parallel branches.
        main = torch.cuda.current_stream()
        side = self._side_stream()
        # the prompt learner's grouped projections head the side branch: the patch-embedding GEMM of the vision tower
        # does not depend on them, only the assembly of the token sequence (shared_ctx) does
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._prompt_learner_fwd()
            self._pl_event = torch.cuda.Event()
            self._pl_event.record(side)
            ft, txs, tstat = self._text_features(True)
        fi, vxs, vstat = self._image_features(img, True, wait_before_assemble=self._pl_event)
        main.wait_stream(side)
        logits = self._buf("head.logits", (B, C), F32)
        loss = loss_out if loss_out is not None else self._buf("head.loss", (1,), F32)
        dfi, dft = self._buf("head.dfi", (B, self.E), F32), self._buf("head.dft", (C, self.E), F32)
        hws = self._buf("head.ws", (ops.head_workspace_floats(B, C, self.E),), F32)
        ops.head_forward_backward(fi, ft, self.logit_scale, label, logits, loss, dfi, dft, hws)
        ln_grads = self.trainable == "reference"

        side.wait_stream(main)
        with torch.cuda.stream(side):
            dt = self._tower_bwd(self.txt, dft, self.tproj, txs, tstat, self.txt_rows, "text_encoder.ln_final", C, 1)
            ops.prompt_splice_bwd(self.txt.ws["g"], None, self.d_ctx_t, C, self.Te, 1, n, False, False)
            self._ln_reduce(self.txt)
        dv = self._tower_bwd(self.vis, dfi, self.vproj, vxs, vstat, self.cls_rows, "image_encoder.ln_post", B,
                             self.Tv - n)
        vws = self.vis.ws
        self._ln_bwd(self.vis, vws["g"], self.vx0, self.vstat0[0], self.vstat0[1], p["image_encoder.ln_pre.weight"],
                          g_out=vws["g"], dgamma=G["image_encoder.ln_pre.weight"] if ln_grads else None,
                          dbeta=G["image_encoder.ln_pre.bias"] if ln_grads else None)
        ops.prompt_splice_bwd(vws["g"], None, self.d_shared, B, self.Tv, self.Tv - n, n, True, False)
        # tail: the grouped dgamma / dbeta reduction of the vision tower (20 us, small CTAs) runs on the by now idle side
        # branch next to the prompt learner's backward instead of in front of it; both join before the optimiser
        main.wait_stream(side)              # text backward (and its reduction) done: d(ctx), d(deep text prompts)
        side.wait_stream(main)              # every vision LayerNorm backward has left its partials
        with torch.cuda.stream(side):
            self._ln_reduce(self.vis)

        # ---- prompt learner backward: dW, db and dx = dy W + d(prompt) for all projections in
        # two grouped launches (the problem table was built with the forward one)
        ops.linear_small_bwd_grouped(self._pl_table, n, max(self.vis.D, self.txt.D), max(self.vis.D, self.txt.D))
        main.wait_stream(side)
        self.last = dict(image_features=fi, text_features=ft, dfi=dfi, dft=dft)
        return loss, logits

    # ------------------------------------------------------------------ optimizer
    @torch.no_grad()
    def sgd_step(self, lr: float, momentum: float = 0.9, weight_decay: float = 5e-4, dampening: float = 0.0,
                 nesterov: bool = False, max_norm: float = 1.0, hyper: Optional[torch.Tensor] = None,
                 loss_dev: Optional[torch.Tensor] = None, flag_dev: Optional[torch.Tensor] = None):
        """clip_grad_norm_(max_norm) over all gradients + SGD on the updated region of the arena
        (trainers/maple.py:592-598), then refresh bf16 copies of trainable block weights."""
        n = self.n_update
        ws = self._buf("opt.ws", (296,), F32)
        norm = self._buf("opt.norm", (1,), F32)
        if hyper is None:
            hyper = torch.tensor([lr, momentum, dampening, weight_decay, max_norm, float(nesterov),
                                  0.0 if self.mom_initialized else 1.0], device=self.dev, dtype=F32)
        ops.grad_norm(self.grads[:n], ws, norm)
        ops.sgd_step(self.params[:n], self.grads[:n], self.momentum[:n], hyper, norm, n, loss_dev, flag_dev)
        self.mom_initialized = True
        self.repack_trainable()
        return norm

    def reset_optimizer_state(self):
        """broadcast_weights deletes every optimizer state entry (trainers/maple_fed.py:332-335)."""
        self.momentum.zero_()
        self.mom_initialized = False

    def round_frozen_to_fp16(self):
        """The reference's FedAvg casts EVERY averaged tensor to fp16 (trainers/maple_fed.py:314), so after the
        first aggregation its fp32 frozen tensors (class/positional embeddings, logit_scale) hold fp16-rounded
        values. Idempotent; shared tensors of co-located clients are rounded once."""
        for t in (self.cls, self.vpos, self.tpos, self.logit_scale):
            t.copy_(t.half().float())
        self._text_cache_valid = False

    # ------------------------------------------------------------------ state exchange
    def trainable_state(self) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((k, v) for k, v in self.p.items())

    def load_trainable(self, tensors: Dict[str, torch.Tensor]):
        for k, v in tensors.items():
            if k in self.p:
                self.p[k].copy_(v.to(self.dev, F32))
        self.repack_trainable()

    @torch.no_grad()
    def reload_state_dict(self, state_dict: Dict[str, torch.Tensor]):
        """CustomCLIP.load_state_dict(strict=True) replaces EVERY tensor (trainers/maple_fed.py:327-331), not only the
        trainable ones: refresh, in place, the arena, logit_scale, the prompt prefix / suffix embeddings and every
        packed frozen tensor (bf16 + transposed block weights, biases, conv1, class / positional embeddings, the two
        projections). Addresses are kept, so captured CUDA graphs stay valid; the text-feature cache and the fp32-mode
        split-weight caches are dropped. Frozen packs are shared by co-located clients (share_from): loading the same
        global state_dict into each of them, as broadcast_weights does, rewrites the shared copy with equal values."""
        sd = {k: v for k, v in state_dict.items() if not k.startswith("clip_model2.")}
        dev = self.dev
        f32 = lambda k: sd[k].detach().to(dev, F32)
        for k, dst in self.p.items():
            if k in sd:
                dst.copy_(f32(k).view(dst.shape))
        self.logit_scale.copy_(f32("logit_scale").reshape(1))
        self.prefix.copy_(f32("prompt_learner.token_prefix"))
        self.suffix.copy_(f32("prompt_learner.token_suffix"))
        for tw in (self.vis, self.txt):
            for l in range(tw.L):
                pre = f"{tw.name}.transformer.resblocks.{l}."
                w = tw.w[l]
                for lin in _LIN:
                    w.pop(lin + ".w3", None)
                    w.pop(lin + ".w3T", None)
                    if lin + ".master" in w:
                        continue  # lives in the arena (copied above); bf16 copies come from repack_trainable()
                    wk, bk = _lin_keys(lin)
                    W = f32(pre + wk)
                    w[lin + ".w"].copy_(W)
                    w[lin + ".wT"].copy_(W.t())
                    w[lin + ".b"].copy_(f32(pre + bk))
        v, t = "image_encoder.", "text_encoder."
        conv = f32(v + "conv1.weight")
        self.conv_w.copy_(conv.reshape(conv.shape[0], -1))
        self.cls.copy_(f32(v + "class_embedding"))
        self.vpos.copy_(f32(v + "positional_embedding"))
        proj = f32(v + "proj")
        self.vproj.copy_(proj)
        self.vproj_T.copy_(self._split_b(proj.t()))
        self.tpos.copy_(f32(t + "positional_embedding"))
        tp = f32(t + "text_projection")
        self.tproj.copy_(tp)
        self.tproj_T.copy_(self._split_b(tp.t()))
        self._conv_w3 = None
        self._proj3 = None
        self._sd_src.update(sd)  # references (fp32 mode re-reads the exact frozen weights)
        self.repack_trainable()  # also invalidates the text-feature cache

    def flops_per_step(self, B: int) -> float:
        """Algorithmic FLOPs of one fwd+bwd step for the work actually performed."""
        def tower(D, L, T, N, causal, wg):
            lin = 2 * N * T * D * (3 * D + D + 4 * D + 4 * D)            # one full layer, forward
            att = 4 * N * T * T * D * ((T + 1) / (2 * T) if causal else 1.0)
            # last layer: in-proj on all rows; attention for the ONE consumed query row of each sequence (att / T);
            # out-proj + MLP only on the N consumed rows
            last = 2 * N * T * D * 3 * D + 2 * N * D * (D + 4 * D + 4 * D)
            att1 = att / T
            fwd = (L - 1) * (lin + att) + last + att1
            bwd = (L - 1) * (lin + 2 * att) + last + 2 * att1 + (last if wg else 0)
            return fwd, bwd
        vf, vb = tower(self.vis.D, self.vis.L, self.Tv, B, False, self.wgrad_last)
        tf, tb = tower(self.txt.D, self.txt.L, self.Te, self.C, True, self.wgrad_last)
        patch = 2 * B * self.P * self.vis.D * 3 * self.patch * self.patch
        heads = 2 * 2 * (B * self.vis.D * self.E + self.C * self.txt.D * self.E)
        return vf + vb + tf + tb + patch + heads
