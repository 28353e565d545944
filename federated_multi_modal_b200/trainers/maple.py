"""Drop-in mirror of the reference's ``trainers/maple.py`` hot path on the B200 engine.

Same names / signatures / state_dict keys as the reference:
  load_clip_to_cpu(cfg)                                   trainers/maple.py:21-40
  TextEncoder(clip_model).forward(prompts, tok, deep)     trainers/maple.py:43-79
  MultiModalPromptLearner(cfg, classnames, clip_model)    trainers/maple.py:81-218
  CustomCLIP(cfg, classnames, clip_model).forward(image, label=None, caption=None, return_feature=False)
                                                          trainers/maple.py:220-381
  MaPLe(TrainerX): check_cfg, build_model, parse_batch_train, forward_backward, run_epoch,
                   update_lr, test, load_model             trainers/maple.py:384-716

``CustomCLIP.forward`` in training mode returns a scalar loss that participates in autograd: its
backward hands the gradients the engine already computed to the ``nn.Parameter``s, so
``loss.backward(); optim.step()`` with any torch optimiser works as in the reference.
``MaPLe.forward_backward`` takes the faster route (engine gradients -> fused clip+SGD kernel,
optionally replayed from a CUDA graph) with one host synchronisation per step instead of ~165.
"""
from __future__ import annotations

import math
import os
import os.path as osp
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import ops, synth
from ..clip import model as clip_model_mod
from ..dassl_compat import (TRAINER_REGISTRY, TrainerX, build_lr_scheduler, load_checkpoint)
from ..engine import MapleEngine

F32 = torch.float32


def load_clip_to_cpu(cfg):
    """Reference: downloads the OpenAI checkpoint (trainers/maple.py:21-40). Offline: a checkpoint path in
    ``cfg.MODEL.CLIP_CHECKPOINT`` (a CLIP state_dict), else the seeded synthetic ViT-B/16 of ``synth``."""
    design_details = {"trainer": "MaPLe", "vision_depth": 0, "language_depth": 0, "vision_ctx": 0,
                      "language_ctx": 0, "maple_length": cfg.TRAINER.MAPLE.N_CTX}
    path = getattr(cfg.MODEL, "CLIP_CHECKPOINT", "")
    if path:
        sd = torch.load(path, map_location="cpu")
        sd = sd.state_dict() if hasattr(sd, "state_dict") else sd
    else:
        sd = synth.random_clip_state_dict(getattr(cfg.MODEL, "SYNTHETIC_SEED", 0))
    return clip_model_mod.build_model(sd, design_details)


class TextEncoder(nn.Module):
    def __init__(self, clip_model):
        super().__init__()
        self.transformer = clip_model.transformer
        self.positional_embedding = clip_model.positional_embedding
        self.ln_final = clip_model.ln_final
        self.text_projection = clip_model.text_projection
        self.dtype = clip_model.dtype

    @torch.no_grad()
    def forward(self, prompts, tokenized_prompts, compound_prompts_deeper_text):
        """Standalone hook (inference): [C,77,D] prompts -> [C,E] text features on the CUDA kernels."""
        x = prompts.detach().to(F32) + self.positional_embedding.detach().to(F32)
        x = x.permute(1, 0, 2)
        x = self.transformer([x, compound_prompts_deeper_text, 0])[0].permute(1, 0, 2).to(F32).contiguous()
        C, T, D = x.shape
        rows = (torch.arange(C, device=x.device) * T + tokenized_prompts.to(x.device).argmax(-1)).to(torch.int32)
        y = torch.empty(C, D, device=x.device, dtype=torch.bfloat16)
        ops.layernorm_fwd(x.reshape(C * T, D), self.ln_final.weight.detach().float(),
                          self.ln_final.bias.detach().float(), rowidx=rows.contiguous(), y_bf16=y, M=C)
        feat = torch.empty(C, self.text_projection.shape[1], device=x.device, dtype=F32)
        ops.gemm(y, self.text_projection.detach().t().to(torch.bfloat16).contiguous(), out_f32=feat)
        return feat.to(self.dtype)


class MultiModalPromptLearner(nn.Module):
    def __init__(self, cfg, classnames, clip_model):
        super().__init__()
        n_cls = len(classnames)
        n_ctx = cfg.TRAINER.MAPLE.N_CTX
        ctx_init = cfg.TRAINER.MAPLE.CTX_INIT
        dtype = clip_model.dtype
        ctx_dim = clip_model.ln_final.weight.shape[0]
        vis_dim = clip_model.visual.conv1.weight.shape[0]
        assert cfg.TRAINER.MAPLE.PROMPT_DEPTH >= 1, "For MaPLe, PROMPT_DEPTH should be >= 1"
        self.compound_prompts_depth = depth = cfg.TRAINER.MAPLE.PROMPT_DEPTH
        assert cfg.INPUT.SIZE[0] == clip_model.visual.input_resolution

        if ctx_init and n_ctx <= 4:  # trainers/maple.py:96-102
            ctx_init = ctx_init.replace("_", " ")
            with torch.no_grad():
                emb = clip_model.token_embedding(synth.synthetic_tokenize(ctx_init)).type(dtype)
            ctx_vectors = emb[0, 1:1 + n_ctx, :]
            prompt_prefix = ctx_init
        else:
            ctx_vectors = torch.empty(n_ctx, ctx_dim, dtype=dtype)
            nn.init.normal_(ctx_vectors, std=0.02)
            prompt_prefix = " ".join(["X"] * n_ctx)
        self.proj_lang_to_vis = nn.Linear(ctx_dim, vis_dim).half()
        self.proj_vis_to_lang = nn.Linear(vis_dim, ctx_dim).half()  # defined, never used (as in the reference)
        self.ctx = nn.Parameter(ctx_vectors.clone())
        self.compound_prompts_text_parameters = nn.ParameterList(
            [nn.Parameter(torch.empty(n_ctx, ctx_dim)) for i in range(depth - 1) if i % 2 == 0])
        self.visual_deep_prompts_parameters = nn.ParameterList(
            [nn.Parameter(torch.empty(n_ctx, vis_dim)) for i in range(depth - 1) if i % 2 != 0])
        for p in list(self.compound_prompts_text_parameters) + list(self.visual_deep_prompts_parameters):
            nn.init.normal_(p, std=0.02)
        self.compound_prompt_projections = nn.ModuleList(
            [nn.Linear(ctx_dim, vis_dim) if i % 2 == 0 else nn.Linear(vis_dim, ctx_dim) for i in range(depth - 1)])

        classnames = [name.replace("_", " ") for name in classnames]
        self.name_lens = [len(synth.synthetic_encode(name)) for name in classnames]
        prompts = [prompt_prefix + " " + name + "." for name in classnames]
        tokenized_prompts = synth.synthetic_tokenize(prompts)
        with torch.no_grad():
            embedding = clip_model.token_embedding(tokenized_prompts).type(dtype)
        self.register_buffer("token_prefix", embedding[:, :1, :].clone())
        self.register_buffer("token_suffix", embedding[:, 1 + n_ctx:, :].clone())
        self.n_cls, self.n_ctx = n_cls, n_ctx
        self.tokenized_prompts = tokenized_prompts

    def construct_prompts(self, ctx, prefix, suffix, label=None):
        if label is not None:
            prefix, suffix = prefix[label], suffix[label]
        return torch.cat([prefix, ctx, suffix], dim=1)

    @torch.no_grad()
    def forward(self):
        """-> (prompts [C,77,512], shared_ctx [n,768], deep_text list, deep_vis list), trainers/maple.py:177-218.
        The nine tiny projections run on the fp32 small-linear kernel."""
        ctx = self.ctx
        prompts = self.construct_prompts(ctx.unsqueeze(0).expand(self.n_cls, -1, -1), self.token_prefix,
                                         self.token_suffix)

        def lin(layer, x):
            y = torch.empty(x.shape[0], layer.weight.shape[0], device=x.device, dtype=F32)
            ops.linear_small_fwd(x.detach().float().contiguous(), layer.weight.detach().float().contiguous(),
                                 layer.bias.detach().float().contiguous(), y)
            return y

        deep_text: List[torch.Tensor] = []
        deep_vis: List[torch.Tensor] = []
        for i, layer in enumerate(self.compound_prompt_projections):
            if i % 2 == 0:
                t = self.compound_prompts_text_parameters[i // 2]
                deep_vis.append(lin(layer, t))
                deep_text.append(t)
            else:
                v = self.visual_deep_prompts_parameters[(i - 1) // 2]
                deep_text.append(lin(layer, v))
                deep_vis.append(v)
        shared = lin(self.proj_lang_to_vis, ctx).to(ctx.dtype)
        return prompts, shared, deep_text, deep_vis


class _EngineStep(torch.autograd.Function):
    """loss = engine.forward_backward(...); backward hands the already-computed gradients to autograd."""

    @staticmethod
    def forward(ctx, model, image, label, *params):
        loss, _ = model.engine.forward_backward(image, label)
        ctx.model = model
        return loss.reshape(()).clone()

    @staticmethod
    def backward(ctx, gout):
        model = ctx.model
        grads = []
        for name, p in model._grad_params:
            g = model.engine.g.get(name)
            grads.append(None if g is None or "proj_vis_to_lang" in name else (g * gout).to(p.dtype))
        return (None, None, None, *grads)


class CustomCLIP(nn.Module):
    def __init__(self, cfg, classnames, clip_model):
        super().__init__()
        self.prompt_learner = MultiModalPromptLearner(cfg, classnames, clip_model)
        self.tokenized_prompts = self.prompt_learner.tokenized_prompts
        self.image_encoder = clip_model.visual
        self.text_encoder = TextEncoder(clip_model)
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))  # new, frozen by the trainer
        self.dtype = clip_model.dtype
        self.clip_model2 = clip_model
        self._cfg = cfg
        self.engine: Optional[MapleEngine] = None
        self._versions = None
        self._arena_newer = False
        self.text_truncate = True

    # -------------------------------------------------------------- engine plumbing
    def _named_unique_params(self):
        return [(n, p) for n, p in self.named_parameters() if not n.startswith("clip_model2.")]

    def build_engine(self, share_from: Optional[MapleEngine] = None) -> MapleEngine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("CustomCLIP: move the model to a CUDA device first — the MaPLe hot path runs only "
                               "on libmfk CUDA kernels (no CPU fallback)")
        named = dict(self._named_unique_params())
        ref_set = any(p.requires_grad for n, p in named.items()
                      if not n.startswith("prompt_learner.") and n != "logit_scale")
        sd = {k: v for k, v in nn.Module.state_dict(self).items()}
        self.engine = MapleEngine(sd, self.tokenized_prompts, n_ctx=self.prompt_learner.n_ctx,
                                  depth=self.prompt_learner.compound_prompts_depth, device=str(dev),
                                  trainable="reference" if ref_set else "prompt_only",
                                  text_truncate=self.text_truncate, share_from=share_from)
        # PREC == "fp32": the reference trains the fp32 model (trainers/maple.py:438-439) -> fp32 training mode
        if str(getattr(self._cfg.TRAINER.MAPLE, "PREC", "bf16")) == "fp32":
            self.engine.train_precision = "fp32"
        self._grad_params = [(n, p) for n, p in named.items() if n in self.engine.p and p.requires_grad]
        self._versions = self._param_versions()
        return self.engine

    def _param_versions(self):
        return tuple(p._version for _, p in self._grad_params)

    def _sync_to_engine(self):
        if self.engine is None:
            self.build_engine()
            return
        if self._arena_newer:
            return  # engine holds the newest values (fused optimiser path)
        v = self._param_versions()
        if v != self._versions:
            self.engine.load_trainable({n: p.detach() for n, p in self._grad_params})
            self._versions = v

    def sync_from_engine(self):
        """Write the engine's fp32 master parameters back into the nn.Parameters (after fused SGD steps)."""
        if self.engine is not None and self._arena_newer:
            with torch.no_grad():
                for n, p in self._grad_params:
                    p.copy_(self.engine.p[n].to(p.dtype))
            self._versions = self._param_versions()
            self._arena_newer = False

    def state_dict(self, *args, **kwargs):
        self.sync_from_engine()
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._arena_newer = False
        if self.engine is not None:
            # every key reaches the kernels, not only the tensors with requires_grad: logit_scale, prompt
            # prefix / suffix, embeddings, frozen block weights, and in prompt_only mode the LN / resblocks.11 weights
            self.engine.reload_state_dict(nn.Module.state_dict(self))
            self._versions = self._param_versions()
        return out

    # -------------------------------------------------------------- forward
    def forward(self, image, label=None, caption=None, return_feature=False):
        if caption is not None and isinstance(caption, list) and len(caption) > 0 and any(c is not None for c in caption):
            raise NotImplementedError("caption branch (trainers/maple.py:307-322, clip/model.py:550-561) is out of "
                                      "scope: it re-draws random weights on every forward (SURVEY.md §2 #11)")
        self._sync_to_engine()
        eng = self.engine
        image = image.to(eng.dev, F32).contiguous()
        if self.training:
            if label is None:
                raise ValueError("CustomCLIP.forward in training mode needs integer labels")
            if label.dtype.is_floating_point:
                raise NotImplementedError("soft-label KL branch (trainers/maple.py:356-360) is out of scope")
            label = label.to(eng.dev, torch.int64).contiguous()
            assert int(label.max()) < eng.C, "Label index out of bounds"  # trainers/maple.py:353
            params = [p for _, p in self._grad_params]
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                loss = _EngineStep.apply(self, image, label, *params)
            else:
                loss = eng.forward_backward(image, label)[0].reshape(()).clone()
            if not bool(torch.isfinite(loss)):
                raise RuntimeError("NaN/Inf in total loss")  # trainers/maple.py:375-376
            return loss
        # PREC == "fp32" (trainers/maple.py:438-439 calls clip_model.float()): inference AND training run the engine's
        # fp32 mode (bf16x3 split-operand GEMMs, fp32 everything else; logits within 1e-3 of the reference's fp32
        # path, gradients within 1e-3 of its fp32 autograd). Every other PREC uses the bf16 tensor-core path with
        # fp32 master weights.
        prec = "fp32" if str(getattr(self._cfg.TRAINER.MAPLE, "PREC", "bf16")) == "fp32" else "bf16"
        logits = eng.logits(image, precision=prec)
        if return_feature:  # kept for signature compatibility with upstream MaPLe
            return logits, eng.last_image_features()
        return logits


def _opt(cfg_optim, name, default):
    return getattr(cfg_optim, name, default)


class _FusedSGD:
    """Holder for the hyper-parameters of the fused clip_grad_norm_ + SGD kernel (Dassl build_optimizer,
    trainers/maple.py:498: sgd, momentum 0.9, weight decay 5e-4, dampening 0, no nesterov by default)."""

    def __init__(self, cfg_optim):
        name = str(_opt(cfg_optim, "NAME", "sgd")).lower()
        if name != "sgd":
            raise NotImplementedError(f"fused optimiser supports sgd only (got {name}); the reference yaml uses sgd")
        self.lr = float(cfg_optim.LR)
        self.momentum = float(_opt(cfg_optim, "MOMENTUM", 0.9))
        self.weight_decay = float(_opt(cfg_optim, "WEIGHT_DECAY", 5e-4))
        self.dampening = float(_opt(cfg_optim, "SGD_DAMPNING", 0.0))
        self.nesterov = bool(_opt(cfg_optim, "SGD_NESTEROV", False))
        self.param_groups = [self.__dict__]  # so `optim.param_groups[0]['lr']` reads as in the reference
        self.state: Dict = {}


@TRAINER_REGISTRY.register()
class MaPLe(TrainerX):
    def __init__(self, cfg, client_id=None, classnames=None, share_engine: Optional[MapleEngine] = None,
                 use_cuda_graph: Optional[bool] = None):
        self.cfg = cfg
        self.client_id = client_id
        self.nan_count = 0
        self.total_batches = 0
        self.classnames = classnames
        self.lr_history: List[float] = []
        self.grad_norms: List[float] = []
        self._share_engine = share_engine
        self._graph = None
        self._use_graph = use_cuda_graph if use_cuda_graph is not None else bool(getattr(cfg, "USE_CUDA_GRAPH", True))
        super().__init__(cfg)

    def check_cfg(self, cfg):
        assert cfg.TRAINER.MAPLE.PREC in ["fp16", "fp32", "amp", "bf16"], cfg.TRAINER.MAPLE.PREC

    def build_model(self):
        cfg = self.cfg
        if self.classnames is None:
            self.classnames = self.dm.dataset.classnames
        clip_model = load_clip_to_cpu(cfg)
        self.model = CustomCLIP(cfg, self.classnames, clip_model)
        # freeze policy of trainers/maple.py:447-479: LayerNorms, prompt learner and every parameter whose name
        # contains "transformer.resblocks.11" stay trainable; cfg.TRAINER.MAPLE.TRAINABLE="prompt_only" gives the
        # upstream-MaPLe set named by north_star.
        prompt_only = getattr(cfg.TRAINER.MAPLE, "TRAINABLE", "reference") == "prompt_only"
        for p in self.model.parameters():
            p.requires_grad_(False)
        for n, p in self.model.named_parameters():
            if "prompt_learner" in n:
                p.requires_grad_(True)
        if not prompt_only:
            for _, m in self.model.named_modules():
                if isinstance(m, (nn.LayerNorm, nn.BatchNorm1d, nn.BatchNorm2d)):
                    for p in m.parameters():
                        p.requires_grad_(True)
            for n, p in self.model.named_parameters():
                if "transformer.resblocks.11" in n:
                    p.requires_grad_(True)
        self.model.to(self.device)
        self.model.build_engine(share_from=self._share_engine)
        if self.model.engine.train_precision == "fp32":
            self._use_graph = False  # the fp32 parity mode runs eagerly (it rebuilds split weights with torch ops)
        self.optim = _FusedSGD(cfg.OPTIM)
        self.sched = build_lr_scheduler(self.optim, cfg.OPTIM)
        key = f"MultiModalPromptLearner_{self.client_id}"
        if key not in self._models:
            self.register_model(key, self.model, self.optim, self.sched)
        self.scaler = None
        self.lr_history = [self.optim.lr]

    # ------------------------------------------------------------------ step
    def _prefetch_batch(self, batch):
        """H2D copy of an assembled batch on a side stream (run_epoch: while the previous step computes) into one of
        two persistent device staging sets. The batch dict then carries device tensors and the event
        parse_batch_train waits for. A set is rewritten two batches later, after the step that read it was synced."""
        if not isinstance(batch, dict) or os.environ.get("MFK_NO_PREFETCH"):
            return batch
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._pf_bufs, self._pf_slot = {}, 0
        self._pf_slot ^= 1
        # The set being rewritten was last read by work enqueued before the PREVIOUS prefetch call (two batches ago):
        # the copy stream waits for the main-stream position recorded then — not for the step enqueued since.
        prev = getattr(self, "_pf_main_ev", None)
        if prev is not None:
            self._copy_stream.wait_event(prev)
        self._pf_main_ev = torch.cuda.Event()
        self._pf_main_ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            for k in ("img", "img_u8", "label"):
                v = batch.get(k)
                if torch.is_tensor(v) and not v.is_cuda:
                    key = (k, self._pf_slot, tuple(v.shape), v.dtype)
                    dst = self._pf_bufs.get(key)
                    if dst is None:
                        dst = self._pf_bufs[key] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    dst.copy_(v, non_blocking=True)
                    batch[k] = dst
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        batch["_h2d_event"] = ev
        return batch

    def parse_batch_train(self, batch):
        if isinstance(batch, dict) and batch.get("_h2d_event") is not None:
            torch.cuda.current_stream().wait_event(batch.pop("_h2d_event"))
        if isinstance(batch, dict) and "img_u8" in batch:
            # GpuAugment loader (client_datamanager.GpuAugment): raw uint8 images + host-drawn crop boxes / flips;
            # random_resized_crop + flip + normalize run on the device
            raw = batch["img_u8"].to(self.device, non_blocking=True)
            x = batch["augment"].apply(raw, batch["rrc_box"], batch["flip"])
            y = batch["label"].to(self.device, non_blocking=True)
            return x, y, batch.get("caption")
        x = batch["img"].to(self.device, non_blocking=True)
        y = batch["label"].to(self.device, non_blocking=True)
        return x, y, batch.get("caption") if isinstance(batch, dict) else None

    def _hyper(self):
        o, eng = self.optim, self.model.engine
        vals = [o.lr, o.momentum, o.dampening, o.weight_decay, 1.0, float(o.nesterov),
                0.0 if eng.mom_initialized else 1.0]
        if getattr(self, "_hyper_vals", None) != vals:
            if getattr(self, "_hyper_dev", None) is None:
                self._hyper_dev = torch.empty(7, device=self.device, dtype=F32)
            self._hyper_dev.copy_(torch.tensor(vals, dtype=F32), non_blocking=False)
            self._hyper_vals = vals
        return self._hyper_dev

    def _step_kernels(self, img, lab):
        eng = self.model.engine
        # input validity scan on the side branch (it only has to finish before the optimiser; forward_backward joins
        # the side stream before the logits head)
        side, cur = eng._side_stream(), torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            ops.check_finite(img, self._flag)
        eng.forward_backward(img, lab, loss_out=self._readback[0:1])
        # device-side guard: a NaN/Inf image or loss skips the update (the reference raises before optim.step())
        norm = eng.sgd_step(0.0, hyper=self._hyper_dev, loss_dev=self._readback[0:1], flag_dev=self._flag)
        self._readback[1:2].copy_(norm)

    def step_async(self, image, label):
        """Enqueue one training step (fwd + bwd + clip + SGD) for device-resident inputs WITHOUT any host
        synchronisation; loss / grad-norm / validity land in device buffers read by ``read_step_result``."""
        model, eng = self.model, self.model.engine
        model._sync_to_engine()
        B = image.shape[0]
        if getattr(self, "_readback", None) is None:
            self._readback = torch.zeros(2, device=self.device, dtype=F32)
            self._flag = torch.zeros(1, device=self.device, dtype=torch.int32)
            self._host = torch.zeros(3, dtype=F32).pin_memory()
        self._hyper()
        if self._use_graph:
            if self._graph is None or self._graph_B != B or self._graph_gen != eng.buffer_generation:
                self._static_img = torch.empty_like(image)
                self._static_lab = torch.empty_like(label)
                self._static_img.copy_(image); self._static_lab.copy_(label)
                # warm-up outside capture (lazy driver entry points, workspace allocation, smem attributes);
                # parameters / momentum are restored afterwards so warm-up and capture are not training steps
                snap = (eng.params.clone(), eng.momentum.clone(), eng.mom_initialized)
                self._step_kernels(self._static_img, self._static_lab)
                torch.cuda.synchronize()
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._step_kernels(self._static_img, self._static_lab)
                eng.params.copy_(snap[0]); eng.momentum.copy_(snap[1]); eng.mom_initialized = snap[2]
                eng.repack_trainable()
                self._flag.zero_()
                self._graph_B = B
                self._graph_gen = eng.buffer_generation
                self._hyper_vals = None
                self._hyper()
            self._static_img.copy_(image, non_blocking=True)
            self._static_lab.copy_(label, non_blocking=True)
            self._graph.replay()
            eng.mom_initialized = True
            # the replayed step moved ctx / deep prompts / LN / resblocks.11: cached text features are stale (the
            # Python flag writes inside forward_backward ran at capture time only)
            eng._text_cache_valid = False
        else:
            self._step_kernels(image, label)
        model._arena_newer = True

    def read_step_result(self):
        """One D2H read + stream sync: (loss, pre-clip grad norm, input validity flag)."""
        self._host[0:2].copy_(self._readback, non_blocking=True)
        flag = int(self._flag.to("cpu"))  # synchronises the stream
        if flag:
            self._flag.zero_()
        return float(self._host[0]), float(self._host[1]), flag

    def forward_backward(self, batch):
        """fwd + bwd + clip_grad_norm_(1.0) + SGD step (trainers/maple.py:547-627) with a single host sync."""
        self._fb_enqueue(batch)
        return self._fb_finish()

    def _fb_enqueue(self, batch):
        """First half of forward_backward: H2D copy of the batch + the whole step, enqueued without a host sync."""
        image, label, caption = self.parse_batch_train(batch)
        if caption is not None and isinstance(caption, list) and any(c is not None for c in caption):
            raise NotImplementedError("caption branch is out of scope (SURVEY.md §2 #11)")
        self.total_batches += 1
        image = image.to(F32).contiguous()
        label = label.to(torch.int64).contiguous()
        self.step_async(image, label)

    def _fb_finish(self):
        """Second half: the one D2H read / stream sync of the step, then the reference's error behaviour."""
        loss, norm, flag = self.read_step_result()
        if flag & 1:
            raise ValueError("NaN values in input image")  # trainers/maple.py:532-535
        if flag & 2:
            raise ValueError("Inf values in input image")
        if not math.isfinite(loss):
            raise RuntimeError("NaN/Inf in total loss")
        if not self.lr_history or self.optim.lr != self.lr_history[-1]:
            self.lr_history.append(self.optim.lr)
        self.grad_norms.append(min(norm, norm / (norm + 1e-6)))  # norm after clipping to 1.0
        return {"loss": loss}

    def run_epoch(self, epoch):
        """trainers/maple.py:629-658. Software-pipelined: while the GPU runs step k the host assembles batch k + 1
        (stack into pinned staging, augmentation parameters), so the loader's host time hides behind the step
        instead of adding to it; results, error behaviour and the order of updates are those of the plain loop."""
        self.model.train()
        total, steps = 0.0, 0
        it = iter(self.dm.train_loader)
        batch = next(it, None)
        batch_idx = 0
        while batch is not None:
            self.batch_idx = batch_idx
            self._fb_enqueue(batch)
            batch = next(it, None)  # host-side assembly of the next batch overlaps the step just enqueued,
            if batch is not None:   # and so does its H2D copy (side stream)
                batch = self._prefetch_batch(batch)
            out = self._fb_finish()
            total += out.get("loss", 0.0)
            steps += 1
            batch_idx += 1
        self.update_lr()
        acc = self.test().get("accuracy", 0) if getattr(self.dm, "test_loader", None) is not None else 0
        avg = total / max(1, steps)
        print(f"[Client {self.client_id}] Epoch {epoch} done. Loss={avg:.4f}, Acc={acc:.2f}%")
        return {"avg_loss": avg}

    def update_lr(self):
        if self.sched is not None:
            self.sched.step()

    @torch.no_grad()
    def test(self, evaluate_train=False):
        """trainers/maple.py:660-681. Pipelined like run_epoch: batch k + 1 is assembled and copied H2D (side stream)
        while batch k is evaluated; hits are counted on the device and read back once per call (the reference
        synchronises per batch). The host never runs more than two batches ahead of the copies (the loader's pinned
        staging rotates over three)."""
        self.model.eval()
        total = 0
        correct = torch.zeros((), device=self.device, dtype=torch.int64)
        it = iter(self.dm.test_loader)
        batch = next(it, None)
        if batch is not None:
            batch = self._prefetch_batch(batch)
        h2d = []
        while batch is not None:
            ev = batch.get("_h2d_event") if isinstance(batch, dict) else None
            x, y, _ = self.parse_batch_train(batch)
            preds = self.model_inference(x).argmax(dim=1)
            correct += (preds == y).sum()
            total += int(y.size(0))
            h2d.append(ev)
            if len(h2d) >= 2 and h2d[-2] is not None:
                h2d[-2].synchronize()
            batch = next(it, None)
            if batch is not None:
                batch = self._prefetch_batch(batch)
        correct = int(correct)
        acc = 100.0 * correct / total if total else 0.0
        print(f"[Client {self.client_id}] Test Accuracy: {acc:.2f}%")
        return {"accuracy": acc}

    def model_inference(self, input):
        return self.model(input)

    def load_model(self, directory, epoch=None):
        if not directory:
            print("Note that load_model() is skipped as no pretrained model is given")
            return
        model_file = "model-best.pth.tar" if epoch is None else f"model.pth.tar-{epoch}"
        for name in self.get_model_names():
            path = osp.join(directory, name, model_file)
            if not osp.exists(path):
                raise FileNotFoundError(f"Model not found at '{path}'")
            ckpt = load_checkpoint(path)
            sd = ckpt["state_dict"]
            for k in ("prompt_learner.token_prefix", "prompt_learner.token_suffix"):
                sd.pop(k, None)  # trainers/maple.py:709-712
            self._models[name].load_state_dict(sd, strict=False)
