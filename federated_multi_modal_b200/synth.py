"""Deterministic synthetic inputs for the MaPLe hot path.

There is no network on the build or GPU boxes, so the CLIP ViT-B/16 checkpoint,
the BPE vocabulary and the datasets of the reference are replaced by seeded
synthetic stand-ins with the same shapes/dtypes:

* ``random_clip_state_dict``  -> a state_dict accepted by the reference's
  ``clip.model.build_model(state_dict, design_details)`` (clip/model.py:750-793)
  and by our own ``build_model``; Linear/conv/proj tensors are fp16 (the
  "fp16 round trip" of a real checkpoint, clip/model.py:726-747), LN /
  embeddings fp32.
* ``synthetic_tokenize``      -> replaces ``clip.tokenize`` (clip/clip.py:185-221)
  with distinct class-token ids; SOT=49406, EOT=49407 (EOT is the arg-max id,
  which trainers/maple.py:76 relies on).
* ``make_batch``              -> seeded images / labels.

Everything is generated on the CPU with an explicit ``torch.Generator`` so the
same bytes are produced here (fixtures) and on the GPU box (tests / bench).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import List, Sequence, Union

import torch

SOT, EOT = 49406, 49407
VOCAB, CTX_LEN = 49408, 77

# ViT-B/16 CLIP geometry (clip/model.py:575-647 as instantiated by build_model)
VIT_B16 = dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768,
               vision_patch_size=16, context_length=77, vocab_size=VOCAB,
               transformer_width=512, transformer_heads=8, transformer_layers=12)


def _word_id(word: str) -> int:
    """Stable id in [1000, 49000) for a word (stand-in for BPE)."""
    return 1000 + (zlib.crc32(word.encode("utf-8")) % 48000)


def synthetic_encode(text: str) -> List[int]:
    text = text.replace(".", " .").lower().split()
    return [_word_id(w) for w in text]


def synthetic_tokenize(texts: Union[str, Sequence[str]], context_length: int = CTX_LEN) -> torch.Tensor:
    """Same contract as clip.tokenize: int64 [n, 77], SOT ... EOT, zero padded."""
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, t in enumerate(texts):
        toks = [SOT] + synthetic_encode(t) + [EOT]
        if len(toks) > context_length:
            raise RuntimeError(f"Input {t} is too long for context length {context_length}")
        out[i, :len(toks)] = torch.tensor(toks)
    return out


def synthetic_classnames(n: int) -> List[str]:
    """Class names of 1-3 words so the EOT position varies (8..10) like real prompts."""
    base = ["forest", "river", "harbor", "runway", "airport terminal", "dense residential",
            "golf course", "parking lot", "storage tank", "sparse residential area"]
    names = []
    for i in range(n):
        b = base[i % len(base)]
        names.append(b if i < len(base) else f"{b} {i}")
    return names


def _randn(g, *shape, std=1.0):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std


def random_clip_state_dict(seed: int = 0, layers: int = 12, dims: dict | None = None) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random CLIP state_dict in checkpoint layout (keys of clip/model.py CLIP).

    LN gains/biases and Linear biases are deliberately non-trivial so that parity
    tests exercise them.
    """
    d = dict(VIT_B16)
    if dims:
        d.update(dims)
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    vw, tw, e = d["vision_width"], d["transformer_width"], d["embed_dim"]
    ps = d["vision_patch_size"]
    grid = d["image_resolution"] // ps
    vl = d["vision_layers"] if layers is None else min(layers, d["vision_layers"])
    tl = d["transformer_layers"] if layers is None else min(layers, d["transformer_layers"])

    def ln(prefix, width):
        sd[prefix + ".weight"] = 1.0 + _randn(g, width, std=0.1)
        sd[prefix + ".bias"] = _randn(g, width, std=0.05)

    def block(prefix, width, nl):
        attn_std = width ** -0.5
        proj_std = (width ** -0.5) * ((2 * nl) ** -0.5)
        fc_std = (2 * width) ** -0.5
        sd[prefix + ".attn.in_proj_weight"] = _randn(g, 3 * width, width, std=attn_std).half()
        sd[prefix + ".attn.in_proj_bias"] = _randn(g, 3 * width, std=0.02).half()
        sd[prefix + ".attn.out_proj.weight"] = _randn(g, width, width, std=proj_std).half()
        sd[prefix + ".attn.out_proj.bias"] = _randn(g, width, std=0.02).half()
        ln(prefix + ".ln_1", width)
        sd[prefix + ".mlp.c_fc.weight"] = _randn(g, 4 * width, width, std=fc_std).half()
        sd[prefix + ".mlp.c_fc.bias"] = _randn(g, 4 * width, std=0.02).half()
        sd[prefix + ".mlp.c_proj.weight"] = _randn(g, width, 4 * width, std=proj_std).half()
        sd[prefix + ".mlp.c_proj.bias"] = _randn(g, width, std=0.02).half()
        ln(prefix + ".ln_2", width)

    scale = vw ** -0.5
    sd["visual.class_embedding"] = _randn(g, vw, std=scale)
    sd["visual.positional_embedding"] = _randn(g, grid * grid + 1, vw, std=scale)
    sd["visual.conv1.weight"] = _randn(g, vw, 3, ps, ps, std=(3 * ps * ps) ** -0.5).half()
    ln("visual.ln_pre", vw)
    for i in range(vl):
        block(f"visual.transformer.resblocks.{i}", vw, vl)
    ln("visual.ln_post", vw)
    sd["visual.proj"] = _randn(g, vw, e, std=scale).half()

    sd["token_embedding.weight"] = _randn(g, d["vocab_size"], tw, std=0.02)
    sd["positional_embedding"] = _randn(g, d["context_length"], tw, std=0.01)
    for i in range(tl):
        block(f"transformer.resblocks.{i}", tw, tl)
    ln("ln_final", tw)
    sd["text_projection"] = _randn(g, tw, e, std=tw ** -0.5).half()
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07), dtype=torch.float32)
    return sd


def make_batch(batch: int, n_cls: int, seed: int = 123, size: int = 224):
    """Seeded images [B,3,S,S] fp32 and labels [B] int64."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(batch, 3, size, size, generator=g, dtype=torch.float32)
    lab = torch.randint(0, n_cls, (batch,), generator=g, dtype=torch.long)
    return img, lab


class _NS:
    """Attribute namespace standing in for a yacs CfgNode."""
    def __init__(self, **kw):
        self.__dict__.update(kw)


def make_cfg(n_ctx: int = 2, depth: int = 9, prec: str = "fp16", ctx_init: str = "a photo of a", size: int = 224):
    """cfg with the keys the hot path reads (train.py:109-113)."""
    return _NS(TRAINER=_NS(MAPLE=_NS(N_CTX=n_ctx, CTX_INIT=ctx_init, PREC=prec, PROMPT_DEPTH=depth)),
               INPUT=_NS(SIZE=(size, size)),
               MODEL=_NS(BACKBONE=_NS(NAME="ViT-B/16"), INIT_WEIGHTS=""),
               OPTIM=_NS(NAME="sgd", LR=0.0026, MAX_EPOCH=2, LR_SCHEDULER="cosine", WARMUP_EPOCH=1,
                         WARMUP_TYPE="constant", WARMUP_CONS_LR=1e-4, MOMENTUM=0.9, WEIGHT_DECAY=5e-4,
                         SGD_DAMPNING=0.0, SGD_NESTEROV=False),
               FED=_NS(NUM_CLIENTS=2, NUM_ROUNDS=30, LOCAL_EPOCHS=10))


def random_prompt_learner_state(seed: int = 1, n_ctx: int = 2, depth: int = 9, ctx_dim: int = 512,
                                vis_dim: int = 768) -> "OrderedDict[str, torch.Tensor]":
    """Seeded values for every *randomly initialised* prompt-learner tensor
    (trainers/maple.py:111-131): deep prompts N(0, 0.02); Linear layers with
    torch's default U(-1/sqrt(in), 1/sqrt(in)) range. ``ctx`` is NOT included
    when CTX_INIT is used (it comes from the token embedding, maple.py:96-102).
    dtypes follow the reference in fp16 mode (Appendix A of SURVEY.md).
    """
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def linear(prefix, fin, fout, dtype):
        bound = 1.0 / math.sqrt(fin)
        sd[prefix + ".weight"] = ((torch.rand(fout, fin, generator=g) * 2 - 1) * bound).to(dtype)
        sd[prefix + ".bias"] = ((torch.rand(fout, generator=g) * 2 - 1) * bound).to(dtype)

    linear("proj_lang_to_vis", ctx_dim, vis_dim, torch.float16)
    linear("proj_vis_to_lang", vis_dim, ctx_dim, torch.float16)
    n_text = len([i for i in range(depth - 1) if i % 2 == 0])
    n_vis = len([i for i in range(depth - 1) if i % 2 != 0])
    for i in range(n_text):
        sd[f"compound_prompts_text_parameters.{i}"] = _randn(g, n_ctx, ctx_dim, std=0.02)
    for i in range(n_vis):
        sd[f"visual_deep_prompts_parameters.{i}"] = _randn(g, n_ctx, vis_dim, std=0.02)
    for i in range(depth - 1):
        if i % 2 == 0:
            linear(f"compound_prompt_projections.{i}", ctx_dim, vis_dim, torch.float32)
        else:
            linear(f"compound_prompt_projections.{i}", vis_dim, ctx_dim, torch.float32)
    return sd
