"""Shared test helpers: build CustomCLIP-layout state dicts from seeds, compare to golden."""
import os
import sys
from collections import OrderedDict

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from federated_multi_modal_b200 import synth  # noqa: E402

GOLDEN = os.path.join(REPO, "tests", "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def customclip_state_dict(C, seed_clip=0, seed_pl=1, n_ctx=2, depth=9, layers=12):
    """CustomCLIP-layout fp32 state_dict + tokenized prompts, built only from seeds
    (no reference needed). Mirrors trainers/maple.py:96-149 for ctx / prefix / suffix."""
    sd = synth.random_clip_state_dict(seed_clip, layers=layers)
    pl = synth.random_prompt_learner_state(seed_pl, n_ctx=n_ctx, depth=depth)
    names = synth.synthetic_classnames(C)
    out = OrderedDict()
    emb = sd["token_embedding.weight"]
    ctx_tok = synth.synthetic_tokenize("a photo of a")
    out["prompt_learner.ctx"] = emb[ctx_tok[0, 1:1 + n_ctx]].clone()
    prompts = ["a photo of a " + n.replace("_", " ") + "." for n in names]
    tok = synth.synthetic_tokenize(prompts)
    e = emb[tok]
    out["prompt_learner.token_prefix"] = e[:, :1, :].clone()
    out["prompt_learner.token_suffix"] = e[:, 1 + n_ctx:, :].clone()
    for k, v in pl.items():
        out["prompt_learner." + k] = v.float()
    for k, v in sd.items():
        if k.startswith("visual."):
            out["image_encoder." + k[len("visual."):]] = v.float()
        elif k.startswith("transformer."):
            out["text_encoder." + k] = v.float()
        elif k in ("positional_embedding", "text_projection"):
            out["text_encoder." + k] = v.float()
        elif k.startswith("ln_final."):
            out["text_encoder." + k] = v.float()
    out["logit_scale"] = sd["logit_scale"].clone()
    return out, tok


def check_grad_against_golden(name, g, packed, rtol, atol_scale=1.0):
    g = g.detach().float().cpu()
    if "full" in packed:
        ref = packed["full"]
        assert g.shape == ref.shape, (name, g.shape, ref.shape)
        got = g
    else:
        assert tuple(g.shape) == tuple(packed["shape"]), (name, g.shape, packed["shape"])
        ref = packed["sample"]
        got = g.reshape(-1)[::packed["stride"]]
        n = g.double().norm().item()
        assert abs(n - packed["norm"]) <= rtol * max(packed["norm"], 1e-12) + 1e-12, (name, n, packed["norm"])
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= rtol * scale * atol_scale + 1e-12, f"{name}: max err {err:.3e} vs scale {scale:.3e}"
    return err / max(scale, 1e-30)
