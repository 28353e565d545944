"""TEST INFRASTRUCTURE — imports the *unmodified* reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference). It is
used by ``tests/golden/make_golden.py`` to generate the committed fixtures that
pin ``oracle/maple_cpu.py``, and by ``bench.py --impl reference`` when the tree
is present. Nothing in the product package imports this file.

Recipe = SURVEY.md Appendix C: stub ``dassl.*``, ``ftfy`` and the BPE tokenizer in
``sys.modules``; register the real ``clip/model.py``; import ``trainers.maple`` and
``trainers.maple_fed`` from the reference tree.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("MAPLE_REFERENCE_ROOT", "/root/reference")
_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.dirname(_here)
if _repo not in sys.path:
    sys.path.insert(0, _repo)

from federated_multi_modal_b200 import synth  # noqa: E402


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "trainers", "maple.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = {}


def load_reference():
    """Returns (clip_model_module, trainers.maple, trainers.maple_fed)."""
    if _loaded:
        return _loaded["model"], _loaded["maple"], _loaded["fed"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")

    class _Registry:
        def register(self, *a, **k):
            return lambda cls: cls

    class TrainerX:
        def __init__(self, cfg=None):
            self._models, self._optims, self._scheds = {}, {}, {}
            self.device = torch.device("cpu")

        def register_model(self, name, model=None, optim=None, sched=None):
            self._models[name], self._optims[name], self._scheds[name] = model, optim, sched

        def get_model_names(self, names=None):
            return list(self._models.keys())

    def _nop(*a, **k):
        return None

    _mod("dassl")
    _mod("dassl.engine", TRAINER_REGISTRY=_Registry(), TrainerX=TrainerX)
    _mod("dassl.metrics", compute_accuracy=_nop)
    _mod("dassl.utils", load_pretrained_weights=_nop, load_checkpoint=_nop, save_checkpoint=_nop,
         mkdir_if_missing=_nop)
    _mod("dassl.optim", build_optimizer=_nop, build_lr_scheduler=_nop)
    _mod("dassl.data", DataManager=object)
    _mod("dassl.data.datasets", Datum=object)
    _mod("dassl.data.data_manager", build_transform=_nop, build_data_loader=_nop)
    _mod("ftfy", fix_text=lambda s: s)

    # fake `clip` package; the real clip/model.py is loaded from the reference tree
    clip_pkg = _mod("clip")
    clip_pkg.__path__ = [os.path.join(REF_ROOT, "clip")]
    spec = importlib.util.spec_from_file_location("clip.model", os.path.join(REF_ROOT, "clip", "model.py"))
    model_mod = importlib.util.module_from_spec(spec)
    sys.modules["clip.model"] = model_mod
    spec.loader.exec_module(model_mod)

    class SimpleTokenizer:
        def encode(self, text):
            return synth.synthetic_encode(text)

    _mod("clip.simple_tokenizer", SimpleTokenizer=SimpleTokenizer)
    clip_clip = _mod("clip.clip", tokenize=synth.synthetic_tokenize, build_model=model_mod.build_model,
                     _MODELS={}, _download=_nop, _tokenizer=SimpleTokenizer())
    clip_pkg.clip = clip_clip
    clip_pkg.model = model_mod

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import trainers.maple as maple  # noqa
    import trainers.maple_fed as fed  # noqa
    _loaded.update(model=model_mod, maple=maple, fed=fed)
    return model_mod, maple, fed


def design_details(n_ctx):
    # trainers/maple.py:33-37
    return {"trainer": "MaPLe", "vision_depth": 0, "language_depth": 0, "vision_ctx": 0,
            "language_ctx": 0, "maple_length": n_ctx}


def apply_freeze_policy(model, nn):
    """Verbatim behaviour of trainers/maple.py:447-479."""
    for p in model.parameters():
        p.requires_grad_(False)
    for _, m in model.named_modules():
        if isinstance(m, (nn.LayerNorm, nn.BatchNorm1d, nn.BatchNorm2d)):
            for p in m.parameters():
                p.requires_grad_(True)
    for n, p in model.named_parameters():
        if "prompt_learner" in n:
            p.requires_grad_(True)
    for n, p in model.named_parameters():
        if "visual.transformer.resblocks.11" in n:
            p.requires_grad_(True)
    for n, p in model.named_parameters():
        if "transformer.resblocks.11" in n:
            p.requires_grad_(True)


def build_reference_customclip(clip_sd, classnames, cfg, pl_state=None, fp32=True):
    """CustomCLIP from the reference, loaded through its own build_model().

    fp32=True is the "fp32-ref" oracle of SURVEY.md §8c: ``clip_model.float()``
    (maple.py:438-439) plus the one documented patch ``model.float()`` (the
    reference crashes in fp32 without it).
    """
    import contextlib
    import io
    model_mod, maple, _ = load_reference()
    sd = {k: v.clone() for k, v in clip_sd.items()}
    with contextlib.redirect_stdout(io.StringIO()):
        clip_model = model_mod.build_model(sd, design_details(cfg.TRAINER.MAPLE.N_CTX))
        if fp32:
            clip_model.float()
        model = maple.CustomCLIP(cfg, classnames, clip_model)
    if pl_state is not None:
        missing = model.prompt_learner.load_state_dict(pl_state, strict=False)
        assert not missing.unexpected_keys, missing
    if fp32:
        model.float()
    apply_freeze_policy(model, torch.nn)
    return model
