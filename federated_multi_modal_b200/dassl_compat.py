"""Boundary shim for Dassl.pytorch (external, un-vendored, absent offline — SURVEY.md §8b).

If the real ``dassl`` package is importable it is used unchanged; otherwise the small subset the
MaPLe hot path touches is restated here from how the reference calls it:
  TRAINER_REGISTRY.register()              trainers/maple.py:384, maple_fed.py:24
  TrainerX.__init__(cfg) -> build_data_loader(), build_model(); register_model / get_model_names
                                           trainers/maple.py:402,502-504,692
  build_lr_scheduler(optim, optim_cfg)     trainers/maple.py:499, maple_fed.py:337 (cosine + constant warm-up)
  save_checkpoint / load_checkpoint        trainers/maple_fed.py:384,401 (``model.pth.tar-<epoch>`` dicts)
The optimiser itself is the fused clip+SGD kernel (``MapleEngine.sgd_step``), configured from the
same ``cfg.OPTIM`` keys as Dassl's ``build_optimizer`` (trainers/maple.py:498).
"""
from __future__ import annotations

import math
import os
import os.path as osp

import torch

try:  # pragma: no cover - real Dassl is not installed in the build/bench images
    from dassl.engine import TRAINER_REGISTRY, TrainerX  # type: ignore
    HAVE_DASSL = True
except Exception:  # noqa: BLE001
    HAVE_DASSL = False

    class _Registry:
        def __init__(self):
            self._obj = {}

        def register(self, *a, **k):
            def deco(cls):
                self._obj[cls.__name__] = cls
                return cls
            return deco

        def get(self, name):
            return self._obj[name]

        def registered_names(self):
            return list(self._obj)

    TRAINER_REGISTRY = _Registry()

    class TrainerX:
        """Minimal restatement of dassl.engine.TrainerX for the methods the reference relies on."""

        def __init__(self, cfg=None):
            self._models, self._optims, self._scheds = {}, {}, {}
            self.cfg = cfg
            self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
            self.start_epoch = self.epoch = 0
            self.max_epoch = getattr(getattr(cfg, "OPTIM", None), "MAX_EPOCH", 0)
            self.check_cfg(cfg)
            self.build_data_loader()
            self.build_model()

        def check_cfg(self, cfg):
            pass

        def build_data_loader(self):
            self.dm = getattr(self, "dm", None)

        def build_model(self):
            raise NotImplementedError

        def register_model(self, name="model", model=None, optim=None, sched=None):
            self._models[name], self._optims[name], self._scheds[name] = model, optim, sched

        def get_model_names(self, names=None):
            return list(self._models.keys()) if names is None else list(names)

        def model_inference(self, input):
            return self.model(input)


def build_trainer(cfg):
    return TRAINER_REGISTRY.get(cfg.TRAINER.NAME)(cfg)


class CosineLR:
    """torch.optim.lr_scheduler.CosineAnnealingLR(T_max=MAX_EPOCH) restated on a plain lr holder."""

    def __init__(self, base_lr: float, t_max: int):
        self.base_lr, self.t_max, self.last_epoch = base_lr, t_max, 0

    def step(self):
        self.last_epoch += 1

    def get_last_lr(self):
        return self.base_lr * (1 + math.cos(math.pi * self.last_epoch / max(self.t_max, 1))) / 2


class ConstantWarmupCosine:
    """Dassl ``build_lr_scheduler`` for LR_SCHEDULER=cosine, WARMUP_TYPE=constant (yaml 15-22 of
    configs/trainers/MaPLeFederated/...): WARMUP_CONS_LR for the first WARMUP_EPOCH epochs, then the
    cosine successor, which only starts stepping once warm-up is over. ``last_epoch`` is a plain
    attribute, so the reference's ``sched.last_epoch = epoch - 1`` after a rebuild
    (trainers/maple_fed.py:337-339) behaves as there: the LR restarts at the warm-up value."""

    def __init__(self, base_lr, max_epoch, warmup_epoch=0, cons_lr=None, holder=None):
        self.successor = CosineLR(base_lr, max_epoch)
        self.warmup_epoch, self.cons_lr = warmup_epoch, cons_lr
        self.last_epoch = -1
        self.holder = holder
        self._lr = base_lr
        self.step()

    def step(self, epoch=None):
        # dassl.optim.lr_scheduler._BaseWarmupScheduler.step: the test uses last_epoch BEFORE the increment,
        # and last_epoch stops advancing once the successor has taken over.
        if self.last_epoch >= self.warmup_epoch:
            self.successor.step()
            self._lr = self.successor.get_last_lr()
        else:
            self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
            self._lr = self.successor.get_last_lr() if self.last_epoch >= self.warmup_epoch else self.cons_lr
        if self.holder is not None:
            self.holder.lr = self._lr

    def get_last_lr(self):
        return [self._lr]


def build_lr_scheduler(optim, optim_cfg):
    warm = getattr(optim_cfg, "WARMUP_EPOCH", 0) or 0
    return ConstantWarmupCosine(optim_cfg.LR, optim_cfg.MAX_EPOCH, warm, getattr(optim_cfg, "WARMUP_CONS_LR", None),
                                holder=optim)


def mkdir_if_missing(d):
    os.makedirs(d, exist_ok=True)


def save_checkpoint(state, save_dir, is_best=False, remove_module_from_keys=True, model_name=""):
    """Dassl wire format: <dir>/model.pth.tar-<epoch> (+ 'checkpoint' pointer file)."""
    mkdir_if_missing(save_dir)
    epoch = state["epoch"]
    fpath = osp.join(save_dir, model_name or f"model.pth.tar-{epoch}")
    torch.save(state, fpath)
    with open(osp.join(save_dir, "checkpoint"), "w") as f:
        f.write(osp.basename(fpath) + "\n")
    if is_best:
        torch.save(state, osp.join(save_dir, "model-best.pth.tar"))
    return fpath


def load_checkpoint(fpath):
    if not osp.exists(fpath):
        raise FileNotFoundError(f'File is not found at "{fpath}"')
    return torch.load(fpath, map_location="cpu", weights_only=False)
