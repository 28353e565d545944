"""``ClientDataManager`` with the reference's interface (trainers/client_datamanager.py:10-156).

The reference builds its loaders with Dassl's ``build_transform`` / ``build_data_loader`` (PIL + DataLoader
workers; file I/O is out of scope here). This version accepts the same pre-split ``Datum``-like lists and
yields the same batch dicts ({"img","label","caption"?}); items may carry an in-memory tensor (``img``) —
the synthetic datasets of the benchmark — or a callable loader. Batches are assembled in pinned host memory
so the trainer's host->device copy is asynchronous.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional

import torch


class Datum:
    """Stand-in for dassl.data.datasets.Datum (impath, label, domain, classname[, caption])."""

    def __init__(self, impath="", label=0, domain=0, classname="", caption=None, img=None):
        self.impath, self.label, self.domain, self.classname, self.caption, self.img = \
            impath, int(label), domain, classname, caption, img


class _Loader:
    def __init__(self, items, batch_size, shuffle, drop_last, tfm, seed=0):
        self.items, self.bs, self.shuffle, self.drop_last, self.tfm = items, batch_size, shuffle, drop_last, tfm
        self.gpu_aug = tfm if isinstance(tfm, GpuAugment) else None
        self._gen = torch.Generator().manual_seed(seed)
        self.dataset = items

    def __len__(self):
        n = len(self.items)
        return n // self.bs if self.drop_last else (n + self.bs - 1) // self.bs

    def __iter__(self):
        n = len(self.items)
        order = torch.randperm(n, generator=self._gen).tolist() if self.shuffle else list(range(n))
        for b in range(len(self)):
            idx = order[b * self.bs:(b + 1) * self.bs]
            if self.gpu_aug is not None:  # raw uint8 batch + host-drawn augmentation parameters; pixels on the GPU
                imgs = [self.items[i].img for i in idx]
                assert imgs[0].dtype == torch.uint8, "GpuAugment expects uint8 [3,H,W] item images"
                same = all(im.shape == imgs[0].shape for im in imgs)
                if torch.cuda.is_available() and same:
                    raw = torch.stack(imgs, out=self._staging(b, len(imgs), imgs[0]))  # straight into pinned memory
                else:
                    raw = torch.stack(imgs)
                boxes, flip = self.gpu_aug.draw(len(idx), raw.shape[-2], raw.shape[-1])
                lab = torch.tensor([self.items[i].label for i in idx], dtype=torch.long)
                if torch.cuda.is_available():
                    raw = raw if raw.is_pinned() else raw.pin_memory()
                    lab, boxes, flip = lab.pin_memory(), boxes.pin_memory(), flip.pin_memory()
                yield {"img_u8": raw, "rrc_box": boxes, "flip": flip, "label": lab, "augment": self.gpu_aug,
                       "impath": [self.items[i].impath for i in idx]}
                continue
            imgs = [self.items[i].img if self.tfm is None else self.tfm(self.items[i]) for i in idx]
            lab = torch.tensor([self.items[i].label for i in idx], dtype=torch.long)
            if torch.cuda.is_available():
                img = torch.stack(imgs, out=self._staging(b, len(imgs), imgs[0]))  # straight into pinned memory
                lab = lab.pin_memory()
            else:
                img = torch.stack(imgs)
            yield {"img": img, "label": lab, "impath": [self.items[i].impath for i in idx]}

    _SLOTS = 3  # batches alive at once: one in flight on the GPU, one being assembled, one spare

    def _staging(self, b, n, like):
        """Rotating pinned staging buffers for assembled batches (a fresh pin_memory() per batch costs more host time
        than the training step it feeds). A yielded batch stays valid until _SLOTS - 1 further batches were drawn."""
        key = (tuple(like.shape), like.dtype)
        if getattr(self, "_stage_key", None) != key:
            self._stage_key = key
            self._stage = [torch.empty((self.bs,) + tuple(like.shape), dtype=like.dtype).pin_memory()
                           for _ in range(self._SLOTS)]
        return self._stage[b % self._SLOTS][:n]


def sample_rrc_params(height: int, width: int, generator: torch.Generator, scale=(0.08, 1.0),
                      ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """Crop box (top, left, h, w) drawn like torchvision.transforms.RandomResizedCrop.get_params — the transform
    Dassl's build_transform uses for "random_resized_crop" (10 attempts, then the centre-crop fallback)."""
    import math
    area = height * width
    log_ratio = (math.log(ratio[0]), math.log(ratio[1]))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1], generator=generator).item()
        aspect = math.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1], generator=generator).item())
        w = int(round(math.sqrt(target_area * aspect)))
        h = int(round(math.sqrt(target_area / aspect)))
        if 0 < w <= width and 0 < h <= height:
            i = int(torch.randint(0, height - h + 1, (1,), generator=generator).item())
            j = int(torch.randint(0, width - w + 1, (1,), generator=generator).item())
            return i, j, h, w
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w, h = width, int(round(width / min(ratio)))
    elif in_ratio > max(ratio):
        h, w = height, int(round(height * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


class GpuAugment:
    """Training transform of the reference's yaml (random_resized_crop bicubic + random_flip + normalize, configs/
    trainers/MaPLe/vit_b16_c2_ep5_batch4_2ctx.yaml:8-13) with the pixel work on the GPU: the loader ships the raw
    uint8 [B,3,H,W] batch (4x fewer host->device bytes than fp32 crops) plus the host-drawn crop boxes / flips, and
    ``apply`` runs libmfk's mfk_rrc_flip_normalize on the device. PIL + DataLoader workers stop being the bottleneck
    at >= 7 k images/s per GPU."""
    MEAN = (0.48145466, 0.4578275, 0.40821073)
    STD = (0.26862954, 0.26130258, 0.27577711)

    def __init__(self, size: int = 224, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0), flip_p: float = 0.5,
                 mean=MEAN, std=STD, seed: int = 0):
        self.size, self.scale, self.ratio, self.flip_p = size, scale, ratio, flip_p
        self.mean, self.std = torch.tensor(mean, dtype=torch.float32), torch.tensor(std, dtype=torch.float32)
        self._gen = torch.Generator().manual_seed(seed)
        self._dev = {}

    def draw(self, n: int, height: int, width: int):
        boxes = torch.tensor([sample_rrc_params(height, width, self._gen, self.scale, self.ratio) for _ in range(n)],
                             dtype=torch.int32)
        flip = (torch.rand(n, generator=self._gen) < self.flip_p).to(torch.uint8)
        return boxes, flip

    def apply(self, img_u8: torch.Tensor, boxes: torch.Tensor, flip: torch.Tensor, out=None) -> torch.Tensor:
        from .. import ops
        dev = img_u8.device
        if dev not in self._dev:
            self._dev[dev] = (self.mean.to(dev), self.std.to(dev))
        mean, std = self._dev[dev]
        if out is None:
            out = torch.empty(img_u8.shape[0], 3, self.size, self.size, device=dev, dtype=torch.float32)
        ops.rrc_flip_normalize(img_u8.contiguous(), boxes.to(dev, non_blocking=True), flip.to(dev, non_blocking=True),
                               mean, std, out)
        return out


class ClientDataManager:
    def __init__(self, train_x, val, test, cfg, custom_tfm_train=None, custom_tfm_test=None, dataset_wrapper=None):
        for name, subset in (("train_x", train_x), ("val", val), ("test", test)):
            self._validate_labels(subset, name)
        self.train_x_list, self.val_list, self.test_list = train_x, val, test
        self.cfg = cfg
        self._classnames = sorted({it.classname for it in (train_x + val + test)})
        self._num_classes = len(self._classnames)
        self._lab2cname = None
        self.tfm_train, self.tfm_test = custom_tfm_train, custom_tfm_test
        dl = getattr(cfg, "DATALOADER", None)
        bs_train = dl.TRAIN_X.BATCH_SIZE if dl is not None else 4
        bs_test = dl.TEST.BATCH_SIZE if dl is not None else 100
        seed = getattr(cfg, "SEED", 0) or 0
        self.train_loader = _Loader(train_x, bs_train, True, True, self.tfm_train, seed) if train_x else None
        self.val_loader = _Loader(val, bs_test, False, False, self.tfm_test) if val else None
        self.test_loader = _Loader(test, bs_test, False, False, self.tfm_test) if test else None

    def _validate_labels(self, subset, name):
        for i, item in enumerate(subset or []):
            if not hasattr(item, "label"):
                raise ValueError(f"Missing 'label' attribute in {name} at index {i}.")
            if not isinstance(item.label, int):
                raise TypeError(f"Invalid label type in {name} at index {i}. Expected int, got {type(item.label)}")

    @property
    def dataset(self):
        return SimpleNamespace(train_x=self.train_x_list, val=self.val_list, test=self.test_list, train_u=[],
                               num_classes=self.num_classes, lab2cname=self.lab2cname, classnames=self._classnames)

    @property
    def num_classes(self):
        return self._num_classes

    @property
    def lab2cname(self):
        if self._lab2cname is None:
            self._lab2cname = {}
            for it in self.train_x_list + self.val_list + self.test_list:
                self._lab2cname.setdefault(it.label, it.classname)
        return self._lab2cname


def synthetic_client_items(num_classes: int, per_class: int, seed: int = 0, size: int = 224,
                           classnames: Optional[List[str]] = None) -> List[Datum]:
    """Synthetic pool for the federated benchmark: per_class seeded images per class."""
    from ..synth import synthetic_classnames
    names = classnames or synthetic_classnames(num_classes)
    g = torch.Generator().manual_seed(seed)
    items = []
    for c in range(num_classes):
        for j in range(per_class):
            items.append(Datum(impath=f"synthetic://{c}/{j}", label=c, classname=names[c],
                               img=torch.randn(3, size, size, generator=g)))
    return items
