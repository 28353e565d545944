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
            imgs = [self.items[i].img if self.tfm is None else self.tfm(self.items[i]) for i in idx]
            img = torch.stack(imgs)
            lab = torch.tensor([self.items[i].label for i in idx], dtype=torch.long)
            if torch.cuda.is_available():
                img, lab = img.pin_memory(), lab.pin_memory()
            yield {"img": img, "label": lab, "impath": [self.items[i].impath for i in idx]}


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
