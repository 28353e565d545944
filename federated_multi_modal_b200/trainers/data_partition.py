"""Client partitioners. ``partition_dataset_iid`` keeps the reference's contract
(trainers/data_partition.py:5-26); ``partition_dataset_dirichlet`` is the label-skew split that
BASELINE configs 3-4 ask for and the reference lacks."""
import random

import numpy as np


def partition_dataset_iid(dataset, num_clients=3):
    """Shuffle dataset.train_x (global `random`, as the reference) and cut it into num_clients chunks; the
    last client takes the remainder; val/test are shared. Returns [(train_i, val, test)]."""
    items = list(dataset.train_x)
    random.shuffle(items)
    size = len(items) // num_clients
    out = []
    for i in range(num_clients):
        hi = (i + 1) * size if i < num_clients - 1 else len(items)
        out.append((items[i * size:hi], dataset.val, dataset.test))
    return out


def dirichlet_label_split(labels, num_clients, alpha=0.5, seed=0, min_size=1):
    """Index lists per client: for every class, its samples are split by proportions ~ Dir(alpha)."""
    rng = np.random.RandomState(seed)
    labels = np.asarray(labels)
    for _ in range(100):
        parts = [[] for _ in range(num_clients)]
        for c in np.unique(labels):
            idx = np.where(labels == c)[0]
            rng.shuffle(idx)
            cuts = (np.cumsum(rng.dirichlet([alpha] * num_clients)) * len(idx)).astype(int)[:-1]
            for k, chunk in enumerate(np.split(idx, cuts)):
                parts[k] += chunk.tolist()
        if min(len(p) for p in parts) >= min_size:
            break
    return [sorted(p) for p in parts]


def partition_dataset_dirichlet(dataset, num_clients=8, alpha=0.5, seed=0):
    items = list(dataset.train_x)
    parts = dirichlet_label_split([it.label for it in items], num_clients, alpha, seed)
    return [([items[i] for i in p], dataset.val, dataset.test) for p in parts]
