"""LEAF json datasets: ``{"users": [...], "num_samples": [...], "user_data": {u: {"x": [...], "y": [...]}}}``.

Functions mirror the reference's scripts (models/utils/sample.py, split_data.py, remove_users.py,
stats.py) as library calls; ``to_fldataset`` bridges a LEAF dataset into this framework."""
from __future__ import annotations

import json
import os
import random
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np

from .util import iid_divide


def load_dir(data_dir: str) -> dict:
    """Merge every ``*.json`` of a LEAF directory."""
    users, num, data, hier = [], [], {}, []
    for f in sorted(os.listdir(data_dir)):
        if not f.endswith('.json'):
            continue
        with open(os.path.join(data_dir, f)) as inf:
            d = json.load(inf)
        users += d['users']
        num += d['num_samples']
        data.update(d['user_data'])
        hier += d.get('hierarchies', [])
    out = {'users': users, 'num_samples': num, 'user_data': data}
    if hier:
        out['hierarchies'] = hier
    return out


def save(dataset: dict, path: str) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, 'w') as f:
        json.dump(dataset, f)


def remove_users(dataset: dict, min_samples: int = 10) -> dict:
    keep = [i for i, n in enumerate(dataset['num_samples']) if n >= min_samples]
    users = [dataset['users'][i] for i in keep]
    return {'users': users, 'num_samples': [dataset['num_samples'][i] for i in keep],
            'user_data': {u: dataset['user_data'][u] for u in users}}


def sample(dataset: dict, fraction: float = 0.1, iid: bool = False, user_fraction: float = 0.01,
           seed: Optional[int] = None) -> dict:
    """niid: keep whole users until ``fraction`` of all samples is reached.
    iid: pool a ``fraction`` of all samples and deal them to ``user_fraction * #users`` pseudo users."""
    rng = random.Random(seed)
    users, nums, data = dataset['users'], dataset['num_samples'], dataset['user_data']
    total = sum(nums)
    target = int(fraction * total)
    if not iid:
        order = list(range(len(users)))
        rng.shuffle(order)
        kept, count = [], 0
        for i in order:
            if count >= target:
                break
            kept.append(i)
            count += nums[i]
        sel = [users[i] for i in kept]
        return {'users': sel, 'num_samples': [nums[i] for i in kept], 'user_data': {u: data[u] for u in sel}}
    pool = [(x, y) for u in users for x, y in zip(data[u]['x'], data[u]['y'])]
    rng.shuffle(pool)
    pool = pool[:target]
    n_users = max(1, int(user_fraction * len(users)))
    groups = iid_divide(pool, n_users)
    out = {'users': [], 'num_samples': [], 'user_data': {}}
    for i, g in enumerate(groups):
        u = str(i)
        out['users'].append(u)
        out['num_samples'].append(len(g))
        out['user_data'][u] = {'x': [p[0] for p in g], 'y': [p[1] for p in g]}
    return out


def split_data(dataset: dict, frac: float = 0.9, by_user: bool = False, seed: Optional[int] = None
               ) -> Tuple[dict, dict]:
    """Train/test split: ``by_user`` holds out whole users, otherwise each user's samples are split."""
    rng = random.Random(seed)
    users, data = dataset['users'], dataset['user_data']

    def pack(sel: Dict[str, dict]) -> dict:
        us = list(sel)
        return {'users': us, 'num_samples': [len(sel[u]['y']) for u in us], 'user_data': sel}

    if by_user:
        order = list(users)
        rng.shuffle(order)
        k = int(frac * len(order))
        return pack(OrderedDict((u, data[u]) for u in order[:k])), pack(OrderedDict((u, data[u]) for u in order[k:]))
    tr, te = OrderedDict(), OrderedDict()
    for u in users:
        n = len(data[u]['y'])
        if n < 2:
            continue
        k = min(max(int(frac * n), 1), n - 1)
        idx = list(range(n))
        rng.shuffle(idx)
        a, b = idx[:k], idx[k:]
        tr[u] = {'x': [data[u]['x'][i] for i in a], 'y': [data[u]['y'][i] for i in a]}
        te[u] = {'x': [data[u]['x'][i] for i in b], 'y': [data[u]['y'][i] for i in b]}
    return pack(tr), pack(te)


def stats(dataset: dict) -> dict:
    n = np.asarray(dataset['num_samples'], dtype=np.float64)
    if len(n) == 0:
        return {'users': 0, 'samples': 0}
    return {'users': int(len(n)), 'samples': int(n.sum()), 'mean': float(n.mean()), 'std': float(n.std()),
            'std/mean': float(n.std() / n.mean()), 'skewness': float(((n - n.mean()) ** 3).mean() / max(n.std() ** 3, 1e-30)),
            'histogram': np.histogram(n, bins=10)[0].tolist()}


def to_fldataset(train: dict, test: Optional[dict] = None, train_bs: int = 32):
    """Bridge: LEAF json -> ``blades_b200.datasets.FLDataset``."""
    import torch

    from ...datasets.basedataset import BatchStream
    from ...datasets.customdataset import CustomTensorDataset
    from ...datasets.dataset import FLDataset
    trains, tests = [], []
    for i, u in enumerate(train['users']):
        x = np.asarray(train['user_data'][u]['x'], dtype=np.float32)
        y = np.asarray(train['user_data'][u]['y'], dtype=np.int64)
        trains.append(BatchStream(x, y, train_bs, seed=i))
        src = test['user_data'].get(u) if test else None
        xt = np.asarray(src['x'], dtype=np.float32) if src else x
        yt = np.asarray(src['y'], dtype=np.int64) if src else y
        tests.append(CustomTensorDataset(torch.from_numpy(xt), torch.from_numpy(yt)))
    return FLDataset(trains, tests)
