"""Data path — mirror of the reference's data_utils.py (data_load :164-213, DataDiffusion :216-226), CSR-native.

`data_load` keeps the reference signature and return value (three scipy CSR float64 matrices + sizes) but builds
them without the per-pair Python loop. `DeviceInteractions` uploads the CSR once (int32 rowptr/col) and hands out
`CsrBatch` objects, replacing the dense n_user x n_item float32 host matrices of main.py:143-156 (7.5 GB at Yelp,
41 GB at Amazon-Book per copy) — the kernels densify rows on the device.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch
from torch.utils.data import Dataset

from .models.gaussian_diffusion import CsrBatch


def data_load(train_path, valid_path, test_path):
    train_list = np.load(train_path, allow_pickle=True)
    valid_list = np.load(valid_path, allow_pickle=True)
    test_list = np.load(test_path, allow_pickle=True)
    n_user = int(train_list[:, 0].max()) + 1
    n_item = int(train_list[:, 1].max()) + 1
    print(f'user num: {n_user}')
    print(f'item num: {n_item}')

    def csr(pairs):
        return sp.csr_matrix((np.ones_like(pairs[:, 0]), (pairs[:, 0], pairs[:, 1])), dtype='float64', shape=(n_user, n_item))

    return csr(train_list), csr(valid_list), csr(test_list), n_user, n_item


class DataDiffusion(Dataset):
    """Reference-compatible dataset: item i -> (row, i). `data` may be a dense tensor (reference) or a scipy CSR
    matrix (rows densified lazily, one at a time, on the host)."""

    def __init__(self, data):
        self.data = data

    def __getitem__(self, index):
        if sp.issparse(self.data):
            return torch.from_numpy(np.asarray(self.data[index].todense(), dtype=np.float32).ravel()), index
        return self.data[index], index

    def __len__(self):
        return self.data.shape[0]


class DeviceInteractions:
    """A scipy CSR interaction matrix resident on the GPU as int32 (rowptr, col), sorted column ids per row."""

    def __init__(self, mat, device="cuda"):
        m = mat.tocsr()
        m.sum_duplicates()
        m.sort_indices()
        self.n_user, self.n_item = m.shape
        self.rowptr_host = m.indptr.astype(np.int64)  # host copy: slice bounds without a device sync
        self.rowptr = torch.from_numpy(m.indptr.astype(np.int32)).to(device)
        self.col = torch.from_numpy(m.indices.astype(np.int32)).to(device)
        self.device = device

    @property
    def csr(self):
        return self.rowptr, self.col

    def batch(self, users) -> CsrBatch:
        if not isinstance(users, torch.Tensor):
            users = torch.as_tensor(np.asarray(users, dtype=np.int32))
        return CsrBatch(self.rowptr, self.col, users.to(self.device, non_blocking=True).to(torch.int32), self.n_item)


def synthetic_interactions(n_user: int, n_item: int, n_pairs: int, seed: int = 0, split=(0.7, 0.1, 0.2)):
    """Synthetic (uid, iid) lists in data_load's .npy format: user degree ~ lognormal(sigma=1)
    rescaled to n_pairs total and clipped to [5, n_item/4]; items ~ Zipf(1.0) over a random permutation;
    de-duplicated; per-pair random 7:1:2 split. Returns (train, valid, test) int64 arrays [n, 2]."""
    rng = np.random.default_rng(seed)
    lo_deg, hi_deg = min(5, n_item), max(min(5, n_item), n_item // 4)
    deg = rng.lognormal(0.0, 1.0, n_user)
    deg = np.clip(deg * (n_pairs / deg.sum()), lo_deg, hi_deg).astype(np.int64)
    # the clip and the floor move the total: hand the difference out one interaction at a time over random users that
    # still have room, so that exactly n_pairs DISTINCT pairs come out (when the bounds allow it at all)
    n_pairs = int(min(max(n_pairs, n_user * lo_deg), n_user * hi_deg))
    while deg.sum() != n_pairs:
        diff = int(n_pairs - deg.sum())
        room = np.flatnonzero(deg < hi_deg) if diff > 0 else np.flatnonzero(deg > lo_deg)
        pick = rng.choice(room, size=min(abs(diff), room.shape[0]), replace=False)
        deg[pick] += 1 if diff > 0 else -1
    cdf = np.cumsum(1.0 / np.arange(1, n_item + 1))
    cdf /= cdf[-1]
    perm = rng.permutation(n_item)
    # draw with replacement, keep each user's first deg[u] distinct items in draw order; users that are still short are
    # topped up with a doubling over-draw (only they are re-examined). One stable sort per round: draws are generated
    # grouped by user, so "first deg[u] distinct, earlier rounds first" needs only a duplicate mask and running counts.
    done = []
    active = np.arange(n_user, dtype=np.int64)
    old = np.empty(0, dtype=np.int64)                     # distinct pairs already held by the active users
    have = np.zeros(n_user, dtype=np.int64)
    mult = 1.25
    while active.size:
        need = deg[active] - have[active]
        extra = (need * mult).astype(np.int64) + 2
        users = np.repeat(active, extra)
        new = users * n_item + perm[np.minimum(np.searchsorted(cdf, rng.random(users.shape[0])), n_item - 1)]
        both = np.concatenate([old, new])
        order = np.argsort(both, kind="stable")
        srt = both[order]
        fresh = np.ones(both.shape[0], dtype=bool)
        fresh[order[1:][srt[1:] == srt[:-1]]] = False     # equal to an earlier entry (held pair or earlier draw)
        fresh = fresh[old.shape[0]:]
        seen = np.cumsum(fresh) - fresh                   # fresh draws before this one ...
        begin = np.cumsum(extra) - extra                  # ... minus those of earlier users = rank within the user
        rank = seen - np.repeat(seen[begin], extra)
        keep = fresh & (have[users] + rank < deg[users])
        new, users = new[keep], users[keep]
        have[active] += np.bincount(np.searchsorted(active, users), minlength=active.size)
        held = np.concatenate([old, new])
        full = have[held // n_item] >= deg[held // n_item]
        done.append(held[full])
        old = held[~full]
        active = active[have[active] < deg[active]]
        mult *= 2.0
    keys = np.concatenate(done)
    key = np.sort(keys)
    users, items = key // n_item, key % n_item
    # every user and the largest ids must appear in train so that n_user/n_item derive like data_load does
    r = rng.random(users.shape[0])
    first = np.ones(users.shape[0], dtype=bool)
    first[1:] = users[1:] != users[:-1]
    part = np.where(first | (r < split[0]), 0, np.where(r < split[0] + split[1], 1, 2))
    part[np.argmax(items)] = 0
    pairs = np.stack([users, items], 1)
    return pairs[part == 0], pairs[part == 1], pairs[part == 2]
