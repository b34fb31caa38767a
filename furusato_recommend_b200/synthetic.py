"""Synthetic Gowalla-shaped bipartite interaction graphs.

Recipe: item popularity ∝ rank^-0.8 under a seeded permutation; user degrees
log-normal(mu=3, sigma=0.9) clipped to [10, 2000] and rescaled to the requested
interaction count; items drawn per user by popularity, de-duplicated; iterated
five-core pruning (reference README.md:3-6); ids compacted; per-user 80/20
train/test split with the lists left in random (not sorted) order so the
sampler's file-order rule is exercised.  Pure torch ops, so the same code runs on
the host for cfg-1/2 and on the device for cfg-3-sized graphs.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch


def _five_core(user: torch.Tensor, item: torch.Tensor, n: int, m: int, core: int = 5):
    while True:
        ic = torch.bincount(item, minlength=m)
        keep = ic[item] >= core
        user, item = user[keep], item[keep]
        uc = torch.bincount(user, minlength=n)
        keep2 = uc[user] >= core
        user, item = user[keep2], item[keep2]
        if bool(keep.all()) and bool(keep2.all()):
            return user, item


def bipartite(n_users: int, m_items: int, n_interactions: int, seed: int = 2020,
              device: str = "cpu", train_frac: float = 0.8, core: int = 5
              ) -> Tuple[int, int, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns (n, m, train_user, train_item, test_user, test_item), int64 tensors on
    `device`; train rows are grouped by user in ascending uid order."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    dev = torch.device(device)
    # user degrees
    deg = torch.exp(3.0 + 0.9 * torch.randn(n_users, generator=g, device=dev)).clamp_(10, 2000)
    deg = (deg * (n_interactions / float(deg.sum()))).clamp_(min=core + 1, max=min(2000, m_items // 2))
    deg = deg.round().long()
    # item popularity CDF under a random permutation
    # built on the HOST: a device cumsum (decoupled look-back scan) associates floating-point sums differently from
    # call to call, and a CDF that differs in its last bits flips a handful of the ~10^9 searchsorted draws below —
    # two ranks generating "the same" cfg-3 graph then disagree about a few edges
    rank = torch.arange(1, m_items + 1, dtype=torch.float64)
    p = rank.pow(-0.8)
    cdf = torch.cumsum(p / p.sum(), 0).to(dev)
    perm = torch.randperm(m_items, generator=g, device=dev)
    # oversampled candidate draws, then de-duplicate (user,item) pairs
    over = (deg.double() * 1.6).ceil().long() + 8
    cu = torch.repeat_interleave(torch.arange(n_users, device=dev), over)
    r = torch.rand(cu.numel(), generator=g, device=dev, dtype=torch.float64)
    ci = perm[torch.searchsorted(cdf, r).clamp_(max=m_items - 1)]
    key = torch.unique(cu * m_items + ci)
    cu, ci = torch.div(key, m_items, rounding_mode="floor"), key % m_items
    # random order inside each user, keep the first deg[u]
    prio = torch.rand(cu.numel(), generator=g, device=dev, dtype=torch.float64)
    order = torch.argsort(cu.double() + prio * 0.999999)
    cu, ci = cu[order], ci[order]
    cnt = torch.bincount(cu, minlength=n_users)
    start = torch.cumsum(cnt, 0) - cnt
    pos = torch.arange(cu.numel(), device=dev) - start[cu]
    keep = pos < deg[cu]
    cu, ci = cu[keep], ci[keep]
    # five-core + id compaction
    cu, ci = _five_core(cu, ci, n_users, m_items, core)
    uu, cu = torch.unique(cu, return_inverse=True)
    ii, ci = torch.unique(ci, return_inverse=True)
    n, m = int(uu.numel()), int(ii.numel())
    # rows are still grouped by user (unique/inverse keeps order of the edge list)
    cnt = torch.bincount(cu, minlength=n)
    start = torch.cumsum(cnt, 0) - cnt
    pos = torch.arange(cu.numel(), device=dev) - start[cu]
    n_train = torch.clamp((cnt.double() * train_frac).ceil().long(), min=1)
    is_train = pos < n_train[cu]
    return n, m, cu[is_train], ci[is_train], cu[~is_train], ci[~is_train]


def make_dataset(n_users: int = 30000, m_items: int = 41000, n_interactions: int = 1_250_000,
                 seed: int = 2020, device: str = "cuda:0", config: dict | None = None):
    """cfg-1/2 of BASELINE.json by default: ~30k x 41k, ~1M train edges."""
    from .dataloader import BasicDataset
    n, m, tu, ti, su, si = bipartite(n_users, m_items, n_interactions, seed=seed, device="cpu")
    cfg = dict(config or {})
    return BasicDataset(n, m, tu.numpy(), ti.numpy(), su.numpy(), si.numpy(), config=cfg, device=device)
