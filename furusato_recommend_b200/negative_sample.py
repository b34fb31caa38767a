"""`UniformSample(dataset)` of the reference (negative_sample.py:98-134) as one
Philox kernel + an order-preserving compaction on device.

The reference draws from numpy's global MT19937 in a Python loop (7e4 samples/s);
here sample i owns the counter-based stream Philox4x32-10(key=seed,
ctr=(i, block, epoch)), so the result is independent of launch geometry and of
how samples are sharded over GPUs.  The decision procedure per
sample (skip users without positives, positive from the FILE-ORDER list,
rejection of negatives by membership) is the reference's.
"""
from __future__ import annotations

import torch

from . import ops

_STATE = {"seed": 2020, "epoch": 0}  # reference parse.py:45 default --seed 2020


def set_seed(seed: int, epoch: int = 0) -> None:
    _STATE["seed"], _STATE["epoch"] = int(seed), int(epoch)


def UniformSample(dataset, neg_ratio: int = 1, *, seed: int | None = None, epoch: int | None = None,
                  count: int | None = None, start: int = 0) -> torch.Tensor:
    """Returns S: int64 CUDA tensor [n_s, 3] = (user, positem, negitem), n_s <= trainDataSize.

    `neg_ratio` = negatives per (user, positive).  The reference accepts and ignores it (:98), i.e.
    always 1; values > 1 produce the flat-triple layout the sampled-softmax variant batches
    (model/lgcnssm.py:141): neg_ratio consecutive rows (u, pos, neg_t) per sample.  Each call
    without an explicit `epoch` advances the module's epoch counter, which plays the
    role of the reference's advancing global RNG state.  `start`/`count` select the
    sub-range of sample indices [start, start+count) (multi-GPU sharding)."""
    if epoch is None:
        epoch = _STATE["epoch"]
        _STATE["epoch"] += 1
    if seed is None:
        seed = _STATE["seed"]
    if count is None:
        count = dataset.trainDataSize  # negative_sample.py:106
    rowptr, file_items, sorted_items = dataset.pos_csr()
    triples, valid = ops.uniform_sample(rowptr, file_items, sorted_items, dataset.n_users, dataset.m_items,
                                        count, seed, epoch, first=start, n_neg=max(1, int(neg_ratio)))
    return ops.compact_triples(triples, valid)


POSITIVE_NUM_LIMIT = 3000   # ddp_lgcn.py:34
TRAIN_ITERATIVE = 3         # ddp_lgcn.py:35


def UniformSampleCapped(dataset, neg_ratio: int = 1, *, limit: int = POSITIVE_NUM_LIMIT,
                        iterative: int = TRAIN_ITERATIVE, seed: int | None = None, epoch: int | None = None,
                        count: int | None = None) -> torch.Tensor:
    """The DDP script's sampler (ddp_lgcn.py:541-582): `trainDataSize * TRAIN_ITERATIVE` draws, and a
    sample is dropped when its positive item was already emitted `limit` times this epoch (:569-570).

    The cap is order dependent in the reference (a Python dict counted along the loop).  With
    per-sample Philox streams a dropped sample never shifts another sample's draws, so the same
    decision is "rank of sample i among the non-empty samples of lower index with the same
    positive < limit": one stable device sort by positive item, no sequential pass."""
    if epoch is None:
        epoch = _STATE["epoch"]
        _STATE["epoch"] += 1
    if seed is None:
        seed = _STATE["seed"]
    if count is None:
        count = dataset.trainDataSize * int(iterative)   # ddp_lgcn.py:549
    rowptr, file_items, sorted_items = dataset.pos_csr()
    triples, valid = ops.uniform_sample(rowptr, file_items, sorted_items, dataset.n_users, dataset.m_items,
                                        count, seed, epoch)
    m = int(dataset.m_items)
    key = torch.where(valid.bool(), triples[:, 1], torch.full_like(triples[:, 1], m))  # empties sort last
    skey, order = torch.sort(key, stable=True)
    pos_in_sorted = torch.arange(count, device=key.device)
    run_start = torch.zeros(m + 2, dtype=torch.int64, device=key.device)
    run_start[1:] = torch.cumsum(torch.bincount(skey, minlength=m + 1), 0)
    rank = pos_in_sorted - run_start[skey]
    keep_sorted = (rank < int(limit)) & (skey < m)
    keep = torch.zeros(count, dtype=torch.uint8, device=key.device)
    keep[order] = keep_sorted.to(torch.uint8)
    return ops.compact_triples(triples, keep)
