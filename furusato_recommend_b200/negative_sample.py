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
