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
    with ops.nvtx("lgcn.uniform_sample"):
        triples, valid = ops.uniform_sample(rowptr, file_items, sorted_items, dataset.n_users, dataset.m_items,
                                            count, seed, epoch, first=start, n_neg=max(1, int(neg_ratio)))
        return ops.compact_triples(triples, valid)


POSITIVE_NUM_LIMIT = 3000   # ddp_lgcn.py:34
TRAIN_ITERATIVE = 3         # ddp_lgcn.py:35


def cap_keep_mask(pos: torch.Tensor, valid: torch.Tensor, m_items: int, limit: int) -> torch.Tensor:
    """uint8[count]: 1 where the sample is valid and fewer than `limit` valid samples of LOWER index
    carry the same positive item (the order-dependent counter of ddp_lgcn.py:569-570 as a rank)."""
    count = pos.numel()
    key = torch.where(valid.bool(), pos, torch.full_like(pos, m_items))  # empties sort last
    skey, order = torch.sort(key, stable=True)
    pos_in_sorted = torch.arange(count, device=key.device)
    run_start = torch.zeros(m_items + 2, dtype=torch.int64, device=key.device)
    run_start[1:] = torch.cumsum(torch.bincount(skey, minlength=m_items + 1), 0)
    rank = pos_in_sorted - run_start[skey]
    keep_sorted = (rank < limit) & (skey < m_items)
    keep = torch.zeros(count, dtype=torch.uint8, device=key.device)
    keep[order] = keep_sorted.to(torch.uint8)
    return keep


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
    keep = cap_keep_mask(triples[:, 1], valid, int(dataset.m_items), int(limit))
    return ops.compact_triples(triples, keep)


class UniformSampling:
    """`UniformSampling(dataset, config).sample()` of the reference (negative_sample.py:12-96): the
    multi-process sampler whose positive is drawn by popularity-powered probabilities when
    `config['sample_pow'] != 0` (`np.random.choice(len(pos), p=self.probs[user])`, :53-56).

    The reference unpickles `self.probs` from `./data/sample_prob/sample_prob_XX.pkl` (private
    files); here `probs` is passed in — a sequence of per-user arrays aligned with `allPos[u]`
    (file order) or one flat array in that order.  When omitted, p_j ∝ popularity(item_j)^(-sample_pow)
    is used (our reading of the file names; say so if you rely on it).  `sample_pow == 0` is the
    uniform pick (:51-52).  The four forked workers of the reference all inherit the same numpy RNG
    state and drop the remainder of count / 4 (:41-42); with per-sample Philox streams there is
    nothing to fork: one launch draws `trainDataSize` samples."""

    def __init__(self, dataset, config, neg_ratio: int = 1, probs=None):
        self.dataset, self.config = dataset, config
        self.sample_pow = float(config.get("sample_pow", 0))
        self.m_items, self.n_users, self.user_num = dataset.m_items, dataset.n_user, dataset.trainDataSize
        self._cdf = None
        if self.sample_pow != 0:
            rowptr, file_items, _ = dataset.pos_csr()
            dev = rowptr.device
            if probs is None:
                pop = torch.bincount(file_items.long(), minlength=self.m_items).to(torch.float64)
                w = pop[file_items.long()].pow(-self.sample_pow)
            elif torch.is_tensor(probs):
                w = probs.to(device=dev, dtype=torch.float64).flatten()
            else:
                import numpy as np
                flat = probs if isinstance(probs, np.ndarray) and probs.ndim == 1 else np.concatenate([np.asarray(p) for p in probs])
                w = torch.as_tensor(flat, dtype=torch.float64, device=dev)
            if w.numel() != file_items.numel():
                raise ValueError("probs must align with the train interactions (allPos order)")
            self._cdf = _segment_cdf(w, rowptr)

    def sample(self, *, seed: int | None = None, epoch: int | None = None, count: int | None = None) -> torch.Tensor:
        if epoch is None:
            epoch = _STATE["epoch"]
            _STATE["epoch"] += 1
        if seed is None:
            seed = _STATE["seed"]
        rowptr, file_items, sorted_items = self.dataset.pos_csr()
        triples, valid = ops.uniform_sample(rowptr, file_items, sorted_items, self.n_users, self.m_items,
                                            self.user_num if count is None else count, seed, epoch, pos_cdf=self._cdf)
        return ops.compact_triples(triples, valid)


def _segment_cdf(w: torch.Tensor, rowptr: torch.Tensor) -> torch.Tensor:
    """Per-user inclusive cumulative sums of w, normalised to end at 1 (numpy: cdf = p.cumsum();
    cdf /= cdf[-1]), in float64, stored as fp32."""
    c = torch.cumsum(w, 0)
    lens = rowptr[1:] - rowptr[:-1]
    start = torch.cat([torch.zeros(1, dtype=c.dtype, device=c.device), c])[rowptr[:-1]]   # sum before each user
    seg = torch.repeat_interleave(torch.arange(lens.numel(), device=c.device), lens)
    local = c - start[seg]
    total = torch.zeros(lens.numel(), dtype=c.dtype, device=c.device).index_add_(0, seg, w)
    return (local / total[seg]).to(torch.float32).contiguous()
