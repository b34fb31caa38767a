"""Train / eval loops over the fused kernels, shaped like the reference's Trainer
(trainer.py:27-280) but restricted to the hot path: no wandb, CSV dumps, product
names or Diversity/Novelty/Coverage (they need the private furusato files).

`Trainer.train()`  = trainer.py:56-81   (sample -> shuffle -> model.OneEpoch)
`Trainer.test()`   = trainer.py:115-187 (batched score -> mask -> top-k -> metrics),
                     with the score matrix never materialised and metrics on device
`get_topk_list(k)` = trainer.py:83-113  (candidate lists for the re-ranker)
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from . import metric
from .negative_sample import UniformSample

DEFAULT_TOPKS = (10, 20)  # reference parse.py:30 --topks "[10,20]"


def shuffle(*tensors: torch.Tensor, generator: torch.Generator | None = None):
    """utils.py:19-38 with a device permutation (the reference permutes with numpy's
    global RNG; on iid rows the permutation is statistically a no-op)."""
    if len({len(t) for t in tensors}) != 1:
        raise ValueError("All inputs to shuffle must have the same length.")
    perm = torch.randperm(len(tensors[0]), device=tensors[0].device, generator=generator)
    out = tuple(t[perm] for t in tensors)
    return out[0] if len(out) == 1 else out


def minibatch(*tensors, batch_size: int):
    """utils.py:7-17"""
    n = len(tensors[0])
    for i in range(0, n, batch_size):
        yield tensors[0][i:i + batch_size] if len(tensors) == 1 else tuple(t[i:i + batch_size] for t in tensors)


class Trainer:
    def __init__(self, config: dict, dataset, model, topks: Sequence[int] = DEFAULT_TOPKS):
        self.config, self.dataset, self.model = config, dataset, model
        self.topks = tuple(int(k) for k in topks)
        self.device = model.device
        self.max_recall = 0.0

    def train(self) -> torch.Tensor:
        """One epoch: trainer.py:56-81."""
        self.model.train()
        S = UniformSample(self.dataset)
        users, pos, neg = shuffle(S[:, 0].contiguous(), S[:, 1].contiguous(), S[:, 2].contiguous())
        return self.model.OneEpoch(users, pos, neg)

    def _eval_users(self) -> torch.Tensor:
        return torch.from_numpy(self.dataset.test_users()).to(self.device)  # list(testDict.keys()), :118

    def _eval_batch(self) -> int:
        """The reference batches users (test_u_batch_size=10000, parse.py:24) only because it
        materialises a [U_b, m] score matrix; the fused kernel does not, so small batches would
        just leave SMs idle (128 users per CTA).  Results do not depend on the batch size."""
        return max(int(self.config.get("test_u_batch_size", 10000)), 1 << 17)

    @torch.no_grad()
    def get_topk_list(self, k: int = 50) -> List[torch.Tensor]:
        """trainer.py:83-113: per user batch, int64 [U_b, k] on the host."""
        self.model.eval()
        users = self._eval_users()
        out = []
        idx_all = torch.cat([self.model.getUsersTopK(bu, k)[0] for bu in minibatch(users, batch_size=self._eval_batch())])
        # same list-of-batches shape as the reference (one int64 [U_b, k] tensor per test_u_batch_size users)
        for s0 in range(0, len(users), int(self.config["test_u_batch_size"])):
            out.append(idx_all[s0:s0 + int(self.config["test_u_batch_size"])].to(torch.int64).cpu())
        return out

    def export_candidates(self, path: str, k: int = 50) -> torch.Tensor:
        """eval.py:35-40: the re-ranker's input file — `torch.save` of one int64 [U, k] tensor
        (train_lgbm.py:113-114 flattens it and assumes k == 50 per user)."""
        cand = torch.cat(self.get_topk_list(k=k), dim=0)
        torch.save(cand, path)
        return cand

    @torch.no_grad()
    def test(self) -> Dict[str, np.ndarray]:
        """trainer.py:115-187 (hot-path metrics only): recall / precision / ndcg / hr @ topks."""
        self.model.eval()
        users = self._eval_users()
        test_rowptr, test_sorted = self.dataset.test_csr()
        kmax = max(self.topks)
        sums = torch.zeros((4, len(self.topks)), dtype=torch.float64, device=self.device)
        for bu in minibatch(users, batch_size=self._eval_batch()):
            idx, _ = self.model.getUsersTopK(bu, kmax)
            metric.batch_metric_sums(idx, bu, test_rowptr, test_sorted, self.topks, sums)
        res = metric.finalize(sums, len(users))
        self.max_recall = max(self.max_recall, float(res["recall"][0]))
        return res

    def train_epoch(self, epochs: int, test_span: int | None = None):
        """trainer.py:237-258 without the logging side effects."""
        span = int(test_span or self.config.get("test_span", 10))
        history = [("test", -1, self.test())]
        for epoch in range(epochs):
            loss = self.train()
            history.append(("loss", epoch, float(loss)))
            if epoch % span == 0:
                history.append(("test", epoch, self.test()))
        return history
