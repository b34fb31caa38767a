"""Ranking metrics with the reference's definitions (metric.py:60-103,
utils.py:40-48) evaluated on device by lgcn_rank_metrics."""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch

from . import ops

METRICS = ("recall", "precision", "hr", "ndcg")


def batch_metric_sums(topk: torch.Tensor, user_ids: torch.Tensor, test_rowptr: torch.Tensor,
                      test_sorted: torch.Tensor, ks: Sequence[int], sums: torch.Tensor | None = None):
    """Adds this batch's per-k SUMS (recall, precision, hr, ndcg) into `sums` [4, len(ks)] fp64."""
    return ops.rank_metrics(topk, user_ids, test_rowptr, test_sorted, ks, sums)


def finalize(sums: torch.Tensor, n_users: int) -> Dict[str, np.ndarray]:
    """trainer.py:169-170: every metric is divided by len(users)."""
    s = (sums / float(n_users)).cpu().numpy()
    return {m: s[i] for i, m in enumerate(METRICS)}
