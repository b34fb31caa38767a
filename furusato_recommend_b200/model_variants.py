"""LightGCN variants of the reference registry that run on the same kernels.

* `rAdjGCN` — model/radj.py:13-156: the propagation weight of edge (j -> i) is
  1 / (deg_j^r * deg_i^(1-r)) instead of the symmetric 1 / sqrt(deg_j deg_i)
  (model/radj.py:32-36,42-44).  That is diag(deg^-(1-r)) A diag(deg^-r): the same CSR SpMM with
  two scale vectors; it is not symmetric, so the backward pass runs the transpose — the same
  launch with the vectors swapped (`lgcn_layer_args_t.src_scale / dst_scale`).
* `RGCN` — model/rgcn.py:45-174: LightGCN propagation over the purchase edges plus the
  "favourite" edges (model/rgcn.py:60-85); the loss, the sampler and the evaluation still use the
  purchase lists only.  `relation_emb` (:86) is created but never read by forward(), as in the
  reference.

* `DDPLightGCN` — the inline model of ddp_lgcn.py:406-538, whose call shape differs from model/lgcn.py:
  `forward(edge_index)`, `stageOne(optimizer, u, p, n)` / `OneEpoch(optimizer, u, p, n)` with a
  caller-owned optimizer (ddp_lgcn.py:664,673), and `getUsersRating() -> (user_x, item_x)` (:535-538).

Everything else (bpr_loss, stageOne, OneEpoch, getUsersRating, getUsersTopK, the fused step and
its CUDA graph) is inherited from `LightGCN`.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from .graph import CsrGraph, build_csr_graph
from .model import LightGCN


class rAdjGCN(LightGCN):
    def __init__(self, config: dict, dataset):
        super().__init__(config, dataset)
        self.r = float(config["r"])  # parse.py:49, world.py:62
        g = self.graph
        deg = (g.rowptr[1:] - g.rowptr[:-1]).to(torch.float32)   # multiplicity-counting, radj.py:29-30
        # isolated nodes: the reference substitutes 1e-6 (radj.py:31) but they have no edges, so the
        # value never reaches an output; 0 keeps their rows exactly zero
        col = torch.where(deg > 0, deg.pow(-self.r), torch.zeros_like(deg))
        row = torch.where(deg > 0, deg.pow(-(1.0 - self.r)), torch.zeros_like(deg))
        self._col_scale, self._row_scale = col.contiguous(), row.contiguous()


class RGCN(LightGCN):
    """`dataset.favoriteUser / favoriteItem` (int arrays, item ids WITHOUT the user offset) or
    `config["favorite_csv"]` (columns cf_customer, cf_product — model/rgcn.py:60-62) name the
    extra edges."""

    def __init__(self, config: dict, dataset):
        super().__init__(config, dataset)
        self.relation_emb = nn.Embedding(2, self.latent_dim, device=self.device)  # model/rgcn.py:86, unused

    def _build_graph(self, dataset) -> CsrGraph:
        fu = getattr(dataset, "favoriteUser", None)
        fi = getattr(dataset, "favoriteItem", None)
        if fu is None and self.config.get("favorite_csv"):
            import pandas as pd
            fav = pd.read_csv(self.config["favorite_csv"])
            fu, fi = fav["cf_customer"].values, fav["cf_product"].values
        if fu is None:
            raise ValueError("RGCN needs dataset.favoriteUser/favoriteItem or config['favorite_csv']")
        dev = torch.device(self.config.get("device", "cuda:0"))
        u = np.concatenate([np.asarray(dataset.trainUser, dtype=np.int64), np.asarray(fu, dtype=np.int64)])
        i = np.concatenate([np.asarray(dataset.trainItem, dtype=np.int64), np.asarray(fi, dtype=np.int64)])
        self.purchaseSize, self.favoriteSize = 2 * len(dataset.trainUser), 2 * len(fu)   # rgcn.py:58,81
        return build_csr_graph(dataset.n_users, dataset.m_items, torch.from_numpy(u).to(dev),
                               torch.from_numpy(i).to(dev))


class DDPLightGCN(LightGCN):
    """The call shape of ddp_lgcn.py's inline `LightGCN` (SURVEY §8b "DDP shape"), for launchers that wrap
    the model in `DistributedDataParallel(model).module` and own the optimizer.

    The reference builds both `edge_index` (train edges) and `inference_edge_index`
    (dataset.inferenceUser / inferenceItem, ddp_lgcn.py:427-441) and only ever propagates over the
    latter (getEmbedding :471, getUsersRating :537): the graph here is built from the inference
    edges when the dataset has them, from the train edges otherwise.  `forward(edge_index)` accepts
    and ignores the argument — the CSR graph was built once, and re-deriving `gcn_norm` from an
    edge list on every call is exactly the work the reference's LGConv wastes.

    With a caller-owned optimizer the step is bpr_loss -> backward -> optimizer.step(), i.e. the
    autograd path through `torch.ops.lgcn_b200.propagate`; pass the model's own `optim` (FusedAdam)
    to get the fused single-graph step instead."""

    def _build_graph(self, dataset) -> CsrGraph:
        iu, ii = getattr(dataset, "inferenceUser", None), getattr(dataset, "inferenceItem", None)
        if iu is None or ii is None:
            return dataset.csr_graph()
        dev = torch.device(self.config.get("device", "cuda:0"))
        return build_csr_graph(dataset.n_users, dataset.m_items, torch.as_tensor(np.asarray(iu, dtype=np.int64)).to(dev),
                               torch.as_tensor(np.asarray(ii, dtype=np.int64)).to(dev))

    def forward(self, edge_index=None):
        """ddp_lgcn.py:466-474"""
        return self.computer()

    def stageOne(self, optimizer, user, pos, neg) -> torch.Tensor:
        """ddp_lgcn.py:498-504"""
        if optimizer is self.optim:
            return super().stageOne(user, pos, neg)
        with torch.enable_grad():
            optimizer.zero_grad()
            loss, reg_loss = self.bpr_loss(user, pos, neg)
            loss = loss + float(self.config["decay"]) * reg_loss
            loss.backward()
            optimizer.step()
        self._invalidate_derived()
        return loss.detach()

    def OneEpoch(self, optimizer, user, pos, neg) -> torch.Tensor:
        """ddp_lgcn.py:506-523: contiguous mini-batches, loss sum / (len // B + 1)."""
        if optimizer is self.optim:
            return super().OneEpoch(user, pos, neg)
        users, pos, neg = self._ids(user), self._ids(pos), self._ids(neg)
        B = int(self.config["bpr_batch_size"])
        aver = torch.zeros((), device=users.device)
        for i in range(0, len(users), B):
            aver = aver + self.stageOne(optimizer, users[i:i + B], pos[i:i + B], neg[i:i + B])
        return aver / (len(users) // B + 1)

    @torch.no_grad()
    def getUsersRating(self, users=None):
        """ddp_lgcn.py:535-538 returns the propagated (user_x, item_x) tables (the launcher does the batched
        matmul / mask / top-k itself, :690-711); with `users` given, the dense scores of model/lgcn.py:120-125."""
        if users is None:
            return self.computer()
        return super().getUsersRating(users)
