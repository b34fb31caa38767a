"""`LightGCNSSM` — the reference's "sampled softmax" variant (model/lgcnssm.py:44-153).

What the reference actually computes: `softmax_loss` (lgcnssm.py:98-118) is byte-for-byte the BPR
softplus loss of model/lgcn.py:98-118, and `OneEpoch` (lgcnssm.py:135-153) batches
`neg_size * bpr_batch_size` flat (user, pos, neg) triples per step (it references an undefined
global `neg_size`, so it raises NameError as shipped).  This class keeps exactly that arithmetic —
so it inherits the BPR parity pins — with `neg_size` read from the config (default 256, cfg-4 of
BASELINE.json), and adds the real sampled-softmax objective of SURVEY §9.7 as an opt-in
(`config["ssm_true_softmax"]`), which has no reference arithmetic behind it (parity unpinned).
"""
from __future__ import annotations

import torch

from .model import LightGCN


class LightGCNSSM(LightGCN):
    def __init__(self, config: dict, dataset):
        super().__init__(config, dataset)
        self.neg_size = int(config.get("neg_size", 256))
        self.tau = float(config.get("ssm_tau", 1.0))
        self.true_softmax = bool(config.get("ssm_true_softmax", False))
        # one fused step now covers neg_size * bpr_batch_size flat triples (lgcnssm.py:141)
        self._step_rows = self.neg_size * int(config["bpr_batch_size"])

    def softmax_loss(self, users, pos, neg):
        """lgcnssm.py:98-118 — identical to bpr_loss (sic)."""
        if self.true_softmax:
            return self._true_softmax_loss(users, pos, neg)
        return self.bpr_loss(users, pos, neg)

    def _true_softmax_loss(self, users, pos, neg):
        """mean_b[logsumexp([s+, s-_1..J]/tau) - s+/tau]; rows of one (user, pos) are the
        neg_size consecutive flat triples.  Autograd over the propagated embeddings (our own spec)."""
        J = self.neg_size
        users, pos, neg = self._ids(users), self._ids(pos), self._ids(neg)
        all_users, all_items = self.computer()
        u = all_users[users[::J]]
        p = all_items[pos[::J]]
        q = all_items[neg].view(-1, J, self.latent_dim)
        s_pos = (u * p).sum(1, keepdim=True)
        s_neg = torch.einsum("bd,bjd->bj", u, q)
        logits = torch.cat([s_pos, s_neg], dim=1) / self.tau
        loss = (torch.logsumexp(logits, dim=1) - logits[:, 0]).mean()
        w = self.all_embedding.weight
        B = u.shape[0]
        reg = 0.5 * (w[users[::J]].pow(2).sum() + w[pos[::J] + self.num_users].pow(2).sum()
                     + w[neg + self.num_users].pow(2).sum()) / float(B)
        return loss, reg

    def stageOne(self, user, pos, neg) -> torch.Tensor:
        """lgcnssm.py:127-133."""
        if not self.true_softmax:
            return super().stageOne(user, pos, neg)
        self.optim.zero_grad()
        loss, reg = self._true_softmax_loss(user, pos, neg)
        total = loss + float(self.config["decay"]) * reg
        total.backward()
        self.optim.step()
        self._eval_cache_valid = False
        return total.detach()

    @torch.no_grad()
    def OneEpoch(self, user, pos, neg) -> torch.Tensor:
        """lgcnssm.py:135-153: batches of neg_size * bpr_batch_size rows, divisor len // B + 1 with
        B = bpr_batch_size (sic)."""
        users, pos, neg = self._ids(user), self._ids(pos), self._ids(neg)
        total_batch = len(users) // int(self.config["bpr_batch_size"]) + 1
        aver = torch.zeros((), device=users.device)
        for i in range(0, len(users), self._step_rows):
            sl = slice(i, i + self._step_rows)
            if self.true_softmax:
                with torch.enable_grad():
                    aver = aver + self.stageOne(users[sl], pos[sl], neg[sl])
            else:
                self._fused_step_eager(users[sl], pos[sl], neg[sl])
                aver = aver + self._buf("loss_out")[2]
        return aver / total_batch
