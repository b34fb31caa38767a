"""`LightGCNSSM` — the reference's "sampled softmax" variant (model/lgcnssm.py:44-153).

What the reference actually computes: `softmax_loss` (lgcnssm.py:98-118) is byte-for-byte the BPR
softplus loss of model/lgcn.py:98-118, and `OneEpoch` (lgcnssm.py:135-153) batches
`neg_size * bpr_batch_size` flat (user, pos, neg) triples per step (it references an undefined
global `neg_size`, so it raises NameError as shipped).  This class keeps exactly that arithmetic —
so it inherits the BPR parity pins — with `neg_size` read from the config (default 256, cfg-4 of
BASELINE.json), and adds the real sampled-softmax objective of SURVEY §9.7 as an opt-in
(`config["ssm_true_softmax"]`), which has no reference arithmetic behind it (parity unpinned): its
fused step runs on `lgcn_ssm_fwd_bwd` (csrc/ssm.cu) + the same Horner backward / Adam epilogue as BPR;
`_true_softmax_loss` is the plain-torch statement of the same formula that the kernel is tested against.
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
        if self.config.get("ssm_autograd", False):   # the torch statement of the objective (tests)
            self.optim.zero_grad()
            loss, reg = self._true_softmax_loss(user, pos, neg)
            total = loss + float(self.config["decay"]) * reg
            total.backward()
            self.optim.step()
            self._eval_cache_valid = False
            return total.detach()
        with torch.no_grad():
            self._fused_ssm_step(self._ids(user), self._ids(pos), self._ids(neg))
            return self._buf("loss_out")[2].clone()

    def _fused_ssm_step(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> None:
        """Propagation, lgcn_ssm_fwd_bwd (loss + gradient seed), Adam tick, Horner backward with Adam in
        the last epilogue — the fused BPR step with the loss kernel swapped."""
        from . import ops
        J = self.neg_size
        w = self.all_embedding.weight
        B = users.numel() // J
        group = self.optim.param_groups[0]
        st = self.optim._init_state(w)
        out = self._buf("OUT")
        self._propagate_into(w.data, out)
        self._reset_seed_buffers()
        G, cnt = self._buf("G"), self._buf("cnt")
        decay = float(self.config["decay"])
        ops.ssm_fwd_bwd(out, w.data, users, pos, neg, J, self.num_users, self.tau, decay, G, cnt, self._buf("loss_out"),
                        self._work(B), self._buf("work_counter"))
        ops.adam_tick(st["step"], st["hp"], group["lr"], group["betas"])
        adam = dict(exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"], hp=st["hp"], betas=group["betas"],
                    eps=group["eps"])
        self._horner_into(G, grad_mode=2, reg_coef=decay / B, cnt=cnt, adam=adam)
        if self.num_layers == 1:
            G.zero_()
        self._eval_cache_valid = False
        if self.prescale_emb and self._col_scale is None:
            self._ze_key = (w.data_ptr(), w._version)

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
                    aver = aver + self.stageOne(users[sl], pos[sl], neg[sl])   # fused kernel path unless ssm_autograd
            else:
                self._fused_step_eager(users[sl], pos[sl], neg[sl])
                aver = aver + self._buf("loss_out")[2]
        return aver / total_batch
