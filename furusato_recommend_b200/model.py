"""LightGCN with the reference's model API on hand-written sm_100a kernels.

Same duck-typed surface as reference model/lgcn.py:44-151 (`forward()`,
`getEmbedding`, `bpr_loss`, `getUsersRating`, `stageOne`, `OneEpoch`, owns
`self.optim`, state-dict key `all_embedding.weight` with users first) plus
`computer()` — the name the same computation has on the legacy torch.sparse class
(model/MF.py:178-210) — and the fused, never-materialised eval entry
`getUsersTopK()`.

Two ways through the training step:
  * `stageOne()` / `OneEpoch()`: the fused path.  2K SpMM launches + 1 BPR launch
    + 1 tiny Adam tick; the layer mean, the Horner backward, the L2 term, the
    G/cnt reset and Adam all live in kernel epilogues.  No autograd involved.
  * `bpr_loss()` + `.backward()` + `optim.step()`: autograd-compatible for
    callers that drive the optimizer themselves (reference ddp_lgcn.py:498-504);
    the same kernels behind `torch.ops.lgcn_b200.propagate` (torch_ops.py, a registered
    custom op with `register_autograd`) and a `torch.autograd.Function` for the loss.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from . import ops
from .dataloader import BasicDataset
from .graph import CsrGraph

_STORAGE = {"fp32": torch.float32, "bf16": torch.bfloat16}


class FusedAdam(torch.optim.Optimizer):
    """Adam (default betas/eps, no weight decay — reference model/lgcn.py:63) whose
    state lives in flat device buffers the fused kernels update in place.  `step()`
    applies lgcn_adam_step to `.grad` for the autograd path."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.on_step = None   # callable: the kernels write the parameters through raw pointers

    def _init_state(self, p: torch.Tensor) -> dict:
        st = self.state[p]
        if not st:
            st["step"] = torch.zeros(1, dtype=torch.int64, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["hp"] = torch.zeros(2, dtype=torch.float32, device=p.device)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self._init_state(p)
                ops.adam_tick(st["step"], st["hp"], group["lr"], group["betas"])
                ops.adam_step(p.data, p.grad.contiguous(), st["exp_avg"], st["exp_avg_sq"], st["hp"],
                              group["betas"], group["eps"])
        if self.on_step is not None:
            self.on_step()
        return loss


class _BprFn(torch.autograd.Function):
    """(loss, reg) = bpr_loss(E; users,pos,neg) with the closed-form backward."""

    @staticmethod
    def forward(ctx, weight: torch.Tensor, model: "LightGCN", users, pos, neg):
        m = model
        m._propagate_into(weight.detach(), m._buf("OUT"))
        m._eval_cache_valid = False
        m._reset_seed_buffers()
        lo = m._buf("loss_out")
        ops.bpr_fwd_bwd(m._buf("OUT"), weight.detach(), users, pos, neg, m.num_users, 0.0,
                        m._buf("G"), m._buf("cnt"), lo, m._work(users.numel()), m._buf("work_counter"))
        m._g_clean = False
        m._seed_generation += 1
        ctx.model, ctx.gen, ctx.batch = m, m._seed_generation, users.numel()
        ctx.drop_bwd = m._drop_bwd
        m._raise_on_bad_ids()   # eager path: one sync, the reference's IndexError (model/lgcn.py:90-95)
        return lo[0].clone(), lo[1].clone()

    @staticmethod
    def backward(ctx, g_loss: torch.Tensor, g_reg: torch.Tensor):
        m: "LightGCN" = ctx.model
        if ctx.gen != m._seed_generation:
            raise RuntimeError("bpr_loss() was called again before backward(): the gradient seed is stale")
        gl, gr = float(g_loss), float(g_reg)
        G = m._buf("G")
        if gl != 1.0:
            G.mul_(gl)
        grad = torch.empty_like(m.all_embedding.weight)
        m._horner_into(G, grad_mode=1, reg_coef=gr / ctx.batch, grad=grad, cnt=m._buf("cnt"), edge_w=ctx.drop_bwd)
        return grad, None, None, None, None


class LightGCN(nn.Module):
    def __init__(self, config: dict, dataset: BasicDataset):
        super().__init__()
        self.config = config
        self.dataset = dataset
        self.num_users = dataset.n_users
        self.num_items = dataset.m_items
        # reference keys (model/lgcn.py:51-52) with the legacy aliases (model/MF.py:126-127)
        self.latent_dim = int(config["recdim"] if "recdim" in config else config["latent_dim_rec"])
        self.num_layers = int(config["layer"] if "layer" in config else config["lightGCN_n_layers"])
        if self.num_layers < 1:
            raise NotImplementedError("layer=0 is plain matrix factorisation, outside the LightGCN hot path")
        self.device = torch.device(config.get("device", "cuda:0"))
        self.storage_dtype = _STORAGE[config.get("storage_dtype", "fp32")]
        # full-rank eval scores on the tcgen05 tensor cores by default (bf16 operands, fp32 accumulate,
        # 2e-2 relative to the fp32 scores, north_star); "fp32" is the exactness mode (CUDA cores)
        self.eval_precision = config.get("eval_precision", "bf16")
        self.graph: CsrGraph = self._build_graph(dataset)
        if self.graph.device != self.device:
            self.graph = self.graph.to(self.device)
        # (column-side, row-side) normalisation vectors of the propagation operator
        # diag(row) A diag(col); None = the graph's dinv on both sides (D^-1/2 A D^-1/2)
        self._col_scale: Optional[torch.Tensor] = None
        self._row_scale: Optional[torch.Tensor] = None
        # edge dropout (model/MF.py:158-192; parse.py:16-17 `--dropout`, `--keepprob`): off by default
        self.dropout = bool(config.get("dropout", 0))
        self.keep_prob = float(config.get("keep_prob", 0.6))
        self._drop_fwd: Optional[torch.Tensor] = None
        self._drop_bwd: Optional[torch.Tensor] = None
        self.dropout_mask_fn = None   # tests: callable(n_entries) -> bool mask, instead of the device RNG
        self.__init_weight()
        self.optim = FusedAdam(self.parameters(), lr=config["lr"])  # model/lgcn.py:63
        self._bufs = {}
        self._g_clean = True
        self._seed_generation = 0
        self._eval_cache_valid = False
        self._eval_cache_key = None
        self.use_cuda_graph = bool(config.get("cuda_graph", True))
        self._graph = None
        self._graph_key = None
        # bf16 storage: the Adam epilogue of a fused step also emits ZE = bf16(dinv (.) E_new), the
        # pre-scaled layer-0 source of the NEXT propagation (push_emb), so every forward layer gathers
        # 2-byte elements.  fp32 storage keeps gathering E itself with dinv[j] applied on the fly
        # (bit-stable against the round-1 goldens).  Symmetric normalisation only.
        self.prescale_emb = bool(config.get("prescale_emb", self.storage_dtype == torch.bfloat16))
        self._ze_key = None
        self.optim.on_step = self._invalidate_derived

    def _build_graph(self, dataset: BasicDataset) -> CsrGraph:
        return dataset.csr_graph()

    def _scales(self, transpose: bool) -> dict:
        """Scale vectors of one layer: forward diag(row) A diag(col), backward its transpose."""
        if self._col_scale is None:
            return {}
        if transpose:
            return dict(src_scale=self._row_scale, dst_scale=self._col_scale)
        return dict(src_scale=self._col_scale, dst_scale=self._row_scale)

    def _sample_dropout(self) -> None:
        """One Bernoulli(keep_prob) draw per coalesced entry of A_hat, rescaled by 1/keep_prob — the
        reference's `__dropout_x` (model/MF.py:158-166: `(rand + keep_prob).int().bool()`), drawn once
        per propagation and shared by its K layers (MF.py:187-205).  The (i, j) and (j, i) entries
        are dropped independently, so the backward pass uses the weights of the reverse entries."""
        ent, n_ent, rev = self.graph.entry_index()
        if self.dropout_mask_fn is not None:
            m = torch.as_tensor(self.dropout_mask_fn(n_ent)).to(device=self.graph.device, dtype=torch.bool)
        else:
            m = (torch.rand(n_ent, device=self.graph.device) + self.keep_prob).int().bool()
        w = m.to(torch.float32) / self.keep_prob
        self._drop_fwd, self._drop_bwd = w[ent].contiguous(), w[rev].contiguous()

    def train(self, mode: bool = True):
        # weights only move in training mode; eval mode may reuse one propagation
        # for every user batch (the reference re-propagates per batch, model/lgcn.py:121)
        self._eval_cache_valid = False
        return super().train(mode)

    def load_state_dict(self, *args, **kwargs):
        self._eval_cache_valid = False
        return super().load_state_dict(*args, **kwargs)

    def __init_weight(self):
        # model/lgcn.py:70-76: one table, users first, N(0, 0.1^2)
        self.all_embedding = nn.Embedding(self.num_users + self.num_items, self.latent_dim,
                                          device=self.device)
        nn.init.normal_(self.all_embedding.weight, std=0.1)

    # ------------------------------------------------------------------ buffers
    def _buf(self, name: str) -> torch.Tensor:
        t = self._bufs.get(name)
        if t is None:
            N, d, dev = self.num_users + self.num_items, self.latent_dim, self.all_embedding.weight.device
            if name in ("Z0", "Z1", "ZE"):
                t = torch.empty((N, d), dtype=self.storage_dtype, device=dev)
            elif name in ("ACC", "OUT"):
                t = torch.empty((N, d), dtype=torch.float32, device=dev)
            elif name == "G":
                t = torch.zeros((N, d), dtype=torch.float32, device=dev)
            elif name in ("cnt", "zero_cnt"):
                t = torch.zeros(N, dtype=torch.int32, device=dev)
            elif name == "loss_out":
                t = torch.zeros(4, dtype=torch.float32, device=dev)
            elif name == "work_counter":   # {CTA arrival counter, skipped out-of-range samples}
                t = torch.zeros(2, dtype=torch.int32, device=dev)
            else:
                raise KeyError(name)
            self._bufs[name] = t
        return t

    def _work(self, batch: int) -> torch.Tensor:
        t = self._bufs.get("work")
        if t is None or t.numel() < 2 * batch:
            t = torch.empty(2 * max(batch, 1), dtype=torch.float32, device=self.all_embedding.weight.device)
            self._bufs["work"] = t
        return t

    def _zero_cnt(self) -> torch.Tensor:
        return self._buf("zero_cnt")

    def _invalidate_derived(self) -> None:
        """The table changed behind torch's version counter (raw-pointer kernels): drop what was
        derived from it."""
        self._eval_cache_valid = False
        self._ze_key = None

    def _ze(self, emb: torch.Tensor) -> Optional[torch.Tensor]:
        """ZE = storage_dtype(dinv (.) E) for the CURRENT table, or None when the mode is off.  Normally
        left behind by the previous fused step's Adam epilogue; rebuilt by one small kernel otherwise."""
        w = self.all_embedding.weight
        if not self.prescale_emb or self._col_scale is not None or emb.data_ptr() != w.data_ptr():
            return None
        ze = self._buf("ZE")
        key = (w.data_ptr(), w._version)
        if self._ze_key != key:
            ops.scale_rows_push(w.detach(), self.graph.dinv, self.storage_dtype, [ze.data_ptr()], 0)
            self._ze_key = key
        return ze

    def _raise_on_bad_ids(self) -> None:
        """The BPR kernel skips (and counts) samples whose ids fall outside the table instead of
        scattering out of bounds; surface them as the reference's IndexError.  One host sync."""
        wc = self._buf("work_counter")
        bad = int(wc[1].item())
        if bad:
            wc[1] = 0
            raise IndexError(f"{bad} (user, pos, neg) sample(s) with an id outside [0, {self.num_users}) / "
                             f"[0, {self.num_items}) were skipped (reference: IndexError at model/lgcn.py:90-95)")

    def check_ids(self) -> None:
        """Raise the reference's IndexError (model/lgcn.py:90-95) if any step since the last call saw a
        (user, pos, neg) id outside the table.  One host sync; OneEpoch and bpr_loss call it themselves."""
        self._raise_on_bad_ids()

    def _check_host_ids(self, users, pos, neg) -> None:
        """Ids that arrive on the host are range-checked there (cheap), before anything is launched."""
        for t, hi, nm in ((users, self.num_users, "user"), (pos, self.num_items, "positive item"),
                          (neg, self.num_items, "negative item")):
            if torch.is_tensor(t) and not t.is_cuda and t.numel() and (int(t.min()) < 0 or int(t.max()) >= hi):
                raise IndexError(f"{nm} id out of range [0, {hi})")

    def _reset_seed_buffers(self) -> None:
        if not self._g_clean:
            self._buf("G").zero_()
            self._buf("cnt").zero_()
            self._g_clean = True

    # ------------------------------------------------------------- propagation
    def _propagate_into(self, emb: torch.Tensor, out: torch.Tensor) -> None:
        """out = (X0 + .. + XK)/(K+1), X0 = emb, X_{k+1} = A_hat X_k  (model/lgcn.py:78-86)."""
        K, g = self.num_layers, self.graph
        z = [self._buf("Z0"), self._buf("Z1")]
        acc = self._buf("ACC")
        if self.dropout and self.training:
            self._sample_dropout()
        else:
            self._drop_fwd = self._drop_bwd = None
        ze = self._ze(emb)
        for k in range(K):
            last = k == K - 1
            ops.propagate_layer(
                g, (emb if ze is None else ze) if k == 0 else z[(k - 1) & 1], scale_src=(k == 0 and ze is None),
                dst=None if last else z[k & 1],
                acc_in=emb if k == 0 else acc, acc_out=out if last else acc,
                acc_scale=1.0 / (K + 1) if last else 1.0, edge_w=self._drop_fwd, **self._scales(False))

    def _horner_into(self, G: torch.Tensor, *, grad_mode: int, reg_coef: float, cnt: torch.Tensor,
                     grad: Optional[torch.Tensor] = None, adam: Optional[dict] = None, edge_w="own") -> None:
        """H0 = G, H_{j+1} = G + A_hat H_j; result H_K/(K+1) + reg_coef*cnt*E goes to
        `grad` (mode 1) or straight into Adam (mode 2)  (SURVEY §8 a-3)."""
        K, g = self.num_layers, self.graph
        z = [self._buf("Z0"), self._buf("Z1")]
        w = self.all_embedding.weight.detach()
        if isinstance(edge_w, str):   # the dropout weights of the latest propagation (fused step)
            edge_w = self._drop_bwd
        for j in range(K):
            last = j == K - 1
            kw = {}
            if last:
                kw = dict(grad_mode=grad_mode, inv_layers=1.0 / (K + 1), reg_coef=reg_coef, cnt=cnt, emb=w,
                          grad=grad)
                if adam is not None:
                    kw.update(adam_m=adam["exp_avg"], adam_v=adam["exp_avg_sq"], adam_hp=adam["hp"],
                              betas=adam["betas"], eps=adam["eps"], zero_base=K > 1)
                    if self.prescale_emb and self._col_scale is None:
                        kw.update(dst=self._buf("ZE"), push_emb=True)   # next step's layer-0 source
            kw.setdefault("dst", None if last else z[j & 1])
            ops.propagate_layer(g, G if j == 0 else z[(j - 1) & 1], scale_src=(j == 0), base=G, edge_w=edge_w, **kw,
                                **self._scales(True))

    def computer(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Propagated (users, items) embeddings — model/MF.py:178-210 naming."""
        w = self.all_embedding.weight
        if torch.is_grad_enabled() and w.requires_grad:
            # the registered custom op (torch_ops.py): torch.ops.lgcn_b200.propagate, backward = the same operator
            from . import torch_ops
            out = torch_ops.propagate(w, torch_ops.register_model(self))
        else:
            out = self._buf("OUT")
            # The eval-mode cache is keyed on the table's storage and autograd version, so in-place
            # edits made through torch (`weight.copy_`, optimizer steps on the autograd path,
            # load_state_dict) invalidate it; the fused kernels write through raw pointers and drop
            # the flag themselves.
            key = (w.data_ptr(), w._version)
            if self.training or not self._eval_cache_valid or self._eval_cache_key != key:
                self._propagate_into(w.detach(), out)
                self._eval_cache_valid = not self.training
                self._eval_cache_key = key
        return out[: self.num_users], out[self.num_users:]

    def forward(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """model/lgcn.py:78-86"""
        return self.computer()

    # ------------------------------------------------------------------ losses
    def getEmbedding(self, users, pos_items, neg_items):
        """model/lgcn.py:88-96"""
        all_users, all_items = self.forward()
        users, pos_items, neg_items = users.long(), pos_items.long(), neg_items.long()
        return (all_users[users], all_items[pos_items], all_items[neg_items],
                self.all_embedding(users), self.all_embedding(pos_items + self.num_users),
                self.all_embedding(neg_items + self.num_users))

    def bpr_loss(self, users, pos, neg):
        """(loss, reg_loss) of model/lgcn.py:98-118, differentiable wrt all_embedding."""
        users, pos, neg = self._ids(users), self._ids(pos), self._ids(neg)
        return _BprFn.apply(self.all_embedding.weight, self, users, pos, neg)

    def _ids(self, t, keep_host: bool = False) -> torch.Tensor:
        if not torch.is_tensor(t):
            t = torch.as_tensor(t)
        if keep_host and not t.is_cuda and self.use_cuda_graph and t.numel() == int(self.config["bpr_batch_size"]):
            return t.to(torch.int64).contiguous()   # copied straight into the graph's static batch buffers
        return t.to(device=self.all_embedding.weight.device, dtype=torch.int64).contiguous()

    @torch.no_grad()
    def stageOne(self, user, pos, neg) -> torch.Tensor:
        """model/lgcn.py:127-133 as one fused step: zero_grad + bpr_loss +
        decay*reg + backward + Adam.  Returns loss + decay*reg (0-dim tensor)."""
        u, p, q = self._ids(user, keep_host=True), self._ids(pos, keep_host=True), self._ids(neg, keep_host=True)
        if self.config.get("check_ids", False):   # six host reductions (~30 us): off on the hot path; the kernel
            self._check_host_ids(u, p, q)         # skips and counts bad ids anyway and the loss comes back NaN
        self._fused_step(u, p, q)
        return self._buf("loss_out")[2].clone()   # NaN if an id was out of range; check_ids() raises the IndexError

    def _fused_step(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> None:
        """One fused train step.  Full-size batches replay a captured CUDA graph (8 kernels, no
        per-launch host work); anything else (last partial batch, cuda_graph=False) launches eagerly."""
        B = users.numel()
        if not self.use_cuda_graph or self.dropout or B != int(self.config["bpr_batch_size"]):
            # (a fresh dropout mask per step is drawn with torch ops: not part of the captured graph)
            return self._fused_step_eager(self._ids(users), self._ids(pos), self._ids(neg))
        self._reset_seed_buffers()  # only dirty after the autograd path; the graph assumes clean G/cnt
        lr = self.optim.param_groups[0]["lr"]
        if self._graph is None or self._graph_key != self._step_graph_key(B):
            self._capture_step_graph(B)
        if not (users.is_cuda or pos.is_cuda or neg.is_cuda):
            # host triples (trainer.py:63 style callers): gather the three id arrays in ONE pinned staging
            # buffer and ship them with one H2D copy instead of three
            if self._hstage_ev is not None:
                self._hstage_ev.synchronize()      # the previous step's copy has left the staging buffer
            for k, src in enumerate((users, pos, neg)):
                self._hstage[k].copy_(src)
            self._gbatch_all.copy_(self._hstage, non_blocking=True)
            if self._hstage_ev is None:
                self._hstage_ev = torch.cuda.Event()
            self._hstage_ev.record()
        else:
            for dst, src in zip(self._gbatch, (users, pos, neg)):
                dst.copy_(src, non_blocking=True)  # device slices
        self._ze(self.all_embedding.weight)        # the captured layer 0 reads ZE: make sure it is current
        self._graph.replay()
        self._eval_cache_valid = False

    def _capture_step_graph(self, B: int) -> None:
        dev = self.all_embedding.weight.device
        self._gbatch_all = torch.zeros((3, B), dtype=torch.int64, device=dev)
        self._gbatch = [self._gbatch_all[k] for k in range(3)]
        self._hstage = torch.zeros((3, B), dtype=torch.int64).pin_memory()
        self._hstage_ev = None
        st = self.optim._init_state(self.all_embedding.weight)
        # the warm-up step below must not move the model: snapshot and restore every mutable buffer
        keep = [t.clone() for t in (self.all_embedding.weight.data, st["exp_avg"], st["exp_avg_sq"], st["step"], st["hp"],
                                    self._buf("loss_out"))]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._fused_step_eager(*self._gbatch)   # allocates every lazily created buffer
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._fused_step_eager(*self._gbatch)
        for t, k in zip((self.all_embedding.weight.data, st["exp_avg"], st["exp_avg_sq"], st["step"], st["hp"],
                         self._buf("loss_out")), keep):
            t.copy_(k)
        self._graph = graph
        self._graph_key = self._step_graph_key(B)

    def _step_graph_key(self, B: int):
        """Everything the captured step bakes in: scalars AND the addresses of the table and of the
        optimizer state (optim.load_state_dict replaces exp_avg / exp_avg_sq)."""
        st = self.optim._init_state(self.all_embedding.weight)
        return (B, self.optim.param_groups[0]["lr"], float(self.config["decay"]), self.num_layers,
                self.all_embedding.weight.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                st["step"].data_ptr(), st["hp"].data_ptr())

    def _fused_step_eager(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> None:
        w = self.all_embedding.weight
        B = users.numel()
        group = self.optim.param_groups[0]
        st = self.optim._init_state(w)
        out = self._buf("OUT")
        with ops.nvtx("lgcn.propagate"):
            self._propagate_into(w.data, out)
        self._reset_seed_buffers()
        G, cnt = self._buf("G"), self._buf("cnt")
        decay = float(self.config["decay"])
        with ops.nvtx("lgcn.bpr"):
            ops.bpr_fwd_bwd(out, w.data, users, pos, neg, self.num_users, decay, G, cnt, self._buf("loss_out"),
                            self._work(B), self._buf("work_counter"))
            ops.adam_tick(st["step"], st["hp"], group["lr"], group["betas"])
        adam = dict(exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"], hp=st["hp"], betas=group["betas"],
                    eps=group["eps"])
        with ops.nvtx("lgcn.backward+adam"):
            self._horner_into(G, grad_mode=2, reg_coef=decay / B, cnt=cnt, adam=adam)
        if self.num_layers == 1:  # the single layer gathers from G, so it cannot clear it in flight
            G.zero_()
        self._eval_cache_valid = False
        if self.prescale_emb and self._col_scale is None:
            self._ze_key = (w.data_ptr(), w._version)   # ZE matches the updated table

    @torch.no_grad()
    def OneEpoch(self, user, pos, neg) -> torch.Tensor:
        """model/lgcn.py:135-151: contiguous mini-batches of bpr_batch_size, the
        epoch loss is sum / (len // B + 1) (sic)."""
        users, pos, neg = self._ids(user), self._ids(pos), self._ids(neg)
        B = int(self.config["bpr_batch_size"])
        total_batch = len(users) // B + 1
        lo = self._buf("loss_out")
        lo[3] = 0.0
        for i in range(0, len(users), B):
            self._fused_step(users[i:i + B], pos[i:i + B], neg[i:i + B])
        self._raise_on_bad_ids()   # one sync per epoch
        return lo[3] / total_batch

    # -------------------------------------------------------------------- eval
    @torch.no_grad()
    def getUsersRating(self, users) -> torch.Tensor:
        """model/lgcn.py:120-125: dense raw scores [U_b, m] (no sigmoid).  Kept for
        Trainer.test compatibility; reuses the cached propagation in eval mode.
        The dense product is a plain library GEMM — the fused path is getUsersTopK."""
        all_users, all_items = self.computer()
        return torch.matmul(all_users[self._ids(users)], all_items.t())

    @torch.no_grad()
    def getUsersTopK(self, users, k: int, mask_train: bool = True, precision: Optional[str] = None):
        """score -> mask train positives with -1024 -> top-k (ties: lowest id), fused
        (model/lgcn.py:120-125 + trainer.py:132-138).  Returns (idx int32[U,k], val fp32[U,k])."""
        all_users, all_items = self.computer()
        rowptr, _, srt = self.dataset.pos_csr()
        if not mask_train:
            rowptr = torch.zeros_like(rowptr)
        with ops.nvtx("lgcn.score_topk"):
            return ops.score_topk(all_users, all_items, self._ids(users), rowptr, srt, k,
                                  precision=precision or self.eval_precision)
