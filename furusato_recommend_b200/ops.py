"""Tensor-level wrappers over the C ABI (ctypes).  PyTorch is plumbing here:
device memory, the current stream and dtype checks.  Every op runs the CUDA
kernels of liblgcn_b200.so on the caller's current stream; CPU tensors are
rejected (there is no fallback)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from .graph import CsrGraph

MASK_VALUE = -float(1 << 10)  # reference trainer.py:137

# NVTX ranges around the phases of the hot path (the reference has only ad-hoc time() prints).
# Off by default — a range costs ~1 us of host time per call; LGCN_NVTX=1 turns them on for nsys / ncu --nvtx.
import contextlib
import os

NVTX = bool(int(os.environ.get("LGCN_NVTX", "0")))


def nvtx(name: str):
    return torch.cuda.nvtx.range(name) if NVTX else contextlib.nullcontext()


class _on:
    """Device guard + stream of one op: the launch goes to the device that holds the tensors (not to
    whatever device happens to be current — `config['device']='cuda:1'` works like in the reference)
    and to torch's current stream on THAT device.  Tensors on mixed devices are rejected."""

    def __init__(self, *tensors):
        devs = {t.device for t in tensors if torch.is_tensor(t)}
        if any(dv.type != "cuda" for dv in devs):
            raise _lib.LgcnLibraryError("CUDA tensors required: the LightGCN hot path has no CPU fallback")
        if len(devs) != 1:
            raise ValueError(f"all tensors of one op must live on one CUDA device, got {sorted(map(str, devs))}")
        self.device = devs.pop()
        self._guard = torch.cuda.device(self.device)

    def __enter__(self) -> int:
        self._guard.__enter__()
        return torch.cuda.current_stream(self.device).cuda_stream

    def __exit__(self, *exc):
        return self._guard.__exit__(*exc)


def _chk(t: Optional[torch.Tensor], dtype, name: str, allow_none: bool = False) -> int:
    if t is None:
        if allow_none:
            return 0
        raise ValueError(f"{name} is required")
    if not t.is_cuda:
        raise _lib.LgcnLibraryError(f"{name} must be a CUDA tensor: the LightGCN hot path has no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t.data_ptr()


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise TypeError(f"unsupported storage dtype {t.dtype}")


def _layer_args(N: int, d: int, src: Optional[torch.Tensor], *, scale_src: bool = False,
                dst: Optional[torch.Tensor] = None,
                base: Optional[torch.Tensor] = None, acc_in: Optional[torch.Tensor] = None,
                acc_out: Optional[torch.Tensor] = None, acc_scale: float = 1.0,
                grad_mode: int = 0, inv_layers: float = 1.0, reg_coef: float = 0.0,
                cnt: Optional[torch.Tensor] = None, emb: Optional[torch.Tensor] = None,
                grad: Optional[torch.Tensor] = None, adam_m: Optional[torch.Tensor] = None,
                adam_v: Optional[torch.Tensor] = None, adam_hp: Optional[torch.Tensor] = None,
                betas=(0.9, 0.999), eps: float = 1e-8, zero_base: bool = False,
                dst_peers: Optional[Sequence[int]] = None, dst_row_offset: int = 0,
                src_scale: Optional[torch.Tensor] = None, dst_scale: Optional[torch.Tensor] = None,
                edge_w: Optional[torch.Tensor] = None, nnz: int = 0, push_emb: bool = False, dst_multicast: int = 0,
                dst_route_rows: int = 0) -> "_lib.LayerArgs":
    """lgcn_layer_args_t for N output rows of width d (shared by lgcn_propagate_layer and lgcn_reduce_rows)."""
    a = _lib.LayerArgs()
    a.d, a.scale_src = d, int(scale_src)
    a.src_dtype = _dt(src) if src is not None else _lib.F32
    a.dst_dtype = _dt(dst) if dst is not None else a.src_dtype
    a.src = _chk(src, src.dtype, "src") if src is not None else 0
    a.dst = _chk(dst, dst.dtype, "dst") if dst is not None else 0
    a.base = _chk(base, torch.float32, "base", True)
    a.acc_in = _chk(acc_in, torch.float32, "acc_in", True)
    a.acc_out = _chk(acc_out, torch.float32, "acc_out", True)
    a.acc_scale = acc_scale
    a.grad_mode, a.inv_layers, a.reg_coef = grad_mode, inv_layers, reg_coef
    a.cnt = _chk(cnt, torch.int32, "cnt", True)
    a.emb = _chk(emb, torch.float32, "emb", True)
    a.grad = _chk(grad, torch.float32, "grad", True)
    a.adam_m = _chk(adam_m, torch.float32, "adam_m", True)
    a.adam_v = _chk(adam_v, torch.float32, "adam_v", True)
    a.adam_hp = _chk(adam_hp, torch.float32, "adam_hp", True)
    a.beta1, a.beta2, a.eps = betas[0], betas[1], eps
    a.zero_base = int(zero_base)
    a.push_emb = int(push_emb)
    a.dst_multicast = int(dst_multicast) or None
    a.dst_route_rows = int(dst_route_rows)
    if dst_multicast:
        a.dst_row_offset = dst_row_offset
    for t, nm in ((src_scale, "src_scale"), (dst_scale, "dst_scale")):
        if t is not None and (t.dim() != 1 or t.shape[0] < N):
            raise ValueError(f"{nm} must be [>={N}], got {tuple(t.shape)}")
    a.src_scale = _chk(src_scale, torch.float32, "src_scale", True)
    a.dst_scale = _chk(dst_scale, torch.float32, "dst_scale", True)
    if edge_w is not None and (edge_w.dim() != 1 or edge_w.numel() != nnz):
        raise ValueError(f"edge_w must be [{nnz}] (one weight per CSR slot), got {tuple(edge_w.shape)}")
    a.edge_w = _chk(edge_w, torch.float32, "edge_w", True)
    if dst_peers:
        if dst is None:
            raise ValueError("dst_peers needs dst (it fixes the dtype and enables the write)")
        a.n_dst_peers, a.dst_row_offset = len(dst_peers), dst_row_offset
        for i, ptr in enumerate(dst_peers):
            a.dst_peers[i] = ptr
    for t, nm in ((None if (dst_peers or dst_multicast) else dst, "dst"), (base, "base"), (acc_in, "acc_in"), (acc_out, "acc_out"), (emb, "emb"),
                  (grad, "grad"), (adam_m, "adam_m"), (adam_v, "adam_v")):
        if t is not None and (t.dim() != 2 or t.shape[0] < N or t.shape[1] != d):
            raise ValueError(f"{nm} must be [>={N}, {d}], got {tuple(t.shape)}")
    return a


def propagate_layer(g: CsrGraph, src: torch.Tensor, *, scale_src: bool, n_src_rows: int = 0, **kw) -> None:
    """lgcn_propagate_layer: one normalised-adjacency SpMM + fused row epilogue (keyword arguments:
    `_layer_args`).  `src_scale` / `dst_scale` ([N] fp32) replace the graph's dinv on the column / row side
    (rAdjGCN's asymmetric normalisation); None keeps D^-1/2 A D^-1/2.  `edge_w` ([nnz] fp32) weights
    every CSR slot (edge dropout).  `n_src_rows`: for a rectangular local graph (rows and columns are different
    node sets) the number of rows `src` must hold; by default the graph's own row count."""
    lib = _lib.load()
    n_src, d = src.shape
    N = g.n_nodes  # rows of this (possibly rank-local) graph; src may hold more rows (all-gathered)
    if n_src < (n_src_rows or N):
        raise ValueError(f"src has {n_src} rows, the graph's columns reach {n_src_rows or N}")
    a = _layer_args(N, d, src, scale_src=scale_src, nnz=g.nnz, **kw)
    with _on(src, kw.get("dst"), kw.get("base"), kw.get("acc_in"), kw.get("acc_out"), kw.get("emb"), kw.get("grad"),
             kw.get("adam_m"), kw.get("adam_v"), g.col) as st:
        _lib.check(lib.lgcn_propagate_layer(C.byref(g.c_struct(d)), C.byref(a), st), "lgcn_propagate_layer")


def reduce_rows(partials: torch.Tensor, n_rows: int, dinv: torch.Tensor, **kw) -> None:
    """lgcn_reduce_rows: s_i = sum_q partials[q, i] for i < n_rows, then lgcn_propagate_layer's row epilogue
    (keyword arguments: `_layer_args`) — the owner-side half of the reduce partition."""
    lib = _lib.load()
    if partials.dim() != 3 or partials.dtype != torch.float32 or partials.shape[1] < n_rows:
        raise ValueError("partials must be fp32 [n_parts, >= n_rows, d]")
    n_parts, part_rows, d = partials.shape
    if dinv.dim() != 1 or dinv.shape[0] < n_rows:
        raise ValueError("dinv must cover the reduced rows")
    a = _layer_args(n_rows, d, None, **kw)
    with _on(partials, dinv, kw.get("dst"), kw.get("base"), kw.get("acc_in"), kw.get("acc_out"), kw.get("emb"),
             kw.get("adam_m"), kw.get("adam_v")) as st:
        _lib.check(lib.lgcn_reduce_rows(_chk(partials, torch.float32, "partials"), n_parts, part_rows, n_rows,
                                        _chk(dinv, torch.float32, "dinv"), C.byref(a), st), "lgcn_reduce_rows")


def scale_rows_push(x: torch.Tensor, dinv: torch.Tensor, dst_dtype: torch.dtype, dst_peers: Sequence[int],
                    dst_row_offset: int) -> None:
    """dst_p[row_offset + i] = dinv[i] * x[i] on every peer buffer (pre-scaled first-layer source)."""
    lib = _lib.load()
    n, d = x.shape
    arr = (C.c_void_p * len(dst_peers))(*dst_peers)
    with _on(x, dinv) as st:
        _lib.check(lib.lgcn_scale_rows_push(_chk(x, torch.float32, "x"), _chk(dinv, torch.float32, "dinv"), n, d,
                                            _lib.BF16 if dst_dtype == torch.bfloat16 else _lib.F32, arr, len(dst_peers),
                                            dst_row_offset, st), "lgcn_scale_rows_push")


def exchange_rows_push(tab_a: torch.Tensor, tab_b: Optional[torch.Tensor], padded_ids: torch.Tensor,
                       rows_per_rank: int, rank: int, dst_peers: Sequence[int]) -> None:
    """Rows of the row-partitioned tables this rank owns -> row i of every peer's [n_ids, width] buffer."""
    lib = _lib.load()
    d = tab_a.shape[1]
    arr = (C.c_void_p * len(dst_peers))(*dst_peers)
    with _on(tab_a, tab_b, padded_ids) as st:
        _lib.check(lib.lgcn_exchange_rows_push(_chk(tab_a, torch.float32, "tab_a"), _chk(tab_b, torch.float32, "tab_b", True),
                                               d, _chk(padded_ids, torch.int64, "padded_ids"), padded_ids.numel(),
                                               rows_per_rank, rank, arr, len(dst_peers), st),
                   "lgcn_exchange_rows_push")


def bpr_fwd_bwd(out: torch.Tensor, emb: torch.Tensor, users: torch.Tensor, pos: torch.Tensor,
                neg: torch.Tensor, n_users: int, decay: float, G: torch.Tensor, cnt: torch.Tensor,
                loss_out: torch.Tensor, work: torch.Tensor, work_counter: torch.Tensor,
                loss_scale: float = 1.0) -> None:
    lib = _lib.load()
    B = users.numel()
    N, d = out.shape
    if work.numel() < 2 * B:
        raise ValueError("work must hold 2*B floats")
    if work_counter.numel() < 2:
        raise ValueError("work_counter must be int32[2]: {CTA arrival counter, skipped out-of-range samples}")
    with _on(out, emb, users, pos, neg, G, cnt, loss_out, work, work_counter) as st:
        _lib.check(lib.lgcn_bpr_fwd_bwd(
            _chk(out, torch.float32, "out"), _chk(emb, torch.float32, "emb"),
            _chk(users, torch.int64, "users"), _chk(pos, torch.int64, "pos"), _chk(neg, torch.int64, "neg"),
            B, n_users, N, d, decay, loss_scale, _chk(G, torch.float32, "G"), _chk(cnt, torch.int32, "cnt"),
            _chk(loss_out, torch.float32, "loss_out"), _chk(work, torch.float32, "work"),
            _chk(work_counter, torch.int32, "work_counter"), st), "lgcn_bpr_fwd_bwd")


def padded_ids(users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor, n_users: int, n_nodes: int,
               cuts: torch.Tensor, side_off: torch.Tensor, world: int, rows_per_rank: int, out: torch.Tensor,
               status: torch.Tensor) -> None:
    """out[3B] = padded partition ids of [users | n+pos | n+neg] (lgcn_padded_ids)."""
    lib = _lib.load()
    B = users.numel()
    if out.numel() < 3 * B:
        raise ValueError("out must hold 3*B ids")
    with _on(users, pos, neg, cuts, side_off, out, status) as st:
        _lib.check(lib.lgcn_padded_ids(_chk(users, torch.int64, "users"), _chk(pos, torch.int64, "pos"),
                                       _chk(neg, torch.int64, "neg"), B, n_users, n_nodes,
                                       _chk(cuts, torch.int64, "cuts"), _chk(side_off, torch.int64, "side_off"), world,
                                       rows_per_rank, _chk(out, torch.int64, "out"), status.data_ptr(), st),
                   "lgcn_padded_ids")


def bpr_fwd_bwd_rows(rows: torch.Tensor, ids: torch.Tensor, batch: int, rows_per_rank: int, rank: int, decay: float,
                     G: torch.Tensor, cnt: torch.Tensor, g0_full: Optional[torch.Tensor],
                     dinv_pad: Optional[torch.Tensor], loss_out: torch.Tensor, work: torch.Tensor,
                     work_counter: torch.Tensor, loss_scale: float = 1.0) -> None:
    """lgcn_bpr_fwd_bwd_rows: BPR on the owner-filled compact table rows[3B, 2d]."""
    lib = _lib.load()
    d = G.shape[1]
    if rows.dim() != 2 or rows.shape[1] != 2 * d or rows.shape[0] < 3 * batch or ids.numel() < 3 * batch:
        raise ValueError("rows must be [>=3B, 2d] and ids [>=3B]")
    if work.numel() < 2 * batch or work_counter.numel() < 2:
        raise ValueError("work must hold 2*B floats, work_counter 2 ints")
    with _on(rows, ids, G, cnt, g0_full, dinv_pad, loss_out, work, work_counter) as st:
        _lib.check(lib.lgcn_bpr_fwd_bwd_rows(
            _chk(rows, torch.float32, "rows"), _chk(ids, torch.int64, "ids"), batch, d, rows_per_rank, rank, decay,
            loss_scale, _chk(G, torch.float32, "G"), _chk(cnt, torch.int32, "cnt"),
            _chk(g0_full, torch.float32, "g0_full", True), _chk(dinv_pad, torch.float32, "dinv_pad", True),
            _chk(loss_out, torch.float32, "loss_out"), _chk(work, torch.float32, "work"),
            _chk(work_counter, torch.int32, "work_counter"), st), "lgcn_bpr_fwd_bwd_rows")


def zero_rows(table: torch.Tensor, ids: torch.Tensor) -> None:
    lib = _lib.load()
    with _on(table, ids) as st:
        _lib.check(lib.lgcn_zero_rows(_chk(table, torch.float32, "table"), table.shape[1],
                                      _chk(ids, torch.int64, "ids"), ids.numel(), st), "lgcn_zero_rows")


def ssm_fwd_bwd(out: torch.Tensor, emb: torch.Tensor, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor,
                n_neg: int, n_users: int, tau: float, decay: float, G: torch.Tensor, cnt: torch.Tensor,
                loss_out: torch.Tensor, work: torch.Tensor, work_counter: torch.Tensor, loss_scale: float = 1.0) -> None:
    """lgcn_ssm_fwd_bwd on flat triples (n_neg consecutive rows per (user, positive))."""
    lib = _lib.load()
    rows = users.numel()
    if rows % n_neg or pos.numel() != rows or neg.numel() != rows:
        raise ValueError(f"flat triples must hold a multiple of n_neg={n_neg} rows")
    B = rows // n_neg
    N, d = out.shape
    if work.numel() < 2 * B or work_counter.numel() < 2:
        raise ValueError("work must hold 2*B floats, work_counter 2 ints")
    with _on(out, emb, users, pos, neg, G, cnt, loss_out, work, work_counter) as st:
        _lib.check(lib.lgcn_ssm_fwd_bwd(
            _chk(out, torch.float32, "out"), _chk(emb, torch.float32, "emb"),
            _chk(users, torch.int64, "users"), _chk(pos, torch.int64, "pos"), _chk(neg, torch.int64, "neg"),
            B, n_neg, n_users, N, d, tau, decay, loss_scale, _chk(G, torch.float32, "G"), _chk(cnt, torch.int32, "cnt"),
            _chk(loss_out, torch.float32, "loss_out"), _chk(work, torch.float32, "work"),
            _chk(work_counter, torch.int32, "work_counter"), st), "lgcn_ssm_fwd_bwd")


def adam_tick(step: torch.Tensor, hp: torch.Tensor, lr: float, betas=(0.9, 0.999)) -> None:
    lib = _lib.load()
    with _on(step, hp) as st:
        _lib.check(lib.lgcn_adam_tick(_chk(step, torch.int64, "step"), _chk(hp, torch.float32, "adam_hp"),
                                      lr, betas[0], betas[1], st), "lgcn_adam_tick")


def adam_step(param: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, hp: torch.Tensor,
              betas=(0.9, 0.999), eps: float = 1e-8) -> None:
    lib = _lib.load()
    with _on(param, grad, m, v, hp) as st:
        _lib.check(lib.lgcn_adam_step(_chk(param, torch.float32, "param"), _chk(grad, torch.float32, "grad"),
                                      _chk(m, torch.float32, "m"), _chk(v, torch.float32, "v"), param.numel(),
                                      _chk(hp, torch.float32, "adam_hp"), betas[0], betas[1], eps, st),
                   "lgcn_adam_step")


def uniform_sample(pos_rowptr: torch.Tensor, pos_file: torch.Tensor, pos_sorted: torch.Tensor,
                   n_users: int, m_items: int, count: int, seed: int, epoch: int, first: int = 0,
                   n_neg: int = 1, pos_cdf: Optional[torch.Tensor] = None):
    """Returns (triples int64[count,3], valid uint8[count]) — not yet compacted.  `pos_cdf` (fp32,
    aligned with pos_file: per-user normalised inclusive cumulative probabilities) switches the
    positive pick from uniform to weighted (lgcn_uniform_sample_weighted)."""
    lib = _lib.load()
    dev = pos_rowptr.device
    triples = torch.empty((count * n_neg, 3), dtype=torch.int64, device=dev)
    valid = torch.empty(count * n_neg, dtype=torch.uint8, device=dev)
    if pos_cdf is not None and pos_cdf.numel() != pos_file.numel():
        raise ValueError("pos_cdf must align with pos_file")
    with _on(pos_rowptr, pos_file, pos_sorted, pos_cdf) as st:
        _lib.check(lib.lgcn_uniform_sample_weighted(
            _chk(pos_rowptr, torch.int64, "pos_rowptr"), _chk(pos_file, torch.int32, "pos_file"),
            _chk(pos_sorted, torch.int32, "pos_sorted"), _chk(pos_cdf, torch.float32, "pos_cdf", True),
            n_users, m_items, first, count, n_neg,
            seed & 0xFFFFFFFFFFFFFFFF, epoch & 0xFFFFFFFF, triples.data_ptr(), valid.data_ptr(), st),
            "lgcn_uniform_sample_weighted")
    return triples, valid


def compact_triples(triples: torch.Tensor, valid: torch.Tensor) -> torch.Tensor:
    """Order-preserving compaction; one host sync to learn the row count."""
    lib = _lib.load()
    count = valid.numel()
    dev = triples.device
    out = torch.empty_like(triples)
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)
    scratch = torch.empty((count + 1023) // 1024 + 1, dtype=torch.int64, device=dev)
    with _on(triples, valid) as st:
        _lib.check(lib.lgcn_compact_triples(_chk(triples, torch.int64, "triples"), _chk(valid, torch.uint8, "valid"),
                                            count, out.data_ptr(), n_out.data_ptr(), scratch.data_ptr(), st),
                   "lgcn_compact_triples")
    return out[: int(n_out.item())]


_WORKSPACES = {}


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """Grow-only scratch per device (bf16 operand tiles of the tensor-core scorer)."""
    t = _WORKSPACES.get(dev)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        _WORKSPACES[dev] = t
    return t


def score_topk(user_emb: torch.Tensor, item_emb: torch.Tensor, user_ids: torch.Tensor,
               pos_rowptr: torch.Tensor, pos_sorted: torch.Tensor, k: int,
               mask_value: float = MASK_VALUE, precision: str = "fp32", return_scores: bool = False):
    """Fused score + mask + top-k.  Returns (idx int32[U,k], val fp32[U,k]) and, with
    return_scores (tests, small shapes), the dense scores the selection saw."""
    lib = _lib.load()
    U = user_ids.numel()
    m, d = item_emb.shape
    dev = item_emb.device
    idx = torch.empty((U, k), dtype=torch.int32, device=dev)
    val = torch.empty((U, k), dtype=torch.float32, device=dev)
    prec = {"fp32": _lib.F32, "bf16": _lib.BF16, "f16": _lib.F16}[precision]
    nbytes = int(lib.lgcn_score_topk_workspace_bytes(U, m, d, prec))
    ws = _workspace(dev, nbytes) if nbytes else None
    args = [_chk(user_emb, torch.float32, "user_emb"), _chk(item_emb, torch.float32, "item_emb"),
            _chk(user_ids, torch.int64, "user_ids"), U, m, d, _chk(pos_rowptr, torch.int64, "pos_rowptr"),
            _chk(pos_sorted, torch.int32, "pos_sorted"), k, mask_value, prec, idx.data_ptr(), val.data_ptr(),
            ws.data_ptr() if ws is not None else 0, nbytes]
    with _on(user_emb, item_emb, user_ids, pos_rowptr, pos_sorted) as st:
        if return_scores:
            dense = torch.empty((U, m), dtype=torch.float32, device=dev)
            _lib.check(lib.lgcn_score_topk_debug(*args, dense.data_ptr(), st), "lgcn_score_topk_debug")
            return idx, val, dense
        _lib.check(lib.lgcn_score_topk(*args, st), "lgcn_score_topk")
    return idx, val


def score_dense_f32(user_emb: torch.Tensor, item_emb: torch.Tensor, user_ids: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    U = user_ids.numel()
    m, d = item_emb.shape
    out = torch.empty((U, m), dtype=torch.float32, device=item_emb.device)
    with _on(user_emb, item_emb, user_ids) as st:
        _lib.check(lib.lgcn_score_dense_f32(_chk(user_emb, torch.float32, "user_emb"),
                                            _chk(item_emb, torch.float32, "item_emb"),
                                            _chk(user_ids, torch.int64, "user_ids"), U, m, d, out.data_ptr(),
                                            st), "lgcn_score_dense_f32")
    return out


def rank_metrics(topk: torch.Tensor, user_ids: torch.Tensor, test_rowptr: torch.Tensor,
                 test_sorted: torch.Tensor, ks: Sequence[int], sums: Optional[torch.Tensor] = None,
                 want_hits: bool = False):
    """Adds the recall/precision/hr/ndcg SUMS of this batch into sums[4, len(ks)] (fp64)."""
    lib = _lib.load()
    U, k = topk.shape
    dev = topk.device
    ks = [int(x) for x in ks]
    order = sorted(range(len(ks)), key=lambda i: ks[i])
    ks_sorted = [ks[i] for i in order]
    tmp = torch.zeros((4, len(ks)), dtype=torch.float64, device=dev)
    hits = torch.empty((U, k), dtype=torch.uint8, device=dev) if want_hits else None
    arr = (C.c_int32 * len(ks))(*ks_sorted)
    with _on(topk, user_ids, test_rowptr, test_sorted) as st:
        _lib.check(lib.lgcn_rank_metrics(_chk(topk, torch.int32, "topk"), U, k, _chk(user_ids, torch.int64, "user_ids"),
                                         _chk(test_rowptr, torch.int64, "test_rowptr"),
                                         _chk(test_sorted, torch.int32, "test_sorted"), arr, len(ks),
                                         tmp.data_ptr(), hits.data_ptr() if hits is not None else 0, st),
                   "lgcn_rank_metrics")
    inv = torch.empty(len(ks), dtype=torch.long)
    for pos_sorted_i, orig_i in enumerate(order):
        inv[orig_i] = pos_sorted_i
    tmp = tmp[:, inv.to(dev)]
    if sums is None:
        sums = tmp
    else:
        sums += tmp
    return (sums, hits) if want_hits else sums


def ingest_text(text: torch.Tensor):
    """lgcn_ingest_*: raw bytes of a `uid item item ...\\n` file (uint8 CUDA tensor) -> (user int64[n],
    item int64[n]) in file order, duplicates kept (reference dataloader.py:93-124).  Two host syncs:
    the interaction count (to size the outputs) and the error flag."""
    lib = _lib.load()
    n_bytes = text.numel()
    dev = text.device
    n_tiles = int(lib.lgcn_ingest_tiles(n_bytes))
    tile_tok = torch.zeros(max(n_tiles, 1), dtype=torch.int64, device=dev)
    tile_uid = torch.zeros(max(n_tiles, 1), dtype=torch.int64, device=dev)
    totals = torch.zeros(2, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    with _on(text, tile_tok, tile_uid, totals, err) as st:
        _lib.check(lib.lgcn_ingest_count(_chk(text, torch.uint8, "text"), n_bytes, tile_tok.data_ptr(), tile_uid.data_ptr(),
                                         totals.data_ptr(), err.data_ptr(), st), "lgcn_ingest_count")
        n_tok, n_uid = (int(x) for x in totals.tolist())
        n = n_tok - n_uid
        line_uid = torch.empty(max(n_uid, 1), dtype=torch.int64, device=dev)
        user = torch.empty(n, dtype=torch.int64, device=dev)
        item = torch.empty(n, dtype=torch.int64, device=dev)
        line_of = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        _lib.check(lib.lgcn_ingest_emit(text.data_ptr(), n_bytes, tile_tok.data_ptr(), tile_uid.data_ptr(),
                                        line_uid.data_ptr(), n_uid, user.data_ptr(), item.data_ptr(), line_of.data_ptr(),
                                        n, err.data_ptr(), st), "lgcn_ingest_emit")
    flag = int(err.item())
    if flag & 1:
        raise ValueError("interaction file holds a byte that is not a digit or whitespace (the reference's int() raises)")
    if flag & 6:
        raise ValueError(f"malformed interaction file (ingest error flag {flag})")
    return user, item
