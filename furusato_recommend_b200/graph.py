"""CSR graph of the bipartite adjacency + the SpMM work decomposition.

Replaces `Loader.getSparseGraph` (reference dataloader.py:215-258): instead of a
coalesced COO FloatTensor holding D^-1/2 A D^-1/2 we keep the structure only
(int64 rowptr, int32 col) and fold the normalisation into `dinv`
(dataloader.py:236-238).  Built with torch ops on whatever device the edge list
lives on (sort + bincount + cumsum), so cfg-3-sized graphs never touch the host.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import _lib


@dataclass
class CsrGraph:
    n_users: int
    m_items: int
    rowptr: torch.Tensor      # int64 [N+1]
    col: torch.Tensor         # int32 [nnz]
    dinv: torch.Tensor        # fp32  [N]
    light_rows: torch.Tensor  # int32 [n_light], degree-descending
    light_desc: torch.Tensor  # int32 [n_light, 4] = {row, degree, first edge lo, hi}
    seg_row: torch.Tensor     # int32 [n_seg]
    seg_begin: torch.Tensor   # int64 [n_seg]
    seg_len: torch.Tensor     # int32 [n_seg]
    seg_hub: torch.Tensor     # int32 [n_seg]
    hub_seg0: torch.Tensor    # int32 [n_hub]
    hub_nseg: torch.Tensor    # int32 [n_hub]
    hub_counter: torch.Tensor  # int32 [n_hub], zero
    partial: Optional[torch.Tensor] = None  # fp32 [n_seg * d]
    _struct: Optional[_lib.GraphStruct] = field(default=None, repr=False)

    @property
    def n_nodes(self) -> int:
        return self.n_users + self.m_items

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    @property
    def device(self) -> torch.device:
        return self.col.device

    def to(self, device) -> "CsrGraph":
        kw = {}
        for f in ("rowptr", "col", "dinv", "light_rows", "light_desc", "seg_row", "seg_begin", "seg_len",
                  "seg_hub", "hub_seg0", "hub_nseg", "hub_counter"):
            kw[f] = getattr(self, f).to(device)
        return CsrGraph(self.n_users, self.m_items, **kw)

    def entry_index(self):
        """(entry_of_slot int64[nnz], n_entries, rev_entry int64[nnz]).  An ENTRY is a coalesced
        (row, col) pair of A_hat — the unit the reference's edge dropout keeps or drops
        (model/MF.py:158-166 works on the coalesced COO values, row-major like this CSR); repeated
        slots of a multi-edge share their entry.  rev_entry[e] is the entry of (col, row): the
        weight the transposed (backward) gather needs at slot e."""
        if getattr(self, "_entry_index", None) is None:
            N = self.n_nodes
            deg = self.rowptr[1:] - self.rowptr[:-1]
            row = torch.repeat_interleave(torch.arange(N, device=self.device), deg)
            col = self.col.to(torch.int64)
            uniq, inverse = torch.unique_consecutive(row * N + col, return_inverse=True)
            rev = torch.searchsorted(uniq, col * N + row)   # exists: A is symmetric
            self._entry_index = (inverse.contiguous(), int(uniq.numel()), rev.contiguous())
        return self._entry_index

    def c_struct(self, d: int) -> _lib.GraphStruct:
        """lgcn_graph_t for embedding width d (allocates the hub scratch once)."""
        need = max(1, int(self.seg_row.numel()) * d)
        if self.partial is None or self.partial.numel() < need:
            self.partial = torch.empty(need, dtype=torch.float32, device=self.device)
            self._struct = None
        if self._struct is None:
            s = _lib.GraphStruct()
            s.n_nodes, s.nnz = self.n_nodes, self.nnz
            s.rowptr, s.col, s.dinv = self.rowptr.data_ptr(), self.col.data_ptr(), self.dinv.data_ptr()
            s.light_desc, s.n_light = self.light_desc.data_ptr(), int(self.light_rows.numel())
            s.seg_row, s.seg_begin = self.seg_row.data_ptr(), self.seg_begin.data_ptr()
            s.seg_len, s.seg_hub = self.seg_len.data_ptr(), self.seg_hub.data_ptr()
            s.n_seg = int(self.seg_row.numel())
            s.hub_seg0, s.hub_nseg = self.hub_seg0.data_ptr(), self.hub_nseg.data_ptr()
            s.hub_counter, s.n_hub = self.hub_counter.data_ptr(), int(self.hub_seg0.numel())
            s.partial = self.partial.data_ptr()
            self._struct = s
        return self._struct


def decompose_rows(rowptr: torch.Tensor, hub_deg: int = _lib.HUB_DEG, seg_edges: int = _lib.SEG_EDGES,
                   interleave: int = 0):
    """Light rows (degree-descending) and CTA segments of the hub rows.

    `interleave` > 0 (row-partitioned graphs whose SpMM epilogue pushes every output row to the peers):
    the degree-sorted light rows are cut into chunks of `interleave` rows and the chunks are dealt
    heaviest, lightest, 2nd heaviest, 2nd lightest, ...  A purely degree-descending order finishes
    few rows per microsecond at the start of the launch and very many at the end, so the NVLink
    stores of the fused all-gather pile up in the tail; with the interleaved order the bytes pushed
    per edge processed are roughly constant over the launch and the exchange hides behind the gather.
    Rows of one chunk (a few CTAs) still have similar degrees, so warps stay balanced."""
    dev = rowptr.device
    deg = rowptr[1:] - rowptr[:-1]
    n = deg.numel()
    order = torch.argsort(deg, descending=True, stable=True)
    sdeg = deg[order]
    n_hub = int((sdeg > hub_deg).sum())
    hub_rows = order[:n_hub]
    light_rows = order[n_hub:]
    if interleave > 0 and light_rows.numel() > 2 * interleave:
        n_l = light_rows.numel()
        n_chunks = (n_l + interleave - 1) // interleave
        c = torch.arange(n_chunks, device=dev)
        half = (n_chunks + 1) // 2
        chunk_order = torch.empty(n_chunks, dtype=torch.int64, device=dev)
        chunk_order[0::2] = c[:half]                                   # heaviest first ...
        chunk_order[1::2] = torch.flip(c[half:], dims=(0,))            # ... alternating with the lightest
        pos = (chunk_order[:, None] * interleave + torch.arange(interleave, device=dev)[None, :]).reshape(-1)
        light_rows = light_rows[pos[pos < n_l]]
    light_rows = light_rows.to(torch.int32)
    hub_deg_t = sdeg[:n_hub]
    hub_nseg = (hub_deg_t + seg_edges - 1) // seg_edges
    hub_seg0 = torch.cumsum(hub_nseg, 0) - hub_nseg
    n_seg = int(hub_nseg.sum()) if n_hub else 0
    if n_seg:
        seg_hub = torch.repeat_interleave(torch.arange(n_hub, device=dev), hub_nseg)
        within = torch.arange(n_seg, device=dev) - hub_seg0[seg_hub]
        seg_row = hub_rows[seg_hub]
        seg_begin = rowptr[seg_row] + within * seg_edges
        seg_len = torch.minimum(hub_deg_t[seg_hub] - within * seg_edges,
                                torch.full_like(within, seg_edges))
    else:
        seg_hub = torch.zeros(0, dtype=torch.int64, device=dev)
        seg_row = seg_hub.clone()
        seg_begin = seg_hub.clone()
        seg_len = seg_hub.clone()
    lr64 = light_rows.to(torch.int64)
    begin = rowptr[lr64]
    light_desc = torch.stack([lr64, deg[lr64], begin & 0xFFFFFFFF, begin >> 32], dim=1)
    # low word may exceed int31: wrap to the signed representation of the same 32 bits
    light_desc = torch.where(light_desc >= 2 ** 31, light_desc - 2 ** 32, light_desc).to(torch.int32)
    return dict(
        light_rows=light_rows.contiguous(),
        light_desc=light_desc.contiguous(),
        seg_row=seg_row.to(torch.int32).contiguous(),
        seg_begin=seg_begin.to(torch.int64).contiguous(),
        seg_len=seg_len.to(torch.int32).contiguous(),
        seg_hub=seg_hub.to(torch.int32).contiguous(),
        hub_seg0=hub_seg0.to(torch.int32).contiguous(),
        hub_nseg=hub_nseg.to(torch.int32).contiguous(),
        hub_counter=torch.zeros(n_hub, dtype=torch.int32, device=dev),
    )


def build_csr_graph(n_users: int, m_items: int, train_user: torch.Tensor, train_item: torch.Tensor,
                    hub_deg: int = _lib.HUB_DEG, seg_edges: int = _lib.SEG_EDGES) -> CsrGraph:
    """A = [[0,R],[R^T,0]] as CSR; R[u,i] = multiplicity (duplicates stay as
    repeated columns, matching csr_matrix's summing at dataloader.py:164);
    dinv = deg^-1/2 in fp32 with 0 for isolated nodes (dataloader.py:236-238)."""
    N = n_users + m_items
    if N >= 2 ** 31:
        raise ValueError("node ids must fit int32")
    tu = train_user.to(torch.int64)
    ti = train_item.to(torch.int64) + n_users
    if tu.numel():
        if int(tu.min()) < 0 or int(tu.max()) >= n_users or int(ti.min()) < n_users or int(ti.max()) >= N:
            raise ValueError("train ids out of range")
    rows = torch.cat([tu, ti])
    cols = torch.cat([ti, tu])
    key, _ = torch.sort(rows * N + cols)
    srow = torch.div(key, N, rounding_mode="floor")
    col = (key - srow * N).to(torch.int32)
    deg = torch.bincount(srow, minlength=N)
    rowptr = torch.zeros(N + 1, dtype=torch.int64, device=key.device)
    rowptr[1:] = torch.cumsum(deg, 0)
    # dinv = np.power(rowsum_fp32, -0.5), inf -> 0 (dataloader.py:236-238).  Degrees are small
    # integers, so the reference's own host powf is evaluated once per distinct degree
    # value and gathered: dinv is bit-identical to the reference on any device.
    max_deg = int(deg.max()) if N else 0
    with np.errstate(divide="ignore"):
        table = np.power(np.arange(max_deg + 1, dtype=np.float32), np.float32(-0.5)).astype(np.float32)
    table[0] = 0.0
    dinv = torch.from_numpy(table).to(key.device)[deg]
    parts = decompose_rows(rowptr, hub_deg, seg_edges)
    return CsrGraph(n_users, m_items, rowptr.contiguous(), col.contiguous(), dinv.contiguous(), **parts)


def graph_to_sparse_coo(g: CsrGraph) -> torch.Tensor:
    """The reference-format graph: coalesced COO FloatTensor with values
    fl32(fl32(dinv_i * mult) * dinv_j) (dataloader.py:207-213,242-243)."""
    N = g.n_nodes
    deg = g.rowptr[1:] - g.rowptr[:-1]
    row = torch.repeat_interleave(torch.arange(N, device=g.device), deg)
    col = g.col.to(torch.int64)
    key = row * N + col
    uniq, mult = torch.unique_consecutive(key, return_counts=True)
    r = torch.div(uniq, N, rounding_mode="floor")
    c = uniq - r * N
    val = (g.dinv[r] * mult.float()) * g.dinv[c]
    return torch.sparse_coo_tensor(torch.stack([r, c]), val, (N, N)).coalesce()


def build_pos_csr(n_users: int, user: torch.Tensor, item: torch.Tensor):
    """(rowptr int64[n+1], items in FILE order int32, items sorted per user int32).

    `user` must be non-decreasing-grouped as in the train file (one line per uid);
    we only rely on a stable sort by user so the file order inside a user survives
    (reference dataloader.py:118: allPos[u] = np.array(items) of the train line)."""
    u = user.to(torch.int64)
    order = torch.argsort(u, stable=True)
    us = u[order]
    file_items = item[order].to(torch.int32).contiguous()
    cnt = torch.bincount(us, minlength=n_users)
    rowptr = torch.zeros(n_users + 1, dtype=torch.int64, device=u.device)
    rowptr[1:] = torch.cumsum(cnt, 0)
    m = int(item.max()) + 1 if item.numel() else 1
    skey, _ = torch.sort(us * m + item[order].to(torch.int64))
    sorted_items = (skey - torch.div(skey, m, rounding_mode="floor") * m).to(torch.int32).contiguous()
    return rowptr.contiguous(), file_items, sorted_items
