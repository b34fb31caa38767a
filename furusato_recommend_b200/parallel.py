"""Row-partitioned LightGCN over the GPUs of one box.

The reference's own "multi-GPU" mode (ddp_lgcn.py:625-746) is replicas that never
synchronise gradients (it trains through `ddp_model.module.OneEpoch`, so DDP's
reducer never runs).  The partition used here is the reference's *single-device*
memory trick promoted to ranks: `_split_A_hat` (dataloader.py:195-205) cuts A_hat
into contiguous row folds; rank r owns fold r (balanced by nnz, not rows), the
matching rows of E, Adam state and every X_k.

One exchange per layer: all-gather of the pre-scaled activations dinv (.) X_k
(forward) or dinv (.) H_j (backward; A_hat is symmetric so no transpose / reduce-
scatter is needed), then the local SpMM with the usual fused epilogue.  Blocks are
padded to a common row count R so the collective is a plain all-gather: padded id
= rank * R + local row, CSR columns are remapped once at partition time.

The BPR step needs <= 3B rows from arbitrary owners: every rank sees the same B
triples, owners contribute their rows to a [3B, d] buffer (one small all-reduce),
every rank runs the fused BPR kernel on that compact table and adds the gradient
rows it owns into its G shard.  The loss is identical on all ranks by construction.

`RowPartition` and `DistPropagator` are plain index / orchestration logic: they are
unit-tested on CPU with the gloo backend (tests/test_distributed_cpu.py) with the
local SpMM injected; on GPUs the local op is lgcn_propagate_layer and the
collectives run over NCCL / NVLink.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


class RowPartition:
    """nnz-balanced row blocks padded to a common size R.

    With `n_users` given, every rank owns one contiguous range of USER rows and one contiguous
    range of ITEM rows, each cut so that it carries 1/W of that side's edges: both the edges and
    the row count (epilogue + exchange volume) are balanced.  A single contiguous cut over
    [users | items] balances edges only — user blocks then hold 5x more rows than item blocks at
    cfg-3 and their exchange traffic dominates (measured: 134 ms vs 463 ms on 8 GPUs).

    `side_split=True` (even world, n_users given) uses the bipartite structure instead: the first
    W/2 ranks own ONLY user rows, the other W/2 ONLY item rows, each side cut into W/2 ranges of
    equal edge count (both sides carry nnz/2 edges, so the SpMM work is still 1/W per rank).  A
    user row's neighbours are all items and vice versa, so a rank never reads a row of its own
    side: it needs just the OTHER side's table, and the per-layer exchange sends every row to W/2
    ranks instead of W-1.  Ranks simply own an empty range on the other side (repeated cut
    points), so owner / padded-id / shard logic is unchanged."""

    def __init__(self, rowptr: torch.Tensor, world: int, n_users: Optional[int] = None, side_split: bool = False):
        N = rowptr.numel() - 1
        dev = rowptr.device
        self.world, self.n_nodes = world, N
        sides = [(0, N)] if n_users is None or n_users <= 0 or n_users >= N else [(0, n_users), (n_users, N)]
        self.side_split = bool(side_split) and len(sides) == 2 and world % 2 == 0
        cuts = []
        for si, (lo, hi) in enumerate(sides):
            pieces = world // 2 if self.side_split else world
            e0, e1 = int(rowptr[lo]), int(rowptr[hi])
            targets = (e0 + torch.arange(1, pieces, device=dev, dtype=torch.float64) * ((e1 - e0) / pieces)).to(rowptr.dtype)
            c = torch.searchsorted(rowptr[lo:hi + 1].contiguous(), targets, right=False).clamp_(0, hi - lo) + lo
            c = torch.cat([torch.full((1,), lo, dtype=c.dtype, device=dev), c, torch.full((1,), hi, dtype=c.dtype, device=dev)])
            if self.side_split:   # empty ranges for the ranks of the other side
                pad = torch.full((world // 2,), hi if si == 0 else lo, dtype=c.dtype, device=dev)
                c = torch.cat([c, pad]) if si == 0 else torch.cat([pad, c])
            cuts.append(torch.cummax(c, 0)[0])
        self.cuts = cuts                                   # per side: [world + 1] global row boundaries
        self.side_lo = [lo for lo, _ in sides]
        sizes = sum((c[1:] - c[:-1]) for c in cuts)         # rows per rank
        self.rows = [int(x) for x in sizes.tolist()]
        self.R = max(max(self.rows) if N else 0, 1)
        # local offset of each side's range inside a rank's block
        self.side_off = [torch.zeros(world, dtype=torch.int64, device=dev)]
        for c in cuts[:-1]:
            self.side_off.append(self.side_off[-1] + (c[1:] - c[:-1]))
        self.starts = cuts[0]                               # kept for the single-range callers / tests

    def readers_of(self, rank: int):
        """Ranks whose local SpMM reads rows owned by `rank` (the targets of its per-layer push)."""
        if not self.side_split:
            return list(range(self.world))
        half = self.world // 2
        return list(range(half, self.world)) if rank < half else list(range(half))

    def _side(self, ids: torch.Tensor):
        if len(self.cuts) == 1:
            return [torch.ones_like(ids, dtype=torch.bool)]
        first = ids < self.side_lo[1]
        return [first, ~first]

    def owner(self, ids: torch.Tensor) -> torch.Tensor:
        out = torch.zeros_like(ids)
        for msk, c in zip(self._side(ids), self.cuts):
            r = torch.searchsorted(c[1:].contiguous(), ids, right=True).clamp_(max=self.world - 1)
            out = torch.where(msk, r, out)
        return out

    def to_padded(self, ids: torch.Tensor) -> torch.Tensor:
        out = torch.zeros_like(ids)
        for msk, c, off in zip(self._side(ids), self.cuts, self.side_off):
            r = torch.searchsorted(c[1:].contiguous(), ids, right=True).clamp_(max=self.world - 1)
            out = torch.where(msk, r * self.R + off[r] + (ids - c[r]), out)
        return out

    def ranges(self, rank: int):
        """[(lo, hi), ...] global row ranges owned by `rank`, in local order."""
        return [(int(c[rank]), int(c[rank + 1])) for c in self.cuts]

    def block(self, rank: int):
        """Single-range partitions only: the (lo, hi) of the rank's block."""
        (lo, hi), = self.ranges(rank)
        return lo, hi

    def local_csr(self, rank: int, rowptr: torch.Tensor, col: torch.Tensor, dinv: torch.Tensor):
        """(rowptr_local int64[rows+1], col_padded int32[nnz_local], dinv_local fp32[R]); only the
        rank's real rows get CSR entries (the padding up to R exists in the buffers only)."""
        rps, cols, dls = [], [], []
        base = 0
        for lo, hi in self.ranges(rank):
            e0, e1 = int(rowptr[lo]), int(rowptr[hi])
            rps.append(rowptr[lo:hi] - e0 + base)
            cols.append(col[e0:e1])
            dls.append(dinv[lo:hi])
            base += e1 - e0
        rp = torch.cat(rps + [torch.full((1,), base, dtype=torch.int64, device=rowptr.device)])
        colp = self.to_padded(torch.cat(cols).to(torch.int64)).to(torch.int32)
        dl = torch.zeros(self.R, dtype=dinv.dtype, device=dinv.device)
        dcat = torch.cat(dls)
        dl[: dcat.numel()] = dcat
        return rp.contiguous(), colp.contiguous(), dl.contiguous()

    def shard(self, rank: int, full: torch.Tensor) -> torch.Tensor:
        """Rows of a global [N, ...] tensor owned by `rank`, zero-padded to R rows."""
        out = torch.zeros((self.R,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
        o = 0
        for lo, hi in self.ranges(rank):
            out[o:o + hi - lo] = full[lo:hi]
            o += hi - lo
        return out

    def unshard(self, gathered: torch.Tensor) -> torch.Tensor:
        """[world*R, ...] padded layout -> global [N, ...]."""
        parts = []
        for si in range(len(self.cuts)):
            for r in range(self.world):
                lo, hi = self.ranges(r)[si]
                o = r * self.R + int(self.side_off[si][r])
                parts.append(gathered[o:o + (hi - lo)])
        return torch.cat(parts)


def build_g0(full: torch.Tensor, dinv_pad: torch.Tensor, ids: torch.Tensor, G_c: torch.Tensor) -> None:
    """full[W*R, d] = dinv (.) G for a gradient seed G that is non-zero only in the rows `ids` (padded
    ids, duplicates add) — the pre-scaled layer-0 source of the backward pass, built LOCALLY from the
    compact BPR table every rank already holds, instead of exchanging the (almost empty) G shards."""
    full.zero_()
    full.index_add_(0, ids, (dinv_pad[ids][:, None] * G_c).to(full.dtype))


# local_spmm(src_full [W*R, d], *, dst, base, acc_in, acc_out, acc_scale, last_kwargs) -> None
LocalSpmm = Callable[..., None]


class DistPropagator:
    """K-layer propagation / Horner backward over a RowPartition.

    `local_spmm(src_full, dst=..., base=..., acc_in=..., acc_out=..., acc_scale=..., **kw)` must
    implement lgcn_propagate_layer's contract with scale_src=0 on the rank's local CSR."""

    def __init__(self, part: RowPartition, rank: int, dinv_local: torch.Tensor, n_layers: int,
                 local_spmm: LocalSpmm, group=None, storage_dtype: torch.dtype = torch.float32):
        self.part, self.rank, self.K = part, rank, n_layers
        self.dinv = dinv_local
        self.spmm = local_spmm
        self.group = group
        self.storage_dtype = storage_dtype
        self._full: Optional[torch.Tensor] = None
        self._z = [None, None]

    def _buffers(self, d: int, device):
        R, W = self.part.R, self.part.world
        if self._full is None or self._full.shape != (W * R, d):
            self._full = torch.empty((W * R, d), dtype=self.storage_dtype, device=device)
            self._z = [torch.empty((R, d), dtype=self.storage_dtype, device=device) for _ in range(2)]
        return self._full, self._z

    def _gather(self, z_local: torch.Tensor) -> torch.Tensor:
        full, _ = self._buffers(z_local.shape[1], z_local.device)
        if self.part.world == 1:
            full.copy_(z_local)
        else:
            dist.all_gather_into_tensor(full, z_local.contiguous(), group=self.group)
        return full

    def forward(self, emb_local: torch.Tensor, acc: torch.Tensor, out: torch.Tensor) -> None:
        """out = (X0 + ... + XK)/(K+1) for the local rows."""
        K = self.K
        _, z = self._buffers(emb_local.shape[1], emb_local.device)
        z0 = (self.dinv[:, None] * emb_local).to(self.storage_dtype)
        src = z0
        for k in range(K):
            last = k == K - 1
            full = self._gather(src)
            self.spmm(full, dst=None if last else z[k & 1], base=None,
                      acc_in=emb_local if k == 0 else acc, acc_out=out if last else acc,
                      acc_scale=1.0 / (K + 1) if last else 1.0)
            src = z[k & 1]

    def backward(self, G_local: torch.Tensor, g0=None, **last_kwargs) -> None:
        """H0 = G, H_{j+1} = G + A_hat H_j; the last layer's epilogue consumes H_K
        (grad_mode 1 or 2 keyword arguments are passed through to the local op).
        g0 = (dinv_pad [W*R], ids [n], G_c [n, d]): build the first layer's gathered source locally
        (`build_g0`) instead of all-gathering dinv (.) G."""
        K = self.K
        full0, z = self._buffers(G_local.shape[1], G_local.device)
        src = None if g0 is not None else (self.dinv[:, None] * G_local).to(self.storage_dtype)
        for j in range(K):
            last = j == K - 1
            if j == 0 and g0 is not None:
                build_g0(full0, *g0)
                full = full0
            else:
                full = self._gather(src)
            kw = last_kwargs if last else {}
            self.spmm(full, dst=None if last else z[j & 1], base=G_local, acc_in=None, acc_out=None,
                      acc_scale=1.0, **kw)
            src = z[j & 1]


class PushPropagator:
    """Same maths as DistPropagator, but every per-layer all-gather is FUSED into the producing
    kernel: the SpMM epilogue stores each pre-scaled output row straight into the gathered buffers
    of the ranks that read it, through NVLink peer mappings (torch symmetric memory supplies the
    mapped pointers and the cross-GPU barrier).  No NCCL call and no staging copy on the layer
    path; the transfer overlaps the gather/FMA work of the same kernel row by row.

    Exchanges per train step: 2K - 1, all fused.
      * forward layer k reads buf[k & 1] and pushes X_{k+1} into buf[(k + 1) & 1] of its readers;
      * the backward pass needs no exchange for its first layer: the gradient seed has <= 3B
        non-zero rows that every rank already holds (the compact BPR table), so its pre-scaled
        gathered copy `g0` is built locally by lgcn_bpr_fwd_bwd_rows;
      * the LAST backward layer's Adam epilogue pushes dinv (.) E_new — the next step's layer-0
        source — into buf[0] (`push_emb`), so the updated table is never exchanged separately.
    One barrier separates the writers of a buffer from its readers: K + 1 per pass pair
    (start of forward, after every non-final layer)."""

    def __init__(self, part: RowPartition, rank: int, graph, dinv_local: torch.Tensor, n_layers: int,
                 ops, group, d: int, storage_dtype: torch.dtype, device, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm
        self.part, self.rank, self.K, self.ops, self.graph = part, rank, n_layers, ops, graph
        self.dinv, self.storage_dtype = dinv_local, storage_dtype
        R, W = part.R, part.world
        # zero-filled: rows nobody pushes (padding up to R; with the side-split partition also the
        # whole own side) stay finite — the SpMM's padding slots re-read row 0 with weight 0
        self.bufs = [symm.empty((W * R, d), dtype=storage_dtype, device=device).zero_() for _ in range(2)]
        self.hdl = [symm.rendezvous(t, group) for t in self.bufs]
        readers = part.readers_of(rank)   # side-split partition: only the other side's ranks
        self.peers = [[int(h.buffer_ptrs[r]) for r in readers] for h in self.hdl]
        self.row0 = rank * R
        # NVSwitch multicast (NVLS): when every rank reads every row (two_sided partition) and the
        # symmetric allocation has a multicast mapping, a row is stored ONCE and the switch replicates
        # it into all W gathered buffers — per-rank NVLink egress per layer drops from W shards to one.
        self.mcast = [0, 0]
        if use_multicast and not part.side_split:
            mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.hdl]
            if all(mc):
                self.mcast = mc
        self.x0_valid = False   # buf[0] holds dinv (.) E of the CURRENT table on every reader

    def _barrier(self):
        self.hdl[0].barrier(channel=0)

    def _push_target(self, b: int) -> dict:
        """Where the epilogue of a layer stores its rows for buffer b: the multicast mapping, or the readers' peer pointers."""
        return dict(dst_multicast=self.mcast[b]) if self.mcast[b] else dict(dst_peers=self.peers[b])

    def push_x0(self, emb_local: torch.Tensor) -> None:
        """Explicit exchange of the pre-scaled table (first step, after load_global_embedding, after
        an eval-only propagation): every later train step gets it from the Adam epilogue."""
        self._barrier()  # every peer is done reading buf[0]
        self.ops.scale_rows_push(emb_local, self.dinv, self.storage_dtype, self.peers[0], self.row0)
        self.x0_valid = True

    def forward(self, emb_local: torch.Tensor, acc: torch.Tensor, out: torch.Tensor) -> None:
        ops, K = self.ops, self.K
        if not self.x0_valid:
            self.push_x0(emb_local)
        self._barrier()  # the X0 rows (previous step's Adam epilogue, or push_x0) are visible everywhere
        for k in range(K):
            last = k == K - 1
            nxt = (k + 1) & 1
            ops.propagate_layer(self.graph, self.bufs[k & 1], scale_src=False,
                                dst=None if last else self.bufs[nxt], dst_row_offset=self.row0,
                                **({} if last else self._push_target(nxt)),
                                acc_in=emb_local if k == 0 else acc, acc_out=out if last else acc,
                                acc_scale=1.0 / (K + 1) if last else 1.0)
            if not last:
                self._barrier()
        if K > 1:
            self.x0_valid = False   # layer 1 pushed X2 over X0

    def backward(self, G_local: torch.Tensor, g0_src: torch.Tensor, **last_kwargs) -> None:
        """H0 = G, H_{j+1} = G + A_hat H_j.  `g0_src` fp32 [W*R, d] = dinv (.) G in the padded layout
        (local).  Layer j reads buf[(a + j) & 1] (j > 0) and pushes into buf[(a + j + 1) & 1] with
        a = K & 1, so the last layer's push — the updated, pre-scaled table — lands in buf[0].  The
        caller guarantees a cross-GPU barrier between the forward pass and this call (the BPR row
        exchange has one), so no peer still reads the buffer layer 0 pushes into."""
        ops, K = self.ops, self.K
        a = K & 1
        for j in range(K):
            last = j == K - 1
            nxt = (a + j + 1) & 1
            kw = dict(last_kwargs, push_emb=True) if last else {}
            ops.propagate_layer(self.graph, g0_src if j == 0 else self.bufs[(a + j) & 1], scale_src=False,
                                dst=self.bufs[nxt], dst_row_offset=self.row0, **self._push_target(nxt),
                                base=G_local, **kw)
            if not last:
                self._barrier()
        self.x0_valid = True


def reduce_local_graphs(part: RowPartition, rank: int, rowptr: torch.Tensor, col: torch.Tensor):
    """The two local CSRs of the "reduce" partition (two-sided RowPartition of a bipartite graph).

    A  rows = this rank's users (local order), columns = PADDED ids of their items: the user rows gather
       from the gathered item table exactly like the all-gather partition does.
    B  rows = ALL items in owner-block numbering b * R_i + j (R_i = largest item block; padding rows are
       empty), columns = LOCAL index of this rank's users: the partial sum of every item over the users
       that live here; block b of the rows belongs to rank b.
    Returns (rowptr_A, col_A, rowptr_B, col_B, R_i, item_rows_per_rank)."""
    dev = rowptr.device
    W = part.world
    (u_lo, u_hi), (i_lo, i_hi) = part.ranges(rank)
    # ---- A
    e0, e1 = int(rowptr[u_lo]), int(rowptr[u_hi])
    rp_a = (rowptr[u_lo:u_hi + 1] - e0).contiguous()
    col_a = part.to_padded(col[e0:e1].to(torch.int64)).to(torch.int32).contiguous()
    # ---- B: every item row, neighbours restricted to [u_lo, u_hi)
    item_cuts = part.cuts[1]                                  # [W + 1] global item-row boundaries
    n_first, n_last = int(item_cuts[0]), int(item_cuts[-1])
    sizes = (item_cuts[1:] - item_cuts[:-1])
    R_i = max(int(sizes.max()), 1)
    f0, f1 = int(rowptr[n_first]), int(rowptr[n_last])
    deg = rowptr[n_first + 1:n_last + 1] - rowptr[n_first:n_last]
    row_of = torch.repeat_interleave(torch.arange(n_last - n_first, device=dev), deg)      # item index (0-based) per entry
    c = col[f0:f1].to(torch.int64)
    keep = (c >= u_lo) & (c < u_hi)
    row_of, c = row_of[keep], (c[keep] - u_lo).to(torch.int32)
    del keep
    blk = torch.searchsorted(item_cuts[1:].contiguous(), row_of + n_first, right=True).clamp_(max=W - 1)
    prow = blk * R_i + (row_of + n_first - item_cuts[blk])                                  # owner-block numbering
    cnt = torch.bincount(prow, minlength=W * R_i)
    rp_b = torch.zeros(W * R_i + 1, dtype=torch.int64, device=dev)
    rp_b[1:] = torch.cumsum(cnt, 0)
    # entries are already grouped by item in ascending item order and blocks ascend with the item id, so prow is sorted
    return rp_a, col_a, rp_b.contiguous(), c.contiguous(), R_i, [int(x) for x in sizes.tolist()]


class ReducePropagator:
    """"reduce" partition for bipartite graphs with many more users than items (cfg-3: 10 M x 2 M).

    The all-gather partition makes every rank ingest the WHOLE table per layer, and 5/6 of that is user rows
    that only the item owners need.  Here user rows never travel: every rank sums its own users' pre-scaled
    rows per item (graph B, raw fp32 partial sums pushed to the item's owner, `dst_route_rows`), the owner
    adds the W partials and runs the usual row epilogue (`lgcn_reduce_rows`), and only the ITEM rows are
    gathered (pushed from that epilogue, NVSwitch multicast when available).  Per layer and rank: item table in
    + one item-table-sized partial out, instead of the whole table in.  Same buffers and padded ids as the push
    partition: buf[b] holds the gathered items and, in its own block, this rank's users."""

    def __init__(self, part: RowPartition, rank: int, rowptr, col, dinv_local: torch.Tensor, n_layers: int,
                 ops, group, d: int, storage_dtype: torch.dtype, device, interleave: int = 32, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm
        from .graph import CsrGraph, decompose_rows
        if part.side_split or len(part.cuts) != 2:
            raise ValueError("the reduce partition needs the two-sided RowPartition")
        self.part, self.rank, self.K, self.ops = part, rank, n_layers, ops
        self.storage_dtype = storage_dtype
        R, W = part.R, part.world
        rp_a, col_a, rp_b, col_b, R_i, _ = reduce_local_graphs(part, rank, rowptr, col)
        (u_lo, u_hi), (i_lo, i_hi) = part.ranges(rank)
        self.nu, self.ni, self.R_i = u_hi - u_lo, i_hi - i_lo, R_i
        self.dinv_u = dinv_local[:self.nu].contiguous()
        self.dinv_i = dinv_local[self.nu:self.nu + self.ni].contiguous()
        self.graph_a = CsrGraph(self.nu, 0, rp_a, col_a, self.dinv_u, **decompose_rows(rp_a))
        self.ones = torch.ones(W * R_i, dtype=torch.float32, device=device)
        self.graph_b = CsrGraph(W * R_i, 0, rp_b, col_b, self.ones, **decompose_rows(rp_b, interleave=interleave))
        self.local_nnz = int(col_a.numel() + col_b.numel())
        self.bufs = [symm.empty((W * R, d), dtype=storage_dtype, device=device).zero_() for _ in range(2)]
        self.hdl = [symm.rendezvous(t, group) for t in self.bufs]
        self.peers = [[int(h.buffer_ptrs[r]) for r in range(W)] for h in self.hdl]
        self.mcast = [0, 0]
        if use_multicast:
            mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.hdl]
            if all(mc):
                self.mcast = mc
        # partial sums at the owner: slot q = the partial of rank q for my item block
        self.partials = symm.empty((W, R_i, d), dtype=torch.float32, device=device).zero_()
        self.phdl = symm.rendezvous(self.partials, group)
        self.ppeers = [int(self.phdl.buffer_ptrs[r]) for r in range(W)]
        self.row0 = rank * R
        self.x0_valid = False

    def _barrier(self):
        self.hdl[0].barrier(channel=0)

    def _own(self, b: int) -> torch.Tensor:
        """This rank's user rows inside gathered buffer b (they are read locally by graph B and never pushed)."""
        return self.bufs[b][self.row0:self.row0 + self.nu]

    def _item_push(self, b: int) -> dict:
        t = dict(dst_multicast=self.mcast[b]) if self.mcast[b] else dict(dst_peers=self.peers[b])
        return dict(dst=self.bufs[b], dst_row_offset=self.row0 + self.nu, **t)

    def push_x0(self, emb_local: torch.Tensor) -> None:
        ops = self.ops
        self._barrier()  # every peer is done reading buf[0]
        ops.scale_rows_push(emb_local[:self.nu], self.dinv_u, self.storage_dtype, [self.peers[0][self.rank]], self.row0)
        ops.scale_rows_push(emb_local[self.nu:self.nu + self.ni], self.dinv_i, self.storage_dtype, self.peers[0],
                            self.row0 + self.nu)
        self.x0_valid = True

    def _layer(self, src_buf, src_own, nxt, users_kw, items_kw, last_push: bool):
        """One layer: user rows from the gathered items (A), item partials from the local users (B) pushed to
        their owners, barrier, owner-side reduce + epilogue (item rows pushed to everyone when `nxt` is set)."""
        ops = self.ops
        ops.propagate_layer(self.graph_a, src_buf, scale_src=False,
                            dst=None if nxt is None else self._own(nxt), **users_kw)
        ops.propagate_layer(self.graph_b, src_own, scale_src=False, n_src_rows=self.nu, dst=self.partials[0],
                            dst_peers=self.ppeers, dst_row_offset=self.rank * self.R_i, dst_route_rows=self.R_i,
                            src_scale=self.ones, dst_scale=self.ones)
        self.phdl.barrier(channel=2)          # every rank's partials for my items have landed
        ops.reduce_rows(self.partials, self.ni, self.dinv_i, **({} if nxt is None else self._item_push(nxt)), **items_kw)

    def forward(self, emb_local: torch.Tensor, acc: torch.Tensor, out: torch.Tensor) -> None:
        K, nu, ni = self.K, self.nu, self.ni
        if not self.x0_valid:
            self.push_x0(emb_local)
        for k in range(K):
            last = k == K - 1
            self._barrier()  # the item rows of layer k (or X0) are visible; everyone is done with the partials of layer k-1
            b = k & 1
            sc = 1.0 / (K + 1) if last else 1.0
            ukw = dict(acc_in=emb_local[:nu] if k == 0 else acc[:nu], acc_out=(out if last else acc)[:nu], acc_scale=sc)
            ikw = dict(acc_in=emb_local[nu:nu + ni] if k == 0 else acc[nu:nu + ni], acc_out=(out if last else acc)[nu:nu + ni],
                       acc_scale=sc)
            self._layer(self.bufs[b], self._own(b), None if last else (k + 1) & 1, ukw, ikw, not last)
        if K > 1:
            self.x0_valid = False

    def backward(self, G_local: torch.Tensor, g0_src: torch.Tensor, **last_kwargs) -> None:
        """Horner backward; layer j reads buf[(a + j) & 1] (j > 0) or the local seed g0, writes buf[(a + j + 1) & 1],
        a = K & 1 so that the last layer's Adam epilogue leaves dinv (.) E_new in buf[0] (users locally, items pushed)."""
        K, nu, ni = self.K, self.nu, self.ni
        a = K & 1

        def rows(kw, lo, hi):   # slice the per-row tensors of the Adam / gradient epilogue
            out = dict(kw)
            for key in ("cnt", "emb", "grad", "adam_m", "adam_v"):
                if out.get(key) is not None:
                    out[key] = out[key][lo:hi]
            return out

        for j in range(K):
            last = j == K - 1
            self._barrier()
            nxt = (a + j + 1) & 1
            src = g0_src if j == 0 else self.bufs[(a + j) & 1]
            src_own = g0_src[self.row0:self.row0 + nu] if j == 0 else self._own((a + j) & 1)
            kw = dict(last_kwargs, push_emb=True) if last else {}
            self._layer(src, src_own, nxt, dict(base=G_local[:nu], **rows(kw, 0, nu)),
                        dict(base=G_local[nu:nu + ni], **rows(kw, nu, nu + ni)), True)
        self.x0_valid = True


def exchange_rows(part: RowPartition, rank: int, local, padded_ids: torch.Tensor, group=None):
    """rows[i] = TABLE[padded_ids[i]] where TABLE is row-partitioned: owners fill, one all-reduce.
    `local` may be a list of tables (their rows are concatenated along dim 1 AFTER the gather, so
    only 3B rows are ever copied).  Sync-free (no boolean indexing); where(), not a 0/1 multiply:
    a non-owner's dummy row may be uninitialised padding (NaN * 0 = NaN)."""
    tables = list(local) if isinstance(local, (list, tuple)) else [local]
    mine = (padded_ids // part.R) == rank
    loc = padded_ids % part.R
    rows = torch.cat([t[loc] for t in tables], dim=1) if len(tables) > 1 else tables[0][loc]
    buf = torch.where(mine[:, None], rows, torch.zeros((), dtype=rows.dtype, device=rows.device))
    if part.world > 1:
        dist.all_reduce(buf, group=group)
    return buf, mine


class DistLightGCN:
    """Row-partitioned LightGCN training step + user-sharded evaluation on CUDA ranks."""

    def __init__(self, config: dict, dataset, rank: int, world: int, group=None, seed: int = 2020):
        from . import ops
        from .graph import CsrGraph, decompose_rows
        self.ops = ops
        self.config, self.dataset, self.rank, self.world, self.group = config, dataset, rank, world, group
        self.n, self.m = dataset.n_users, dataset.m_items
        self.d = int(config["recdim"])
        self.K = int(config["layer"])
        self.device = torch.device(config["device"])
        g = dataset.csr_graph()
        # bipartite side split (users on the first W/2 ranks, items on the rest) halves the exchange
        # when the two sides have comparable row counts; cfg "dist_partition": "two_sided" keeps
        # every rank on both sides
        mode_p = config.get("dist_partition", "auto")   # auto | side_split | two_sided | reduce
        mode = config.get("dist_exchange", "auto")
        reduce_optional = False
        if mode_p == "auto":
            if max(self.n, self.m) <= 2 * min(self.n, self.m):
                mode_p = "side_split"   # comparable sides (cfg-2 x N): a rank reads only the OTHER side's table
            elif self.n > 2 * self.m and world > 1 and mode in ("auto", "push") and self.device.type == "cuda":
                # user-heavy graph (cfg-3: 10 M users x 2 M items): with a two-sided cut every rank ingests the
                # whole table per layer and 5/6 of it is user rows; the reduce partition keeps them home.
                # Measured on 8 B200s: 54.7 vs 66.5 ms/step (7.5x vs 6.2x the single-GPU step).  Falls back to the
                # two-sided cut when the peers cannot map each other's memory.
                mode_p, reduce_optional = "reduce", True
            else:
                mode_p = "two_sided"
        self.part = RowPartition(g.rowptr, world, n_users=self.n, side_split=(mode_p == "side_split"))
        self.reduce_mode = mode_p == "reduce"
        self.prop = None
        storage = torch.bfloat16 if config.get("storage_dtype") == "bf16" else torch.float32
        if self.reduce_mode:
            if world < 2 or mode == "nccl":
                raise ValueError("dist_partition='reduce' needs the push exchange on >= 2 ranks")
            try:
                self.prop = ReducePropagator(self.part, rank, g.rowptr, g.col, self.part.shard(rank, g.dinv), self.K, ops,
                                             group if group is not None else dist.group.WORLD, self.d, storage,
                                             self.device, interleave=int(config.get("dist_row_interleave", 32)),
                                             use_multicast=bool(config.get("dist_multicast", True)))
            except Exception as e:
                if not reduce_optional:
                    raise
                import warnings
                warnings.warn(f"reduce partition unavailable ({e!r}); using the two-sided partition")
                self.reduce_mode = False
        if self.reduce_mode:
            # "reduce": user rows never travel (ReducePropagator); only the local deg^-1/2 shard is needed here
            dl = self.part.shard(rank, g.dinv)
            self.local_graph, self.local_nnz = None, 0
        else:
            rp, colp, dl = self.part.local_csr(rank, g.rowptr, g.col, g.dinv)
            # the push exchange stores every output row to the peers from the SpMM epilogue: interleave heavy and
            # light row chunks so those stores are spread over the launch (graph.decompose_rows)
            self.local_graph = CsrGraph(self.part.rows[rank], 0, rp, colp, dl,
                                        **decompose_rows(rp, interleave=int(config.get("dist_row_interleave", 32))))
            self.local_nnz = int(colp.numel())
        R, d, dev = self.part.R, self.d, self.device
        gen = torch.Generator(device=dev).manual_seed(seed + rank)
        self.emb = torch.randn((R, d), generator=gen, device=dev) * 0.1      # model/lgcn.py:75, local rows
        self.emb[self.part.rows[rank]:] = 0
        self.m1, self.v1 = torch.zeros_like(self.emb), torch.zeros_like(self.emb)
        self.acc, self.out = torch.zeros_like(self.emb), torch.zeros_like(self.emb)
        self.G = torch.zeros_like(self.emb)
        self.cnt = torch.zeros(R, dtype=torch.int32, device=dev)
        self.step_t = torch.zeros(1, dtype=torch.int64, device=dev)
        self.hp = torch.zeros(2, dtype=torch.float32, device=dev)
        self.loss_out = torch.zeros(4, dtype=torch.float32, device=dev)
        self.work_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        # "push": all-gather fused into the SpMM epilogue over NVLink peer memory (default when the
        # ranks can map each other's memory); "nccl": ncclAllGather per layer (the baseline).
        self.exchange = "nccl"
        if self.reduce_mode:
            self.local_nnz = self.prop.local_nnz
            self.exchange = "push"
        elif world > 1 and mode in ("auto", "push"):
            try:
                self.prop = PushPropagator(self.part, rank, self.local_graph, dl, self.K, ops,
                                           group if group is not None else dist.group.WORLD, d, storage, dev,
                                           use_multicast=bool(config.get("dist_multicast", True)))
                self.exchange = "push"
            except Exception as e:  # no P2P / symmetric memory on this box
                if mode == "push":
                    raise
                import warnings
                warnings.warn(f"symmetric-memory push exchange unavailable ({e!r}); using NCCL all-gather")
        if self.prop is None:
            self.prop = DistPropagator(self.part, rank, dl, self.K, self._local_spmm, group, storage)
        self.collectives_per_step = 2 * self.K + 1
        # NCCL baseline only (CPU-tested index logic): build the backward pass's layer-0 source locally
        # from the compact BPR table instead of all-gathering the G shards (fp32 storage only).  The
        # push exchange always does this, in its BPR kernel.
        self.sparse_g0 = (bool(config.get("dist_sparse_g0", False)) and storage == torch.float32 and world > 1
                          and self.exchange == "nccl")
        self._dinv_pad = None
        if self.sparse_g0 or self.exchange == "push":
            self._dinv_pad = torch.zeros(world * R, dtype=torch.float32, device=dev)
            self._dinv_pad[self.part.to_padded(torch.arange(self.n + self.m, device=dev))] = g.dinv
        if self.exchange == "push":
            # partition tables of lgcn_padded_ids and the local, pre-scaled gradient seed dinv (.) G
            self._cuts = torch.stack([c.to(torch.int64) for c in self.part.cuts]).contiguous()
            self._side_off = torch.stack([o.to(torch.int64) for o in self.part.side_off[:2]]).contiguous()
            self._g0 = torch.zeros((world * R, d), dtype=torch.float32, device=dev)
        self._cap = 0           # batch size the scratch below was allocated for (grow-only)
        self._xbuf = None
        # Capturing NCCL collectives into a CUDA graph deadlocked on the 2-GPU box (round 1), so the
        # NCCL exchange stays eager unless asked.  The push exchange has no NCCL call on the step at
        # all (layers AND the 3B-row exchange go through peer stores + symmetric-memory barriers,
        # which are plain kernels), so its step is captured by default.
        self.use_cuda_graph = bool(config.get("dist_cuda_graph", self.exchange == "push"))
        self._graph = None
        self._graph_cap = 0

    def load_global_embedding(self, E: torch.Tensor) -> None:
        """Take this rank's rows of a global [N, d] table (checkpoint key all_embedding.weight)."""
        self.emb.copy_(self.part.shard(self.rank, E.to(self.device)))
        if self.exchange == "push":
            self.prop.x0_valid = False

    def gather_embedding(self) -> torch.Tensor:
        full = torch.empty((self.world * self.part.R, self.d), dtype=torch.float32, device=self.device)
        if self.world == 1:
            full.copy_(self.emb)
        else:
            dist.all_gather_into_tensor(full, self.emb, group=self.group)
        return self.part.unshard(full)

    def _local_spmm(self, src_full, **kw):
        self.ops.propagate_layer(self.local_graph, src_full, scale_src=False, **kw)

    def computer_local(self) -> torch.Tensor:
        self.prop.forward(self.emb, self.acc, self.out)
        return self.out

    def fused_step(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
        """One training step on the same B triples on every rank (global ids).  Full-size batches
        replay one CUDA graph (push exchange: kernels and symmetric-memory barriers only); anything
        else — the last, partial batch of an epoch — runs the same kernels eagerly on the SAME
        grow-only scratch, so the captured graph never sees a dangling pointer."""
        B = users.numel()
        if not self.use_cuda_graph or B != int(self.config["bpr_batch_size"]):
            users, pos, neg = (t.to(self.device, non_blocking=True) for t in (users, pos, neg))
            return self._fused_step_eager(users, pos, neg)
        if self._graph is None or self._graph_cap != self._cap or self._cap < B:
            self._capture(B)
        if self.exchange == "push" and not self.prop.x0_valid:
            self.prop.push_x0(self.emb)   # an eval propagation (or a restore) ran since the last step
        for dst, src in zip(self._gbatch, (users, pos, neg)):
            dst.copy_(src, non_blocking=True)
        self._graph.replay()
        if self.exchange == "push":
            self.prop.x0_valid = True
        return self.loss_out[2]

    def _ensure_scratch(self, B: int) -> None:
        """Per-step scratch, allocated once for the largest batch seen (normally bpr_batch_size) and
        sliced for smaller ones.  Growing re-allocates (and re-rendezvouses the symmetric exchange
        buffer on every rank: all ranks see the same batch sizes) and invalidates the graph."""
        if B <= self._cap:
            return
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("step scratch must exist before graph capture")
        cap = max(B, int(self.config["bpr_batch_size"]))
        dev, d = self.device, self.d
        self._work = torch.empty(2 * cap, dtype=torch.float32, device=dev)
        self._ids = torch.zeros(3 * cap, dtype=torch.int64, device=dev)
        if self.exchange == "push":
            import torch.distributed._symmetric_memory as symm
            self._xbuf = symm.empty((3 * cap, 2 * d), dtype=torch.float32, device=dev)
            self._xhdl = symm.rendezvous(self._xbuf, self.group if self.group is not None else dist.group.WORLD)
            self._xpeers = [int(p) for p in self._xhdl.buffer_ptrs]
        else:
            self._ar = torch.arange(cap, device=dev, dtype=torch.int64)
            self._G_c = torch.empty((3 * cap, d), dtype=torch.float32, device=dev)
            self._cnt_c = torch.empty(3 * cap, dtype=torch.int32, device=dev)
        self._cap = cap
        self._graph = None

    def _capture(self, B: int) -> None:
        dev = self.device
        self._ensure_scratch(B)
        self._gbatch = [torch.zeros(B, dtype=torch.int64, device=dev) for _ in range(3)]
        state = (self.emb, self.m1, self.v1, self.step_t, self.hp, self.loss_out)
        keep = [t.clone() for t in state]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):                      # NCCL channels and every lazy buffer exist before capture
                self._fused_step_eager(*self._gbatch)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self._fused_step_eager(*self._gbatch)
        for t, k in zip(state, keep):
            t.copy_(k)
        if self.exchange == "push":
            self.prop.x0_valid = False   # buf[0] holds the warm-up steps' table, not the restored one
        self._graph = graph
        self._graph_cap = self._cap

    def _fused_step_eager(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
        if self.exchange == "push":
            return self._step_push(users, pos, neg)
        return self._step_nccl(users, pos, neg)

    def _step_push(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
        """2K SpMM launches (2K - 1 of them carry a fused exchange) + id mapping + row exchange + BPR +
        Adam tick + row clear; K + 2 ... barriers.  No torch op touches the data path."""
        ops, part, B = self.ops, self.part, users.numel()
        self._ensure_scratch(B)
        self.prop.forward(self.emb, self.acc, self.out)
        ids = self._ids[:3 * B]
        ops.padded_ids(users, pos, neg, self.n, self.n + self.m, self._cuts, self._side_off, self.world, part.R, ids,
                       self.work_counter[1:])
        # owners store [out | emb] rows straight into every rank's [3B, 2d] buffer over NVLink.  No barrier is
        # needed BEFORE the stores: a peer can only be here after the barriers of this step's forward pass,
        # which every rank enters after it finished reading the previous step's rows.
        ops.exchange_rows_push(self.out, self.emb, ids, part.R, self.rank, self._xpeers)
        self._xhdl.barrier(channel=1)
        decay = float(self.config["decay"])
        ops.bpr_fwd_bwd_rows(self._xbuf, ids, B, part.R, self.rank, decay, self.G, self.cnt, self._g0, self._dinv_pad,
                             self.loss_out, self._work, self.work_counter)
        ops.adam_tick(self.step_t, self.hp, float(self.config["lr"]))
        self.prop.backward(self.G, self._g0, grad_mode=2, inv_layers=1.0 / (self.K + 1), reg_coef=decay / B,
                           cnt=self.cnt, emb=self.emb, adam_m=self.m1, adam_v=self.v1, adam_hp=self.hp, zero_base=True)
        ops.zero_rows(self._g0, ids)
        return self.loss_out[2]

    def _step_nccl(self, users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor) -> torch.Tensor:
        """The baseline: ncclAllGather per layer, one [3B, 2d] all-reduce, torch glue."""
        ops, part, B = self.ops, self.part, users.numel()
        self._ensure_scratch(B)
        self.prop.forward(self.emb, self.acc, self.out)
        ids = part.to_padded(torch.cat([users, pos + self.n, neg + self.n]))
        both, mine = exchange_rows(part, self.rank, [self.out, self.emb], ids, self.group)
        out_c, emb_c = both[:, :self.d].contiguous(), both[:, self.d:].contiguous()
        ar, G_c, cnt_c = self._ar[:B], self._G_c[:3 * B], self._cnt_c[:3 * B]
        G_c.zero_()
        cnt_c.zero_()
        decay = float(self.config["decay"])
        ops.bpr_fwd_bwd(out_c, emb_c, ar, ar, ar + B, B, decay, G_c, cnt_c, self.loss_out, self._work,
                        self.work_counter)
        loc = ids % part.R                      # owners add their rows, everyone else adds zeros to a valid row
        self.G.index_add_(0, loc, torch.where(mine[:, None], G_c, torch.zeros((), dtype=G_c.dtype, device=G_c.device)))
        self.cnt.index_add_(0, loc, cnt_c * mine.to(cnt_c.dtype))
        ops.adam_tick(self.step_t, self.hp, float(self.config["lr"]))
        g0 = (self._dinv_pad, ids, G_c) if self.sparse_g0 else None
        self.prop.backward(self.G, g0=g0, grad_mode=2, inv_layers=1.0 / (self.K + 1), reg_coef=decay / B, cnt=self.cnt,
                           emb=self.emb, adam_m=self.m1, adam_v=self.v1, adam_hp=self.hp, zero_base=False)
        self.G.zero_()
        return self.loss_out[2]

    def check_ids(self) -> None:
        """Raise the reference's IndexError if a step since the last call saw an id outside the table
        (one host sync; the kernels count and skip them)."""
        bad = int(self.work_counter[1].item())
        if bad:
            self.work_counter[1] = 0
            raise IndexError(f"{bad} (user, pos, neg) id(s) outside [0, {self.n}) / [0, {self.m})")

    def gather_out(self) -> torch.Tensor:
        """Global light_out [N, d] on every rank (eval: the item table is replicated)."""
        full = torch.empty((self.world * self.part.R, self.d), dtype=torch.float32, device=self.device)
        if self.world == 1:
            full.copy_(self.out)
        else:
            dist.all_gather_into_tensor(full, self.out, group=self.group)
        return self.part.unshard(full)

    def topk_user_shard(self, users: torch.Tensor, k: int, precision: str = "bf16"):
        """Evaluation sharded by user: this rank scores users[rank::world] against all items."""
        full = self.gather_out()
        mine = users[self.rank::self.world].contiguous()
        rowptr, _, srt = self.dataset.pos_csr()
        idx, val = self.ops.score_topk(full[: self.n].contiguous(), full[self.n:].contiguous(), mine, rowptr, srt, k,
                                       precision=precision)
        return mine, idx, val
