"""CPU oracle for the LightGCN train + full-rank-eval hot path.

TEST INFRASTRUCTURE ONLY.  This module is a CPU restatement (numpy + torch-CPU)
of what `HiromasaYamanishi/furusato_recommend` computes on the path named by
BASELINE.json.  Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline /
`--impl reference` legs of `bench.py` may import it; the product package
`furusato_recommend_b200` never does (it fails loudly without its CUDA library).

Parity status: PINNED.  `oracle/make_golden.py` imports the live reference from
/root/reference (dataloader.Loader -> model.MF.LightGCN -> trainer.Trainer with
the shims of SURVEY.md §8c), checks every function below against it and freezes
the outputs in `tests/golden/*.npz`; `tests/test_oracle_golden.py` replays them.
The only unpinned piece is the Philox counter->draw mapping of the GPU sampler,
which is our own specification (the reference draws from numpy's MT19937); the
*decision procedure* that consumes the draws is pinned by `uniform_sample_mt`.

Every function cites the reference lines it restates (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np
import torch

MASK_VALUE = -float(1 << 10)  # trainer.py:137  rating[...] = -(1<<10)


# --------------------------------------------------------------------------
# 9.1 graph  (dataloader.py:215-258)
# --------------------------------------------------------------------------
def degree_inv_sqrt(n_users: int, m_items: int, train_user, train_item) -> np.ndarray:
    """dinv[v] = deg[v]^-0.5 in fp32, 0 where deg == 0.

    dataloader.py:236-238: rowsum of the fp32 adjacency (multi-edges add up),
    np.power(rowsum, -0.5), inf -> 0.
    """
    tu = np.asarray(train_user, dtype=np.int64)
    ti = np.asarray(train_item, dtype=np.int64)
    deg = np.zeros(n_users + m_items, dtype=np.float32)
    np.add.at(deg, tu, np.float32(1.0))
    np.add.at(deg, ti + n_users, np.float32(1.0))
    with np.errstate(divide="ignore"):
        dinv = np.power(deg, np.float32(-0.5)).astype(np.float32)
    dinv[np.isinf(dinv)] = 0.0
    return dinv


def norm_adj_coo(n_users: int, m_items: int, train_user, train_item):
    """Coalesced, row-major sorted COO of  D^-1/2 [[0,R],[R^T,0]] D^-1/2.

    dataloader.py:226-244 builds it through scipy dok/lil; the values are
    fl32(fl32(dinv[i] * mult_ij) * dinv[j]) (two fp32 roundings, :242-243) and
    R[u,i] is the multiplicity of (u,i) in the train file (csr_matrix sums
    duplicates, :164).  Returns (row int64, col int64, val fp32).
    """
    N = n_users + m_items
    tu = np.asarray(train_user, dtype=np.int64)
    ti = np.asarray(train_item, dtype=np.int64) + n_users
    r = np.concatenate([tu, ti])
    c = np.concatenate([ti, tu])
    key = r * N + c
    uniq, mult = np.unique(key, return_counts=True)  # sorted => row-major order
    row = uniq // N
    col = uniq % N
    dinv = degree_inv_sqrt(n_users, m_items, train_user, train_item)
    left = (dinv[row] * mult.astype(np.float32)).astype(np.float32)
    val = (left * dinv[col]).astype(np.float32)
    return row, col, val


def norm_adj_scipy_as_shipped(n_users: int, m_items: int, train_user, train_item):
    """The SAME matrix through the scipy route the reference takes as shipped (dataloader.py:226-244:
    an empty dok -> lil, block assignment of R and R^T, back to dok, row sums, D^-1/2 as sp.diags, two
    sparse products, csr).  Only used to TIME the reference's graph build (bench.py's `ingest` leg) and
    to check that `norm_adj_coo` above yields the identical values; R = csr_matrix(ones, (u, i)) is the
    commented-out `UserItemNet` of dataloader.py:164-165.  Returns scipy csr fp32."""
    import scipy.sparse as sp
    tu = np.asarray(train_user, dtype=np.int64)
    ti = np.asarray(train_item, dtype=np.int64)
    N = n_users + m_items
    R = sp.csr_matrix((np.ones(len(tu)), (tu, ti)), shape=(n_users, m_items)).tolil()
    adj = sp.dok_matrix((N, N), dtype=np.float32).tolil()
    adj[:n_users, n_users:] = R
    adj[n_users:, :n_users] = R.T
    adj = adj.todok()
    rowsum = np.array(adj.sum(axis=1))
    with np.errstate(divide="ignore"):
        d_inv = np.power(rowsum, -0.5).flatten()
    d_inv[np.isinf(d_inv)] = 0.0
    d_mat = sp.diags(d_inv)
    return d_mat.dot(adj).dot(d_mat).tocsr()


def sparse_graph(n_users: int, m_items: int, train_user, train_item, folds: int = 0):
    """torch sparse graph as `Loader.getSparseGraph` returns it.

    dataloader.py:207-213,251-257: coalesced COO FloatTensor; with A_split a list
    of `folds` row slices of N//folds rows, the last one taking the remainder
    (:195-205).
    """
    N = n_users + m_items
    row, col, val = norm_adj_coo(n_users, m_items, train_user, train_item)
    if not folds:
        idx = torch.from_numpy(np.stack([row, col]))
        return torch.sparse_coo_tensor(idx, torch.from_numpy(val), (N, N)).coalesce()
    out = []
    fold_len = N // folds
    for f in range(folds):
        lo = f * fold_len
        hi = N if f == folds - 1 else (f + 1) * fold_len
        sel = (row >= lo) & (row < hi)
        idx = torch.from_numpy(np.stack([row[sel] - lo, col[sel]]))
        out.append(torch.sparse_coo_tensor(idx, torch.from_numpy(val[sel]), (hi - lo, N)).coalesce())
    return out


# --------------------------------------------------------------------------
# 9.2 propagation  (model/MF.py:178-210 == model/lgcn.py:78-86)
# --------------------------------------------------------------------------
def computer(all_emb: torch.Tensor, graph, n_layers: int, n_users: int):
    """X0=E; X_{k+1} = A_hat X_k; OUT = mean_k X_k; split users/items.

    model/MF.py:196-208 (`torch.sparse.mm` per layer, per fold when the graph is
    a list, then stack+mean).  Differentiable (torch autograd) like the
    reference.
    """
    embs = [all_emb]
    x = all_emb
    for _ in range(n_layers):
        if isinstance(graph, (list, tuple)):
            x = torch.cat([torch.sparse.mm(g, x) for g in graph], dim=0)
        else:
            x = torch.sparse.mm(graph, x)
        embs.append(x)
    out = torch.mean(torch.stack(embs, dim=1), dim=1)
    return out[:n_users], out[n_users:]


# --------------------------------------------------------------------------
# 9.3 BPR step  (model/lgcn.py:88-133)
# --------------------------------------------------------------------------
def bpr_loss(all_emb: torch.Tensor, graph, n_layers: int, n_users: int,
             users: torch.Tensor, pos: torch.Tensor, neg: torch.Tensor):
    """(loss, reg_loss) exactly as model/lgcn.py:98-118.

    reg = 0.5*(|u0|^2+|p0|^2+|n0|^2)/len(users) on the EGO rows (:102-110);
    loss = mean softplus(neg_score - pos_score) on propagated rows (:111-116).
    """
    all_users, all_items = computer(all_emb, graph, n_layers, n_users)
    users, pos, neg = users.long(), pos.long(), neg.long()
    u, p, q = all_users[users], all_items[pos], all_items[neg]
    u0, p0, q0 = all_emb[users], all_emb[pos + n_users], all_emb[neg + n_users]
    reg = 0.5 * (u0.norm(2).pow(2) + p0.norm(2).pow(2) + q0.norm(2).pow(2)) / float(len(users))
    pos_scores = torch.sum(u * p, dim=1)
    neg_scores = torch.sum(u * q, dim=1)
    loss = torch.mean(torch.nn.functional.softplus(neg_scores - pos_scores))
    return loss, reg


class OracleModel:
    """Holds E (N x d, users first), the graph and an Adam, like model/lgcn.py:44-68."""

    def __init__(self, n_users, m_items, train_user, train_item, emb: torch.Tensor,
                 n_layers: int, lr: float, decay: float, folds: int = 0):
        self.n_users, self.m_items = n_users, m_items
        self.n_layers, self.decay = n_layers, decay
        self.graph = sparse_graph(n_users, m_items, train_user, train_item, folds)
        self.weight = emb.detach().clone().float().requires_grad_(True)
        self.optim = torch.optim.Adam([self.weight], lr=lr)  # model/lgcn.py:63

    def computer(self):
        return computer(self.weight, self.graph, self.n_layers, self.n_users)

    def stage_one(self, users, pos, neg) -> torch.Tensor:
        """model/lgcn.py:127-133: zero_grad, loss + decay*reg, backward, Adam step."""
        self.optim.zero_grad()
        loss, reg = bpr_loss(self.weight, self.graph, self.n_layers, self.n_users, users, pos, neg)
        total = loss + self.decay * reg
        total.backward()
        self.optim.step()
        return total.detach()

    def one_epoch(self, users, pos, neg, batch_size: int) -> torch.Tensor:
        """model/lgcn.py:135-151: contiguous slices; divisor len//B + 1 (sic)."""
        total_batch = len(users) // batch_size + 1
        aver = torch.zeros(())
        for i in range(0, len(users), batch_size):
            aver = aver + self.stage_one(users[i:i + batch_size], pos[i:i + batch_size], neg[i:i + batch_size])
        return aver / total_batch

    def users_rating(self, users: torch.Tensor) -> torch.Tensor:
        """model/lgcn.py:120-125: raw scores, no sigmoid."""
        with torch.no_grad():
            au, ai = self.computer()
            return torch.matmul(au[users.long()], ai.t())


def closed_form_grad(all_emb, graph, n_layers, n_users, users, pos, neg, decay):
    """Horner-form gradient of (loss + decay*reg) wrt E.

    Not in the reference (it uses autograd); kept here so tests can check the
    identity the CUDA backward relies on:  H0=G, H_{j+1}=G+A_hat H_j,
    grad = H_K/(K+1) + decay/B * scatter(ego rows).
    """
    with torch.no_grad():
        B = len(users)
        au, ai = computer(all_emb, graph, n_layers, n_users)
        out = torch.cat([au, ai])
        u, p, q = users.long(), pos.long() + n_users, neg.long() + n_users
        x = (out[u] * out[q]).sum(1) - (out[u] * out[p]).sum(1)
        sig = torch.sigmoid(x) / B
        G = torch.zeros_like(all_emb)
        G.index_add_(0, u, sig[:, None] * (out[q] - out[p]))
        G.index_add_(0, p, -sig[:, None] * out[u])
        G.index_add_(0, q, sig[:, None] * out[u])
        g = sparse_graph_full(graph)
        H = G.clone()
        for _ in range(n_layers):
            H = G + torch.sparse.mm(g, H)
        grad = H / (n_layers + 1)
        ego = torch.zeros_like(all_emb)
        for idx in (u, p, q):
            ego.index_add_(0, idx, all_emb[idx])
        return grad + (decay / B) * ego


def sparse_graph_full(graph):
    if isinstance(graph, (list, tuple)):
        return torch.cat(list(graph), dim=0).coalesce()
    return graph


# --------------------------------------------------------------------------
# 9.4 sampler  (negative_sample.py:98-134)
# --------------------------------------------------------------------------
def uniform_sample_core(all_pos: Sequence[np.ndarray], sample_users: np.ndarray,
                        next_int: Callable[[int, int], int], n_neg: int = 1,
                        pos_index: Callable[[int, int, int], int] = None,
                        max_tries: int = 0) -> np.ndarray:
    """The reference's per-sample decision procedure, RNG factored out.

    negative_sample.py:113-130: for each drawn user IN ORDER: empty positive
    list -> emit nothing (:116-117); positive = allPos[user][randint(len)] in
    FILE ORDER, duplicates weigh the draw (:119-120); negative = first
    randint(m_items) whose value is not contained in allPos[user] (:121-126).
    `next_int(i, k)` returns the next draw in [0,k) for sample i.  `pos_index(i, user, len)`
    replaces the uniform positive pick (negative_sample.py:53-56, weighted by probs[user]).
    `max_tries` > 0 (the GPU spec, LGCN_MAX_NEG_TRIES): after that many rejected candidates in a
    row the sample and the rest of its negatives are dropped — the reference's `while True` never
    returns for a user whose positives cover every item.
    """
    S = []
    for i, user in enumerate(sample_users):
        P = all_pos[int(user)]
        if len(P) == 0:
            continue
        positem = P[next_int(i, len(P)) if pos_index is None else pos_index(i, int(user), len(P))]
        for _ in range(n_neg):  # n_neg > 1: flat (u, pos, neg_t) rows, the lgcnssm.py:141 batch layout
            tries, neg = 0, None
            while True:
                if max_tries and tries == max_tries:
                    neg = None
                    break
                neg = next_int(i, -1)
                tries += 1
                if neg in P:
                    continue
                break
            if neg is None:
                break
            S.append([int(user), int(positem), int(neg)])
    return np.array(S, dtype=np.int64).reshape(-1, 3)


def uniform_sample_mt(all_pos, n_users: int, m_items: int, count: int) -> np.ndarray:
    """`UniformSample` with the reference's own RNG (numpy legacy global MT19937).

    negative_sample.py:106-107: `count = trainDataSize` users in ONE vectorised
    randint, then the sequential loop.  Seed with np.random.seed(s) before the
    call to compare with the live reference bit for bit.
    """
    sample_users = np.random.randint(0, n_users, count)

    def next_int(_i, k):
        return np.random.randint(0, m_items if k < 0 else k)

    return uniform_sample_core(all_pos, sample_users, next_int)


MAX_NEG_TRIES = 256   # include/lgcn_b200.h LGCN_MAX_NEG_TRIES

_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11; Random123 v1.09 constants).

    ctr: uint32[..., 4], key: uint32[..., 2] -> uint32[..., 4].  The reference
    has no counter-based RNG (it uses MT19937); this is the generator the CUDA
    sampler is specified on.  Checked against the Random123
    known-answer vectors in tests/test_oracle_golden.py.
    """
    c = [ctr[..., i].astype(np.uint32).copy() for i in range(4)]
    k0 = key[..., 0].astype(np.uint32).copy()
    k1 = key[..., 1].astype(np.uint32).copy()
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = _PH_M0 * c[0].astype(np.uint64)
        p1 = _PH_M1 * c[2].astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        with np.errstate(over="ignore"):
            k0 = (k0 + _PH_W0).astype(np.uint32)
            k1 = (k1 + _PH_W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def philox_randint(r: int, k: int) -> int:
    """SURVEY §9.4:  ri(k, r) := (u64(r) * k) >> 32  stands in for randint(0, k)."""
    return (int(r) * int(k)) >> 32


def uniform_sample_philox(all_pos: Sequence[np.ndarray], n_users: int, m_items: int,
                          count: int, seed: int, epoch: int, n_neg: int = 1,
                          pos_cdf: Sequence[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Sampler with the "same uniform draws" contract of SURVEY §9.4.

    Draw j of sample i is word (j % 4) of Philox4x32-10(key=(seed_lo, seed_hi),
    ctr=(i_lo, i_hi, j // 4, epoch)).  Draw 0 -> user, 1 -> positive index,
    2.. -> negative candidates.  Decisions are `uniform_sample_core`'s.
    Returns (S int64[n_s,3], valid bool[count]).
    """
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    idx = np.arange(count, dtype=np.uint64)
    ctr0 = np.zeros((count, 4), dtype=np.uint32)
    ctr0[:, 0] = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr0[:, 1] = (idx >> np.uint64(32)).astype(np.uint32)
    ctr0[:, 3] = np.uint32(epoch & 0xFFFFFFFF)
    first = philox4x32_10(ctr0, np.broadcast_to(key, (count, 2)))
    sample_users = np.array([philox_randint(first[i, 0], n_users) for i in range(count)], dtype=np.int64)
    cursor: Dict[int, int] = {}
    block_cache: Dict[Tuple[int, int], np.ndarray] = {}

    def draw(i: int) -> int:
        j = cursor.get(i, 1)
        cursor[i] = j + 1
        blk = j // 4
        if blk == 0:
            return int(first[i, j])
        ck = (i, blk)
        if ck not in block_cache:
            c = ctr0[i].copy()
            c[2] = np.uint32(blk)
            block_cache[ck] = philox4x32_10(c[None, :], key[None, :])[0]
        return int(block_cache[ck][j % 4])

    def next_int(i, k):
        return philox_randint(draw(i), m_items if k < 0 else k)

    def pos_index(i, user, n):
        # weighted pick (our spec of np.random.choice's inverse-CDF step): u = (word >> 8) * 2^-24 as
        # fp32, first j with cdf[j] > u, clamped; pos_cdf[user] is the fp32 normalised cumulative table
        u = np.float32((draw(i) >> 8) * 5.9604644775390625e-08)
        j = int(np.searchsorted(pos_cdf[user], u, side="right"))
        return min(j, n - 1)

    valid = np.array([len(all_pos[int(u)]) > 0 for u in sample_users], dtype=bool)
    S = uniform_sample_core(all_pos, sample_users, next_int, n_neg, pos_index if pos_cdf is not None else None,
                            max_tries=MAX_NEG_TRIES)
    return S, valid


# --------------------------------------------------------------------------
# 9.5 full-rank eval  (model/lgcn.py:120-125, trainer.py:125-138)
# --------------------------------------------------------------------------
def masked_topk(rating: torch.Tensor, all_pos_batch: Sequence[np.ndarray], k: int):
    """Mask train positives with -1024 (assignment, trainer.py:132-137) and take
    the top-k.  trainer.py:138 uses torch.topk whose tie order is unspecified;
    north_star pins "ties broken by item id", i.e. a stable descending sort.
    Returns (values fp32[U,k], indices int64[U,k]).
    """
    rating = rating.clone()
    rows, cols = [], []
    for r, items in enumerate(all_pos_batch):
        rows.extend([r] * len(items))
        cols.extend(int(i) for i in items)
    if rows:
        rating[rows, cols] = MASK_VALUE
    vals, idx = torch.sort(rating, dim=1, descending=True, stable=True)
    return vals[:, :k].contiguous(), idx[:, :k].contiguous()


# --------------------------------------------------------------------------
# 9.6 metrics  (utils.py:40-48, metric.py:60-103, trainer.py:162-170,263-280)
# --------------------------------------------------------------------------
def get_label(ground_true: Sequence[Sequence[int]], topk: np.ndarray) -> np.ndarray:
    """utils.py:40-48: r[b,j] = topk[b,j] in groundTrue[b], as float64."""
    r = np.zeros(topk.shape, dtype=np.float64)
    for b in range(len(ground_true)):
        gt = set(int(x) for x in ground_true[b])
        r[b] = [float(int(x) in gt) for x in topk[b]]
    return r


def batch_metrics(ground_true, topk: np.ndarray, ks: Sequence[int]) -> Dict[str, np.ndarray]:
    """Per-batch SUMS of recall / precision / hr / ndcg for each k.

    metric.py:60-72 (recall divides by len(gt)+1e-6, precision by k, hr counts
    rows with >=1 hit) and metric.py:84-103 (DCG with 1/log2(j+2), IDCG over
    min(k, len(gt)), idcg==0 -> 1, nan -> 0); trainer.py:269-280 loops ks.
    """
    r = get_label(ground_true, topk)
    n_gt = np.array([len(g) for g in ground_true])
    res = {m: [] for m in ("recall", "precision", "hr", "ndcg")}
    for k in ks:
        right = r[:, :k].sum(1)
        res["recall"].append(np.sum(right / (n_gt + 1e-6)))
        res["precision"].append(np.sum(right) / k)
        res["hr"].append(float(np.sum(right >= 1)))
        disc = 1.0 / np.log2(np.arange(2, k + 2))
        ideal = np.zeros((len(r), k))
        for i, g in enumerate(ground_true):
            ideal[i, :min(k, len(g))] = 1
        idcg = (ideal * disc).sum(1)
        dcg = (r[:, :k] * disc).sum(1)
        idcg[idcg == 0.0] = 1.0
        nd = dcg / idcg
        nd[np.isnan(nd)] = 0.0
        res["ndcg"].append(np.sum(nd))
    return {m: np.array(v, dtype=np.float64) for m, v in res.items()}


def evaluate(model: OracleModel, all_pos, test_dict: Dict[int, List[int]], ks: Sequence[int],
             u_batch_size: int):
    """trainer.py:115-170 restricted to the hot path: users = testDict keys in
    insertion order, batches of test_u_batch_size, rating -> mask -> top max(ks),
    metric sums / len(users).  Returns (results dict, list of top-k index arrays).
    """
    users = list(test_dict.keys())
    kmax = max(ks)
    tot = {m: np.zeros(len(ks)) for m in ("recall", "precision", "hr", "ndcg")}
    tops = []
    for s in range(0, len(users), u_batch_size):
        bu = users[s:s + u_batch_size]
        rating = model.users_rating(torch.tensor(bu, dtype=torch.long))
        _, idx = masked_topk(rating, [all_pos[u] for u in bu], kmax)
        tops.append(idx.numpy())
        bm = batch_metrics([test_dict[u] for u in bu], idx.numpy(), ks)
        for m in tot:
            tot[m] += bm[m]
    return {m: v / float(len(users)) for m, v in tot.items()}, tops


# --------------------------------------------------------------------------
# f-4 variants on the same kernels: rAdjGCN, RGCN, PyG-LGConv form,
# capped sampler.  Pinned by oracle/make_golden_variants.py against the live
# reference classes (their un-vendored imports torch_geometric / torch_scatter
# are stubbed with the semantics restated below).
# --------------------------------------------------------------------------
def directed_edges(n_users: int, train_user, train_item, extra_user=None, extra_item=None):
    """train_edge of model/radj.py:21-27 / edge_index of model/rgcn.py:54-85:
    [users -> items+n | items+n -> users], then the favourite edges the same way."""
    tu = torch.as_tensor(np.asarray(train_user), dtype=torch.int64)
    ti = torch.as_tensor(np.asarray(train_item), dtype=torch.int64) + n_users
    e = torch.cat([torch.stack([tu, ti]), torch.stack([ti, tu])], dim=1)
    if extra_user is not None:
        fu = torch.as_tensor(np.asarray(extra_user), dtype=torch.int64)
        fi = torch.as_tensor(np.asarray(extra_item), dtype=torch.int64) + n_users
        e = torch.cat([e, torch.stack([fu, fi]), torch.stack([fi, fu])], dim=1)
    return e


def scatter_sum(src: torch.Tensor, index: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """torch_scatter.scatter(src, index, out=out, dim=0) with the default reduce='sum'
    (un-vendored dependency, unpinned version; call sites model/radj.py:44, model/lgcn.py:41)."""
    return out.index_add(0, index, src)


def radj_conv(x: torch.Tensor, edge: torch.Tensor, n_nodes: int, r: float) -> torch.Tensor:
    """rAdjConv.forward, model/radj.py:28-45: deg = multiplicity-counting out-degree of
    train_edge[0] (0 -> 1e-6), div_e = deg[src]^r * deg[dst]^(1-r), out[dst] += x[src] / div_e."""
    all_div = torch.zeros(n_nodes)
    value, count = torch.unique(edge[0], return_counts=True)
    all_div[value] = count.float()
    all_div[all_div == 0] = 1e-6
    div = (all_div[edge[0]] ** r) * (all_div[edge[1]] ** (1 - r))
    msg = x[edge[0]] / div.unsqueeze(1)
    return scatter_sum(msg, edge[1], torch.zeros_like(x))


def radj_forward(all_emb: torch.Tensor, edge: torch.Tensor, n_layers: int, n_users: int, r: float):
    """rAdjGCN.forward, model/radj.py:84-92."""
    x = all_emb
    x_out = x
    for _ in range(n_layers):
        x = radj_conv(x, edge, all_emb.shape[0], r)
        x_out = x_out + x
    x_out = x_out / (1 + n_layers)
    return x_out[:n_users], x_out[n_users:]


def lgconv_pyg(x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """PyG `LGConv(normalize=True)` = gcn_norm(add_self_loops=False) + sum aggregation
    (un-vendored torch_geometric, version unpinned; call sites model/lgcn.py:64-66,82,
    model/rgcn.py:97,112): deg = scatter_add(1, col); w_e = deg[row]^-1/2 deg[col]^-1/2
    (inf -> 0); out_i = sum_{e: col_e = i} w_e x[row_e].  SURVEY §8c."""
    row, col = edge_index[0], edge_index[1]
    n = x.shape[0]
    deg = torch.zeros(n, dtype=x.dtype).index_add(0, col, torch.ones(col.numel(), dtype=x.dtype))
    dis = deg.pow(-0.5)
    dis[dis == float("inf")] = 0
    w = dis[row] * dis[col]
    return torch.zeros_like(x).index_add(0, col, w.unsqueeze(1) * x[row])


def lgconv_forward(all_emb: torch.Tensor, edge_index: torch.Tensor, n_layers: int, n_users: int):
    """LightGCN.forward of model/lgcn.py:78-86 (running sum, then / (1 + K)); the same loop is
    RGCN.forward (model/rgcn.py:108-116) on the purchase + favourite edge_index."""
    x = all_emb
    x_out = x
    for _ in range(n_layers):
        x = lgconv_pyg(x, edge_index)
        x_out = x_out + x
    x_out = x_out / (1 + n_layers)
    return x_out[:n_users], x_out[n_users:]


def bpr_loss_from(all_users, all_items, all_emb, n_users, users, pos, neg):
    """bpr_loss body shared by model/lgcn.py:98-118, model/radj.py:105-126, model/rgcn.py:129-150."""
    users, pos, neg = users.long(), pos.long(), neg.long()
    u, p, q = all_users[users], all_items[pos], all_items[neg]
    u0, p0, q0 = all_emb[users], all_emb[pos + n_users], all_emb[neg + n_users]
    reg = 0.5 * (u0.norm(2).pow(2) + p0.norm(2).pow(2) + q0.norm(2).pow(2)) / float(len(users))
    loss = torch.mean(torch.nn.functional.softplus(torch.sum(u * q, dim=1) - torch.sum(u * p, dim=1)))
    return loss, reg


def capped_sample_core(all_pos: Sequence[np.ndarray], sample_users: np.ndarray,
                       next_int: Callable[[int, int], int], limit: int) -> np.ndarray:
    """The DDP script's sampler, ddp_lgcn.py:541-582: as uniform_sample_core, but a sample whose
    positive item was already emitted `limit` (POSITIVE_NUM_LIMIT = 3000) times this epoch is
    dropped BEFORE its negative is drawn (:569-570); order dependent."""
    S = []
    oc: Dict[int, int] = {}
    for i, user in enumerate(sample_users):
        P = all_pos[int(user)]
        if len(P) == 0:
            continue
        positem = int(P[next_int(i, len(P))])
        if oc.get(positem, 0) >= limit:
            continue
        oc[positem] = oc.get(positem, 0) + 1
        while True:
            neg = next_int(i, -1)
            if neg in P:
                continue
            break
        S.append([int(user), positem, int(neg)])
    return np.array(S, dtype=np.int64).reshape(-1, 3)


def capped_sample_mt(all_pos, m_items: int, count: int, limit: int) -> np.ndarray:
    """ddp_lgcn.py:549-551 with the reference's RNG: users from randint(0, len(allPos), count),
    count = trainDataSize * TRAIN_ITERATIVE."""
    sample_users = np.random.randint(0, len(all_pos), count)

    def next_int(_i, k):
        return np.random.randint(0, m_items if k < 0 else k)

    return capped_sample_core(all_pos, sample_users, next_int, limit)


def capped_sample_philox(all_pos, n_users: int, m_items: int, count: int, seed: int, epoch: int,
                         limit: int) -> np.ndarray:
    """Capped sampler on the Philox draw mapping of `uniform_sample_philox` (our spec): every
    sample owns its counter stream, so dropping a sample never shifts another sample's draws and
    the cap is "rank of sample i among the earlier non-empty samples with the same positive"."""
    S, _ = uniform_sample_philox(all_pos, n_users, m_items, count, seed, epoch)
    keep = np.ones(len(S), dtype=bool)
    oc: Dict[int, int] = {}
    for t, p in enumerate(S[:, 1].tolist()):
        c = oc.get(p, 0)
        if c >= limit:
            keep[t] = False
        else:
            oc[p] = c + 1
    return S[keep]


def weighted_sample_mt(all_pos, probs, n_users: int, m_items: int, count: int) -> np.ndarray:
    """`UniformSampling.sample` / `sample_parallel` with `sample_pow != 0` under the reference's RNG
    (negative_sample.py:39-69,73-74): users from one vectorised randint, positive index by
    np.random.choice(len(pos), p=probs[user]) (:56), negatives by rejection."""
    sample_users = np.random.randint(0, n_users, count)

    def next_int(_i, k):
        return np.random.randint(0, m_items if k < 0 else k)

    def pos_index(_i, user, n):
        return int(np.random.choice(n, p=probs[user]))

    return uniform_sample_core(all_pos, sample_users, next_int, 1, pos_index)


def normalised_cdf(p: np.ndarray) -> np.ndarray:
    """np.random.choice's table: cdf = p.cumsum(); cdf /= cdf[-1] (float64), stored as fp32."""
    c = np.cumsum(np.asarray(p, dtype=np.float64))
    return (c / c[-1]).astype(np.float32)


def dropout_graph(graph: torch.Tensor, mask: torch.Tensor, keep_prob: float) -> torch.Tensor:
    """`__dropout_x` of model/MF.py:158-166 with the Bernoulli draw factored out: entries of the
    coalesced graph where `mask` is False are removed, the others divided by keep_prob.  The
    reference's own draw is `(torch.rand(len(values)) + keep_prob).int().bool()` (:162-163)."""
    idx, val = graph.indices(), graph.values()
    mask = torch.as_tensor(mask, dtype=torch.bool)
    return torch.sparse_coo_tensor(idx[:, mask], val[mask] / keep_prob, graph.shape)
