"""Dataset contract of the reference's `dataloader.Loader`, CSR-backed.

Mirrors reference dataloader.py:66-297 for the attributes the hot path reads
(`n_users, m_items, n_user, m_item, trainUser, trainItem, trainDataSize, allPos,
testDict, getUserPosItems(), getSparseGraph()`, reference model/lgcn.py:49-54,
negative_sample.py:106-108, trainer.py:86,93-94) and adds the device-side
structures the kernels consume (`csr_graph()`, `pos_csr()`, `test_csr()`).

Differences that are deliberate:
  * `allPos` is a CSR-backed sequence (a Python list of 10 M numpy arrays does not
    scale, SURVEY §7 hard part 7); `allPos[u]` still returns the user's train line
    in file order as an int64 numpy array (dataloader.py:118).
  * `getSparseGraph()` does not need the commented-out `UserItemNet`
    (dataloader.py:163-165) and never reads a stale `s_pre_adj_mat.npz`
    (dataloader.py:218-221): the graph is rebuilt from the edge list on device.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .graph import CsrGraph, build_csr_graph, build_pos_csr, graph_to_sparse_coo


class PosLists:
    """Sequence view over a host CSR: obj[u] -> np.int64 array (file order)."""

    def __init__(self, rowptr: np.ndarray, items: np.ndarray):
        self.rowptr = rowptr
        self.items = items

    def __len__(self) -> int:
        return len(self.rowptr) - 1

    def __getitem__(self, u):
        u = int(u)
        if u < 0:
            u += len(self)
        return self.items[self.rowptr[u]:self.rowptr[u + 1]]

    def __iter__(self):
        for u in range(len(self)):
            yield self[u]


class BasicDataset:
    """Interactions held as flat arrays; everything else is derived lazily."""

    def __init__(self, n_users: int, m_items: int, train_user, train_item, test_user, test_item,
                 config: Optional[dict] = None, device: Optional[str] = None):
        self.config = dict(config or {})
        self.n_user, self.m_item = int(n_users), int(m_items)
        self.trainUser = np.asarray(train_user, dtype=np.int64)
        self.trainItem = np.asarray(train_item, dtype=np.int64)
        self.testUser = np.asarray(test_user, dtype=np.int64)
        self.testItem = np.asarray(test_item, dtype=np.int64)
        self.traindataSize = int(len(self.trainUser))
        self.testDataSize = int(len(self.testUser))
        self.split = bool(self.config.get("A_split", False))
        self.folds = int(self.config.get("A_n_fold", 1))
        self.device = torch.device(device or self.config.get("device", "cuda:0"))
        self.Graph = None
        self._csr: Optional[CsrGraph] = None
        self._pos_dev = None
        self._test_dev = None
        self._allPos: Optional[PosLists] = None
        self._testDict: Optional[Dict[int, List[int]]] = None

    # ---- reference property names (dataloader.py:175-193) ----
    @property
    def n_users(self) -> int:
        return self.n_user

    @property
    def m_items(self) -> int:
        return self.m_item

    @property
    def trainDataSize(self) -> int:
        return self.traindataSize

    @property
    def allPos(self) -> PosLists:
        if self._allPos is None:
            order = np.argsort(self.trainUser, kind="stable")
            cnt = np.bincount(self.trainUser, minlength=self.n_user)
            rowptr = np.zeros(self.n_user + 1, dtype=np.int64)
            np.cumsum(cnt, out=rowptr[1:])
            self._allPos = PosLists(rowptr, self.trainItem[order])
        return self._allPos

    @property
    def testDict(self) -> Dict[int, List[int]]:
        """dataloader.py:260-272: {user: [items]} in first-appearance order."""
        if self._testDict is None:
            d: Dict[int, List[int]] = {}
            for u, i in zip(self.testUser.tolist(), self.testItem.tolist()):
                d.setdefault(u, []).append(i)
            self._testDict = d
        return self._testDict

    def getUserPosItems(self, users: Sequence[int]):
        """dataloader.py:286-291"""
        ap = self.allPos
        return [ap[u] for u in users]

    # ---- graph ----
    def csr_graph(self) -> CsrGraph:
        if self._csr is None:
            tu = torch.from_numpy(self.trainUser).to(self.device)
            ti = torch.from_numpy(self.trainItem).to(self.device)
            self._csr = build_csr_graph(self.n_user, self.m_item, tu, ti)
        return self._csr

    def getSparseGraph(self):
        """dataloader.py:215-258 contract: coalesced COO FloatTensor of
        D^-1/2 A D^-1/2 on the device, or the list of `A_n_fold` row slices when
        `A_split` (dataloader.py:195-205)."""
        if self.Graph is None:
            full = graph_to_sparse_coo(self.csr_graph())
            if not self.split:
                self.Graph = full
            else:
                N = self.n_user + self.m_item
                idx, val = full.indices(), full.values()
                fold_len = N // self.folds
                out = []
                for f in range(self.folds):
                    lo = f * fold_len
                    hi = N if f == self.folds - 1 else (f + 1) * fold_len
                    sel = (idx[0] >= lo) & (idx[0] < hi)
                    sub = torch.stack([idx[0][sel] - lo, idx[1][sel]])
                    out.append(torch.sparse_coo_tensor(sub, val[sel], (hi - lo, N)).coalesce())
                self.Graph = out
        return self.Graph

    # ---- device CSRs for the sampler / eval kernels ----
    def pos_csr(self):
        """(rowptr int64[n+1], file-order items int32, sorted items int32) on device."""
        if self._pos_dev is None:
            tu = torch.from_numpy(self.trainUser).to(self.device)
            ti = torch.from_numpy(self.trainItem).to(self.device)
            self._pos_dev = build_pos_csr(self.n_user, tu, ti)
        return self._pos_dev

    def test_csr(self):
        """(rowptr int64[n+1], sorted test items int32) on device."""
        if self._test_dev is None:
            tu = torch.from_numpy(self.testUser).to(self.device)
            ti = torch.from_numpy(self.testItem).to(self.device)
            rp, _, srt = build_pos_csr(self.n_user, tu, ti)
            self._test_dev = (rp, srt)
        return self._test_dev

    def test_users(self) -> np.ndarray:
        """list(testDict.keys()) order (trainer.py:118) without building the dict."""
        _, first = np.unique(self.testUser, return_index=True)
        return self.testUser[np.sort(first)]


class DeviceDataset(BasicDataset):
    """Same contract, but the interactions live on the GPU and never visit the host (cfg-3-sized
    graphs: 500 M edges).  Host-side views (`allPos`, `testDict`, `trainUser`) are materialised
    lazily and only make sense for small cases."""

    def __init__(self, n_users: int, m_items: int, train_user: torch.Tensor, train_item: torch.Tensor,
                 test_user: torch.Tensor, test_item: torch.Tensor, config: Optional[dict] = None):
        self.config = dict(config or {})
        self.n_user, self.m_item = int(n_users), int(m_items)
        self._tu, self._ti, self._su, self._si = train_user, train_item, test_user, test_item
        self.traindataSize, self.testDataSize = int(train_user.numel()), int(test_user.numel())
        self.split = bool(self.config.get("A_split", False))
        self.folds = int(self.config.get("A_n_fold", 1))
        self.device = train_user.device
        self.Graph = None
        self._csr = self._pos_dev = self._test_dev = self._allPos = self._testDict = None

    trainUser = property(lambda self: self._tu.cpu().numpy())
    trainItem = property(lambda self: self._ti.cpu().numpy())
    testUser = property(lambda self: self._su.cpu().numpy())
    testItem = property(lambda self: self._si.cpu().numpy())

    def csr_graph(self) -> CsrGraph:
        if self._csr is None:
            self._csr = build_csr_graph(self.n_user, self.m_item, self._tu, self._ti)
        return self._csr

    def pos_csr(self):
        if self._pos_dev is None:
            self._pos_dev = build_pos_csr(self.n_user, self._tu, self._ti)
        return self._pos_dev

    def test_csr(self):
        if self._test_dev is None:
            rp, _, srt = build_pos_csr(self.n_user, self._su, self._si)
            self._test_dev = (rp, srt)
        return self._test_dev


class Loader(BasicDataset):
    """`Loader(config, path)`: parses `{path}/{suffix}/train{suffix}.txt` and
    `test{suffix}.txt` ("uid item item ..." per line, dataloader.py:93-150).

    n_user / m_item = 1 + max id over both files (:119-120,145-146,151-152); the
    train file must list uids 0..n-1 ascending (allPos is appended per line but
    indexed by uid, :118 vs :289) — checked here instead of failing later.
    `for_lgbm` / `cold_start` splits (:100-113) are reference product features
    outside the hot path and are rejected explicitly."""

    def __init__(self, config: dict, path: str = "./data/cf", device: Optional[str] = None, ingest: str = "auto"):
        if config.get("for_lgbm") or config.get("cold_start"):
            raise NotImplementedError("for_lgbm / cold_start splits are outside the LightGCN hot path")
        suffix = config.get("suffix", "")
        self.path = path
        dev = torch.device(device or config.get("device", "cuda:0"))
        # "device": the file's bytes are parsed by the lgcn_ingest_* kernels; "host": the Python loop
        # (the only one that honours --test's early stop, dataloader.py:122-124); "auto": device when
        # the dataset lives on a GPU and --test is off
        if ingest == "auto":
            ingest = "device" if dev.type == "cuda" and torch.cuda.is_available() and not config.get("test") else "host"
        parse = (lambda f: self._parse_device(f, dev)) if ingest == "device" else (lambda f: self._parse(f, bool(config.get("test"))))
        self.ingest = ingest
        tr_u, tr_i = parse(f"{path}/{suffix}/train{suffix}.txt")
        te_u, te_i = parse(f"{path}/{suffix}/test{suffix}.txt")
        n = int(max(tr_u.max(initial=-1), te_u.max(initial=-1))) + 1
        m = int(max(tr_i.max(initial=-1), te_i.max(initial=-1))) + 1
        uniq = np.unique(tr_u)
        if len(uniq) and (np.any(np.diff(tr_u) < 0) or uniq[0] != 0 or uniq[-1] != len(uniq) - 1):
            raise ValueError("train file must list uids 0..n-1 in ascending order without gaps")
        super().__init__(n, m, tr_u, tr_i, te_u, te_i, config=config, device=device)

    @staticmethod
    def _parse_device(fname: str, device):
        """The same (users, items) arrays from the GPU text parser (csrc/ingest.cu)."""
        from . import ops
        raw = np.fromfile(fname, dtype=np.uint8)
        if raw.size == 0:
            return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
        u, i = ops.ingest_text(torch.from_numpy(raw).to(device))
        return u.cpu().numpy(), i.cpu().numpy()

    @staticmethod
    def _parse(fname: str, truncate: bool):
        users, items = [], []
        with open(fname) as f:
            for line in f:
                line = line.strip("\n")
                if not line:
                    continue
                tok = line.split(" ")
                uid = int(tok[0])
                its = [int(t) for t in tok[1:] if t != ""]
                users.extend([uid] * len(its))
                items.extend(its)
                if truncate and uid == 100:  # dataloader.py:122-124 (--test)
                    break
        return np.asarray(users, dtype=np.int64), np.asarray(items, dtype=np.int64)


def write_reference_files(ds: BasicDataset, path: str, suffix: str = "") -> None:
    """Emit train/test txt in the reference's format (for oracle / golden runs)."""
    import os
    os.makedirs(f"{path}/{suffix}", exist_ok=True)
    for name, users, lists in (("train", range(ds.n_user), ds.allPos),
                               ("test", list(ds.testDict.keys()), None)):
        with open(f"{path}/{suffix}/{name}{suffix}.txt", "w") as f:
            for u in users:
                its = lists[u] if lists is not None else ds.testDict[u]
                f.write(" ".join([str(u)] + [str(int(i)) for i in its]) + "\n")
