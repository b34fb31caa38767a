"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the
CSR/decomposition builders agree with the oracle graph, the dataset contract and
the synthetic generator.  No compute call into the CUDA library happens here."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from furusato_recommend_b200 import _lib, graph as G
from furusato_recommend_b200.dataloader import BasicDataset, Loader, write_reference_files
from furusato_recommend_b200.synthetic import bipartite
from oracle import lgcn_oracle as orc

REPO = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    header = (REPO / "include" / "lgcn_b200.h").read_text()
    declared = set(re.findall(r"\b(lgcn_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()  # raises if the .so is missing or a symbol is absent
    for name in declared:
        assert hasattr(lib, name)
    assert lib.lgcn_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header(tmp_path):
    """ctypes mirrors vs the real C layout: compile a probe against include/lgcn_b200.h with gcc."""
    import subprocess
    src = tmp_path / "probe.c"
    fields_g = [f[0] for f in _lib.GraphStruct._fields_]
    fields_a = [f[0] for f in _lib.LayerArgs._fields_]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "lgcn_b200.h"', 'int main(void){',
             'printf("%zu %zu\\n", sizeof(lgcn_graph_t), sizeof(lgcn_layer_args_t));']
    lines += [f'printf("%zu\\n", offsetof(lgcn_graph_t, {f}));' for f in fields_g]
    lines += [f'printf("%zu\\n", offsetof(lgcn_layer_args_t, {f}));' for f in fields_a]
    lines += ['return 0;}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", str(REPO / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(_lib.GraphStruct) and int(out[1]) == ctypes.sizeof(_lib.LayerArgs)
    offs = [int(x) for x in out[2:]]
    want = [getattr(_lib.GraphStruct, f).offset for f in fields_g] + [getattr(_lib.LayerArgs, f).offset for f in fields_a]
    assert offs == want


def test_ops_reject_cpu_tensors():
    from furusato_recommend_b200 import ops
    x = torch.zeros(4, 32)
    with pytest.raises(_lib.LgcnLibraryError):
        ops.adam_tick(torch.zeros(1, dtype=torch.int64), torch.zeros(2), 1e-3)
    with pytest.raises(_lib.LgcnLibraryError):
        ops.score_dense_f32(x, x, torch.zeros(1, dtype=torch.int64))


def test_csr_matches_oracle_graph(golden):
    n, m = int(golden["n_users"]), int(golden["m_items"])
    tu, ti = torch.from_numpy(golden["train_user"]), torch.from_numpy(golden["train_item"])
    g = G.build_csr_graph(n, m, tu, ti)
    assert g.nnz == 2 * len(golden["train_user"])
    dinv = orc.degree_inv_sqrt(n, m, golden["train_user"], golden["train_item"])
    assert np.array_equal(g.dinv.numpy(), dinv)
    coo = G.graph_to_sparse_coo(g)
    assert np.array_equal(coo.indices()[0].numpy(), golden["adj_row"])
    assert np.array_equal(coo.indices()[1].numpy(), golden["adj_col"])
    assert np.array_equal(coo.values().numpy(), golden["adj_val"])  # bit-identical to the live reference
    # CSR SpMM semantics == oracle propagation (fp64 check of the folding identity)
    E = torch.from_numpy(golden["E0"]).double()
    z = g.dinv.double()[:, None] * E
    deg = (g.rowptr[1:] - g.rowptr[:-1])
    rows = torch.repeat_interleave(torch.arange(n + m), deg)
    s = torch.zeros_like(E).index_add_(0, rows, z[g.col.long()])
    x1 = g.dinv.double()[:, None] * s
    ref = torch.sparse.mm(orc.sparse_graph(n, m, golden["train_user"], golden["train_item"]).double(), E)
    assert torch.allclose(x1, ref, rtol=1e-6, atol=1e-7)


def test_decompose_rows_covers_every_edge_once():
    torch.manual_seed(0)
    deg = torch.cat([torch.randint(0, 40, (500,)), torch.tensor([257, 1024, 1025, 5000, 0, 256])])
    rowptr = torch.zeros(len(deg) + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(deg, 0)
    p = G.decompose_rows(rowptr)
    light = p["light_rows"].long()
    assert (deg[light] <= _lib.HUB_DEG).all()
    assert (deg[light][:-1] >= deg[light][1:]).all()  # degree-descending
    seen = torch.zeros(int(rowptr[-1]), dtype=torch.int32)
    for r in light.tolist():
        seen[rowptr[r]:rowptr[r + 1]] += 1
    for b, l, r, h in zip(p["seg_begin"].tolist(), p["seg_len"].tolist(), p["seg_row"].tolist(), p["seg_hub"].tolist()):
        assert 0 < l <= _lib.SEG_EDGES and rowptr[r] <= b and b + l <= rowptr[r + 1]
        seen[b:b + l] += 1
    assert (seen == 1).all()
    assert p["hub_nseg"].tolist() == [5, 2, 1, 1]  # 5000, 1025, 1024, 257
    assert p["hub_seg0"].tolist() == [0, 5, 7, 8]
    assert len(light) + 4 == len(deg)


def test_pos_csr_file_order_and_sorted(golden, tiny_lists):
    train, _ = tiny_lists
    n = int(golden["n_users"])
    rp, file_items, sorted_items = G.build_pos_csr(n, torch.from_numpy(golden["train_user"]),
                                                   torch.from_numpy(golden["train_item"]))
    for u in (0, 17, 29, 299):
        a, b = int(rp[u]), int(rp[u + 1])
        assert file_items[a:b].tolist() == train[u].tolist()
        assert sorted_items[a:b].tolist() == sorted(train[u].tolist())


def test_loader_contract_roundtrip(tmp_path, golden, tiny_lists):
    train, test = tiny_lists
    n, m = int(golden["n_users"]), int(golden["m_items"])
    ds = BasicDataset(n, m, golden["train_user"], golden["train_item"], golden["test_user"], golden["test_item"],
                      config={"A_split": False}, device="cpu")
    write_reference_files(ds, str(tmp_path), "x")
    ld = Loader({"suffix": "x", "A_split": True, "A_n_fold": 7}, path=str(tmp_path), device="cpu")
    assert (ld.n_users, ld.m_items, ld.trainDataSize) == (n, m, len(golden["train_user"]))
    assert np.array_equal(ld.trainUser, golden["train_user"]) and np.array_equal(ld.trainItem, golden["train_item"])
    assert list(ld.testDict.keys()) == list(test.keys()) and ld.testDict[3] == test[3]
    assert ld.test_users().tolist() == list(test.keys())
    assert ld.allPos[17].tolist() == train[17].tolist() and len(ld.allPos) == n
    assert [p.tolist() for p in ld.getUserPosItems([0, 5])] == [train[0].tolist(), train[5].tolist()]
    folds = ld.getSparseGraph()  # A_split contract: list of row slices (dataloader.py:195-205)
    ref = orc.sparse_graph(n, m, golden["train_user"], golden["train_item"], folds=7)
    assert len(folds) == 7
    for a, b in zip(folds, ref):
        assert a.shape == b.shape and torch.equal(a.indices(), b.indices()) and torch.equal(a.values(), b.values())


def test_loader_rejects_gappy_train_file(tmp_path):
    d = tmp_path / "s"
    d.mkdir()
    (d / "trains.txt").write_text("0 1 2\n2 3\n")
    (d / "tests.txt").write_text("0 4\n")
    with pytest.raises(ValueError):
        Loader({"suffix": "s"}, path=str(tmp_path), device="cpu")


def test_synthetic_generator_shape():
    n, m, tu, ti, su, si = bipartite(2000, 3000, 60000, seed=1)
    assert bool((tu[1:] >= tu[:-1]).all()) and int(tu.max()) == n - 1 and len(torch.unique(tu)) == n
    key = torch.cat([tu, su]) * m + torch.cat([ti, si])
    assert len(torch.unique(key)) == len(key)  # no duplicate interactions
    both_i = torch.bincount(torch.cat([ti, si]), minlength=m)
    both_u = torch.bincount(torch.cat([tu, su]), minlength=n)
    assert int(both_i.min()) >= 5 and int(both_u.min()) >= 5  # five-core
    frac = len(tu) / (len(tu) + len(su))
    assert 0.78 < frac < 0.86
    n2, m2, tu2, *_ = bipartite(2000, 3000, 60000, seed=1)
    assert (n2, m2) == (n, m) and torch.equal(tu, tu2)


def test_cap_keep_mask_is_the_sequential_counter():
    """ddp_lgcn.py:569-570 (a dict counted along the loop) == rank among earlier samples with the same positive."""
    from furusato_recommend_b200.negative_sample import cap_keep_mask
    rng = np.random.default_rng(3)
    m, count, limit = 17, 500, 4
    pos = rng.integers(0, m, count)
    valid = rng.random(count) > 0.2
    pos_t = torch.from_numpy(np.where(valid, pos, -1))
    keep = cap_keep_mask(pos_t, torch.from_numpy(valid.astype(np.uint8)), m, limit).numpy().astype(bool)
    want, oc = np.zeros(count, dtype=bool), {}
    for i in range(count):
        if valid[i] and oc.get(pos[i], 0) < limit:
            oc[pos[i]] = oc.get(pos[i], 0) + 1
            want[i] = True
    assert np.array_equal(keep, want) and keep.sum() == sum(min(limit, int((pos[valid] == v).sum())) for v in range(m))


def test_segment_cdf_matches_numpy_choice_table():
    from furusato_recommend_b200.negative_sample import _segment_cdf
    rng = np.random.default_rng(4)
    lens = np.array([3, 0, 1, 7, 0, 5])
    rowptr = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]))
    w = rng.random(lens.sum()) + 0.01
    cdf = _segment_cdf(torch.from_numpy(w), rowptr).numpy()
    o = 0
    for n in lens:
        if n:
            want = orc.normalised_cdf(w[o:o + n])       # np.random.choice: cumsum / cumsum[-1]
            assert np.abs(cdf[o:o + n] - want).max() < 1e-6 and abs(cdf[o + n - 1] - 1.0) < 1e-6
            assert np.all(np.diff(cdf[o:o + n]) > 0)
        o += n


def test_entry_index_groups_multi_edges_and_finds_reverse_entries(golden):
    n, m = int(golden["n_users"]), int(golden["m_items"])
    g = G.build_csr_graph(n, m, torch.from_numpy(golden["train_user"]), torch.from_numpy(golden["train_item"]))
    ent, n_ent, rev = g.entry_index()
    assert n_ent == len(golden["adj_val"])                      # one entry per coalesced COO value of the reference
    assert int(ent.max()) == n_ent - 1 and g.nnz > n_ent         # the tiny graph has multi-edges
    N = n + m
    deg = g.rowptr[1:] - g.rowptr[:-1]
    row = torch.repeat_interleave(torch.arange(N), deg)
    col = g.col.long()
    er = torch.zeros(n_ent, dtype=torch.int64).scatter_(0, ent, row)
    ec = torch.zeros(n_ent, dtype=torch.int64).scatter_(0, ent, col)
    assert torch.equal(er, torch.from_numpy(golden["adj_row"])) and torch.equal(ec, torch.from_numpy(golden["adj_col"]))
    assert torch.equal(er[rev], col) and torch.equal(ec[rev], row)   # rev[e] is the entry (col, row) of slot e


def test_custom_op_is_registered_with_fake_and_autograd():
    """north_star's "thin C-ABI torch custom-op layer": the op exists, has a shape rule (FakeTensor) and an
    autograd formula — checked without a GPU (no kernel is launched under FakeTensorMode)."""
    from torch._subclasses.fake_tensor import FakeTensorMode
    from furusato_recommend_b200 import torch_ops  # noqa: F401  (registers on import)
    assert hasattr(torch.ops.lgcn_b200, "propagate") and hasattr(torch.ops.lgcn_b200, "propagate_backward")
    with FakeTensorMode():
        w = torch.empty(10, 64, requires_grad=True)
        out = torch.ops.lgcn_b200.propagate(w, 7)
        assert out.shape == (10, 64) and out.requires_grad
        (g,) = torch.autograd.grad(out.sum(), w)
        assert g.shape == (10, 64)


def test_interleaved_row_order_is_a_permutation_of_the_light_rows():
    from furusato_recommend_b200.graph import decompose_rows
    torch.manual_seed(0)
    deg = torch.randint(0, 400, (5000,))
    rp = torch.zeros(5001, dtype=torch.int64)
    rp[1:] = torch.cumsum(deg, 0)
    a, b = decompose_rows(rp), decompose_rows(rp, interleave=32)
    assert sorted(a["light_rows"].tolist()) == sorted(b["light_rows"].tolist())
    assert torch.equal(a["seg_row"], b["seg_row"]) and torch.equal(a["hub_nseg"], b["hub_nseg"])
    ld = b["light_desc"]
    assert torch.equal(ld[:, 0].long(), b["light_rows"].long()) and torch.equal(ld[:, 1].long(), deg[b["light_rows"].long()])
    assert torch.equal(ld[:, 2].long(), rp[b["light_rows"].long()])
    d = deg[b["light_rows"].long()].float()
    # heavy and light chunks alternate: the two halves of the launch carry about the same number of edges
    h = len(d) // 2
    assert abs(float(d[:h].sum() - d[h:].sum())) < 0.1 * float(d.sum())
