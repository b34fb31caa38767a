"""Device-side ingest (csrc/ingest.cu) against the host parser that restates the reference's loop
(dataloader.py:93-124): bit-exact (user, item) arrays in file order, duplicates kept."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from furusato_recommend_b200 import Loader  # noqa: E402
from furusato_recommend_b200.dataloader import BasicDataset, write_reference_files  # noqa: E402
from furusato_recommend_b200.synthetic import bipartite  # noqa: E402

DEV = "cuda:0"


def _same(fname):
    hu, hi = Loader._parse(str(fname), False)
    du, di = Loader._parse_device(str(fname), torch.device(DEV))
    assert du.dtype == np.int64 and di.dtype == np.int64
    assert np.array_equal(hu, du) and np.array_equal(hi, di), (len(hu), len(du))
    return hu, hi


def test_ingest_edge_cases(tmp_path):
    cases = {
        "plain.txt": "0 5 7 9\n1 3\n2 10 11 12 13\n",
        "no_final_newline.txt": "0 5 7 9\n1 3\n2 10 11",
        "blank_lines.txt": "\n0 5 7\n\n\n1 3 3 3\n\n",                    # duplicates kept; blank lines skipped
        "uid_only_line.txt": "0 1 2\n1\n2 4\n",                           # a uid without items emits nothing
        "big_ids.txt": "123456 9876543 12\n123457 1\n",
        "crlf.txt": "0 1 2\r\n1 3\r\n",
        "single.txt": "7 8",
    }
    for name, text in cases.items():
        f = tmp_path / name
        f.write_text(text)
        _same(f)
    u, i = _same(tmp_path / "blank_lines.txt")
    assert u.tolist() == [0, 0, 1, 1, 1] and i.tolist() == [5, 7, 3, 3, 3]
    (tmp_path / "empty.txt").write_text("")
    du, di = Loader._parse_device(str(tmp_path / "empty.txt"), torch.device(DEV))
    assert len(du) == 0 and len(di) == 0
    (tmp_path / "bad.txt").write_text("0 1 -2\n")
    with pytest.raises(ValueError):
        Loader._parse_device(str(tmp_path / "bad.txt"), torch.device(DEV))


def test_ingest_long_lines_cross_tiles(tmp_path):
    """Lines far longer than the 4 KiB tile and the 16-byte per-thread chunk, ids of every width."""
    rng = np.random.default_rng(0)
    lines = []
    for u in range(300):
        n = int(rng.choice([1, 2, 7, 300, 2500]))
        its = rng.integers(0, 10 ** int(rng.integers(1, 8)), n)
        lines.append(" ".join([str(u)] + [str(int(x)) for x in its]))
    f = tmp_path / "long.txt"
    f.write_text("\n".join(lines) + "\n")
    u, i = _same(f)
    assert len(u) == sum(len(l.split(" ")) - 1 for l in lines)


def test_loader_device_ingest_matches_host(tmp_path):
    n, m, tu, ti, su, si = bipartite(3000, 2000, 80000, seed=3)
    ds = BasicDataset(n, m, tu.numpy(), ti.numpy(), su.numpy(), si.numpy(), config={}, device=DEV)
    write_reference_files(ds, str(tmp_path), suffix="s")
    cfg = dict(suffix="s", device=DEV)
    a = Loader(cfg, path=str(tmp_path), device=DEV, ingest="host")
    b = Loader(cfg, path=str(tmp_path), device=DEV)           # auto -> device
    assert a.ingest == "host" and b.ingest == "device"
    for name in ("trainUser", "trainItem", "testUser", "testItem"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert (a.n_users, a.m_items, a.trainDataSize) == (b.n_users, b.m_items, b.trainDataSize) == (n, m, len(tu))
