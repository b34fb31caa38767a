"""Round-1 advisor findings, as tests: ids outside the table (the reference's IndexError, model/lgcn.py:90-95)
must not scatter out of bounds; a user whose positives cover every item (the reference's `while True`,
negative_sample.py:121-126, never returns) must not wedge the device; the eval-mode propagation cache must
notice weight edits; ops follow the tensors' device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from furusato_recommend_b200 import LightGCN, UniformSample, ops  # noqa: E402
from furusato_recommend_b200.dataloader import BasicDataset  # noqa: E402

DEV = "cuda:0"


def _tiny(n=40, m=30, seed=0, **cfg):
    rng = np.random.default_rng(seed)
    tu, ti = [], []
    for u in range(n):
        its = rng.choice(m, size=int(rng.integers(2, 8)), replace=False)
        tu += [u] * len(its); ti += its.tolist()
    c = dict(recdim=32, layer=2, lr=1e-3, decay=1e-4, bpr_batch_size=16, device=DEV, test_u_batch_size=64, **cfg)
    ds = BasicDataset(n, m, np.array(tu), np.array(ti), np.array(tu[:10]), np.array(ti[:10]), config=c, device=DEV)
    return c, ds


def test_out_of_range_ids_raise_like_the_reference():
    cfg, ds = _tiny()
    model = LightGCN(cfg, ds).train()
    u = torch.arange(16, device=DEV); p = torch.zeros(16, dtype=torch.int64, device=DEV)
    q = torch.ones(16, dtype=torch.int64, device=DEV)
    assert torch.isfinite(model.stageOne(u, p, q))
    model.check_ids()                                           # nothing to report
    w0 = model.all_embedding.weight.detach().clone()
    bad = q.clone(); bad[3] = ds.m_items                        # one past the last item
    loss = model.stageOne(u, p, bad)                            # no fault, no out-of-bounds scatter
    assert torch.isnan(loss)
    with pytest.raises(IndexError):
        model.check_ids()
    model.check_ids()                                           # the flag is cleared once reported
    assert torch.isfinite(model.all_embedding.weight).all() and not torch.equal(w0, model.all_embedding.weight)
    with pytest.raises(IndexError):                             # eager autograd path: raised at the call
        model.bpr_loss(u, p, bad)
    neg_user = u.clone(); neg_user[0] = -1
    with pytest.raises(IndexError):                             # epoch loop: raised at the end of the epoch
        model.OneEpoch(neg_user, p, q)
    strict = LightGCN(dict(cfg, check_ids=True), ds).train()
    with pytest.raises(IndexError):                             # host ids can be validated before any launch
        strict.stageOne(u.cpu(), p.cpu(), bad.cpu())


def test_sampler_gives_up_on_a_user_without_negatives():
    n, m = 6, 5
    tu = np.array([0] * m + [1, 1, 2, 3, 4, 5]); ti = np.array(list(range(m)) + [0, 1, 2, 3, 4, 0])   # user 0 owns every item
    ds = BasicDataset(n, m, tu, ti, tu[:2], ti[:2], config={}, device=DEV)
    S = UniformSample(ds, seed=1, epoch=0, count=4000)         # returns (the reference's loop would not)
    torch.cuda.synchronize()
    S = S.cpu().numpy()
    assert 0 < len(S) < 4000 and not (S[:, 0] == 0).any()       # user 0's samples are dropped like an empty user's
    pos = {u: set(ti[tu == u].tolist()) for u in range(n)}
    assert all(int(b) in pos[int(a)] and int(c) not in pos[int(a)] for a, b, c in S)


def test_eval_cache_follows_weight_edits():
    cfg, ds = _tiny()
    model = LightGCN(cfg, ds).eval()
    u0, _ = model.computer()
    u0 = u0.clone()
    with torch.no_grad():
        model.all_embedding.weight.mul_(2.0)                    # in-place edit in eval mode (ADVICE: stale cache)
    u1, _ = model.computer()
    assert torch.allclose(u1, 2.0 * u0, rtol=1e-5, atol=1e-7)
    users = torch.arange(8, device=DEV)
    r0 = model.getUsersRating(users).clone()
    model.train()
    model.stageOne(users, torch.zeros(8, dtype=torch.int64, device=DEV), torch.ones(8, dtype=torch.int64, device=DEV))
    model.eval()
    assert not torch.equal(r0, model.getUsersRating(users))


def test_ops_reject_mixed_devices_and_follow_the_tensor_device():
    x = torch.zeros(4, 32, device=DEV)
    with pytest.raises(Exception):
        ops.score_dense_f32(x, x.cpu(), torch.zeros(1, dtype=torch.int64, device=DEV))
    if torch.cuda.device_count() > 1:                           # launches go to the tensors' device, not the current one
        cfg, ds = _tiny()
        cfg1 = dict(cfg, device="cuda:1")
        ds1 = BasicDataset(ds.n_users, ds.m_items, ds.trainUser, ds.trainItem, ds.testUser, ds.testItem, config=cfg1, device="cuda:1")
        torch.manual_seed(0); a = LightGCN(cfg, ds).eval()
        b = LightGCN(cfg1, ds1).eval()
        with torch.no_grad():
            b.all_embedding.weight.copy_(a.all_embedding.weight)
        assert torch.cuda.current_device() == 0
        assert torch.allclose(a.computer()[0].cpu(), b.computer()[0].cpu(), rtol=1e-6, atol=1e-8)
