"""Drop-in proof: the reference's UNMODIFIED `trainer.Trainer` (trainer.py:27-280) drives this
repo's `Loader`, `LightGCN` and `UniformSample` exactly as INTEGRATION.md §1 shows — swap three
names, change nothing else — through `train()` (trainer.py:56-81: sampler -> torch.tensor ->
utils.shuffle -> model.OneEpoch) and `test()` (trainer.py:115-187: getUserPosItems ->
getUsersRating -> exclude-list mask -> torch.topk -> metrics -> save_model / save_result).

Needs a copy of the reference (never committed): /root/reference in the dev container, or
baseline/_ref on a GPU box (`python tools/stage_reference.py` puts the few .py files there before a
`gpurun` call).  Skips cleanly otherwise.  Shims are the ones of SURVEY §8c: argv before `import
world`, the data fixtures `Trainer` np.loads from fixed relative paths, WANDB_MODE=disabled, and the
hard-coded checkpoint directory (trainer.py:222)."""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = Path(__file__).resolve().parents[1]
REF = next((p for p in (REPO / "baseline" / "_ref", Path("/root/reference")) if (p / "trainer.py").exists()), None)
DEV = "cuda:0"


@pytest.mark.skipif(REF is None, reason="no copy of the reference (baseline/_ref or /root/reference)")
def test_reference_trainer_drives_our_classes_unchanged(monkeypatch):
    from oracle.make_golden import make_tiny, write_files, N_USERS, M_ITEMS
    import furusato_recommend_b200 as ours

    train, test = make_tiny()
    tmp = Path(tempfile.mkdtemp(prefix="lgcn_boundary_"))
    write_files(tmp, train, test)
    monkeypatch.chdir(tmp)
    monkeypatch.setenv("WANDB_MODE", "disabled")
    monkeypatch.setattr(sys, "argv", ["main.py", "--model", "lgn", "--recdim", "32", "--layer", "3", "--suffix", "t",
                                      "--bpr_batch", "64", "--lr", "1e-3", "--decay", "1e-4", "--testbatch", "128",
                                      "--topks", "[10, 20]", "--wandb", "boundary", "--epochs", "1"])
    monkeypatch.syspath_prepend(str(REF))
    for name in ("world", "parse", "trainer", "utils", "metric", "negative_sample", "dataloader"):
        sys.modules.pop(name, None)
    import world                      # the reference's (parses argv at import, world.py:14)
    world.device = DEV
    world.config["device"] = DEV
    import trainer as ref_trainer     # unmodified reference module
    assert Path(ref_trainer.__file__).resolve().parent == REF.resolve()

    # ---- swap the three names ----
    ref_trainer.UniformSample = ours.UniformSample
    monkeypatch.setattr(ref_trainer.Trainer, "checkpoint_save_path", staticmethod(lambda config: str(tmp / "ckpt.pth")))
    ds = ours.Loader(world.config, path=str(tmp / "data" / "cf"), device=DEV)
    assert (ds.n_users, ds.m_items) == (N_USERS, M_ITEMS)
    torch.manual_seed(2020)
    model = ours.LightGCN(world.config, ds)

    tr = ref_trainer.Trainer(world.config, ds, model)         # Metric, UniformSampling, model.to(device), print(model)
    res0 = tr.test()                                          # dense getUsersRating + the reference's own mask/top-k/metrics
    # our fused evaluation of the same weights must agree with what the reference computed from our scores
    mine = ours.Trainer(world.config, ds, model, topks=world.topks)
    model.eval_precision = "fp32"
    own = mine.test()
    for name in ("recall", "precision", "ndcg", "hr"):
        assert np.allclose(res0[name], own[name], rtol=0, atol=2e-3), (name, res0[name], own[name])
    assert (tmp / "ckpt.pth").exists()                        # recall improved over 0 -> save_model(state_dict)
    sd = torch.load(tmp / "ckpt.pth")
    assert list(sd.keys()) == ["all_embedding.weight"] and sd["all_embedding.weight"].shape == (N_USERS + M_ITEMS, 32)
    assert (tmp / "data" / "result" / "lgn" / "lgn_32_3_boundary.csv").exists()   # save_result ran

    w0 = model.all_embedding.weight.detach().clone()
    np.random.seed(0)
    loss = tr.train()                                         # reference loop: our sampler, its shuffle, our OneEpoch
    assert torch.is_tensor(loss) and torch.isfinite(loss).item() and 0.0 < float(loss) < 1.0
    assert not torch.equal(w0, model.all_embedding.weight.detach())
    res1 = tr.test()                                          # eval-mode cache must notice the training step
    assert any(not np.array_equal(res0[k], res1[k]) for k in ("recall", "ndcg", "precision"))
    own1 = mine.test()
    for name in ("recall", "precision", "ndcg", "hr"):
        assert np.allclose(res1[name], own1[name], rtol=0, atol=2e-3), (name, res1[name], own1[name])
