"""Shared checks of the GPU parity tests."""
import torch


def gap_rule(idx: torch.Tensor, ref_scores: torch.Tensor, k: int, tol: float):
    """SURVEY §9.5 / north_star, for a selection made on scores that may differ from `ref_scores` by
    up to `tol` per entry:
      (a) every selected item scores, in the reference, within 2*tol of the reference's k-th value;
      (b) wherever a reference rank is separated from BOTH neighbours by more than 2*tol the ids
          must agree with the reference's stable sort (ties -> lowest id).
    `ref_scores` is the dense [U, m] reference score block with the mask already applied.
    Returns (rows violating (a), positions compared under (b), mismatches under (b))."""
    idx = idx.long().to(ref_scores.device)
    ov, oi = torch.sort(ref_scores, dim=1, descending=True, stable=True)
    ov, oi = ov[:, :k + 1], oi[:, :k + 1]
    if ov.shape[1] == k:   # k == m: no rank k+1
        ov = torch.cat([ov, torch.full_like(ov[:, :1], -float("inf"))], dim=1)
    got = torch.gather(ref_scores, 1, idx)
    bad_rows = int((got < (ov[:, k - 1:k] - 2 * tol)).any(dim=1).sum())
    gap_dn = ov[:, :k] - ov[:, 1:k + 1]
    gap_up = torch.cat([torch.full_like(gap_dn[:, :1], float("inf")), gap_dn[:, :-1]], dim=1)
    decided = (gap_dn > 2 * tol) & (gap_up > 2 * tol)
    mism = decided & (idx != oi[:, :k])
    return bad_rows, int(decided.sum()), int(mism.sum())
