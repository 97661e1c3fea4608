"""Checker of Model.detect_records against Model.forward's [B, A, no] tensor (used by the GPU tests and by bench.py's gate).

The record's box / obj / class columns must be bit-identical to the tensor's.  The theta index is the first maximum of the 180
LOGITS; the tensor holds tanh.approx sigmoids (relative error ~2^-11) of those logits, so where two logits are closer than
that error the tensor's argmax may name the other one.  Such a row is accepted only if the tensor's value at the record's index
is within 2^-9 (relative) of the row maximum - anything else is an error.  `patched` is the tensor with those rows' maxima
moved to the record's index, so that non_max_suppression_obb(patched) must equal non_max_suppression_obb(records) EXACTLY."""
import torch

THETA_REL_TOL = 2.0 ** -9


def check_records(pred: torch.Tensor, rec_data: torch.Tensor, nc: int):
    """pred [B, A, 5+nc+180] (activated), rec_data [B, A, >= 6+nc].  Returns (patched_pred, n_near_tie_rows); raises
    AssertionError on any disagreement beyond the near-tie rule."""
    nfix = 5 + nc
    assert torch.equal(rec_data[..., :nfix], pred[..., :nfix]), "record box/obj/class columns differ from the tensor"
    theta = pred[..., nfix:]
    idx_t = theta.argmax(-1)
    idx_r = rec_data[..., nfix].long()
    assert int(idx_r.min()) >= 0 and int(idx_r.max()) < theta.shape[-1]
    diff = idx_t != idx_r
    n = int(diff.sum())
    patched = pred
    if n:
        top = theta.max(-1).values
        at_rec = theta.gather(-1, idx_r.unsqueeze(-1)).squeeze(-1)
        bad = diff & (at_rec < top * (1.0 - THETA_REL_TOL))
        assert not bool(bad.any()), f"{int(bad.sum())} rows: record theta index is not a (near-)maximum of the tensor's bins"
        patched = pred.clone()
        th = patched[..., nfix:]
        bump = torch.where(diff, top * (1.0 + 2.0 ** -20) + 1e-30, at_rec)
        th.scatter_(-1, idx_r.unsqueeze(-1), bump.unsqueeze(-1))
        assert torch.equal(th.argmax(-1), idx_r)
    return patched, n
