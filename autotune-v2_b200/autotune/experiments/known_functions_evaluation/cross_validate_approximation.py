"""k-fold cross-validation of the closest-known-loss-function approximation (reference
experiments/known_functions_evaluation/cross_validate_approximation.py): every recorded arm is looked up in the database
minus its own fold and the error between its real curve and the approximating curve is reported per epoch, normalised
by the range of the recorded losses at that epoch.  All folds run as ONE batched nearest-arm launch with a per-query
held-out range (torch.ops.at2.nn_lookup_excluding) instead of rebuilding the problem per fold."""
import numpy as np
import torch


def get_normalization_factors(min_loss_fn_len, all_known_fns):
    cols = np.asarray([list(fn[:min_loss_fn_len]) for fn in all_known_fns.values()], dtype=np.float64)
    return cols.max(axis=0) - cols.min(axis=0)


def get_errors(expected_loss_fn, actual_loss_fn, min_loss_fn_len):
    return np.abs(np.asarray(actual_loss_fn[:min_loss_fn_len], dtype=np.float64) -
                  np.asarray(expected_loss_fn[:min_loss_fn_len], dtype=np.float64))


def cross_validate(n_folds, all_known_fns, real_problem):
    """-> (res [n_arms, min_len] normalised errors, min_len), rows in the order of all_known_fns."""
    from autotune.b200.engine import KnownFnSpec
    arms = list(all_known_fns.keys())
    n = len(arms)
    fold_size = n // n_folds
    min_len = min(len(all_known_fns[a]) for a in arms)
    norm = get_normalization_factors(min_len, all_known_fns)
    db = KnownFnSpec(all_known_fns, real_problem.domain, real_problem.hyperparams_to_opt)
    starts = (np.arange(n) // fold_size) * fold_size
    exclude = torch.as_tensor(np.stack([starts, np.minimum(starts + fold_size, n)], axis=1).astype(np.int32)).to(db.device)
    idx, _ = torch.ops.at2.nn_lookup_excluding(db.domain_bytes, db.db, db.db, exclude.contiguous())
    curves = db.curves[:, :min_len]
    res = (curves[idx.long()] - curves).abs() / torch.as_tensor(norm, device=db.device)
    return res.cpu().numpy(), min_len
