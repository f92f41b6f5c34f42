"""Stand-alone successive-halving selection (at2_segmented_topk_f32 / _f64) against python's own
sorted(..., key=loss, reverse=...)[:keep] - the reference's selection (optimisers/sequential/hyperband_optimiser.py:46-57)
- and against the survivor indices stored in the reference-generated goldens."""
import numpy as np
import pytest
import torch

from tests.helpers import hb_case_names, load_hb_case

pytestmark = pytest.mark.gpu


def _python_topk(losses, keep, descending):
    order = sorted(range(len(losses)), key=lambda i: losses[i], reverse=descending)   # the reference's call, on positions
    return order[:keep]


def _run(losses_list, keeps, dtype, descending):
    from autotune.b200 import ops
    offs = np.zeros(len(losses_list) + 1, dtype=np.int64)
    offs[1:] = np.cumsum([len(x) for x in losses_list])
    flat = np.concatenate([np.asarray(x, dtype=np.float64) for x in losses_list]) if offs[-1] else np.zeros(0)
    t = torch.tensor(flat, dtype=dtype, device='cuda')
    out = ops.segmented_topk(t, torch.tensor(offs, device='cuda'), torch.tensor(keeps, dtype=torch.int32, device='cuda'),
                             max(1, max((len(x) for x in losses_list), default=1)), descending).cpu().numpy()
    return [out[offs[s]:offs[s + 1]] for s in range(len(losses_list))], t.cpu().numpy(), offs


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
@pytest.mark.parametrize('descending', [False, True])
def test_segmented_topk_matches_pythons_stable_sorted(dtype, descending):
    rng = np.random.default_rng(5)
    # every code path: counting rank (<= 32), register networks (<= 128, <= 256), shared-memory network (wider); ragged
    # lengths incl. 1 and exact powers of two; heavy ties (few distinct values), signed zeros, keep = 0 and keep > n
    lens = [1, 2, 3, 9, 27, 31, 32, 33, 64, 81, 100, 128, 129, 200, 243, 256, 257, 500, 729, 1024, 1500]
    segs, keeps = [], []
    for rep in range(3):
        for n in lens:
            kind = (rep + n) % 3
            if kind == 0:
                x = rng.standard_normal(n) * 100
            elif kind == 1:
                x = rng.integers(-2, 3, n).astype(np.float64)            # five distinct values: ties everywhere
                x[x == 0] = rng.choice([0.0, -0.0], size=(x == 0).sum())    # python compares -0.0 == 0.0
            else:
                x = np.round(rng.standard_normal(n), 1)
            segs.append(x)
            keeps.append(int(rng.integers(0, n + 3)) if rep else n // 3)
    got, flat, offs = _run(segs, keeps, dtype, descending)
    for s, (x, k) in enumerate(zip(segs, keeps)):
        xs = flat[offs[s]:offs[s + 1]].tolist()                        # the values as the device saw them (float32 rounding)
        want = _python_topk(xs, k, descending)
        assert got[s][:len(want)].tolist() == want, (s, len(x), k)
        assert (got[s][len(want):] == -1).all()


@pytest.mark.parametrize('name', hb_case_names())
def test_segmented_topk_reproduces_reference_survivors(name):
    """The rounds of the reference-generated golden repetitions (tests/golden/hb_*.npz: the losses the unmodified reference
    saw in each round, in evaluation order, and the evaluators it kept): the stand-alone selection, run on the float64
    losses with every round of every repetition as one segment, returns the reference's survivors, in its order."""
    from autotune.b200 import ops
    cfg, reps = load_hb_case(name)
    segs, keeps, ids, want = [], [], [], []
    for rep in reps:
        ro, so = rep['round_off'], rep['surv_off']
        for r in range(len(rep['keep'])):
            segs.append(rep['eval_fvals'][ro[r]:ro[r + 1]])
            ids.append(rep['eval_ids'][ro[r]:ro[r + 1]])
            keeps.append(int(rep['keep'][r]))
            want.append(rep['surv_ids'][so[r]:so[r + 1]])
    offs = np.zeros(len(segs) + 1, dtype=np.int64)
    offs[1:] = np.cumsum([len(x) for x in segs])
    out = ops.segmented_topk(torch.tensor(np.concatenate(segs), device='cuda'), torch.tensor(offs, device='cuda'),
                             torch.tensor(keeps, dtype=torch.int32, device='cuda'), max(len(x) for x in segs),
                             cfg['min_or_max'] == 'max').cpu().numpy()
    for s in range(len(segs)):
        got = out[offs[s]:offs[s + 1]]
        k = len(want[s])
        assert ids[s][got[:k]].tolist() == want[s].tolist(), (name, s)
        assert (got[k:] == -1).all()


def test_segmented_topk_errors_and_empty():
    from autotune.b200 import ops
    from autotune.b200._native import At2Error
    x = torch.arange(40, dtype=torch.float64, device='cuda')
    offs = torch.tensor([0, 40], device='cuda')
    keep = torch.tensor([3], dtype=torch.int32, device='cuda')
    with pytest.raises(At2Error, match='longer than max_seg_len'):
        ops.segmented_topk(x, offs, keep, 32, False)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.segmented_topk(x.cpu(), offs, keep, 40, False)
    out = ops.segmented_topk(x[:0], torch.tensor([0], device='cuda'), keep[:0], 1, False)
    assert out.numel() == 0
    # empty segments between full ones
    out = ops.segmented_topk(x, torch.tensor([0, 0, 40, 40], device='cuda'),
                             torch.tensor([1, 2, 1], dtype=torch.int32, device='cuda'), 40, True)
    assert out[:2].tolist() == [39, 38] and (out[2:] == -1).all()
