"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0, checked against the oracle."""
import os
import sys

import numpy as np
import torch


def run():
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import hyperband as ohb          # the checker (test infrastructure), never the thing shipped
    from oracle.kde import epdf

    from .engine import HyperbandSimSpec, host_start_and_target, run_hyperband_predrawn, run_hyperband_repetitions
    from .epdf import density

    if not torch.cuda.is_available():
        raise RuntimeError('smoke() needs cuda:0 (no CPU fallback)')
    torch.cuda.set_device(0)
    fams = [(1.5, 10, 15, False, 0, 200), (0.5, 7, 10, False, 0, 200), (0.2, 4, 7, True, 0, 200)]
    func, R, eta, noise = 'rastrigin', 27, 3, 10
    spec = HyperbandSimSpec(func, fams, R, eta, noise)
    A = spec.n_arms
    rng = np.random.default_rng(0)
    xy = rng.uniform(-5.12, 5.12, (2, A, 2))
    fam = rng.integers(0, len(fams), (2, A)).astype(np.int32)
    z = rng.standard_normal((2, A))
    t = np.arange(1, R)
    beta = (2 + np.sqrt(4 + 4 * (R + 1 - t))) / (2 * (R + 1 - t))
    gamma = rng.gamma(2 * beta + 1, 1 / beta, size=(2, A, R - 1))
    f1, fn = host_start_and_target(spec, xy, fam, z)
    out = run_hyperband_predrawn(spec, f1, fn, fam, gamma, xy)
    for i in range(2):
        want = ohb.run_predrawn(func, fams, R, eta, noise, xy[i], fam[i], z[i], gamma[i])
        assert np.array_equal(out['survivors'][i].cpu().numpy(), want['surv_ids']), 'survivor sets differ from oracle'
        assert int(out['ofe_arm'][i]) == int(want['ofe_id'])
        assert abs(float(out['ofe'][i]) - float(want['ofe'])) <= 1e-9 * max(1.0, abs(float(want['ofe'])))
    # production path: a few thousand Philox repetitions + the EPDF
    ofe = run_hyperband_repetitions(spec, 4096, seed=1)['ofe']
    assert torch.isfinite(ofe).all()
    dens = density(ofe, float(ofe.min()) - 5, float(ofe.max()) + 5, 3.0, 256)
    want = epdf(ofe.cpu().numpy().astype(np.float64), float(ofe.min()) - 5, float(ofe.max()) + 5, 3.0, 256)
    assert np.max(np.abs(dens.cpu().numpy() - want)) < 1e-5 * want.max()
    torch.cuda.synchronize()
    print(f'smoke ok: predrawn parity on 2 repetitions (R={R}), 4096 Philox repetitions, OFE mean {ofe.mean().item():.3f}')
