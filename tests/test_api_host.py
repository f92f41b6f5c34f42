"""CPU-only checks of the host-side mirror of the reference API (no GPU compute)."""
import numpy as np
import pytest

from tests.helpers import hb_case_names, load_hb_case


@pytest.mark.parametrize('name', hb_case_names())
def test_host_start_and_target_bit_exact_vs_reference(name):
    """f_1 (first trajectory point) and f_n computed by the product's host code equal the reference's float64 values."""
    from autotune.b200.engine import HyperbandSimSpec, host_start_and_target
    cfg, reps = load_hb_case(name)
    spec = HyperbandSimSpec(cfg['func'], cfg['families'], cfg['R'], cfg['eta'], cfg['noise'], 'uniform',
                            cfg['min_or_max'])
    for rep in reps:
        f1, fn = host_start_and_target(spec, rep['xy'], rep['fam'], rep['z'])
        assert np.array_equal(f1, rep['traj_raw'][:, 0])
        # the last raw point is f_n up to the final step's rounding (asserted < 1e-4 by the reference itself)
        assert np.max(np.abs(rep['traj_raw'][:, -1] - fn)) < 1e-4
