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


def test_reference_import_lines_resolve():
    """The import lines of the reference's drivers (experiments/run_simulation.py:8-17,
    experiments/run_closest_loss_fn_approximation.py:13-23) resolve against the mirror package.  Out-of-scope names exist as
    domain-only problems (MNIST/CIFAR/SVHN/MRBI) or placeholders that refuse construction (SigOpt)."""
    import io
    from contextlib import redirect_stdout

    import autotune.benchmarks as b
    import autotune.optimisers as o
    for name in ('CifarProblem', 'MnistProblem', 'MrbiProblem', 'SvhnProblem', 'OptFunctionSimulationProblem',
                 'OptFunctionProblem', 'AVAILABLE_OPT_FUNCTIONS', 'KnownFnProblem'):          # benchmarks/__init__.py:9-12
        assert hasattr(b, name) and name in b.__all__
    for name in ('HybridHyperbandTpeOptimiser', 'HybridHyperbandSigoptOptimiser', 'HyperbandOptimiser', 'RandomOptimiser',
                 'SigOptimiser', 'TpeOptimiser', 'HybridHyperbandTpeTransferAllOptimiser',
                 'HybridHyperbandTpeTransferLongestOptimiser', 'HybridHyperbandTpeTransferThresholdOptimiser',
                 'HybridHyperbandTpeNoTransferOptimiser', 'HybridHyperbandTpeTransferSameOptimiser'):  # optimisers/__init__.py:13-19
        assert hasattr(o, name) and name in o.__all__
    with redirect_stdout(io.StringIO()):
        mnist, cifar = b.MnistProblem('in', 'out'), b.CifarProblem('in', 'out', dataset_loader=None)
    assert tuple(mnist.domain.hyperparams_names()) == ('learning_rate', 'weight_decay', 'momentum', 'batch_size')
    assert cifar.hyperparams_to_opt == ('learning_rate', 'n_units_1', 'n_units_2', 'n_units_3', 'batch_size')
    assert len(tuple(cifar.domain.hyperparams_names())) == 9
    with pytest.raises(NotImplementedError):
        mnist.get_evaluator()
    with pytest.raises(NotImplementedError):
        o.SigOptimiser(n_resources=3, max_iter=9)


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use it."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'autotune-v2_b200')
    pat = re.compile(r'^\s*(from\s+oracle\b|import\s+oracle\b)', re.M)
    offenders = []
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                path = os.path.join(dirpath, f)
                if pat.search(open(path, encoding='utf-8', errors='ignore').read()):
                    offenders.append(path)
    assert offenders == []
