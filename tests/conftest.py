import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'autotune-v2_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_addoption(parser):
    parser.addoption('--runslow', action='store_true', default=False,
                     help='also run the tests marked slow (minutes of CPU work each; AT2_RUN_SLOW=1 does the same)')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'slow: minutes of CPU work; skipped unless --runslow / AT2_RUN_SLOW=1')


def pytest_collection_modifyitems(config, items):
    if config.getoption('--runslow') or os.environ.get('AT2_RUN_SLOW') == '1':
        return
    skip = pytest.mark.skip(reason='slow test: run with --runslow or AT2_RUN_SLOW=1')
    for item in items:
        if 'slow' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
