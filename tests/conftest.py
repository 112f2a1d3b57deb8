import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def oracle_build():
    """Compile the C oracle (test infrastructure) once per session."""
    subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle'), 'all'], stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, 'oracle', '_build')
