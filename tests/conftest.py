import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")
    # native pieces are built artefacts; build whatever is missing (nvcc cross-compiles without a GPU)
    for d, so in (("oracle", "libfe_oracle.so"), ("feature_extraction_b200/synth", "libfe_synth.so"),
                  ("feature_extraction_b200/csrc", "libfe_b200.so")):
        if not os.path.exists(os.path.join(ROOT, d, so)):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, d)])


@pytest.fixture(scope="session")
def ob():
    from oracle import oracle_binding
    return oracle_binding


@pytest.fixture(scope="session")
def synth():
    from feature_extraction_b200 import synth as s
    return s
