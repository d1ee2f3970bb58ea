import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    import golden_data
    return golden_data.Golden()


@pytest.fixture(scope="session")
def golden128():
    """frozen dict_freq_index files built by the reference with constants::block_size = 128
    (BASELINE.json configs[1]; oracle/make_fixtures.py --bs 128)"""
    import golden_data
    return golden_data.Golden("bs128")
