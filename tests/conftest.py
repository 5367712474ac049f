import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


_SD_CACHE = {}


@pytest.fixture(scope="session")
def synthetic_state_dict():
    """seed, sharp -> encoder-path state_dict (cached; ~0.85 GB fp32 each, keep at most 2)."""
    from oracle import weights as W

    def get(seed, sharp=1.0, outlier=False, decoder_layers=0):
        key = (seed, sharp, outlier, decoder_layers)
        if key not in _SD_CACHE:
            if len(_SD_CACHE) >= 2:
                _SD_CACHE.pop(next(iter(_SD_CACHE)))
            _SD_CACHE[key] = W.make_state_dict(seed, sharp, outlier, decoder_layers=decoder_layers)
        return _SD_CACHE[key]
    return get
