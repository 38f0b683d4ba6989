import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KEY_SEED = 42  # every test and fixture uses the key set generated from this seed


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def keys(oracle):
    """TFHE default 128-bit gate parameters (n=630, N=1024), keys from KEY_SEED."""
    K = oracle.Keys.generate(oracle.default_params(128), KEY_SEED)
    yield K
    K.close()


@pytest.fixture(scope="session")
def engine(keys):
    """GPU engine with the oracle's key material loaded (gpu tests only)."""
    import peba1_b200 as P
    e = P.Engine(P.default_params(128), device=0)
    e.load_keys(keys.bk, keys.ksk)
    yield e
    e.close()


def random_bits(seed, count):
    return np.random.default_rng(seed).integers(0, 2, size=count)
