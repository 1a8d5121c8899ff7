import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    import bulletproofspp_b200 as bp
    c = bp.Context(0)        # raises loudly without a GPU / without the built library
    yield c
    c.close()


@pytest.fixture(scope="session")
def gens():
    """generators from basisSeed = "test points" (app/Main.hs:68-72), cached for the session"""
    from oracle.curve import Secp256k1 as G
    from oracle.transcript import get_points
    cache = {}

    def get(n):
        if len(cache.get("p", [])) < n:
            cache["p"] = get_points(G, "test points", n)
        return cache["p"][:n]
    return get
