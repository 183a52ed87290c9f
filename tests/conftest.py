import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One spring_b200 context on cuda:0 for the whole GPU session."""
    from spring_b200 import capi
    c = capi.Context(0)
    c.set_stitch(0)  # contig stitching off: the parity tests compare the encoder with the oracle's on the same reorder stream
    yield c
    c.close()
