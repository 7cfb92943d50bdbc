import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def table_dir():
    from relxill_b200.tables import synth
    return synth.generate(synth.default_table_dir("test"), "test")


@pytest.fixture(scope="session")
def oracle(table_dir):
    from oracle.pyoracle import Oracle
    return Oracle(table_dir)


@pytest.fixture(scope="session")
def ref(table_dir):
    """The unmodified reference (oracle/_ref), if it was built (needs /root/reference at build time)."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref/librelxill_ref.so not built")
    os.environ.pop("RELXILL_NUM_RZONES", None)
    return pyref.Ref(table_dir)


@pytest.fixture(scope="session")
def rx(table_dir):
    import relxill_b200 as rx
    rx.init(table_dir)
    return rx
