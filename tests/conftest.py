import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ncc_golden.npz"))


@pytest.fixture(scope="session")
def pkg():
    import font_ocr_b200

    return font_ocr_b200


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.port_lib()
    return O


@pytest.fixture(scope="session")
def font(pkg):
    return pkg.raster.Font()


@pytest.fixture(scope="session")
def built_lib():
    """The product library, built in-tree (nvcc cross-compiles without a GPU)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("focr_build", os.path.join(ROOT, "font-ocr_b200", "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.build()


@pytest.fixture(scope="session")
def ctx(pkg, built_lib):
    from font_ocr_b200 import ncc

    c = ncc.Context(0)
    yield c
    c.close()
