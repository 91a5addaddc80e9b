import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")
REFERENCE = os.environ.get("VQA_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_meta():
    import json
    with open(os.path.join(GOLDEN, "golden_meta.json")) as f:
        return json.load(f)


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "models"))


@pytest.fixture(scope="session")
def reference_modules(tmp_path_factory):
    """Import the live reference (build container only); skipped where it is absent."""
    if not have_reference():
        pytest.skip("reference sources not present on this machine")
    os.chdir(tmp_path_factory.mktemp("refcwd"))  # utils.config mkdirs relative to CWD
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import importlib
    return importlib
