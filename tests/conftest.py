import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def corpus(manifest):
    """[(name, compressed bytes, original_size, original_sha256)] for the 100 decodecorpus files."""
    out = []
    for e in manifest["files"]:
        with open(os.path.join(GOLDEN, "decodecorpus", e["name"]), "rb") as f:
            out.append((e["name"], f.read(), e["original_size"], e["original_sha256"]))
    return out
