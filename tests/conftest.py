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


def _cuda_device_present() -> bool:
    """False only when the engine itself says there is no device (-67); a missing or broken libszb200.so still fails loudly."""
    from sparkzstd_b200.decompression import Context, SzbError

    try:
        Context(0).close()
        return True
    except SzbError as e:
        if getattr(e, "code", None) == -67:
            return False
        raise


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped, not failed, on a box without a CUDA device (szb_ctx_create: SZB_ERR_NO_DEVICE)."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the engine has no CPU fallback")
    for it in gpu_items:
        it.add_marker(skip)


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
