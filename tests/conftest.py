import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _built(*paths):
    return all(os.path.exists(os.path.join(ROOT, p)) for p in paths)


@pytest.fixture(scope="session")
def libs():
    """Product + oracle libraries; built on demand where a toolchain exists (the GPU box uses the prebuilt files)."""
    need = ["rnaseqc_b200/lib/librsq_b200.so", "rnaseqc_b200/lib/librsq_host.so", "rnaseqc_b200/bin/rnaseqc",
            "oracle/_build/librsq_oracle.so", "oracle/_build/rnaseqc_oracle"]
    if not _built(*need):
        subprocess.run(["make", "-C", ROOT, "all", "oracle"], check=True, stdout=subprocess.DEVNULL)
    from rnaseqc_b200 import api
    return {
        "product": os.path.join(ROOT, need[0]),
        "host": os.path.join(ROOT, need[1]),
        "cli": os.path.join(ROOT, need[2]),
        "oracle": os.path.join(ROOT, need[3]),
        "oracle_cli": os.path.join(ROOT, need[4]),
        "ref_cli": os.path.join(ROOT, "oracle/_ref/rnaseqc"),
        "api": api,
    }


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
