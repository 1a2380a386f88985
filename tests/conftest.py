import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# RSQ_EMULATE=1: replay the -m gpu suite WITHOUT a GPU — the product library and CLI are replaced by their CPU emulation
# (oracle/_build/librsq_emu.so and rnaseqc_emu: the CUDA sources compiled for the host through tests/cudaemu/cuda_emu.h).  The
# emulation runs a few thousand records per second, so the record counts of the synthetic fixtures are scaled down (RSQ_TEST_SCALE,
# default 0.05 under emulation) and the tests that need device pointers are skipped.
EMULATE = os.environ.get("RSQ_EMULATE", "0") == "1"
if EMULATE:
    os.environ["RSQ_B200_LIB"] = os.path.join(ROOT, "oracle", "_build", "librsq_emu.so")     # api.PRODUCT_LIB / default_backend()
SCALE = float(os.environ.get("RSQ_TEST_SCALE", "0.05" if EMULATE else "1"))


def scaled(n, floor=2000):
    """record count of a synthetic fixture under RSQ_TEST_SCALE"""
    return max(floor, int(n * SCALE)) if SCALE != 1 else n


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
    if EMULATE and not _built("oracle/_build/librsq_emu.so", "oracle/_build/rnaseqc_emu"):
        subprocess.run(["make", "-C", ROOT, "oracle"], check=True, stdout=subprocess.DEVNULL)
    return {
        "product": os.path.join(ROOT, "oracle/_build/librsq_emu.so" if EMULATE else need[0]),
        "host": os.path.join(ROOT, need[1]),
        "cli": os.path.join(ROOT, "oracle/_build/rnaseqc_emu" if EMULATE else need[2]),
        "oracle": os.path.join(ROOT, need[3]),
        # RSQ_EMULATE_CLI=1: the CLI-level reference tests (goldens, live-reference fuzz) run the PRODUCT sources in emulation instead of
        # the oracle harness: the strongest check of the device code that is possible without a device
        "oracle_cli": os.path.join(ROOT, "oracle/_build/rnaseqc_emu" if os.environ.get("RSQ_EMULATE_CLI") == "1" else need[4]),
        "ref_cli": os.path.join(ROOT, "oracle/_ref/rnaseqc"),
        "api": api,
    }


def needs_device_pointers():
    """skip marker for tests that hand DEVICE pointers to the library (meaningless under the CPU emulation)"""
    return pytest.mark.skipif(EMULATE, reason="device pointers: not under RSQ_EMULATE")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
