"""CPU tests: both libraries load and export every symbol include/rsq_b200.h declares; the product refuses to run
without a GPU instead of falling back."""
import ctypes
import os
import re

import pytest

from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rsq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsq_[a-z_]+)\s*\(", text)))


def test_header_symbols_known(libs):
    assert declared_symbols() == sorted(libs["api"].ABI_SYMBOLS)


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_library_exports_every_declared_symbol(libs, which):
    lib = ctypes.CDLL(libs[which])
    for sym in declared_symbols():
        assert hasattr(lib, sym), f"{libs[which]} does not export {sym}"
    lib.rsq_backend.restype = ctypes.c_char_p
    lib.rsq_abi_version.restype = ctypes.c_uint32
    assert lib.rsq_abi_version() == 2
    assert lib.rsq_backend().decode() == {"product": "cuda-sm_100a", "oracle": "oracle-cpu"}[which]


def test_struct_sizes(libs):
    api = libs["api"]
    assert api.READ_DTYPE.itemsize == 48 and api.BLOCK_DTYPE.itemsize == 8
    assert ctypes.sizeof(api.RsqConfig) == 80


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_product_fails_loudly_without_gpu(libs):
    api = libs["api"]
    with pytest.raises(api.RsqError) as e:
        api.default_backend().create(api.make_config())
    assert e.value.code == api.RSQ_ERR_NO_DEVICE


def test_product_does_not_link_the_oracle(libs):
    import subprocess
    for f in (libs["product"], libs["cli"], libs["host"]):
        out = subprocess.run(["ldd", f], stdout=subprocess.PIPE, text=True).stdout
        assert "oracle" not in out, out
