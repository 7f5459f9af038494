"""CPU: libgrimb200.so loads and exports every symbol include/grimb200.h declares; the Python
binding lists the same set; and without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from grim.imputation import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "grimb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(grimb_[a-z_]+)\s*\(", text)))


@pytest.fixture(scope="module", params=[1, 2], ids=["keys64", "keys128"])
def lib(request):
    kw = request.param
    if not os.path.exists(_lib.lib_path(kw)):
        subprocess.run(["sh", os.path.join(ROOT, "py-graph-imputation_b200", "csrc", "build.sh")], check=True)
    return ctypes.CDLL(_lib.lib_path(kw))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "libgrimb200.so does not export %s" % n


def test_binding_lists_header_symbols():
    assert sorted(_lib.EXPORTED) == _declared()


def test_abi_version_and_struct_sizes(lib):
    import numpy as np
    lib.grimb_abi_version.restype = ctypes.c_int
    assert lib.grimb_abi_version() == 5
    assert np.dtype(_lib.SUBJECT_DTYPE).itemsize == 48
    assert np.dtype(_lib.COMPACT_DTYPE).itemsize == 16
    assert ctypes.sizeof(_lib.Results) == 10 * 8
    assert np.dtype(_lib.hap_row_dtype(1)).itemsize == 24
    assert np.dtype(_lib.hap_row_dtype(2)).itemsize == 40
    assert np.dtype(_lib.POP_ROW_DTYPE).itemsize == 16


def test_no_cpu_fallback_without_gpu(lib):
    """On a box without a CUDA device the product must raise, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import goldenlib
    from grim.imputation.networkx_graph import Graph
    from grim.run_impute_def import load_config
    _, conf, _, _ = goldenlib.load_case("g1_readme_donor")
    with pytest.raises(RuntimeError):
        Graph(load_config(conf)).build_graph()


def test_product_never_reaches_the_oracle_or_the_emulation():
    pkg = os.path.join(ROOT, "py-graph-imputation_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "grim_oracle" not in src and "emu" not in src.lower(), f
    build = open(os.path.join(pkg, "csrc", "build.sh")).read()
    assert "GRIMB_EMU" not in build
