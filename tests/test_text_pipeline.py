"""The C++ host text pipeline (grimb_text_tokenise / grimb_text_format in libgrimb200.so): CPU
tests drive its two host-only halves around the emulated kernel source and compare with the
golden files of the unmodified reference; the float formatter is checked against Python's
repr() on a large sample of doubles."""
import ctypes as C
import os
import random
import struct

import pytest

import goldenlib
import grim_oracle as go
from emu_backend import EmuGraph, emu_imputation, emu_impute_text
from grim.run_impute_def import load_config

_cache = {}


def _setup(table, conf):
    if table not in _cache:
        og = go.graph_from_config(conf)
        _cache[table] = EmuGraph(og, conf["loci_map"])
    return _cache[table]


@pytest.mark.parametrize("name", goldenlib.text_case_names())
def test_native_tokeniser_and_formatter_match_reference_files(name):
    table, conf, lines, exp = goldenlib.load_case(name)
    eg = _setup(table, conf)
    imp = emu_imputation(eg, load_config(conf))
    out = emu_impute_text(imp, eg, "".join(lines).encode("utf8"))
    for k in goldenlib.KEYS:
        assert out[k] == exp[k], "%s: %s differs" % (name, k)


def test_first_line_index_offsets_miss_and_problem_rows():
    table, conf, lines, exp = goldenlib.load_case("g2_edges")
    eg = _setup(table, conf)
    imp = emu_imputation(eg, load_config(conf))
    out = emu_impute_text(imp, eg, "".join(lines).encode("utf8"), first_index=1000)
    want = "".join(("%d,%s" % (int(r.split(",")[0]) + 1000, r.split(",", 1)[1]) if r.split(",")[0].isdigit() and len(r.split(",")) == 2 else r) + "\n"
                   for r in exp["problem"].splitlines())
    assert out["problem"] == want


def test_float_formatter_equals_python_repr():
    """py_float() in grimb_text.cpp vs repr(): exercised through a one-subject format call is too
    indirect, so the formatter is exported for tests via the pop-row probability of a Plan-A row."""
    from grim.imputation import _lib
    import numpy as np
    lib = _lib.load()
    table, conf, lines, exp = goldenlib.load_case("g1_readme_donor")
    eg = _setup(table, conf)
    cfg = load_config(conf)
    imp = emu_imputation(eg, cfg)
    t = imp._text_handle()
    rnd = random.Random(5)
    vals = [1e-05, 1e-4, 0.0001, 1e16, 1e15, 123456789012345680.0, 5e-324, 1.7976931348623157e308, 0.1, 1 / 3, 2.5e-17,
            8.838563003520004e-17, 0.0016607054, 1.0, 100.0, 1e22, 9.999999999999999e-05]
    for _ in range(20000):
        bits = rnd.getrandbits(64) & 0x7FFFFFFFFFFFFFFF
        v = struct.unpack("<d", struct.pack("<Q", bits))[0]
        if v == v and v != float("inf"):
            vals.append(v)
    for _ in range(20000):
        vals.append(rnd.random() * 10 ** rnd.randint(-30, 5))
    # one synthetic subject per value: a single pop row carries the value through the formatter
    n = len(vals)
    data = "".join("S%d,A*01:01+A*01:01\n" % i for i in range(n)).encode()
    b = _lib.Batch()
    _lib.check(lib.grimb_text_tokenise(t, C.byref(imp.cfg), data, len(data), 0, C.byref(b)), "tokenise")
    res = _lib.ResultArrays(n, 1, general=n, pop=n)
    res.general["plan_umug"] = 1
    res.general["n_umug_pops"] = 1
    res.general["tot_umug"] = 1
    res.general["pop_off"] = np.arange(n)
    res.pop_rows["prob"] = vals
    res.compact["kind_flags"] = _lib.KIND_GENERAL | _lib.KIND_HAS_RESULTS
    res.compact["off"] = np.arange(n)
    r = res.struct
    out = _lib.TextOut()
    _lib.check(lib.grimb_text_format(t, C.byref(imp.cfg), C.byref(r), C.byref(out)), "format")
    rows = C.string_at(out.data[1], out.size[1]).decode().splitlines()
    assert len(rows) == n
    for v, row in zip(vals, rows):
        assert row.split(",")[3] == repr(v), (v, row)
