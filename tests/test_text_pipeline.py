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


def _tokenise_arrays(imp, data, fast):
    """Batch arrays of grimb_text_tokenise for `data` with the fast path of the tokeniser on or off."""
    import numpy as np
    from grim.imputation import _lib
    os.environ["GRIMB_TEXT_FAST"] = "1" if fast else "0"
    try:
        imp._text = None
        t = imp._text_handle()          # the switch is read when the GrimbText is created
    finally:
        os.environ.pop("GRIMB_TEXT_FAST", None)
    lib = _lib.load()
    b = _lib.Batch()
    _lib.check(lib.grimb_text_tokenise(t, C.byref(imp.cfg), data, len(data), 7, C.byref(b)), "tokenise")
    S, L = b.n_subjects, imp.L

    def arr(ptr, n, ct):
        return None if not ptr else np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()

    off = arr(b.allele_off, S + 1, C.c_uint32)
    counts = arr(b.counts, S * L * 2, C.c_uint16)
    typed = arr(b.typed_mask, S, C.c_uint16)
    if counts is None:   # all ones where typed
        counts = np.repeat(np.array([[(m >> l) & 1 for l in range(L)] for m in typed], np.uint16), 2, axis=1).reshape(-1)
    pri = arr(b.prior_index, S, C.c_uint32)
    priors = arr(b.priors, b.n_priors * imp.P * imp.P, C.c_double).reshape(b.n_priors, -1)
    pm = priors[pri] if pri is not None else np.repeat(priors[:1], S, axis=0)
    # classification of every line, as the formatter sees it: format an all-skipped result
    res = _lib.ResultArrays(S, 1)
    res.compact["status"] = _lib.ST_SKIPPED
    res.compact["off"] = _lib.NO_RECORD
    out = _lib.TextOut()
    _lib.check(lib.grimb_text_format(t, C.byref(imp.cfg), C.byref(res.struct), C.byref(out)), "format")
    texts = [C.string_at(out.data[i], out.size[i]) for i in range(6)]
    return {"typed": typed, "counts": counts, "off": off, "alleles": arr(b.alleles, int(off[S]), C.c_uint16), "priors": pm,
            "texts": texts}


@pytest.mark.parametrize("table", ["cau", "pop3"])
def test_fast_tokeniser_path_equals_the_general_parser(table):
    """The single-pass fast path of the tokeniser accepts only lines for which it does exactly what the general
    parser does; everything else falls through.  Same batch arrays, same per-line classification, on clean,
    ambiguous and deliberately dirty lines (LF and CRLF)."""
    import numpy as np
    import synth
    sys_path = os.path.join(goldenlib.GOLD, "..", "golden")
    import sys
    if sys_path not in sys.path:
        sys.path.insert(0, sys_path)
    from fuzz_random_tables import dirty
    name = {"cau": "g5_messy_cau", "pop3": "g3_pop3_messy"}[table]
    _t, conf, lines, _exp = goldenlib.load_case(name)
    eg = _setup(table, conf)
    imp = emu_imputation(eg, load_config(conf))
    tab = synth.Table(open(conf["freq_file"]).read(), conf["populations"][0])
    rng = np.random.RandomState(3)
    races = synth.race_fields(conf["populations"])
    base = synth.typed_subjects(tab, 400, 1, races) + synth.messy_subjects(tab, 400, 2, races=races) + list(lines)
    cases = base + dirty(base, rng, synth.LOCI5, False) + dirty(base[:300], rng, synth.LOCI5, True)
    cases += ["X1,A*01:01+A*02:01^^B*07:02+B*08:01\n", "X2,+A*01:01+A*02:01\n", "X3,A*01:01+A*02:01+A*03:01\n",
              "X4,A*01:01/A*01:01+A*02:01,CAU,CAU,extra\n", "X5%A*01:01+A*02:01\n", "X6,A*01:01+A*02:01^\n",
              "X7, A*01:01+A*02:01\n", "X8,A*01:01+A*02:01 ,CAU,CAU \n", "X9,B*07:02+B*08:01^A*01:01+A*02:01\n", "\n",
              "X10,A*01:01+B*07:02\n", "X11,A*01:01/B*07:02+A*02:01\n", "X12,A*01:01+A*02:01,CAU\n", "X13"]
    data = "".join(cases).encode("utf8")
    a = _tokenise_arrays(imp, data, True)
    b = _tokenise_arrays(imp, data, False)
    for k in ("typed", "counts", "off", "alleles"):
        assert np.array_equal(a[k], b[k]), k
    live = a["typed"] != 0     # the prior of a line that is not imputed is never read (and its index is arbitrary)
    assert np.array_equal(a["priors"][live], b["priors"][live])
    assert a["texts"] == b["texts"]
    assert (a["typed"] != 0).sum() > 1000


@pytest.mark.parametrize("tail_newline", [True, False])
def test_byte_range_line_counts_tile_the_file(tmp_path, tail_newline):
    """grimb_file_count_lines: what every rank of the sharded public API calls on its byte range -- the ranges,
    adjusted to line starts, tile the file and the counts add up to the file's line count."""
    from grim.imputation import _lib
    lib = _lib.load()
    rnd = random.Random(11)
    lines = ["S%d,%s\n" % (i, "x" * rnd.randint(0, 90)) for i in range(5000)] + ["\n", "last,line\n"]
    text = "".join(lines)
    if not tail_newline:
        text = text[:-1]
    path = str(tmp_path / "in.csv")
    open(path, "w").write(text)
    size = len(text)
    for world in (1, 2, 3, 8, 61):
        prev_hi, total = 0, 0
        for rank in range(world):
            lo, hi = size * rank // world, (size * (rank + 1) // world if rank + 1 < world else -1)
            n, a, b = C.c_int64(), C.c_int64(), C.c_int64()
            _lib.check(lib.grimb_file_count_lines(path.encode(), lo, hi, 3, C.byref(n), C.byref(a), C.byref(b)), "count")
            assert a.value == prev_hi
            assert n.value == len(text[a.value:b.value].splitlines())
            prev_hi = b.value
            total += n.value
        assert prev_hi == size and total == len(lines)
    # a range inside one long line holds no line start
    n, a, b = C.c_int64(), C.c_int64(), C.c_int64()
    first_len = len(lines[0])
    _lib.check(lib.grimb_file_count_lines(path.encode(), 1, first_len - 1, 2, C.byref(n), C.byref(a), C.byref(b)), "count")
    assert n.value == 0 and a.value == b.value == first_len


def test_write_at_places_bytes(tmp_path):
    from grim.imputation import _lib
    lib = _lib.load()
    path = str(tmp_path / "out.bin")
    with open(path, "wb") as f:
        f.truncate(10)
    assert lib.grimb_file_write_at(path.encode(), 4, b"abc", 3) == 0
    assert lib.grimb_file_write_at(path.encode(), 0, b"zz", 2) == 0
    assert open(path, "rb").read() == b"zz\x00\x00abc\x00\x00\x00"


_FORK_SCRIPT = r"""
import ctypes as C, os, sys, time
import goldenlib, synth
import grim_oracle as go
from emu_backend import EmuGraph, emu_imputation
from grim.imputation import _lib
from grim.run_impute_def import load_config
table, conf, _l, _e = goldenlib.load_case("g1_readme_donor")
eg = EmuGraph(go.graph_from_config(conf), conf["loci_map"])
imp = emu_imputation(eg, load_config(conf))
tab = synth.Table(open(conf["freq_file"]).read())
data = "".join(synth.typed_subjects(tab, 3000, 5, ["CAU,CAU"])).encode()
lib = _lib.load()
t = imp._text_handle()
def tokenise():
    b = _lib.Batch()
    _lib.check(lib.grimb_text_tokenise(t, C.byref(imp.cfg), data, len(data), 0, C.byref(b)), "tokenise")
    return b.n_subjects
assert tokenise() == 3000                               # the pool's workers exist now
pid = os.fork()
if pid == 0:
    import faulthandler
    faulthandler.dump_traceback_later(60, exit=True)    # a hang becomes a failure, not a stuck test run
    assert tokenise() == 3000                           # the child: same pool object, no threads
    faulthandler.cancel_dump_traceback_later()
    sys.exit(0)                                         # a NORMAL exit: thread_local destructors run
t_end = time.time() + 120
status = None
while time.time() < t_end:
    p, st = os.waitpid(pid, os.WNOHANG)
    if p:
        status = st
        break
    time.sleep(0.05)
if status is None:
    os.kill(pid, 9)
    raise SystemExit("the forked child hangs")
assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0, status
assert tokenise() == 3000
print("FORK-OK")
"""


def test_worker_pool_survives_a_fork():
    """The text pipeline parks its worker threads between parallel regions.  A forked child has the pool object but
    none of its threads (and condition variables with the parent's waiters recorded): it must start afresh, use the
    pipeline and exit normally, and the parent must carry on."""
    import subprocess
    import sys
    env = dict(os.environ, GRIMB_HOST_THREADS="4", PYTHONPATH=os.pathsep.join(sys.path))
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", _FORK_SCRIPT], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "FORK-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
