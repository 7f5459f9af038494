#!/usr/bin/env python
"""CPU microbenchmark of the C++ tokeniser (no CUDA calls): README table, typed subjects.
    GRIMB_HOST_THREADS=1 python tools/bench_tokenise.py [n_lines]"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "py-graph-imputation_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import goldenlib  # noqa: E402
import grim_oracle as go  # noqa: E402
import synth  # noqa: E402
from emu_backend import EmuGraph, emu_imputation  # noqa: E402
from grim.imputation import _lib  # noqa: E402
from grim.run_impute_def import load_config  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
_, conf, _, _ = goldenlib.load_case("g1_readme_donor")
og = go.graph_from_config(conf)
eg = EmuGraph(og, conf["loci_map"])
imp = emu_imputation(eg, load_config(conf))
tab = synth.Table(open(conf["freq_file"]).read())
data = "".join(synth.typed_subjects(tab, n, 5, ["CAU,CAU"])).encode()
lib = _lib.load()
t = imp._text_handle()
b = _lib.Batch()
for rep in range(4):
    t0 = time.time()
    _lib.check(lib.grimb_text_tokenise(t, C.byref(imp.cfg), data, len(data), 0, C.byref(b)), "tokenise")
    dt = time.time() - t0
    print("tokenise %d lines (%d bytes): %.1f ms, %.0f ns/line, %.0f MB/s, packed=%s" % (
        n, len(data), dt * 1e3, dt / n * 1e9, len(data) / dt / 1e6, bool(b.packed_keys)))

# formatter: results of the emulated kernel source for the same batch, formatted repeatedly
S = b.n_subjects
res = _lib.ResultArrays(S, 1, general=S + 16, hap=4096, pop=4096)
while True:
    r = res.struct
    rc = eg.emu.grimb_emu_impute(C.byref(eg.tables), C.byref(imp.cfg), C.byref(b), C.byref(r), 64 << 20)
    if rc == _lib.E_CAPACITY:
        res.grow()
        continue
    assert rc == 0
    break
out = _lib.TextOut()
for rep in range(4):
    t0 = time.time()
    _lib.check(lib.grimb_text_format(t, C.byref(imp.cfg), C.byref(r), C.byref(out)), "format")
    dt = time.time() - t0
    nb = sum(out.size[i] for i in range(6))
    print("format %d subjects -> %d bytes: %.1f ms, %.0f ns/subject, %.0f MB/s" % (S, nb, dt * 1e3, dt / S * 1e9, nb / dt / 1e6))
