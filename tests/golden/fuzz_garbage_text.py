"""Crash-safety fuzz of the C++ text pipeline (grimb_text.cpp: tokeniser + formatter) around the emulated
kernel source: valid subject lines with random byte insertions / deletions / replacements (separators, NUL,
non-UTF-8 bytes, duplicated prefixes), very long lines, a missing final newline.  Nothing is compared: the run
must not crash, hang or trip a sanitizer (tests/tools/run_sanitizers.sh runs it under ASan + UBSan).

    python tests/golden/fuzz_garbage_text.py [seed] [iterations]
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(HERE, "..", "..", "py-graph-imputation_b200"), os.path.join(HERE, "..", "..", "oracle"),
          os.path.join(HERE, "..")):
    sys.path.insert(0, p)
import numpy as np
import grim_oracle as go, synth, goldenlib
from emu_backend import EmuGraph, emu_imputation, emu_impute_text
from grim.run_impute_def import load_config
conf = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
hpf = synth.zipf_table(120, [4,5,3,4,6], 11)
counts = "CAU,1000.0,1.0\n"
og = go.OracleGraph(hpf.splitlines(True), conf["populations"], conf["loci_map"], conf["freq_trim_threshold"], counts.splitlines(True))
eg = EmuGraph(og, conf["loci_map"])
tab = synth.Table(hpf)
cbp = np.array([1.0])
base = synth.typed_subjects(tab, 40, 1, ["CAU,CAU"]) + synth.messy_subjects(tab, 40, 2)
special = b",^+/*:~%;\t \r|-_gLU0123456789ABCDQRP\x00\xff\xc3\x28"
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 200):
    lines = []
    for ln in base:
        b = bytearray(ln.rstrip("\n").encode())
        for _ in range(int(rng.randint(0, 4))):
            op = int(rng.randint(0, 4)); pos = int(rng.randint(0, len(b) + 1))
            ch = special[int(rng.randint(len(special)))]
            if op == 0: b.insert(pos, ch)
            elif op == 1 and len(b): del b[min(pos, len(b) - 1)]
            elif op == 2 and len(b): b[min(pos, len(b) - 1)] = ch
            elif op == 3: b[pos:pos] = bytes(b[: int(rng.randint(0, 40))])   # duplicate a prefix
        b = bytes(b).replace(b"\n", b"")
        lines.append(b + b"\n")
    data = b"".join(lines)
    if it % 5 == 0:
        data = data + b"A*01:01" * 3000 + b"\n"       # one very long line
    if it % 7 == 0:
        data = data[:-1]                              # no trailing newline
    imp = emu_imputation(eg, load_config(conf), cbp)
    try:
        out = emu_impute_text(imp, eg, data)
    except UnicodeDecodeError:
        out = None   # the helper decodes for comparison; the pipeline itself returned
    if it % 20 == 0:
        print("iter", it, "ok", {k: v.count("\n") for k, v in out.items()} if out else "non-utf8 echoed", flush=True)
print("DONE no crash")
