#!/usr/bin/env python
"""Where does the time of a C4 (messy) batch go: the whole batch, prefixes of it, and the slowest single subjects
(each timed alone through grimb_impute_text).    python tools/c4_tail.py [n_subjects] [n_singles]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "py-graph-imputation_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import goldenlib  # noqa: E402
import synth  # noqa: E402
from grim.imputation.impute import Imputation  # noqa: E402
from grim.imputation.networkx_graph import Graph  # noqa: E402
from grim.run_impute_def import load_config  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
singles = int(sys.argv[2]) if len(sys.argv) > 2 else 400
_t, conf, _l, _e = goldenlib.load_case("g1_readme_donor")
cfg = load_config(conf)
g = Graph(cfg).build_graph()
tab = synth.Table(open(conf["freq_file"]).read())
lines = synth.messy_subjects(tab, n, 4, max_amb=6)
imp = Imputation(g, cfg)


def run(ls):
    t = time.time()
    out = imp.impute_text("".join(ls).encode())
    return time.time() - t, out


if os.environ.get("C4_ONLY"):       # one subject, for an ncu capture of its launch
    i = int(os.environ["C4_ONLY"])
    for _ in range(2):
        dt, _o = run(lines[i:i + 1])
        print("subject #%d alone: %.2f ms" % (i, dt * 1e3))
    sys.exit(0)
run(lines[:8])
for m in (n, n // 2, n // 4, n // 8, n // 16):
    dt, _ = run(lines[:m])
    print("first %5d subjects: %.1f ms" % (m, dt * 1e3))
times = []
for i in range(min(singles, n)):
    dt, out = run(lines[i:i + 1])
    times.append((dt, i, len(out["pmug"]), len(out["umug"])))
times.sort(reverse=True)
print("slowest single subjects (ms, index, pmug bytes, umug bytes):")
for dt, i, a, b in times[:8]:
    print("  %.2f ms  #%d  pmug %d B umug %d B   %s" % (dt * 1e3, i, a, b, lines[i][:160].strip()))
print("median single subject: %.2f ms" % (sorted(t[0] for t in times)[len(times) // 2] * 1e3))
