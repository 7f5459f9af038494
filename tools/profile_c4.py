#!/usr/bin/env python
"""Driver for profiling the general kernel on BASELINE config 4: `heavy` (6-40 alleles per side, products on
both sides of the 100,000-option threshold) or `messy` subjects on the README table, one batch through
grimb_impute_text.    python tools/profile_c4.py heavy|messy [n_subjects]"""
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

kind = sys.argv[1] if len(sys.argv) > 1 else "heavy"
n = int(sys.argv[2]) if len(sys.argv) > 2 else (400 if kind == "heavy" else 4000)
_t, conf, _l, _e = goldenlib.load_case("g1_readme_donor")
cfg = load_config(conf)
g = Graph(cfg).build_graph()
tab = synth.Table(open(conf["freq_file"]).read())
lines = synth.heavy_subjects(tab, n, 44, races=["CAU,CAU"]) if kind == "heavy" else synth.messy_subjects(tab, n, 4, max_amb=6)
imp = Imputation(g, cfg)
if kind == "heavy":
    imp.workspaces = [512 << 20, 4 << 30]
data = "".join(lines).encode()
imp.impute_text("".join(lines[:8]).encode())
t = time.time()
imp.impute_text(data)
dt = time.time() - t
eng = g.engine(imp.workspaces[0])
print(kind, n, "subjects", round(n / dt), "subj/s; abi", imp.stats.get("abi_seconds"), "retries", imp.stats["workspace_retries"],
      "plans", imp.stats["plan"])
