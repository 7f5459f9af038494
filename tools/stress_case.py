#!/usr/bin/env python
"""Runs golden cases repeatedly on the GPU and reports every run that differs from the reference's
files (race detector: compute-sanitizer is not available on the GPU pool).
    [GRIMB_KEY_WORDS=2] python tools/stress_case.py <repeats> <case> [<case> ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "py-graph-imputation_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import goldenlib  # noqa: E402
from grim.imputation.impute import Imputation  # noqa: E402
from grim.imputation.networkx_graph import Graph  # noqa: E402
from grim.run_impute_def import load_config  # noqa: E402


def main():
    reps = int(sys.argv[1])
    graphs = {}
    bad = 0
    total = 0
    for r in range(reps):
        for name in sys.argv[2:]:
            table, conf, lines, exp = goldenlib.load_case(name)
            cfg = load_config(conf)
            if table not in graphs:
                graphs[table] = Graph(cfg).build_graph()
            imp = Imputation(graphs[table], cfg)
            out = {k: "".join(v) for k, v in imp.impute_lines(lines, em_mr=conf["_hap_pop_pair"]).items()}
            total += 1
            for k in goldenlib.KEYS:
                if out[k] != exp[k]:
                    bad += 1
                    a, b = exp[k].split("\n"), out[k].split("\n")
                    msg = "lengths %d vs %d" % (len(a), len(b))
                    for i, (x, y) in enumerate(zip(a, b)):
                        if x != y:
                            msg = "line %d\n   want %s\n   got  %s" % (i, x[:160], y[:160])
                            break
                    print("run %d %s %s: %s" % (r, name, k, msg), flush=True)
    print("runs", total, "differing texts", bad)


if __name__ == "__main__":
    main()
