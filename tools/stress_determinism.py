#!/usr/bin/env python
"""Repeats one mixed batch (typed + messy subjects, 3 populations) through both host front ends and
reports any run whose six output texts differ from the first (race detector for the kernels).
    python tools/stress_determinism.py [repeats]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "py-graph-imputation_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import goldenlib  # noqa: E402
import synth  # noqa: E402
from grim.imputation.impute import Imputation  # noqa: E402
from grim.imputation.networkx_graph import Graph  # noqa: E402
from grim.run_impute_def import load_config  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    _, conf, _, _ = goldenlib.load_case("g3_pop3_typed")
    hpf = open(conf["freq_file"]).read()
    tab = synth.Table(hpf, "AAA")
    races = synth.race_fields(conf["populations"])
    lines = synth.typed_subjects(tab, 20000, 77, races) + synth.messy_subjects(tab, 300, 78, races=races)
    cfg = load_config(conf)
    g = Graph(cfg).build_graph()
    data = "".join(lines).encode("utf8")
    first = None
    bad = 0
    for r in range(reps):
        imp = Imputation(g, cfg)
        out = imp.impute_text(data) if r % 2 == 0 else {k: "".join(v).encode("utf8") for k, v in imp.impute_lines(lines).items()}
        if first is None:
            first = out
            continue
        for k in goldenlib.KEYS:
            if out[k] != first[k]:
                bad += 1
                a, b = first[k].split(b"\n"), out[k].split(b"\n")
                for i, (x, y) in enumerate(zip(a, b)):
                    if x != y:
                        print("run %d (%s) %s line %d:\n  first: %s\n  now:   %s" % (r, "text" if r % 2 == 0 else "python", k, i, x[:200], y[:200]))
                        break
                else:
                    print("run %d %s: lengths %d vs %d" % (r, k, len(a), len(b)))
    print("runs", reps, "differing texts", bad)


if __name__ == "__main__":
    main()
