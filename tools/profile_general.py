#!/usr/bin/env python
"""Small driver for profiling the general kernel (k_impute): C3-style typed subjects on 21
populations, or C4-style messy subjects on the CAU table.
    python tools/profile_general.py c3|c4 [n_subjects]"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "py-graph-imputation_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import goldenlib  # noqa: E402
import synth  # noqa: E402
from grim.imputation.impute import Imputation  # noqa: E402
from grim.imputation.networkx_graph import Graph  # noqa: E402
from grim.run_impute_def import load_config  # noqa: E402


def main():
    kind = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    tmp = tempfile.mkdtemp()
    base = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    cau = open(os.path.join(goldenlib.GOLD, "data", "cau_hpf.csv")).read()
    conf = dict(base)
    if kind == "c3":
        pops = ["P%02d" % i for i in range(21)]
        hpf, cnt = synth.multipop_hpf(cau, pops, 21)
        conf.update({"populations": pops, "UNK_priors": "MR", "number_of_pop_results": 100})
        lines = synth.typed_subjects(synth.Table(hpf, "P00"), n, 3, synth.race_fields(pops))
    else:
        hpf, cnt = cau, open(os.path.join(goldenlib.GOLD, "data", "cau_pop_counts.txt")).read()
        lines = synth.messy_subjects(synth.Table(cau), n, 4, max_amb=6)
    open(tmp + "/hpf.csv", "w").write(hpf)
    open(tmp + "/cnt.txt", "w").write(cnt)
    conf["freq_file"], conf["pops_count_file"] = tmp + "/hpf.csv", tmp + "/cnt.txt"
    cfg = load_config(conf)
    g = Graph(cfg).build_graph()
    imp = Imputation(g, cfg)
    data = "".join(lines).encode()
    imp.impute_text(data[: len(data) // 50 + 200].rsplit(b"\n", 1)[0] + b"\n")
    imp = Imputation(g, cfg)
    t = time.time()
    imp.impute_text(data)
    dt = time.time() - t
    print(kind, n, "subjects", round(n / dt), "subj/s total; abi", imp.stats.get("abi_seconds"), "retries", imp.stats["workspace_retries"])
    eng = g.engine(imp.workspaces[0])
    print("last call: k_impute_fast %.3f ms, k_impute %.3f ms, k_impute_typed %.3f ms, handed to k_impute: %d subjects"
          % tuple(g.lib.grimb_engine_kernel_ms(eng, w) for w in (0, 1, 2, 3)))


if __name__ == "__main__" and sys.argv[1] not in ("each", "heavy"):
    main()


def each():
    """time every subject of a C4 batch on its own (finds the heavy tail)"""
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    tmp = tempfile.mkdtemp()
    base = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    cau = open(os.path.join(goldenlib.GOLD, "data", "cau_hpf.csv")).read()
    open(tmp + "/hpf.csv", "w").write(cau)
    open(tmp + "/cnt.txt", "w").write(open(os.path.join(goldenlib.GOLD, "data", "cau_pop_counts.txt")).read())
    conf = dict(base)
    conf["freq_file"], conf["pops_count_file"] = tmp + "/hpf.csv", tmp + "/cnt.txt"
    cfg = load_config(conf)
    g = Graph(cfg).build_graph()
    imp = Imputation(g, cfg)
    lines = synth.messy_subjects(synth.Table(cau), n, 4, max_amb=6)
    imp.impute_text("".join(lines[:50]).encode())
    res = []
    for ln in lines:
        t = time.time()
        imp.impute_text(ln.encode())
        res.append((time.time() - t, ln))
    res.sort(reverse=True)
    tot = sum(r[0] for r in res)
    print("total", round(tot, 3), "s; top 10 =", round(sum(r[0] for r in res[:10]), 3), "s; median", round(res[len(res) // 2][0] * 1e3, 3), "ms")
    for dt, ln in res[:8]:
        gl = ln.split(",")[1]
        loci = gl.split("^")
        print(round(dt * 1e3, 2), "ms", len(loci), "loci", [tuple(len(x.split("/")) for x in l.split("+")) for l in loci])


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "each":
    each()


def heavy():
    """phase 1 (no argument file): find the slowest subjects and store their lines;
    phase 2 (file exists): impute only those in one batch (the launch to profile)."""
    path = os.path.join(ROOT, "gpurun_out", "heavy_lines.json")
    tmp = tempfile.mkdtemp()
    base = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    cau = open(os.path.join(goldenlib.GOLD, "data", "cau_hpf.csv")).read()
    open(tmp + "/hpf.csv", "w").write(cau)
    open(tmp + "/cnt.txt", "w").write(open(os.path.join(goldenlib.GOLD, "data", "cau_pop_counts.txt")).read())
    conf = dict(base)
    conf["freq_file"], conf["pops_count_file"] = tmp + "/hpf.csv", tmp + "/cnt.txt"
    cfg = load_config(conf)
    g = Graph(cfg).build_graph()
    imp = Imputation(g, cfg)
    if not os.path.exists(path):
        lines = synth.messy_subjects(synth.Table(cau), 400, 4, max_amb=6)
        imp.impute_text("".join(lines[:50]).encode())
        res = []
        for ln in lines:
            t = time.time()
            imp.impute_text(ln.encode())
            res.append((time.time() - t, ln))
        res.sort(reverse=True)
        json.dump([r[1] for r in res[:8]], open(path, "w"))
        print("stored", [round(r[0] * 1e3) for r in res[:8]])
        return
    lines = json.load(open(path))
    imp.workspaces = [imp.workspaces[-2]]
    imp.impute_text(lines[7].encode())
    t = time.time()
    imp.impute_text("".join(lines).encode())
    print("heavy batch", round((time.time() - t) * 1e3, 1), "ms", imp.stats)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "heavy":
    heavy()
