"""Golden graph CSV files (nodes/edges/top_links/info_node) written by the UNMODIFIED reference
for a 40-haplotype, 2-population table (build container only).  Used to check the CSV writer and
reader of the B200 build.

    python tests/golden/make_graph_csv.py
"""
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)

import synth  # noqa: E402
from refrun import RefSession  # noqa: E402


def main():
    base = json.load(open(os.path.join(HERE, "data", "base_conf.json")))
    rows = synth.parse_hpf(open(os.path.join(HERE, "data", "pop3_hpf.csv")).read())
    keep = []
    seen = []
    for h, p, f in rows:
        if p not in ("AAA", "BBB"):
            continue
        if h not in seen:
            if len(seen) >= 40:
                continue
            seen.append(h)
        keep.append("%s,%s,%s\n" % (h, p, f))
    hpf = "hap,pop,freq\n" + "".join(keep)
    conf = dict(base)
    conf["populations"] = ["AAA", "BBB"]
    counts = "AAA,1000.0,0.6\nBBB,600.0,0.4\n"
    s = RefSession(conf, hpf, counts)
    out = os.path.join(HERE, "data", "graph40")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    open(os.path.join(out, "hpf.csv"), "w").write(hpf)
    open(os.path.join(out, "pop_counts_file.txt"), "w").write(counts)
    for f in ("nodes.csv", "edges.csv", "top_links.csv", "info_node.csv"):
        shutil.copy(os.path.join(s.dir, "csv", f), os.path.join(out, f))
    s.close()
    print("wrote", out, os.listdir(out))


if __name__ == "__main__":
    main()
