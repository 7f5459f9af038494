"""Store seam: the host-side view over the table arrays answers the reference's Graph
queries (networkx_graph.py:215-321) exactly like the oracle's restatement of them.  CPU: the arrays
are laid out from the oracle graph by the test helper (same layout as Graph.export())."""
import json
import os

import numpy as np
import pytest

import goldenlib
import grim_oracle as go
import synth
from emu_backend import arrays_from_oracle
from grim.imputation.networkx_graph import loci_in_order
from grim.imputation.store_view import StoreView


def _view(og, loci_map):
    loci = loci_in_order(loci_map)
    arr = arrays_from_oracle(og, loci)
    return StoreView(arr, loci, loci_map, arr["alleles"])


def _tables():
    conf = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    cau = open(os.path.join(goldenlib.GOLD, "data", "cau_hpf.csv")).read()
    yield "cau", conf, cau, "CAU,3380.0,1.0\n", conf["populations"]
    pops = ["AAA", "BBB", "CCC"]
    hpf3, cnt3 = synth.multipop_hpf(cau, pops, 7)
    yield "pop3", conf, hpf3, cnt3, pops
    dense = synth.zipf_table(150, [3, 4, 2, 5, 3], 5, pops=("AAA", "BBB"))
    yield "dense", conf, dense, "AAA,100.0,0.5\nBBB,100.0,0.5\n", ["AAA", "BBB"]


@pytest.mark.parametrize("case", list(_tables()), ids=lambda c: c[0])
def test_store_view_matches_oracle_queries(case):
    _name, conf, hpf, counts, pops = case
    og = go.OracleGraph(hpf.splitlines(True), pops, conf["loci_map"], conf["freq_trim_threshold"], counts.splitlines(True))
    sv = _view(og, conf["loci_map"])
    rng = np.random.RandomState(3)
    labels = list(og.by_label.keys())
    # haps_by_label / haps_with_probs_by_label: every label, order included
    for lab in labels:
        assert sv.haps_by_label(lab) == og.haps_by_label(lab), lab
    for lab in [labels[i] for i in rng.choice(len(labels), size=6, replace=False)]:
        a, b = sv.haps_with_probs_by_label(lab), og.haps_with_probs_by_label(lab)
        assert list(a) == list(b) and all(list(a[k]) == [float(x) for x in b[k]] for k in a), lab
    assert sv.haps_by_label("99") == [] and sv.haps_with_probs_by_label("") == {}
    # adjs_query / node_probs on names drawn from every label, plus absent and malformed names
    names = []
    for lab in labels:
        lst = og.by_label[lab]
        names += [lst[i] for i in rng.choice(len(lst), size=min(6, len(lst)), replace=False)]
    names += ["A*99:99", "A*01:01~ZZZ*01:01", "", "B*07:02~A*01:01", og.names[-1], og.names[og.n_nodes - 1]]
    rng.shuffle(names)

    def same(a, b):
        assert list(a) == list(b)
        for k in a:
            assert list(a[k]) == [float(x) for x in b[k]], k

    def call(f, *args):
        try:
            return f(*args)
        except IndexError:
            return "IndexError"

    for i in range(0, len(names), 7):
        chunk = names[i:i + 7]
        x, y = call(sv.adjs_query, chunk), call(og.adjs_query, chunk)
        if "IndexError" in (x, y):
            assert x == y
        else:
            same(x, y)
        same(sv.node_probs(chunk, "12345"), og.node_probs(chunk))
    # adjs_query_by_color: children of every label against every label (valid and invalid parents)
    some = [labels[i] for i in rng.choice(len(labels), size=8, replace=False)]
    for la in some:
        kids = og.by_label[la][:5] + [og.by_label[la][-1]]
        for lb in labels:
            x, y = call(sv.adjs_query_by_color, kids, la, lb), call(og.adjs_query_by_color, kids, la, lb)
            if "IndexError" in (x, y):
                assert x == y, (la, lb)
            else:
                same(x, y)
