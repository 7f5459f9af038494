"""GPU: Graph.adjs_query / adjs_query_by_color / node_probs / haps_by_label answered from the DEVICE
tables (Graph.export() -> StoreView) equal the oracle's restatement of networkx_graph.py:215-321,
including SURVEY trap T1 (the last node's empty adjacency).  The query code itself is covered on the
CPU by test_store_view.py over arrays laid out from the oracle graph."""
import numpy as np
import pytest

import goldenlib
import grim_oracle as go

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["g1_readme_donor", "g3_pop3_typed"])
def test_graph_store_queries_match_oracle(name):
    from grim.imputation.networkx_graph import Graph
    from grim.run_impute_def import load_config
    _, conf, _, _ = goldenlib.load_case(name)
    g = Graph(load_config(conf)).build_graph()
    og = go.graph_from_config(conf)
    rng = np.random.RandomState(5)
    labels = list(og.by_label.keys())

    def same(a, b):
        assert list(a) == list(b)
        for k in a:
            assert list(a[k]) == [float(x) for x in b[k]], k

    try:
        for lab in labels:
            assert g.haps_by_label(lab) == og.haps_by_label(lab), lab
        same(g.haps_with_probs_by_label(labels[3]), og.haps_with_probs_by_label(labels[3]))
        assert len(g.adjs_query(["DRB1*15:01"])) == len(og.adjs_query(["DRB1*15:01"])) > 0
        same(g.adjs_query([og.names[og.n_nodes - 1]]), og.adjs_query([og.names[og.n_nodes - 1]]))   # trap T1
        names = []
        for lab in labels:
            lst = og.by_label[lab]
            names += [lst[i] for i in rng.choice(len(lst), size=min(4, len(lst)), replace=False)]
        names += ["A*99:99", ""]
        same(g.adjs_query(names), og.adjs_query(names))
        same(g.node_probs(names, "12345"), og.node_probs(names))
        for la in [labels[i] for i in rng.choice(len(labels), size=5, replace=False)]:
            kids = og.by_label[la][:4]
            for lb in labels:
                same(g.adjs_query_by_color(kids, la, lb), og.adjs_query_by_color(kids, la, lb))
    finally:
        g.close()
