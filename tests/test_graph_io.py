"""Graph file formats.  CPU: the nodes.csv reader against the reference-written
golden directory.  GPU: CSV files written from the device tables equal the reference's files
(top_links as a row set, see graph_io.py), tables rebuilt from nodes.csv equal tables built from
hpf.csv, and the binary cache round-trips."""
import json
import os

import numpy as np
import pytest

import goldenlib
from grim.imputation.graph_io import read_nodes_csv
from grim.run_impute_def import load_config

G40 = os.path.join(goldenlib.GOLD, "data", "graph40")


def _conf():
    conf = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    conf["populations"] = ["AAA", "BBB"]
    conf["freq_file"] = os.path.join(G40, "hpf.csv")
    conf["pops_count_file"] = os.path.join(G40, "pop_counts_file.txt")
    return conf


def test_read_nodes_csv_recovers_full_haplotypes():
    alleles, fa, ff = read_nodes_csv(os.path.join(G40, "nodes.csv"), ["A", "B", "C", "DQB1", "DRB1"], "12345")
    assert fa.shape == (40, 5) and ff.shape == (40, 2)
    rows = [l.split(",") for l in open(os.path.join(G40, "nodes.csv")).read().splitlines()[1:41]]
    for i, r in enumerate(rows):
        assert "~".join(alleles[l][fa[i, l] - 1] for l in range(5)) == r[1]
        assert [float(x) for x in r[3].split(";")] == list(ff[i])


@pytest.mark.gpu
def test_written_csv_equals_reference_files(tmp_path):
    from grim.imputation.networkx_graph import Graph
    g = Graph(load_config(_conf())).build_graph()
    g.write_csv(str(tmp_path))
    for f in ("nodes.csv", "edges.csv", "info_node.csv"):
        assert open(os.path.join(str(tmp_path), f), newline="").read() == open(os.path.join(G40, f), newline="").read(), f
    mine = sorted(open(os.path.join(str(tmp_path), "top_links.csv"), newline="").read().split("\r\n"))
    ref = sorted(open(os.path.join(G40, "top_links.csv"), newline="").read().split("\r\n"))
    assert mine == ref


@pytest.mark.gpu
def test_tables_from_nodes_csv_and_cache_round_trip(tmp_path):
    from grim.imputation.networkx_graph import Graph
    cfg = load_config(_conf())
    a = Graph(cfg).build_graph().export()
    cfg2 = dict(cfg)
    cfg2["freq_file"] = os.path.join(str(tmp_path), "missing.csv")
    g2 = Graph(cfg2).build_graph(os.path.join(G40, "nodes.csv"))
    b = g2.export()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    path = os.path.join(str(tmp_path), "tables.bin")
    g2.save_cache(path)
    g3 = Graph(cfg).load_cache(path)
    c = g3.export()
    for k in a:
        assert np.array_equal(a[k], c[k]), k
    assert g3.alleles == g2.alleles
