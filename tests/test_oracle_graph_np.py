"""Pins oracle/oracle_graph_np.py (the array-backed oracle store used for the BASELINE-size
tables) against oracle/grim_oracle.OracleGraph, which is itself pinned against the unmodified
reference by the golden fixtures: every query, every label, incl. the CSR sentinel quirk on the
last node / last connector; then every golden case end to end through the numpy store.  CPU only."""
import os
import random

import pytest

import goldenlib
import grim_oracle as go
from oracle_graph_np import NumpyOracleGraph

_CASE_OF_TABLE = {"cau": "g1_readme_donor", "pop3": "g3_pop3_typed", "nine": "g6_nine_loci"}
_pairs = {}


def _graphs(table):
    if table not in _pairs:
        _t, conf, _lines, _exp = goldenlib.load_case(_CASE_OF_TABLE[table])
        og = go.graph_from_config(conf)
        hpf = open(conf["freq_file"]).readlines()
        pc = open(conf["pops_count_file"]).readlines()
        ng = NumpyOracleGraph.from_hpf(hpf, conf["populations"], conf["loci_map"], conf["freq_trim_threshold"], pc)
        _pairs[table] = (og, ng)
    return _pairs[table]


def _same(a, b, order=True):
    assert list(a.keys()) == list(b.keys()) if order else set(a) == set(b)
    for k in a:
        assert a[k] == b[k], k


@pytest.mark.parametrize("table", ["cau", "pop3", "nine"])
def test_queries_equal_oracle_graph(table):
    og, ng = _graphs(table)
    assert ng.n_nodes == og.n_nodes and ng.n_full == og.n_full and ng.labels == og.labels
    rng = random.Random(5)
    labels = og.labels if len(og.labels) <= 40 else [og.labels[0]] + rng.sample(og.labels[1:], 24) + og.labels[-3:]
    for lab in labels:
        names = og.haps_by_label(lab)
        assert ng.haps_by_label(lab) == names
        _same(og.haps_with_probs_by_label(lab), ng.haps_with_probs_by_label(lab))
        pick = names if len(names) <= 60 else rng.sample(names, 50) + names[-5:] + names[:5]
        pick = pick + ["A*99:99", "A*99:99~B*07:02"]
        _same(og.node_probs(pick), ng.node_probs(pick))
        # top links one name at a time (the reference's IndexError must surface identically)
        for n in pick[:30] + pick[-8:]:
            try:
                a = og.adjs_query([n])
            except IndexError:
                with pytest.raises(IndexError):
                    ng.adjs_query([n])
                continue
            _same(a, ng.adjs_query([n]))
        _same(og.adjs_query(pick[5:25]), ng.adjs_query(pick[5:25]))
        # connectors towards every label one locus longer
        for lab_b in og.labels:
            if len(lab_b) != len(lab) + 1 or not set(lab) <= set(lab_b):
                continue
            for n in pick[:12] + pick[-8:]:
                try:
                    a = og.adjs_query_by_color([n], lab, lab_b)
                except IndexError:
                    with pytest.raises(IndexError):
                        ng.adjs_query_by_color([n], lab, lab_b)
                    continue
                _same(a, ng.adjs_query_by_color([n], lab, lab_b))
        _same(og.adjs_query_by_color(pick[:10], lab, lab), ng.adjs_query_by_color(pick[:10], lab, lab))


def test_last_node_and_last_connector_quirk():
    og, ng = _graphs("cau")
    last = og.names[-1]
    assert og.adjs_query([last]) == ng.adjs_query([last]) == {}     # SURVEY trap T1: DRB1*15:03 has no top links
    full = og.full_label
    for lab_b in og.labels:
        if len(lab_b) == 2 and full[-1] in lab_b:
            _same(og.adjs_query_by_color([last], full[-1], lab_b), ng.adjs_query_by_color([last], full[-1], lab_b))


# a subset that covers every plan, table and query kind (the whole set takes minutes in pure Python and
# runs through OracleGraph in test_oracle_golden.py; GOLDEN_ALL=1 runs every default-mode case here too)
_SUBSET = ["g1_readme_donor", "g2_edges", "g2_t1_last_node", "g3_pop3_typed", "g3_pop3_priority",
           "g4_amb12_over_threshold", "g4_low_threshold", "g4_save_space", "g5_typed_cau", "g6_nine_loci"]


@pytest.mark.parametrize("name", [n for n in goldenlib.case_names() if not goldenlib.is_special(n)
                                  and (n in _SUBSET or os.environ.get("GOLDEN_ALL") == "1")])
def test_golden_cases_through_numpy_store(name):
    table, conf, lines, exp = goldenlib.load_case(name)
    _og, ng = _graphs(table)
    out, _ = go.impute_file(conf, graph=ng, lines=lines)
    for k in goldenlib.KEYS:
        assert out[k] == exp[k], "%s: %s differs" % (name, k)


@pytest.mark.parametrize("name", ["g8_plan_a_blocks", "g8_plan_a_blocks_planb_off", "g8_plan_a_last_node", "g8_plan_a_nine"])
def test_plan_a_matrix_cases_through_numpy_store(name):
    """The array-backed store under a Plan_A_Matrix (what bench.py's C5_matrix record is checked with): the
    reference's files for the g8 cases."""
    _t, conf, lines, exp = goldenlib.load_case(name)
    hpf = open(conf["freq_file"]).readlines()
    pc = open(conf["pops_count_file"]).readlines()
    ng = NumpyOracleGraph.from_hpf(hpf, conf["populations"], conf["loci_map"], conf["freq_trim_threshold"], pc,
                                   plan_a_matrix=conf["Plan_A_Matrix"])
    out, _ = go.impute_file(conf, graph=ng, lines=lines)
    for k in goldenlib.KEYS:
        assert out[k] == exp[k], "%s: %s differs" % (name, k)
