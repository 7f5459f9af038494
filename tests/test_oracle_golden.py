"""Pins oracle/grim_oracle.py against every golden fixture produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import pytest

import goldenlib
import grim_oracle as go

_graphs = {}


def _graph(table, conf):
    if table not in _graphs:
        _graphs[table] = go.graph_from_config(conf)
    return _graphs[table]


@pytest.mark.parametrize("name", goldenlib.case_names())
def test_oracle_matches_reference_files(name):
    table, conf, lines, exp = goldenlib.load_case(name)
    out, _ = go.impute_file(conf, graph=_graph(table, conf), lines=lines, em_mr=conf["_hap_pop_pair"])
    for k in goldenlib.KEYS:
        assert out[k] == exp[k], "%s: %s differs" % (name, k)


def test_readme_counts():
    # README.md:123-124 of the reference: 8400 PMUG pairs / 6028 UMUG genotypes for donor D1
    table, conf, lines, _ = goldenlib.load_case("g1_readme_donor")
    imp = go.OracleImputation(_graph(table, conf), go.load_config(conf),
                              go.count_by_prob_from_file(1, conf["pops_count_file"]))
    res_muugs, res_haps = imp.impute_one(lines[0].split(",")[1], "CAU", "CAU")
    assert len(res_haps["Haps"]) == 8400
    assert len(res_muugs["Haps"]) == 6028
