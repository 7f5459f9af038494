"""GPU parity of the 128-bit-key build (libgrimb200w.so, GRIMB_KEY_WORDS = 2): every golden case
written by the unmodified reference is run again with the wide key layout forced, through both
host front ends (numpy/Python formatter and the C++ text pipeline).  Same bar: files identical."""
import numpy as np
import pytest

import goldenlib

pytestmark = pytest.mark.gpu

_graphs = {}


@pytest.fixture(autouse=True)
def _force_wide(monkeypatch):
    monkeypatch.setenv("GRIMB_KEY_WORDS", "2")


def _graph(table, conf):
    from grim.imputation.networkx_graph import Graph
    from grim.run_impute_def import load_config
    if table not in _graphs:
        if "+matrix:" in table:   # of the Plan_A_Matrix graphs only the one in use stays (engine workspaces are GBs)
            for k in [k for k in _graphs if "+matrix:" in k]:
                _graphs.pop(k).close()
        g = Graph(load_config(conf)).build_graph()
        assert g.kw == 2 and sum(g.key_bits) > 63
        _graphs[table] = g
    return _graphs[table]


@pytest.mark.parametrize("name", goldenlib.case_names())
def test_wide_key_build_matches_reference_files(name):
    from grim.imputation.impute import Imputation
    from grim.run_impute_def import load_config
    table, conf, lines, exp = goldenlib.load_case(name)
    imp = Imputation(_graph(table, conf), load_config(conf))
    out = {k: "".join(v) for k, v in imp.impute_lines(lines, em_mr=conf["_hap_pop_pair"]).items()}
    for k in goldenlib.KEYS:
        assert out[k] == exp[k], "%s: %s differs (python host)" % (name, k)
    if goldenlib.is_special(name):
        return
    txt = imp.impute_text("".join(lines).encode("utf8"))
    for k in goldenlib.KEYS:
        assert txt[k].decode("utf8") == exp[k], "%s: %s differs (text pipeline)" % (name, k)


def test_wide_tables_equal_narrow_tables():
    """Same node ids, sums and adjacency whichever key width packs the haplotypes."""
    import os
    from grim.imputation.networkx_graph import Graph
    from grim.run_impute_def import load_config
    _, conf, _, _ = goldenlib.load_case("g3_pop3_typed")
    wide = _graph("pop3", conf).export()
    os.environ["GRIMB_KEY_WORDS"] = "1"
    g1 = Graph(load_config(conf)).build_graph()
    assert g1.kw == 1
    narrow = g1.export()
    for k in ("node_freq", "tl_cnt", "label_first", "label_count", "cn_cnt"):
        assert np.array_equal(wide[k], narrow[k]), k
    n = len(narrow["tl_cnt"])
    for i in range(0, n, max(1, n // 3000)):
        a = wide["tl_adj"][wide["tl_start"][i]: wide["tl_start"][i] + wide["tl_cnt"][i]]
        b = narrow["tl_adj"][narrow["tl_start"][i]: narrow["tl_start"][i] + narrow["tl_cnt"][i]]
        assert np.array_equal(a, b)
    g1.close()
