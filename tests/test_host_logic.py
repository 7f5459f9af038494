"""CPU: host-side logic of the product (tokeniser, config, priors, sharding) against the oracle's
restatement of the reference behaviour."""
import numpy as np
import pytest

import goldenlib
import grim_oracle as go
from emu_backend import EmuGraph, emu_imputation
from grim.imputation import impute as gi
from grim.imputation.multi_gpu import shard_range
from grim.imputation.networkx_graph import key_layout, loci_in_order, read_hpf
from grim.run_impute_def import load_config

_state = {}


def _cau():
    if "cau" not in _state:
        _, conf, _, _ = goldenlib.load_case("g1_readme_donor")
        og = go.graph_from_config(conf)
        _state["cau"] = (conf, og, EmuGraph(og, conf["loci_map"]))
    return _state["cau"]


def test_clean_up_gl_matches_oracle():
    for gl in ["A*01:01g+A*02:01L^B*07:02+B*08:01", "A*01:01+A*02:01^C*UUUU+C*UUUU", "UUUU", "",
               "A*UUUU+A*01:01^B*07:02+B*08:01", "A*01:01+A*UUUU", "gL", "A*01:01+A*02:01^^B*07:02+B*08:01"]:
        assert gi.clean_up_gl(gl) == go.clean_gl(gl)


def test_read_hpf_matches_oracle_full_nodes():
    conf, og, eg = _cau()
    cfg = load_config(conf)
    alleles, fa, ff = read_hpf(cfg["freq_file"], cfg["pops"], loci_in_order(cfg["loci_map"]),
                               cfg["freq_trim_threshold"], cfg["pops_count_file"])
    assert fa.shape[0] == og.n_full
    idx = range(0, og.n_full, 97)
    names = ["~".join(alleles[l][fa[i, l] - 1] for l in range(5)) for i in idx]
    assert names == [og.names[i] for i in idx]
    assert [float(ff[i, 0]) for i in idx] == [og.node[og.names[i]][1][0] for i in idx]


def test_loci_map_must_be_alphabetical():
    assert loci_in_order({"A": 1, "B": 2, "C": 3, "DQB1": 4, "DRB1": 5}) == ["A", "B", "C", "DQB1", "DRB1"]
    with pytest.raises(NotImplementedError):
        loci_in_order({"A": 1, "B": 3, "C": 2, "DQB1": 4, "DRB1": 5})
    with pytest.raises(ValueError):
        loci_in_order({"A": 1, "B": 3})


def test_key_layout_fits_and_leaves_room_for_unknown_alleles():
    n = [54, 85, 47, 16, 44]
    bits = key_layout(n)
    assert sum(bits) <= 63 and all((1 << b) - 1 > k for b, k in zip(bits, n))
    with pytest.raises(NotImplementedError):
        key_layout([40000] * 9)


class _Net:
    loci = ["A", "B", "C", "DQB1", "DRB1"]
    alleles = [[] for _ in range(5)]
    allele_id = [{} for _ in range(5)]
    key_bits = [12] * 5
    shift = [0, 12, 24, 36, 48]


@pytest.mark.parametrize("r1,r2", [("AAA", "BBB"), ("AAA;CCC", "BBB"), ("XXX", "CCC"), ("", ""),
                                   ("AAA", "XXX;YYY"), ("BBB", "BBB"), (None, None)])
@pytest.mark.parametrize("unk", ["MR", "SR"])
def test_prior_matrix_bit_identical_to_oracle(r1, r2, unk):
    _, conf, _, _ = goldenlib.load_case("g3_pop3_typed")
    conf = dict(conf)
    conf["UNK_priors"] = unk
    cbp = go.count_by_prob_from_file(3, conf["pops_count_file"])
    imp = gi.Imputation(_Net(), load_config(conf), cbp)
    mine = imp._priors[imp._prior_for(r1, r2)]
    o = go.OracleImputation(None, go.load_config(conf), cbp)
    o.impute_one("", r1, r2)       # empty GL: only the prior is computed
    assert np.array_equal(mine, np.asarray(o.M))


def test_line_classification_matches_reference_fixture():
    """problem / miss classification of the edge-case fixture, through the tokeniser."""
    table, conf, lines, exp = goldenlib.load_case("g2_edges")
    _, og, eg = _cau()
    imp = emu_imputation(eg, load_config(conf))
    out = {k: "".join(v) for k, v in imp.impute_lines(lines).items()}
    assert out["problem"] == exp["problem"]
    assert out["miss"] == exp["miss"]


def test_shard_ranges_partition_input():
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_config_defaults_match_reference():
    cfg = load_config({"populations": ["CAU"], "priority": {}, "graph_files_path": "x"})
    assert cfg["epsilon"] == 1e-3 and cfg["number_of_results"] == 1000 and cfg["number_of_pop_results"] == 100
    assert cfg["output_MUUG"] is True and cfg["output_haplotypes"] is False and cfg["planb"] is True
    assert cfg["factor_missing_data"] == 0.01 and cfg["number_of_options_threshold"] == 100000
    assert cfg["max_haplotypes_number_in_phase"] == 100 and cfg["UNK_priors"] == "MR" and cfg["save_mode"] is False
    assert len(cfg["matrix_planb"]) == 6 and cfg["full_loci"] == "12345"


def test_open_gl_string_matches_reference_fixture():
    """EM helper: Imputation.open_gl_string vs the fixture written by the unmodified
    reference (tests/golden/make_open_gl.py)."""
    import json
    import os
    import goldenlib
    from grim.imputation.impute import Imputation
    imp = object.__new__(Imputation)          # host-only helper: needs no tables
    cases = json.load(open(os.path.join(goldenlib.GOLD, "data", "open_gl_string.json")))
    assert len(cases) >= 7
    for c in cases:
        assert imp.open_gl_string(c["gl"], c["cutoff"]) == c["phases"], c["gl"]


def test_produce_hpf_matches_reference_output(tmp_path):
    """README step 1: produce_hpf on the example CAU.freqs.gz (fixture) must write the same hpf.csv and
    pop_counts_file.txt the reference's produce_hpf wrote (tests/golden/data/cau_*)."""
    import json
    import os
    import goldenlib
    from graph_generation import generate_hpf
    pkg = os.path.dirname(os.path.dirname(os.path.abspath(generate_hpf.__file__)))
    conf = json.load(open(os.path.join(pkg, "conf", "minimal-configuration.json")))
    assert conf == json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    d = str(tmp_path)
    conf["freq_data_dir"] = os.path.join(goldenlib.GOLD, "data")     # CAU.freqs.gz: the README example's input
    conf["graph_files_path"] = d + "/csv/"
    conf["freq_file"] = d + "/hpf.csv"
    conf["pops_count_file"] = d + "/pop_counts_file.txt"
    json.dump(conf, open(d + "/conf.json", "w"))
    generate_hpf.produce_hpf(d + "/conf.json")
    assert open(d + "/hpf.csv", newline="").read() == open(os.path.join(goldenlib.GOLD, "data", "cau_hpf.csv"), newline="").read()
    assert open(d + "/pop_counts_file.txt").read() == open(os.path.join(goldenlib.GOLD, "data", "cau_pop_counts.txt")).read()


# ---- Plan_A_Matrix: label restriction, subject filter, the Plan B guard
def test_plan_a_matrix_validation():
    from grim.run_impute_def import plan_a_label_masks
    masks, store = plan_a_label_masks([[1, 2, 3, 4, 5], [1, 2, 3], [4, 5], [2]], 5)
    assert masks == [31, 7, 24, 2]
    assert store == [31, 7, 24, 2, 1, 4, 8, 16]      # + the single-locus labels the matrix does not name
    for bad in ([[1, 2, 3], [1, 2, 3, 4, 5]],          # full label not first: vertex positions != node ids
                [[1, 2, 3, 4, 5], [3, 2]],             # not ascending
                [[1, 2, 3, 4, 5], [1, 2], [1, 2]],     # repeated
                [[1, 2, 3, 4, 5], [6]],                # no such locus
                [[1, 2, 3, 4, 5], []]):
        with pytest.raises(NotImplementedError):
            plan_a_label_masks(bad, 5)
        with pytest.raises(NotImplementedError):
            go.plan_a_labels(bad, "12345")
    with pytest.raises(ValueError):                    # the reference's graph load: np.vstack of no edges
        plan_a_label_masks([[1, 2, 3, 4, 5]], 5)
    with pytest.raises(ValueError):
        go.plan_a_labels([[1, 2, 3, 4, 5]], "12345")


def test_plan_b_under_a_matrix_is_refused_not_guessed():
    """A subject that leaves Plan A empty-handed while Plan B is on: the oracle, the numpy front end and the C++ text
    pipeline all refuse (the reference reads adjacencies of unrelated nodes there); with Plan B off the same subject
    is a .miss row everywhere."""
    from emu_backend import emu_impute_text
    _, conf, lines, exp = goldenlib.load_case("g8_plan_a_blocks_planb_off")
    miss_ids = {r.split(",")[1] for r in exp["miss"].splitlines()}
    assert miss_ids
    pick = [ln for ln in lines if ln.split(",")[0] in miss_ids][:3] + lines[:5]
    conf_on = dict(conf, planb=True)
    og = go.graph_from_config(conf_on)
    with pytest.raises(go.PlanBUnderMatrix):
        go.impute_file(conf_on, graph=og, lines=pick)
    eg = EmuGraph(og, conf["loci_map"])
    with pytest.raises(NotImplementedError):
        emu_imputation(eg, load_config(conf_on)).impute_lines(pick)
    with pytest.raises(NotImplementedError):
        emu_impute_text(emu_imputation(eg, load_config(conf_on)), eg, "".join(pick).encode("utf8"))
    off = emu_imputation(eg, load_config(conf)).impute_lines(pick)
    ref, _ = go.impute_file(conf, graph=og, lines=pick)
    assert "".join(off["miss"]) == ref["miss"] != ""


def test_graph_and_configuration_must_name_the_same_matrix():
    _, conf, _, _ = goldenlib.load_case("g8_plan_a_blocks")
    conf_c, og_c, eg_c = _cau()
    with pytest.raises(ValueError):
        emu_imputation(eg_c, load_config(conf))         # unrestricted graph, matrix in the configuration
    eg = EmuGraph(go.graph_from_config(conf), conf["loci_map"])
    with pytest.raises(ValueError):
        emu_imputation(eg, load_config(conf_c))


# ---- the per-subject seam (impute.py:1940-1983): un-merged lists / dicts in traversal order
from impute_one_check import check_impute_one as _check_impute_one  # noqa: E402


def test_impute_one_returns_the_reference_structures():
    import synth
    _, conf, lines, _ = goldenlib.load_case("g2_edges")
    assert _check_impute_one(conf, lines) >= 8
    _, conf3, lines3, _ = goldenlib.load_case("g3_pop3_messy")
    assert _check_impute_one(conf3, lines3[:25]) >= 15
    assert _check_impute_one(conf3, lines3[25:32], binary=[1, 0, 1, 0]) >= 3
    _, conf1, lines1, _ = goldenlib.load_case("g1_readme_donor")
    assert _check_impute_one(conf1, lines1[:1]) == 1          # README donor: 8400 un-merged pairs
