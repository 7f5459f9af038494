"""GPU parity at the BASELINE.json configurations' own shapes: the product path
(text in -> grimb_impute_text -> six texts out) against the CPU oracle, byte for byte.

  C2  1M-haplotype Zipf table (bench.py's own table), 20,000 subjects incl. homozygous, unknown-allele and
      recombinant ones: k_fast_probe + k_fast_score, their overflow list and the hand-over to k_impute
  C3  the same 1M haplotypes x 21 populations (168-byte frequency vectors), race-field shapes, top-100
      population rows: k_impute_typed and its hand-over; messy subjects on the big and on the README table
  P   32 and 33 populations: the lane-mapping edge of k_impute_typed (P <= 32) and the general kernel beyond
  C4  >= 200 highly ambiguous subjects straddling the 100,000-option threshold, full Plan-B matrix
  C5  nine loci, 128-bit keys, a table large enough that Plan A is not the only plan

The oracle store for the big tables is oracle/oracle_graph_np.py (pinned against OracleGraph by
tests/test_oracle_graph_np.py); the oracle runs over forked workers (tests/oracle_par.py).  Sizes can be
scaled with GRIMB_TEST_SCALE (default 1)."""
import json
import os

import numpy as np
import pytest

import goldenlib
import grim_oracle as go
import oracle_par
import synth
from oracle_graph_np import NumpyOracleGraph

pytestmark = pytest.mark.gpu

SCALE = float(os.environ.get("GRIMB_TEST_SCALE", "1"))
LM5 = {"A": 1, "B": 2, "C": 3, "DQB1": 4, "DRB1": 5}
_cache = {}


def _base():
    return json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))


def _big_table():
    if "big" not in _cache:
        import bench
        _cache["big"] = bench.make_table(1000000)
    return _cache["big"]


def _gpu_graph(conf, names, fa, ff):
    from grim.imputation.networkx_graph import Graph
    from grim.run_impute_def import load_config
    cfg = load_config(conf)
    return Graph(cfg).from_arrays(names, fa, ff), cfg


def _gpu_texts(g, cfg, lines, cbp):
    from grim.imputation.impute import Imputation
    imp = Imputation(g, cfg, cbp)
    out = imp.impute_text_stream("".join(lines).encode("utf8"))
    return {k: v.decode("utf8") for k, v in out.items()}, imp


def _compare(mine, ref, what):
    for k in oracle_par.KEYS:
        if mine[k] != ref[k]:
            a, b = mine[k].splitlines(), ref[k].splitlines()
            i = next((j for j in range(min(len(a), len(b))) if a[j] != b[j]), min(len(a), len(b)))
            raise AssertionError("%s: %s differs at row %d: %r vs oracle %r (%d / %d rows)" % (
                what, k, i, a[i] if i < len(a) else None, b[i] if i < len(b) else None, len(a), len(b)))


def test_c2_one_million_haplotype_table():
    names, fa, ff = _big_table()
    conf = _base()
    conf.update({"populations": ["CAU"], "loci_map": LM5, "number_of_results": 10})
    n = int(20000 * SCALE)
    p = ff[:, 0] / ff[:, 0].sum()
    lines = synth.array_subject_lines(names, fa, p, n, 11, ["CAU,CAU"])
    g, cfg = _gpu_graph(conf, names, fa, ff)
    try:
        mine, imp = _gpu_texts(g, cfg, lines, np.ones(1))
    finally:
        g.close()
    og = NumpyOracleGraph(names, fa, ff, ["CAU"], LM5)
    ref, evals = oracle_par.oracle_texts(og, conf, lines, np.ones(1))
    _compare(mine, ref, "C2")
    assert imp.stats["pair_evals"] == evals
    assert imp.stats["plan"][2] > 0 and imp.stats["plan"][1] > n * 0.8     # both the warp kernels and k_impute ran
    assert mine["umug"].count("\n") >= n * 0.95


def test_c3_twenty_one_populations_full_table():
    names, fa, ff1 = _big_table()
    pops = ["P%02d" % i for i in range(21)]
    ff = synth.multipop_freqs(ff1[:, 0], 21, 21)
    ctext, ratio = synth.pop_counts(pops)
    conf = _base()
    conf.update({"populations": pops, "loci_map": LM5, "UNK_priors": "MR", "number_of_pop_results": 100,
                 "number_of_results": 10})
    races = synth.race_fields(pops)
    p = ff.mean(axis=1)
    p = p / p.sum()
    lines = synth.array_subject_lines(names, fa, p, int(2000 * SCALE), 3, races)
    # ambiguous / unknown-allele subjects without missing loci (a missing locus on a table of this size
    # sends the oracle through whole-label scans of ~10^5 nodes)
    at = synth.ArrayTable(names, fa, p)
    lines += synth.messy_subjects(at, int(150 * SCALE), 5, max_amb=3, p_missing=0.0, races=races)
    g, cfg = _gpu_graph(conf, names, fa, ff)
    try:
        mine, imp = _gpu_texts(g, cfg, lines, ratio)
    finally:
        g.close()
    og = NumpyOracleGraph(names, fa, ff, pops, LM5)
    ref, evals = oracle_par.oracle_texts(og, conf, lines, ratio)
    _compare(mine, ref, "C3 (1M haplotypes x 21 populations)")
    assert imp.stats["pair_evals"] == evals
    assert max(int(r.split(",")[-1]) for r in mine["umug_pops"].splitlines()) > 10    # long population lists


@pytest.mark.parametrize("n_pops", [21, 32, 33])
def test_many_populations_on_the_readme_table(n_pops):
    """P = 32 is the last population count k_impute_typed serves (lane = population), P = 33 the first one
    the general kernel takes whole; 21 adds missing-locus subjects to C3."""
    cau = open(os.path.join(goldenlib.GOLD, "data", "cau_hpf.csv")).read()
    pops = ["P%02d" % i for i in range(n_pops)]
    hpf, ctext = synth.multipop_hpf(cau, pops, 100 + n_pops)
    conf = _base()
    conf.update({"populations": pops, "UNK_priors": "MR", "number_of_pop_results": 100})
    tab = synth.Table(hpf, "P00")
    races = synth.race_fields(pops)
    lines = synth.typed_subjects(tab, int(600 * SCALE), 3, races)
    lines += synth.messy_subjects(tab, int((200 if n_pops == 21 else 60) * SCALE), 6, races=races)
    og = NumpyOracleGraph.from_hpf(hpf.splitlines(), pops, conf["loci_map"], conf["freq_trim_threshold"],
                                   ctext.splitlines())
    ratio = np.array([float(l.split(",")[2]) for l in ctext.splitlines()])
    g, cfg = _gpu_graph(conf, og.allele_names, og.fa, og.ff)
    try:
        mine, imp = _gpu_texts(g, cfg, lines, ratio)
    finally:
        g.close()
    ref, evals = oracle_par.oracle_texts(og, conf, lines, ratio)
    _compare(mine, ref, "P = %d" % n_pops)
    assert imp.stats["pair_evals"] == evals


def test_c4_heavy_subjects_straddling_the_options_threshold():
    _t, conf, _l, _e = goldenlib.load_case("g1_readme_donor")
    tab = synth.Table(open(conf["freq_file"]).read())
    n = int(208 * SCALE)
    lines = synth.heavy_subjects(tab, n, 44, races=["CAU,CAU"])
    # the generator's list sizes put subjects on both sides of number_of_options_threshold = 100000
    over = 0
    for ln in lines:
        gl = ln.split(",")[1]
        prod = 1
        for loc in gl.split("^"):
            prod *= len(loc.split("+")[0].split("/"))
        over += prod >= 100000
    assert 0.15 * n < over < 0.85 * n
    from grim.imputation.networkx_graph import Graph
    from grim.run_impute_def import load_config
    cfg = load_config(conf)
    g = Graph(cfg).build_graph()
    try:
        mine, imp = _gpu_texts(g, cfg, lines, None)
    finally:
        g.close()
    og = go.graph_from_config(conf)
    ref, evals = oracle_par.oracle_texts(og, conf, lines, go.count_by_prob_from_file(1, conf["pops_count_file"]))
    _compare(mine, ref, "C4 heavy")
    assert imp.stats["pair_evals"] == evals


def test_c5_nine_loci_wide_keys_beyond_plan_a():
    loci = ["A", "B", "C", "DPA1", "DPB1", "DQA1", "DQB1", "DRB1", "DRBX"]
    n_all = [700, 1200, 600, 40, 300, 60, 250, 700, 100]
    pops = ["Q%d" % i for i in range(5)]
    lm = {l: i + 1 for i, l in enumerate(loci)}
    names, fa, base_f = synth.zipf_arrays(int(30000 * SCALE), n_all, 20261018, loci)
    ff = synth.multipop_freqs(base_f, 5, 9, zero_frac=0.2)
    ctext, ratio = synth.pop_counts(pops)
    conf = _base()
    conf.update({"populations": pops, "UNK_priors": "MR", "number_of_pop_results": 100, "loci_map": lm,
                 "Plan_B_Matrix": [[[1, 2, 3, 4, 5, 6, 7, 8, 9]], [[1, 2, 3], [4, 5], [6, 7, 8, 9]],
                                   [[1], [2, 3], [4, 5], [6, 7], [8, 9]],
                                   [[1], [2], [3], [4], [5], [6], [7], [8], [9]]]})
    p = ff.mean(axis=1)
    p = p / p.sum()
    lines = synth.array_subject_lines(names, fa, p, int(1200 * SCALE), 9, synth.race_fields(pops))
    g, cfg = _gpu_graph(conf, names, fa, ff)
    try:
        assert g.kw == 2                                          # the 128-bit-key build serves this table
        mine, imp = _gpu_texts(g, cfg, lines, ratio)
    finally:
        g.close()
    og = NumpyOracleGraph(names, fa, ff, pops, lm)
    assert og.n_words == 2
    ref, evals = oracle_par.oracle_texts(og, conf, lines, ratio)
    _compare(mine, ref, "C5")
    assert imp.stats["pair_evals"] == evals
    assert imp.stats["plan"][2] + imp.stats["plan"][3] > 0 and imp.stats["plan"][1] > 0


def test_c4_wide_subjects_through_the_cooperative_slot_pass():
    """Fully typed subjects with 7-9 alleles per locus side: Cartesian products of tens of thousands of candidates
    per phase and side, below the options threshold -- the subjects the cooperative slot pass (k_impute mode 1)
    spreads over one CTA per (phase, side).  Same files as the oracle, with the pass on and off."""
    _t, conf, _l, _e = goldenlib.load_case("g1_readme_donor")
    tab = synth.Table(open(conf["freq_file"]).read())
    lines = synth.wide_subjects(tab, int(16 * SCALE), 45, races=["CAU,CAU"])
    from grim.imputation.networkx_graph import Graph
    from grim.run_impute_def import load_config
    cfg = load_config(conf)
    og = go.graph_from_config(conf)
    ref, evals = oracle_par.oracle_texts(og, conf, lines, go.count_by_prob_from_file(1, conf["pops_count_file"]), warm=0)
    for group in ("4096", "0"):
        os.environ["GRIMB_GROUP_SUBJECTS"] = group      # read when an engine is created
        try:
            g = Graph(cfg).build_graph()
            try:
                mine, imp = _gpu_texts(g, cfg, lines, None)
                eng = g.engine(imp.workspaces[0])
            finally:
                g.close()
        finally:
            os.environ.pop("GRIMB_GROUP_SUBJECTS", None)
        _compare(mine, ref, "C4 wide (slot pass %s)" % ("on" if group != "0" else "off"))
        assert imp.stats["pair_evals"] == evals
