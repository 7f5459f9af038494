"""GPU parity tests (B200): the product path -- Python host -> ctypes -> libgrimb200.so ->
CUDA kernels -- against (a) the golden files written by the unmodified reference and (b) the
CPU oracle on seeded synthetic inputs.  Integer work (enumeration, classification, top-N
membership and order) must be bit-exact; probabilities are compared as printed text, i.e.
exactly (the north star allows 1e-9 relative; we hold 0)."""
import numpy as np
import pytest

import goldenlib
import grim_oracle as go
import synth

pytestmark = pytest.mark.gpu

_graphs = {}
_oracles = {}


def _graph(table, conf):
    from grim.imputation.networkx_graph import Graph
    from grim.run_impute_def import load_config
    if table not in _graphs:
        _evict_matrix_graphs(_graphs, table)
        _graphs[table] = Graph(load_config(conf)).build_graph()
    return _graphs[table]


def _evict_matrix_graphs(cache, table):
    """A cached Graph keeps its engines' workspaces (GBs of HBM): of the Plan_A_Matrix graphs, used by one or two
    cases each, only the one in use stays."""
    if "+matrix:" in table:
        for k in [k for k in cache if "+matrix:" in k]:
            cache.pop(k).close()


def _oracle_graph(table, conf):
    if table not in _oracles:
        _oracles[table] = go.graph_from_config(conf)
    return _oracles[table]


def _run_gpu(table, conf, lines):
    from grim.imputation.impute import Imputation
    from grim.run_impute_def import load_config
    cfg = load_config(conf)
    imp = Imputation(_graph(table, conf), cfg)
    files = imp.impute_lines(lines, em_mr=conf.get("_hap_pop_pair", False))
    return {k: "".join(v) for k, v in files.items()}, imp


@pytest.mark.parametrize("table", ["cau", "pop3"])
def test_device_table_build_matches_oracle_graph(table):
    """K0: node ids, keys, sequential marginal sums, top links, connectors (incl. trap T1)."""
    from emu_backend import arrays_from_oracle
    name = {"cau": "g1_readme_donor", "pop3": "g3_pop3_typed"}[table]
    _, conf, _, _ = goldenlib.load_case(name)
    g = _graph(table, conf)
    og = _oracle_graph(table, conf)
    want = arrays_from_oracle(og, g.loci)
    got = g.export()
    assert g.alleles == want["alleles"]
    assert g.key_bits == want["bits"]
    n = og.n_nodes
    assert g.info()["n_nodes"] == n
    assert np.array_equal(got["node_key"], want["node_key"])
    assert np.array_equal(got["node_freq"], want["freq"])          # bit-exact FP64 sums
    assert np.array_equal(got["label_first"], want["label_first"])
    assert np.array_equal(got["label_count"], want["label_count"])
    assert np.array_equal(got["tl_cnt"], want["tl_cnt"])
    for i in range(og.n_full, n, max(1, n // 4000)):
        a = got["tl_adj"][got["tl_start"][i]: got["tl_start"][i] + got["tl_cnt"][i]]
        b = want["tl_adj"][want["tl_start"][i]: want["tl_start"][i] + want["tl_cnt"][i]]
        assert np.array_equal(a, b)
    i = n - 1  # the last node (trap T1)
    assert np.array_equal(got["tl_adj"][got["tl_start"][i]: got["tl_start"][i] + got["tl_cnt"][i]],
                          want["tl_adj"][want["tl_start"][i]: want["tl_start"][i] + want["tl_cnt"][i]])
    assert np.array_equal(got["cn_cnt"], want["cn_cnt"])
    L = len(g.loci)
    for i in list(range(og.n_full, n, max(1, n // 4000))) + [n - 1]:
        for l in range(L):
            c = int(got["cn_cnt"][i, l])
            if c and c != 0xFFFFFFFF:
                a = got["cn_adj"][got["cn_start"][i, l]: got["cn_start"][i, l] + c]
                b = want["cn_adj"][want["cn_start"][i, l]: want["cn_start"][i, l] + c]
                assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["g8_plan_a_blocks", "g8_plan_a_four_pop3", "g8_plan_a_nine"])
def test_device_table_build_with_a_plan_a_matrix(name):
    """K0 over a label list (Plan_A_Matrix): node ids in matrix order, sums, top links, the CSR sentinel on the last
    Plan-A node, no connectors."""
    from emu_backend import arrays_from_oracle
    table, conf, _, _ = goldenlib.load_case(name)
    g = _graph(table, conf)
    og = _oracle_graph(table, conf)
    want = arrays_from_oracle(og, g.loci)
    got = g.export()
    n = og.n_nodes
    assert g.info()["n_nodes"] == n and g.info()["n_conn_edges"] == 0
    assert np.array_equal(got["node_key"], want["node_key"])
    assert np.array_equal(got["node_freq"], want["freq"])
    assert np.array_equal(got["label_first"], want["label_first"])
    assert np.array_equal(got["label_count"], want["label_count"])
    assert np.array_equal(got["tl_cnt"], want["tl_cnt"])
    assert not got["cn_cnt"].any()
    for i in list(range(og.n_full, n, max(1, n // 4000))) + [og.n_plan_a_nodes - 1, n - 1]:
        c = int(got["tl_cnt"][i])
        if c != 0xFFFFFFFF:
            assert np.array_equal(got["tl_adj"][got["tl_start"][i]: got["tl_start"][i] + c],
                                  want["tl_adj"][want["tl_start"][i]: want["tl_start"][i] + c])


@pytest.mark.parametrize("name", goldenlib.case_names())
def test_cuda_path_matches_reference_files(name):
    table, conf, lines, exp = goldenlib.load_case(name)
    out, _ = _run_gpu(table, conf, lines)
    for k in goldenlib.KEYS:
        assert out[k] == exp[k], "%s: %s differs" % (name, k)


@pytest.mark.parametrize("seed,kind", [(101, "typed"), (102, "messy"), (103, "typed_pop3"), (104, "messy_pop3")])
def test_cuda_path_matches_oracle_on_seeded_inputs(seed, kind):
    table = "pop3" if kind.endswith("pop3") else "cau"
    name = {"cau": "g1_readme_donor", "pop3": "g3_pop3_typed"}[table]
    _, conf, _, _ = goldenlib.load_case(name)
    hpf = open(conf["freq_file"]).read()
    pops = conf["populations"]
    tab = synth.Table(hpf, pops[0])
    races = synth.race_fields(pops) if table == "pop3" else ["CAU,CAU"]
    if kind.startswith("typed"):
        lines = synth.typed_subjects(tab, 3000, seed, races)
    else:
        lines = synth.messy_subjects(tab, 150, seed, races=races)
    out, imp = _run_gpu(table, conf, lines)
    ref, _ = go.impute_file(conf, graph=_oracle_graph(table, conf), lines=lines)
    for k in goldenlib.KEYS:
        assert out[k] == ref[k], "%s differs" % k


def test_pair_eval_counter_matches_oracle():
    _, conf, _, _ = goldenlib.load_case("g1_readme_donor")
    tab = synth.Table(open(conf["freq_file"]).read())
    lines = synth.typed_subjects(tab, 500, 7, ["CAU,CAU"])
    out, imp = _run_gpu("cau", conf, lines)
    o = go.OracleImputation(_oracle_graph("cau", conf), go.load_config(conf),
                            go.count_by_prob_from_file(1, conf["pops_count_file"]))
    o.impute_lines(lines)
    assert imp.stats["pair_evals"] == o.pair_evals


def test_full_size_properties():
    """2^17 synthetic subjects (config 2 shape): size-independent properties -- every subject
    classified, UMUG of a fully typed unambiguous subject is its own genotype with rank 0,
    probabilities positive and non-increasing with rank, results independent of batch split."""
    _, conf, _, _ = goldenlib.load_case("g1_readme_donor")
    tab = synth.Table(open(conf["freq_file"]).read())
    n = 1 << 17
    lines = synth.typed_subjects(tab, n, 11, ["CAU,CAU"])
    out, imp = _run_gpu("cau", conf, lines)
    assert out["problem"] == "" and out["miss"] == ""
    rows = out["umug"].splitlines()
    assert len(rows) == n
    for r, ln in zip(rows[:: n // 512], lines[:: n // 512]):
        sid, geno, prob, rank = r.split(",")
        assert sid == ln.split(",")[0] and rank == "0" and float(prob) > 0
        want = "^".join("+".join(sorted(x.split("+"))) for x in ln.split(",")[1].split("^"))
        assert geno == want
    last = {}
    for r in out["pmug"].splitlines():
        sid, _, prob, rank = r.split(",")
        if rank != "0":
            assert float(prob) <= last[sid]
        last[sid] = float(prob)
    # idempotence / batch-split independence on a slice
    imp2_out, _ = _run_gpu("cau", conf, lines[:5000])
    assert imp2_out["umug"] == "".join(x + "\n" for x in rows[:5000])


@pytest.mark.parametrize("name", goldenlib.text_case_names())
def test_native_text_path_matches_reference_files(name):
    """Same cases through grimb_impute_text (C++ tokeniser / formatter around the kernels)."""
    from grim.imputation.impute import Imputation
    from grim.run_impute_def import load_config
    table, conf, lines, exp = goldenlib.load_case(name)
    imp = Imputation(_graph(table, conf), load_config(conf))
    out = imp.impute_text("".join(lines).encode("utf8"))
    for k in goldenlib.KEYS:
        assert out[k].decode("utf8") == exp[k], "%s: %s differs" % (name, k)


def test_public_api_impute_writes_reference_files(tmp_path):
    """grim.grim.impute(conf_file) end to end: JSON config in, six files out (README flow)."""
    import json
    from grim import grim
    table, conf, lines, exp = goldenlib.load_case("g2_edges")
    conf = dict(conf)
    d = str(tmp_path)
    open(d + "/subjects.csv", "w").writelines(lines)
    conf["imputation_in_file"] = d + "/subjects.csv"
    conf["imputation_out_path"] = d + "/out"
    names = {"umug": "imputation_out_umug_freq_filename", "umug_pops": "imputation_out_umug_pops_filename",
             "pmug": "imputation_out_hap_freq_filename", "pmug_pops": "imputation_out_hap_pops_filename",
             "miss": "imputation_out_miss_filename", "problem": "imputation_out_problem_filename"}
    for k, ck in names.items():
        conf[ck] = "x." + k
    json.dump(conf, open(d + "/conf.json", "w"))
    g = grim.impute(conf_file=d + "/conf.json")
    for k in names:
        assert open(d + "/out/x." + k).read() == exp[k], k
    g2 = grim.impute(conf_file=d + "/conf.json", graph=g)     # graph reuse, as in the reference
    assert g2 is g


def test_native_text_path_large_batch_equals_python_host_path():
    _, conf, _, _ = goldenlib.load_case("g3_pop3_typed")
    hpf = open(conf["freq_file"]).read()
    tab = synth.Table(hpf, "AAA")
    lines = synth.typed_subjects(tab, 20000, 77, synth.race_fields(conf["populations"])) + \
        synth.messy_subjects(tab, 300, 78, races=synth.race_fields(conf["populations"]))
    out_py, imp = _run_gpu("pop3", conf, lines)
    out_native = imp.impute_text("".join(lines).encode("utf8"))
    for k in goldenlib.KEYS:
        assert out_native[k].decode("utf8") == out_py[k], k


def test_homozygous_subjects_multi_population_match_oracle():
    """Both haplotypes equal: one phase, identical side lists, geno_seen drops the mirrored
    (population-swapped) pair (impute.py:508-513) -- the typed warp kernel's `same` path."""
    _, conf, _, _ = goldenlib.load_case("g3_pop3_typed")
    pops = conf["populations"]
    tab = synth.Table(open(conf["freq_file"]).read(), pops[0])
    races = synth.race_fields(pops)
    rng = np.random.RandomState(5)
    lines = []
    for s in range(400):
        h = tab.haps[int(rng.choice(len(tab.haps), p=tab.p))]
        g = list(h)
        if s % 3 == 1:      # one heterozygous locus for contrast
            g2 = list(tab.haps[int(rng.choice(len(tab.haps), p=tab.p))])
            gl = "^".join("%s+%s" % (a, b if l == 2 else a) for l, (a, b) in enumerate(zip(g, g2)))
        else:
            gl = "^".join("%s+%s" % (a, a) for a in g)
        lines.append("H%d,%s,%s\n" % (s, gl, races[s % len(races)]))
    out, imp = _run_gpu("pop3", conf, lines)
    ref, _ = go.impute_file(conf, graph=_oracle_graph("pop3", conf), lines=lines)
    for k in goldenlib.KEYS:
        assert out[k] == ref[k], "%s differs" % k


def test_readme_flow_with_packaged_defaults(tmp_path, monkeypatch):
    """BASELINE config 1, as the reference's README runs it: produce_hpf -> graph_freqs -> impute with
    the packaged minimal configuration, relative paths resolved against the working directory."""
    import json
    import os
    import shutil
    from graph_generation import generate_hpf
    from grim import grim
    pkg = os.path.dirname(os.path.dirname(os.path.abspath(generate_hpf.__file__)))
    d = str(tmp_path)
    os.makedirs(d + "/data/freqs")
    os.makedirs(d + "/data/subjects")
    shutil.copy(os.path.join(goldenlib.GOLD, "data", "CAU.freqs.gz"), d + "/data/freqs/CAU.freqs.gz")
    shutil.copy(os.path.join(goldenlib.GOLD, "data", "donor.csv"), d + "/data/subjects/donor.csv")
    shutil.copytree(os.path.join(pkg, "conf"), d + "/conf")
    monkeypatch.chdir(d)
    generate_hpf.produce_hpf("conf/minimal-configuration.json")
    g = grim.graph_freqs(conf_file="conf/minimal-configuration.json")
    grim.impute(conf_file="conf/minimal-configuration.json", graph=g)
    _, _, _, exp = goldenlib.load_case("g1_readme_donor")
    names = {"umug": "don.umug", "umug_pops": "don.umug.pops", "pmug": "don.pmug", "pmug_pops": "don.pmug.pops",
             "miss": "don.miss", "problem": "don.problem"}
    for k, fn in names.items():
        assert open(os.path.join(d, "output", fn)).read() == exp[k], k
    g.close()


@pytest.mark.parametrize("mode", ["masks", "hap_pop_pair", "both"])
def test_em_modes_match_oracle_on_seeded_inputs(mode, tmp_path):
    """SURVEY 8(f)-4 modes beyond the golden cases: random per-subject phase masks and / or the
    hap_pop_pair output on seeded typed + messy subjects (3 populations), CUDA path vs oracle."""
    import json
    from grim.imputation.impute import Imputation
    from grim.run_impute_def import load_config
    _, conf, _, _ = goldenlib.load_case("g3_pop3_typed")
    conf = dict(conf)
    pops = conf["populations"]
    tab = synth.Table(open(conf["freq_file"]).read(), pops[0])
    races = synth.race_fields(pops)
    lines = synth.typed_subjects(tab, 200, 301, races) + synth.messy_subjects(tab, 120, 302, max_amb=3, races=races)
    if mode in ("masks", "both"):
        rng = np.random.RandomState(303)
        masks = {ln.split(",")[0]: [int(x) for x in rng.randint(0, 2, size=4)] for ln in lines}
        path = str(tmp_path / "masks.json")
        json.dump(masks, open(path, "w"))
        conf["bin_imputation_in_file"] = path
    em_mr = mode in ("hap_pop_pair", "both")
    imp = Imputation(_graph("pop3", conf), load_config(conf))
    out = {k: "".join(v) for k, v in imp.impute_lines(lines, em_mr=em_mr).items()}
    ref, _ = go.impute_file(conf, graph=_oracle_graph("pop3", conf), lines=lines, em_mr=em_mr)
    for k in goldenlib.KEYS:
        assert out[k] == ref[k], "%s differs" % k


@pytest.mark.parametrize("name", ["g5_messy_cau", "g2_edges", "g3_pop3_typed"])
def test_file_pipeline_many_chunks_files_and_memory(name, tmp_path):
    """grimb_impute_file: the input cut into many small chunks (tokenise | GPU | format | write overlapped on
    host threads), outputs streamed to files and, separately, returned in memory for a byte range."""
    import ctypes as C
    from grim.imputation import _lib
    from grim.imputation.impute import Imputation
    from grim.run_impute_def import load_config
    table, conf, lines, exp = goldenlib.load_case(name)
    cfg = load_config(conf)
    imp = Imputation(_graph(table, conf), cfg)
    src = str(tmp_path / "in.csv")
    open(src, "w").write("".join(lines))
    paths = {k: str(tmp_path / ("out." + k)) for k in goldenlib.KEYS}
    _out, st = imp.impute_file_native(src, paths, chunk_bytes=300)
    assert st.n_chunks > 3 and st.n_lines == len(lines)
    for k in goldenlib.KEYS:
        assert open(paths[k]).read() == exp[k], "%s: %s differs" % (name, k)
    # two byte ranges, kept in memory, global line indices: concatenation == whole file
    size = len("".join(lines).encode("utf8"))
    cut = size // 2
    lib = imp.netGraph.lib
    n0 = C.c_int64()
    _lib.check(lib.grimb_file_count_lines(src.encode(), 0, cut, 0, C.byref(n0), None, None), "count")
    parts = []
    for lo, hi, first in ((0, cut, 0), (cut, -1, n0.value)):
        out, _ = imp.impute_file_native(src, None, lo, hi, first, chunk_bytes=700)
        parts.append({k: C.string_at(out.data[i], out.size[i]).decode("utf8") for i, k in enumerate(_lib.OUT_KEYS)})
    for k in goldenlib.KEYS:
        assert parts[0][k] + parts[1][k] == exp[k], "%s: %s differs (byte ranges)" % (name, k)


def test_impute_one_seam_on_the_gpu():
    """Imputation.impute_one (the reference's per-subject seam): traversal-order dicts and un-merged lists from the
    general kernel's encounter-order mode, against the oracle."""
    from impute_one_check import check_impute_one
    from grim.imputation.impute import Imputation
    from grim.run_impute_def import load_config
    for name, sl, binary in (("g2_edges", slice(None), None), ("g3_pop3_messy", slice(0, 40), None),
                             ("g3_pop3_messy", slice(40, 50), [0, 1, 1, 0]), ("g1_readme_donor", slice(0, 1), None)):
        table, conf, lines, _ = goldenlib.load_case(name)
        imp = Imputation(_graph(table, conf), load_config(conf))
        assert check_impute_one(conf, lines[sl], binary, imp=imp, og=_oracle_graph(table, conf)) >= 1
