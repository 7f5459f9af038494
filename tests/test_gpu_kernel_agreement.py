"""GPU, BASELINE-size property: the warp-per-subject kernels (k_impute_fast, k_impute_typed) and the
general CTA-per-subject kernel (k_impute) are independent implementations of the same path.  The
general kernel is pinned against the reference on the golden cases; here both are run on large
seeded batches (2^20 single-population subjects = BASELINE config 2 at full size, 200 k subjects x 21
populations; sizes the CPU oracle cannot reach) and must produce byte-identical files."""
import json
import os

import numpy as np
import pytest

import goldenlib
import synth

pytestmark = pytest.mark.gpu


def _both(conf, hpf, cnt, lines, tmp_path, monkeypatch, wide=False):
    from grim.imputation.impute import Imputation
    from grim.imputation.networkx_graph import Graph
    from grim.run_impute_def import load_config
    if wide:
        monkeypatch.setenv("GRIMB_KEY_WORDS", "2")
    monkeypatch.setenv("GRIMB_HOST_EVENTS", "1")   # kernel timing events in the host-pointer path
    d = str(tmp_path)
    open(d + "/hpf.csv", "w").write(hpf)
    open(d + "/cnt.txt", "w").write(cnt)
    conf = dict(conf)
    conf["freq_file"], conf["pops_count_file"] = d + "/hpf.csv", d + "/cnt.txt"
    cfg = load_config(conf)
    g = Graph(cfg).build_graph()
    data = "".join(lines).encode("utf8")
    imp = Imputation(g, cfg)
    fast = imp.impute_text(data)
    eng = g.engine(imp.workspaces[0])
    handed = g.lib.grimb_engine_kernel_ms(eng, 3)
    assert handed < 0.05 * len(lines), "the warp kernels handed %d of %d subjects on" % (handed, len(lines))
    monkeypatch.setenv("GRIMB_FAST", "0")          # read when an engine is created
    imp2 = Imputation(g, cfg)
    imp2.workspaces = [w + 4096 for w in imp2.workspaces]   # new engines -> general kernel only
    slow = imp2.impute_text(data)
    eng2 = g.engine(imp2.workspaces[0])
    assert g.lib.grimb_engine_kernel_ms(eng2, 0) < 0 and g.lib.grimb_engine_kernel_ms(eng2, 2) < 0
    for k in goldenlib.KEYS:
        assert fast[k] == slow[k], k
    assert fast["umug"].count(b"\n") >= 0.9 * len(lines)
    g.close()


def _homozygous(tab, n, seed, races):
    rng = np.random.RandomState(seed)
    out = []
    for s in range(n):
        h = tab.haps[int(rng.choice(len(tab.haps), p=tab.p))]
        out.append("Z%d,%s,%s\n" % (s, "^".join("%s+%s" % (a, a) for a in h), races[s % len(races)]))
    return out


def test_typed_kernel_equals_general_kernel_21_populations(tmp_path, monkeypatch):
    base = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    cau = open(os.path.join(goldenlib.GOLD, "data", "cau_hpf.csv")).read()
    pops = ["P%02d" % i for i in range(21)]
    hpf, cnt = synth.multipop_hpf(cau, pops, 21)
    conf = dict(base)
    conf.update({"populations": pops, "UNK_priors": "MR", "number_of_pop_results": 100})
    tab = synth.Table(hpf, "P00")
    races = synth.race_fields(pops)
    lines = synth.typed_subjects(tab, 200000, 11, races) + _homozygous(tab, 1500, 12, races)
    _both(conf, hpf, cnt, lines, tmp_path, monkeypatch)


def test_typed_kernel_equals_general_kernel_wide_keys(tmp_path, monkeypatch):
    base = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    hpf = open(os.path.join(goldenlib.GOLD, "data", "pop3_hpf.csv")).read()
    cnt = open(os.path.join(goldenlib.GOLD, "data", "pop3_pop_counts.txt")).read()
    conf = dict(base)
    conf.update({"populations": ["AAA", "BBB", "CCC"], "UNK_priors": "SR"})
    tab = synth.Table(hpf, "AAA")
    races = synth.race_fields(conf["populations"])
    lines = synth.typed_subjects(tab, 30000, 13, races) + _homozygous(tab, 500, 14, races)
    _both(conf, hpf, cnt, lines, tmp_path, monkeypatch, wide=True)


def test_fast_kernel_equals_general_kernel_single_population(tmp_path, monkeypatch):
    base = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    cau = open(os.path.join(goldenlib.GOLD, "data", "cau_hpf.csv")).read()
    cnt = open(os.path.join(goldenlib.GOLD, "data", "cau_pop_counts.txt")).read()
    tab = synth.Table(cau)
    # BASELINE config 2 at full size: 2^20 subjects
    lines = synth.typed_subjects(tab, 1 << 20, 15, ["CAU,CAU"]) + _homozygous(tab, 2000, 16, ["CAU,CAU"])
    _both(dict(base), cau, cnt, lines, tmp_path, monkeypatch)
