"""GPU parity on RANDOM small frequency tables: the CUDA path against the CPU oracle, byte for
byte, on tables unlike the README one -- a few alleles per locus shared by most haplotypes, so
recombinant phases hit (1-16 candidate phases: hand-over records, the overflow list, the typed
kernel's phase cap), top-link lists are long and the top-K cap binds.  One population runs the
split fast path (k_fast_probe + k_fast_score), several populations the typed warp kernel, messy
subjects the general kernel.  The same generator, run three-way against the unmodified
reference on the CPU (tests/golden/fuzz_random_tables.py), pins the oracle on these shapes.

As a script it runs a longer campaign:  python tests/test_gpu_random_tables.py [--nine] [n_tables] [n_subjects] [seed]
(GRIMB_KEY_WORDS=2 in the environment forces the 128-bit-key build)
"""
import json
import os
import sys
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
if __name__ == "__main__":
    for p in (os.path.join(HERE, "..", "py-graph-imputation_b200"), os.path.join(HERE, "..", "oracle"), HERE):
        sys.path.insert(0, p)

import goldenlib  # noqa: E402
import grim_oracle as go  # noqa: E402
import synth  # noqa: E402

pytestmark = pytest.mark.gpu


def random_table(rng, t):
    n_full = int(rng.choice([30, 120, 500]))
    # every other table is dense: 2-5 alleles per locus, so most recombinants exist in the table
    n_alleles = [int(x) for x in (rng.randint(2, 6, size=5) if t % 2 == 0 else rng.randint(3, 13, size=5))]
    pops = [["CAU"], ["AAA", "BBB"], ["AAA", "BBB", "CCC", "DDD"]][t % 3]
    tseed = int(rng.randint(1, 1 << 30))
    hpf = synth.zipf_table(n_full, n_alleles, tseed, pops=tuple(pops))
    cnt = 1000.0 / np.arange(1, len(pops) + 1) ** 1.1
    counts = "".join("%s,%s,%s\n" % (p, repr(float(c)), repr(float(c / cnt.sum()))) for p, c in zip(pops, cnt))
    return hpf, counts, pops, tseed, n_alleles


def cases_for(tab, n, tseed, races):
    kw = {"races": races} if races else {}
    return [
        ("typed", synth.typed_subjects(tab, 4 * n, tseed + 1, races), {}),
        ("typed nres=2 npop=1", synth.typed_subjects(tab, 2 * n, tseed + 5, races),
         {"number_of_results": 2, "number_of_pop_results": 1}),
        ("messy", synth.messy_subjects(tab, n, tseed + 2, **kw), {}),
        ("messy thr=40 topk=7", synth.messy_subjects(tab, n, tseed + 3, max_amb=5, **kw),
         {"number_of_options_threshold": 40, "max_haplotypes_number_in_phase": 7}),
        ("unknown heavy nres=3", synth.messy_subjects(tab, n, tseed + 4, p_unknown=0.3, p_random=0.3, **kw),
         {"number_of_results": 3, "number_of_pop_results": 2}),
    ]


LOCI9 = ["A", "B", "C", "DPA1", "DPB1", "DQA1", "DQB1", "DRB1", "DRBX"]
NINE_OVER = {
    "populations": ["AAA", "BBB"], "UNK_priors": "MR", "freq_trim_threshold": 1e-9,
    "loci_map": {l: i + 1 for i, l in enumerate(LOCI9)},
    "Plan_B_Matrix": [[[1, 2, 3, 4, 5, 6, 7, 8, 9]], [[1, 2, 3], [4, 5], [6, 7, 8, 9]],
                      [[1], [2, 3], [4, 5], [6, 7], [8, 9]], [[1], [2], [3], [4], [5], [6], [7], [8], [9]]],
}


def random_table_nine(rng, t):
    """Nine loci (BASELINE config 5 shape: 256 phases, 510 marginal labels, 9-block Plan-B matrix)."""
    n_full = int(rng.choice([40, 100, 200]))
    n_alleles = [int(x) for x in (rng.randint(2, 5, size=9) if t % 2 == 0 else rng.randint(3, 15, size=9))]
    tseed = int(rng.randint(1, 1 << 30))
    pops = NINE_OVER["populations"]
    hpf = synth.zipf_table(n_full, n_alleles, tseed, loci=LOCI9, pops=tuple(pops))
    return hpf, "AAA,100.0,0.5\nBBB,100.0,0.5\n", pops, tseed, n_alleles


def cases_nine(tab, n, tseed):
    races = ["AAA,BBB", ",", "AAA;BBB,XXX"]
    return [
        ("typed", synth.typed_subjects(tab, 3 * n, tseed + 1, races), {}),
        ("messy", synth.messy_subjects(tab, n, tseed + 2, max_amb=2, p_missing=0.3, races=races[:2]), {}),
        ("messy save_space nres=3", synth.messy_subjects(tab, n, tseed + 3, max_amb=2, p_missing=0.4, races=races[:2]),
         {"save_space_mode": True, "number_of_results": 3}),
    ]


def run_table(t, seed, n, d, verbose=False, nine=False):
    from grim.imputation.impute import Imputation
    from grim.imputation.networkx_graph import Graph
    from grim.run_impute_def import load_config
    rng = np.random.RandomState((seed * 1000 + t) % (1 << 32))
    hpf, counts, pops, tseed, n_alleles = random_table_nine(rng, t) if nine else random_table(rng, t)
    open(os.path.join(d, "hpf.csv"), "w").write(hpf)
    open(os.path.join(d, "cnt.txt"), "w").write(counts)
    conf = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    conf.update({"freq_file": os.path.join(d, "hpf.csv"), "pops_count_file": os.path.join(d, "cnt.txt"),
                 "populations": pops, "UNK_priors": "MR"})
    if nine:
        conf.update(NINE_OVER)
    tab = synth.Table(hpf, pops[0], loci=LOCI9 if nine else synth.LOCI5)
    races = synth.race_fields(pops) if len(pops) > 1 else None
    g = Graph(load_config(conf)).build_graph()
    bad = []
    try:
        for tag, lines, over in (cases_nine(tab, n, tseed) if nine else cases_for(tab, n, tseed, races)):
            c = dict(conf)
            c.update(over)
            imp = Imputation(g, load_config(c))
            out = imp.impute_text("".join(lines).encode("utf8"))
            ref, _ = go.impute_file(c, lines=lines)
            diff = [k for k in goldenlib.KEYS if out[k].decode("utf8") != ref[k]]
            if verbose:
                print("   table %d (%d haplotypes, alleles %s, %d pops) %-22s %4d subj  %s  rows umug=%d pmug=%d problem=%d" % (
                    t, len(tab.haps), n_alleles, len(pops), tag, len(lines), "OK" if not diff else "MISMATCH " + ",".join(diff),
                    ref["umug"].count("\n"), ref["pmug"].count("\n"), ref["problem"].count("\n")), flush=True)
            if diff:
                bad.append((t, tag, diff))
    finally:
        g.close()
    return bad


@pytest.mark.parametrize("t", range(6))
def test_cuda_path_matches_oracle_on_random_tables(t, tmp_path):
    assert run_table(t, 20261018, 12, str(tmp_path)) == []


@pytest.mark.parametrize("t", range(2))
def test_cuda_path_matches_oracle_on_random_nine_locus_tables(t, tmp_path):
    assert run_table(t, 20261018, 3, str(tmp_path), nine=True) == []


@pytest.mark.parametrize("t", range(3))
def test_wide_key_build_matches_oracle_on_random_tables(t, tmp_path, monkeypatch):
    monkeypatch.setenv("GRIMB_KEY_WORDS", "2")
    assert run_table(t, 20261019, 8, str(tmp_path)) == []


if __name__ == "__main__":
    nine = len(sys.argv) > 1 and sys.argv[1] == "--nine"
    if nine:
        del sys.argv[1]
    n_tables = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 7
    bad = []
    for t in range(n_tables):
        with tempfile.TemporaryDirectory() as d:
            bad += run_table(t, seed, n, d, verbose=True, nine=nine)
    print("ALL OK" if not bad else "MISMATCH %r" % bad)
    sys.exit(1 if bad else 0)
