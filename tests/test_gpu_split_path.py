"""GPU parity of the split single-population fast path (k_fast_probe + k_fast_score + the overflow
list served by the fused kernel) against the CPU oracle, on a table built so that EVERY
recombinant of a subject is present: 2 alleles per locus x 5 loci = 32 haplotypes, so a subject
heterozygous at h loci has 2^(h-1) candidate phases (16, 8, 4, 2, 1) -- more than the four a
hand-over record carries for h = 5 and 4 -- under several configurations (epsilon schedule length,
output kinds, result limits, Plan B off)."""
import itertools
import json
import os

import numpy as np
import pytest

import goldenlib
import grim_oracle as go

pytestmark = pytest.mark.gpu

LOCI = ["A", "B", "C", "DQB1", "DRB1"]


def _table(seed):
    rng = np.random.RandomState(seed)
    names = [["%s*01:01" % l, "%s*02:01" % l] for l in LOCI]
    rows = ["hap,pop,freq\n"]
    f = rng.lognormal(0.0, 2.0, size=32)
    f[rng.rand(32) < 0.15] = 0.0            # some recombinants absent
    f = f / f.sum()
    for k, combo in enumerate(itertools.product(range(2), repeat=5)):
        if f[k] > 0:
            rows.append("%s,CAU,%r\n" % ("~".join(names[l][c] for l, c in enumerate(combo)), float(f[k])))
    return names, "".join(rows)


def _subjects(names, n, seed):
    rng = np.random.RandomState(seed)
    out = []
    for s in range(n):
        sides = []
        for l in range(5):
            a, b = rng.randint(0, 2, size=2)
            x, y = names[l][a], names[l][b]
            r = rng.rand()
            if r < 0.03:
                x = "%s*99:01" % LOCI[l]          # allele absent from the table
            sides.append("%s+%s" % (x, y))
        out.append("F%d,%s,CAU,CAU\n" % (s, "^".join(sides)))
    return out


CONFIGS = [
    {},
    {"epsilon": 1e-1},
    {"epsilon": 1e-6},
    {"output_MUUG": False},
    {"output_haplotypes": False},
    {"number_of_results": 1, "number_of_pop_results": 1},
    {"number_of_results": 3},
    {"planb": False},
    {"UNK_priors": "MR"},
]


@pytest.mark.parametrize("over", CONFIGS, ids=[json.dumps(c, sort_keys=True) for c in CONFIGS])
def test_split_fast_path_matches_oracle(over, tmp_path, monkeypatch):
    monkeypatch.setenv("GRIMB_HOST_EVENTS", "1")   # kernel timing events in the host-pointer path (read at engine creation)
    from grim.imputation.impute import Imputation
    from grim.imputation.networkx_graph import Graph
    from grim.run_impute_def import load_config
    names, hpf = _table(3)
    d = str(tmp_path)
    open(d + "/hpf.csv", "w").write(hpf)
    open(d + "/cnt.txt", "w").write("CAU,1000.0,1.0\n")
    conf = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    conf.update({"freq_file": d + "/hpf.csv", "pops_count_file": d + "/cnt.txt", "freq_trim_threshold": 1e-30})
    conf.update(over)
    lines = _subjects(names, 3000, 9)
    cfg = load_config(conf)
    g = Graph(cfg).build_graph()
    imp = Imputation(g, cfg)
    out = imp.impute_text("".join(lines).encode("utf8"))
    eng = g.engine(imp.workspaces[0])
    assert g.lib.grimb_engine_kernel_ms(eng, 4) >= 0, "the split fast path did not run"
    ref, _ = go.impute_file(conf, lines=lines)
    for k in goldenlib.KEYS:
        assert out[k].decode("utf8") == ref[k], "%s differs" % k
    g.close()
