#!/usr/bin/env python
"""Throughput + parity of the BASELINE.json parity-test configurations through the product's
Python API on one GPU (not bench lines: bench.py measures configs[1]).  For each configuration:
subjects/s end to end through Imputation.impute_lines (tokeniser + C ABI + formatter), the share
spent inside the C ABI call, and a byte-for-byte comparison of the six output texts with the CPU
oracle on the first `sample` subjects.  Writes one JSON line per configuration.

    python tests/tools/run_configs.py > profiles/rNN_configs.jsonl
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "py-graph-imputation_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import goldenlib  # noqa: E402
import grim_oracle as go  # noqa: E402
import synth  # noqa: E402
from grim.imputation.impute import Imputation  # noqa: E402
from grim.imputation.networkx_graph import Graph  # noqa: E402
from grim.run_impute_def import load_config  # noqa: E402


def run(name, conf, hpf_text, counts_text, lines, sample, tmp, oracle_marginals=True):
    hpf = os.path.join(tmp, name + "_hpf.csv")
    cnt = os.path.join(tmp, name + "_counts.txt")
    open(hpf, "w").write(hpf_text)
    open(cnt, "w").write(counts_text)
    conf = dict(conf)
    conf["freq_file"], conf["pops_count_file"] = hpf, cnt
    cfg = load_config(conf)
    t = time.time()
    g = Graph(cfg).build_graph()
    t_build = time.time() - t
    imp = Imputation(g, cfg)
    imp.impute_lines(lines[:64])                      # warm-up (engine creation)
    data = "".join(lines).encode("utf8")
    Imputation(g, cfg).impute_text(data)              # untimed pass: sizes the pinned staging buffers
    imp = Imputation(g, cfg)
    t = time.time()
    texts = imp.impute_text(data)
    dt = time.time() - t
    out = {k: v.splitlines() for k, v in texts.items()}
    mine = {k: v.decode("utf8") for k, v in imp.impute_text("".join(lines[:sample]).encode("utf8")).items()}
    og = go.graph_from_config(conf, marginals=oracle_marginals)
    oimp = go.OracleImputation(og, go.load_config(conf), go.count_by_prob_from_file(len(conf["populations"]), cnt))
    t = time.time()
    ref = oimp.impute_lines(lines[:sample])
    t_cpu = time.time() - t
    if not oracle_marginals:
        # full-label-only oracle graph: exact only while every sampled subject is served by Plan A
        only_a = imp.stats["plan"]
        assert only_a[2] == 0 and only_a[3] == 0, "sample left Plan A: %s" % only_a
    info = g.info()
    eng = g.engine(imp.workspaces[0])
    kms = [g.lib.grimb_engine_kernel_ms(eng, w) for w in (0, 1, 2, 3)]   # needs GRIMB_HOST_EVENTS=1
    rec = {
        "config": name, "subjects": len(lines), "gpu_subjects_per_s": len(lines) / dt,
        "gpu_abi_seconds": imp.stats.get("abi_seconds"), "gpu_total_seconds": dt,
        "tokenise_seconds": imp.stats.get("tokenise_seconds"), "format_seconds": imp.stats.get("format_seconds"),
        "pair_evals": imp.stats["pair_evals"], "plans": imp.stats["plan"],
        "workspace_retries": imp.stats["workspace_retries"],
        # device time of the kernels of the LAST ABI call of the first workspace tier (the sample run)
        "last_call_kernel_ms": {"k_impute_fast": kms[0], "k_impute": kms[1], "k_impute_typed": kms[2],
                                "subjects_handed_to_k_impute": kms[3], "subjects": sample},
        "cpu_port_subjects_per_s_1core": sample / t_cpu, "cpu_sample": sample,
        "parity_identical_on_sample": all(mine[k] == ref[k] for k in ref),
        "rows": {k: len(v) for k, v in out.items()},
        "table": {"n_full": info["n_full"], "n_nodes": info["n_nodes"], "pops": len(conf["populations"]), "build_s": t_build,
                  "device_bytes": info["device_bytes"], "key_words": g.kw, "key_bits": sum(g.key_bits)},
    }
    print(json.dumps(rec), flush=True)
    g.close()


def main():
    import tempfile
    tmp = tempfile.mkdtemp()
    base = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    cau = open(os.path.join(goldenlib.GOLD, "data", "cau_hpf.csv")).read()
    cau_cnt = open(os.path.join(goldenlib.GOLD, "data", "cau_pop_counts.txt")).read()
    tab = synth.Table(cau)
    scale = float(os.environ.get("CONFIG_SCALE", "1"))
    # C1: the README example
    run("C1_readme_donor", base, cau, cau_cnt, open(os.path.join(goldenlib.GOLD, "data", "donor.csv")).readlines(), 1, tmp)
    # C2 on the CAU table, through the Python host
    run("C2_cau_typed", base, cau, cau_cnt, synth.typed_subjects(tab, int(1000000 * scale), 1, ["CAU,CAU"]), 2000, tmp)
    # C3: 21 populations, race fields, top-100 population results
    pops = ["P%02d" % i for i in range(21)]
    hpf21, cnt21 = synth.multipop_hpf(cau, pops, 21)
    c3 = dict(base)
    c3.update({"populations": pops, "UNK_priors": "MR", "number_of_pop_results": 100})
    tab21 = synth.Table(hpf21, "P00")
    run("C3_21pops_typed", c3, hpf21, cnt21, synth.typed_subjects(tab21, int(100000 * scale), 3, synth.race_fields(pops)), 300, tmp)
    # C4: ambiguous / missing / unknown, incl. subjects over a lowered options threshold
    run("C4_messy", base, cau, cau_cnt, synth.messy_subjects(tab, int(3000 * scale), 4, max_amb=6), 100, tmp)
    c4 = dict(base)
    c4["number_of_options_threshold"] = 200
    run("C4_messy_threshold200", c4, cau, cau_cnt, synth.messy_subjects(tab, int(3000 * scale), 5, max_amb=6), 100, tmp)
    run("C3_21pops_messy", c3, hpf21, cnt21, synth.messy_subjects(tab21, int(1000 * scale), 6, races=synth.race_fields(pops)), 60, tmp)


def c5(tmp, scale):
    """C5: 9 loci (256 phases, 511 labels), 5 populations, allele dictionaries wide enough that the
    packed key needs the 128-bit build; the table is sized to be HBM-resident and L2-hostile."""
    base = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    loci = ["A", "B", "C", "DPA1", "DPB1", "DQA1", "DQB1", "DRB1", "DRBX"]
    n_all = [700, 1200, 600, 40, 300, 60, 250, 700, 100]
    pops = ["Q%d" % i for i in range(5)]
    n_full = int(float(os.environ.get("C5_HAPLOTYPES", "300000")) * scale)
    hpf = synth.zipf_table(n_full, n_all, 20261018, loci=loci, pops=pops)
    counts = 1000.0 / np.arange(1, 6) ** 1.1
    cnt = "".join("%s,%s,%s\n" % (p, repr(float(c)), repr(float(c / counts.sum()))) for p, c in zip(pops, counts))
    c = dict(base)
    c.update({"populations": pops, "UNK_priors": "MR", "number_of_pop_results": 100,
              "loci_map": {l: i + 1 for i, l in enumerate(loci)},
              "freq_trim_threshold": 1e-30,
              "Plan_B_Matrix": [[[1, 2, 3, 4, 5, 6, 7, 8, 9]], [[1, 2, 3], [4, 5], [6, 7, 8, 9]],
                                [[1], [2, 3], [4, 5], [6, 7], [8, 9]], [[1], [2], [3], [4], [5], [6], [7], [8], [9]]]})
    tab = synth.Table(hpf, pops[0], loci=loci)
    lines = synth.typed_subjects(tab, int(100000 * scale), 9, synth.race_fields(pops))
    run("C5_nine_loci_typed", c, hpf, cnt, lines, 200, tmp, oracle_marginals=False)


def c4_heavy(tmp, scale):
    """C4 proper: list sizes 6-40 per locus side, products on both sides of the 100,000-option
    threshold, 0-3 missing loci, unknown alleles, the full 6-row Plan-B matrix."""
    base = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    cau = open(os.path.join(goldenlib.GOLD, "data", "cau_hpf.csv")).read()
    cau_cnt = open(os.path.join(goldenlib.GOLD, "data", "cau_pop_counts.txt")).read()
    tab = synth.Table(cau)
    n = int(float(os.environ.get("C4_SUBJECTS", "2000")) * scale)
    lines = synth.heavy_subjects(tab, n, 44, races=["CAU,CAU"])
    run("C4_heavy_over_threshold", base, cau, cau_cnt, lines, int(os.environ.get("C4_SAMPLE", "3")), tmp)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "c4heavy":
        import tempfile
        c4_heavy(tempfile.mkdtemp(), float(os.environ.get("CONFIG_SCALE", "1")))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "c5":
        import tempfile
        c5(tempfile.mkdtemp(), float(os.environ.get("CONFIG_SCALE", "1")))
        sys.exit(0)
    main()
