#!/usr/bin/env python
"""Repeats the text -> text call of bench.py's C5 record (nine loci, 128-bit keys) and prints, per call, the time in
the ABI and the subjects re-issued on a bigger workspace tier.    python tools/c5_text_repeat.py [calls]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "py-graph-imputation_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import goldenlib  # noqa: E402
import synth  # noqa: E402
from grim.imputation.impute import Imputation  # noqa: E402
from grim.imputation.networkx_graph import Graph  # noqa: E402
from grim.run_impute_def import load_config  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 6
base = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
loci = ["A", "B", "C", "DPA1", "DPB1", "DQA1", "DQB1", "DRB1", "DRBX"]
n_all = [700, 1200, 600, 40, 300, 60, 250, 700, 100]
pops5 = ["Q%d" % i for i in range(5)]
lm9 = {l: i + 1 for i, l in enumerate(loci)}
names9, fa9, base_f = synth.zipf_arrays(100000, n_all, 20261018, loci)
ff9 = synth.multipop_freqs(base_f, 5, 9, zero_frac=0.2)
_ct, ratio5 = synth.pop_counts(pops5)
conf5 = dict(base)
conf5.update({"populations": pops5, "UNK_priors": "MR", "number_of_pop_results": 100, "loci_map": lm9,
              "Plan_B_Matrix": [[[1, 2, 3, 4, 5, 6, 7, 8, 9]], [[1, 2, 3], [4, 5], [6, 7, 8, 9]],
                                [[1], [2, 3], [4, 5], [6, 7], [8, 9]], [[1], [2], [3], [4], [5], [6], [7], [8], [9]]]})
cfg5 = load_config(conf5)
g5 = Graph(cfg5).from_arrays(names9, fa9, ff9)
p5 = ff9.mean(axis=1)
p5 = p5 / p5.sum()
lines5 = synth.array_subject_lines(names9, fa9, p5, 1 << 16, 9, synth.race_fields(pops5), variants=False)
data = "".join(lines5).encode()
warm = Imputation(g5, cfg5, ratio5)
warm.impute_text(data)
eng = g5.engine(warm.workspaces[0])
for c in range(calls):
    imp = Imputation(g5, cfg5, ratio5)
    imp._text = warm._text            # the same GrimbText (pinned staging sized by the warm-up), as bench_configs does
    t0 = time.time()
    imp.impute_text(data)
    print("call %d: %.1f ms wall, tokenise %.4f abi %.4f format %.4f s, retries %d, k_impute %.3f ms, k_impute_typed %.3f ms, slots %.3f ms" % (
        c, (time.time() - t0) * 1e3, imp.stats.get("tokenise_seconds", 0.0), imp.stats.get("abi_seconds", 0.0),
        imp.stats.get("format_seconds", 0.0), imp.stats["workspace_retries"], g5.lib.grimb_engine_kernel_ms(eng, 1),
        g5.lib.grimb_engine_kernel_ms(eng, 2), g5.lib.grimb_engine_kernel_ms(eng, 5)), flush=True)
