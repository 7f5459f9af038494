"""Differential fuzz on the CPU: the kernel source (csrc/grimb_plan.h compiled for the host,
tests/emu/) driven by the product's host code, against oracle/grim_oracle.py, on random
subjects in many configurations.  The oracle itself is pinned against the real reference by
fuzz_vs_reference.py; this script extends the same random coverage to the CUDA source without a
GPU (general kernel logic only -- the warp kernels are covered by the GPU agreement tests).

    python tests/golden/fuzz_emu_vs_oracle.py [n_subjects_per_case] [seed] [rounds]
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
for p in (os.path.join(ROOT, "py-graph-imputation_b200"), os.path.join(ROOT, "oracle"), os.path.join(HERE, "..")):
    sys.path.insert(0, p)

import grim_oracle as go  # noqa: E402
import synth  # noqa: E402
from emu_backend import EmuGraph, emu_imputation  # noqa: E402
from grim.run_impute_def import load_config  # noqa: E402

BASE_CONF = json.load(open(os.path.join(HERE, "data", "base_conf.json")))
CAU = open(os.path.join(HERE, "data", "cau_hpf.csv")).read()
KEYS = ("umug", "umug_pops", "pmug", "pmug_pops", "miss", "problem")


class Session(object):
    def __init__(self, conf, hpf, counts):
        self.conf = dict(conf)
        self.og = go.OracleGraph(hpf.splitlines(True), conf["populations"], conf["loci_map"],
                                 conf["freq_trim_threshold"], counts.splitlines(True) if counts else None)
        self.eg = EmuGraph(self.og, conf["loci_map"])
        self.cbp = np.array([float(l.split(",")[2]) for l in counts.splitlines()]) if counts else None

    def compare(self, tag, lines, hap_pop_pair=False, phase_masks=None, **over):
        conf = dict(self.conf)
        conf.update(over)
        path = None
        if phase_masks is not None:
            fd, path = tempfile.mkstemp(suffix=".json")
            with os.fdopen(fd, "w") as f:
                json.dump(phase_masks, f)
            conf["bin_imputation_in_file"] = path
        t = time.time()
        ref = go.OracleImputation(self.og, go.load_config(conf), self.cbp).impute_lines(lines, em_mr=hap_pop_pair)
        t_o = time.time() - t
        t = time.time()
        out = emu_imputation(self.eg, load_config(conf), self.cbp).impute_lines(lines, em_mr=hap_pop_pair)
        t_e = time.time() - t
        if path:
            os.unlink(path)
        mine = {k: "".join(v) for k, v in out.items()}
        bad = [k for k in KEYS if ref[k] != mine[k]]
        print("%-40s %5d subj  %s  (oracle %.1fs, emulated kernel %.1fs)  rows umug=%d pmug=%d miss=%d problem=%d" % (
            tag, len(lines), "OK" if not bad else "MISMATCH " + ",".join(bad), t_o, t_e, ref["umug"].count("\n"),
            ref["pmug"].count("\n"), ref["miss"].count("\n"), ref["problem"].count("\n")), flush=True)
        for k in bad:
            a, b = ref[k].splitlines(), mine[k].splitlines()
            for i in range(max(len(a), len(b))):
                x = a[i] if i < len(a) else None
                y = b[i] if i < len(b) else None
                if x != y:
                    print("   first diff in", k, "line", i, "\n     oracle:", x, "\n     kernel:", y, flush=True)
                    break
        return not bad


def one_round(n, seed, s1, s3, tab, tab3, races):
    ok = True
    c = s1.compare
    ok &= c("cau typed", synth.typed_subjects(tab, n, seed))
    ok &= c("cau messy", synth.messy_subjects(tab, n, seed + 1))
    ok &= c("cau messy thr=40", synth.messy_subjects(tab, n, seed + 2, max_amb=5), number_of_options_threshold=40)
    ok &= c("cau messy topk=5 nres=3", synth.messy_subjects(tab, n, seed + 3), max_haplotypes_number_in_phase=5,
            number_of_results=3)
    ok &= c("cau messy save_space", synth.messy_subjects(tab, n, seed + 4), save_space_mode=True)
    ok &= c("cau messy planb off", synth.messy_subjects(tab, n, seed + 5), planb=False)
    ok &= c("cau messy umug only", synth.messy_subjects(tab, n, seed + 6), output_haplotypes=False)
    ok &= c("cau messy pmug only", synth.messy_subjects(tab, n, seed + 7), output_MUUG=False)
    ok &= c("cau heavy unknown", synth.messy_subjects(tab, n, seed + 8, p_unknown=0.4, p_random=0.4))
    ok &= c("cau messy thr=200 topk=20", synth.messy_subjects(tab, n, seed + 9, max_amb=4),
            number_of_options_threshold=200, max_haplotypes_number_in_phase=20)
    c = s3.compare
    ok &= c("pop3 typed races MR", synth.typed_subjects(tab3, n, seed + 10, races))
    ok &= c("pop3 messy races MR", synth.messy_subjects(tab3, n, seed + 11, races=races))
    ok &= c("pop3 messy races SR", synth.messy_subjects(tab3, n, seed + 12, races=races), UNK_priors="SR")
    ok &= c("pop3 messy races thr=40", synth.messy_subjects(tab3, n, seed + 13, max_amb=5, races=races),
            number_of_options_threshold=40)
    ok &= c("pop3 messy eta>0", synth.messy_subjects(tab3, n, seed + 14, races=races),
            priority={"alpha": 0.4, "eta": 0.01, "beta": 1e-3, "gamma": 1e-2, "delta": 0.3})
    ok &= c("pop3 messy npop=2 nres=1", synth.messy_subjects(tab3, n, seed + 18, races=races),
            number_of_pop_results=2, number_of_results=1)
    rng = np.random.RandomState(seed + 20)

    def masks_for(lines):
        return {ln.split(",")[0]: [int(x) for x in rng.randint(0, 2, size=4)] for ln in lines}

    lines = synth.typed_subjects(tab3, n, seed + 15, races) + synth.messy_subjects(tab3, n, seed + 16, races=races)
    ok &= c("pop3 hap_pop_pair", lines, hap_pop_pair=True)
    ok &= c("pop3 hap_pop_pair nres=4", lines, hap_pop_pair=True, number_of_results=4)
    ok &= c("pop3 phase masks", lines, phase_masks=masks_for(lines))
    lines = synth.messy_subjects(tab3, n, seed + 17, max_amb=3, p_missing=0.4, races=races)
    ok &= c("pop3 phase masks missing loci", lines, phase_masks=masks_for(lines))
    ok &= c("pop3 phase masks + hap_pop_pair", lines, hap_pop_pair=True, phase_masks=masks_for(lines))
    return ok


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    tab = synth.Table(CAU)
    s1 = Session(BASE_CONF, CAU, "CAU,3380.0,1.0\n")
    pops = ["AAA", "BBB", "CCC"]
    hpf3, cnt3 = synth.multipop_hpf(CAU, pops, 7)
    conf3 = dict(BASE_CONF)
    conf3["populations"] = pops
    conf3["UNK_priors"] = "MR"
    tab3 = synth.Table(hpf3, "AAA")
    races = synth.race_fields(pops)
    s3 = Session(conf3, hpf3, cnt3)
    ok = True
    for r in range(rounds):
        print("== round %d (seed %d)" % (r, seed + 1000 * r), flush=True)
        ok &= one_round(n, seed + 1000 * r, s1, s3, tab, tab3, races)
    print("ALL OK" if ok else "SOME MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
