"""Differential fuzz: oracle/grim_oracle.py against the real reference (build container only).

    python tests/golden/fuzz_vs_reference.py [n_subjects_per_case] [seed]

Every case runs the same hpf/config/subject file through both and compares the six output
files byte for byte.  Used to pin the oracle before trusting it ("Oracle")."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))

import grim_oracle as go  # noqa: E402
import synth  # noqa: E402
from refrun import RefSession  # noqa: E402

BASE_CONF = json.load(open(os.path.join(HERE, "data", "base_conf.json")))
CAU = open(os.path.join(HERE, "data", "cau_hpf.csv")).read()


def compare(tag, sess, hpf, counts, lines, hap_pop_pair=False, phase_masks=None, **over):
    conf = dict(sess.conf)
    conf.update(over)
    ref = sess.run(lines, hap_pop_pair=hap_pop_pair, phase_masks=phase_masks, **over)
    if phase_masks is not None:
        import tempfile
        fd, path = tempfile.mkstemp(suffix=".json")
        with os.fdopen(fd, "w") as f:
            json.dump(phase_masks, f)
        conf["bin_imputation_in_file"] = path
    t = time.time()
    g = go.OracleGraph(hpf.splitlines(True), conf["populations"], conf["loci_map"],
                       conf["freq_trim_threshold"], counts.splitlines(True) if counts else None)
    cfg = go.load_config(conf)
    cbp = None
    if counts:
        import numpy as np
        cbp = np.array([float(l.split(",")[2]) for l in counts.splitlines()])
    imp = go.OracleImputation(g, cfg, cbp)
    mine = imp.impute_lines(lines, em_mr=hap_pop_pair)
    dt = time.time() - t
    bad = [k for k in ref if ref[k] != mine[k]]
    print("%-40s %5d subj  %s  (%.1fs oracle)  rows umug=%d pmug=%d miss=%d problem=%d" % (
        tag, len(lines), "OK" if not bad else "MISMATCH " + ",".join(bad), dt,
        ref["umug"].count("\n"), ref["pmug"].count("\n"), ref["miss"].count("\n"), ref["problem"].count("\n")))
    if bad:
        for k in bad:
            a, b = ref[k].splitlines(), mine[k].splitlines()
            for i in range(max(len(a), len(b))):
                x = a[i] if i < len(a) else None
                y = b[i] if i < len(b) else None
                if x != y:
                    print("   first diff in", k, "line", i, "\n     ref :", x, "\n     mine:", y)
                    break
    return not bad


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    ok = True
    tab = synth.Table(CAU)
    s1 = RefSession(BASE_CONF, CAU, "CAU,3380.0,1.0\n")
    ok &= compare("cau typed", s1, CAU, "CAU,3380.0,1.0\n", synth.typed_subjects(tab, n, seed))
    ok &= compare("cau messy", s1, CAU, "CAU,3380.0,1.0\n", synth.messy_subjects(tab, n, seed + 1))
    ok &= compare("cau messy thr=40", s1, CAU, "CAU,3380.0,1.0\n",
                  synth.messy_subjects(tab, n, seed + 2, max_amb=5), number_of_options_threshold=40)
    ok &= compare("cau messy topk=5 nres=3", s1, CAU, "CAU,3380.0,1.0\n",
                  synth.messy_subjects(tab, n, seed + 3), max_haplotypes_number_in_phase=5,
                  number_of_results=3)
    ok &= compare("cau messy save_space", s1, CAU, "CAU,3380.0,1.0\n",
                  synth.messy_subjects(tab, n, seed + 4), save_space_mode=True)
    ok &= compare("cau messy planb off", s1, CAU, "CAU,3380.0,1.0\n",
                  synth.messy_subjects(tab, n, seed + 5), planb=False)
    ok &= compare("cau messy umug only", s1, CAU, "CAU,3380.0,1.0\n",
                  synth.messy_subjects(tab, n, seed + 6), output_haplotypes=False)
    ok &= compare("cau messy pmug only", s1, CAU, "CAU,3380.0,1.0\n",
                  synth.messy_subjects(tab, n, seed + 7), output_MUUG=False)
    ok &= compare("cau heavy unknown", s1, CAU, "CAU,3380.0,1.0\n",
                  synth.messy_subjects(tab, n, seed + 8, p_unknown=0.4, p_random=0.4))
    s1.close()

    pops = ["AAA", "BBB", "CCC"]
    hpf3, cnt3 = synth.multipop_hpf(CAU, pops, 7)
    conf3 = dict(BASE_CONF)
    conf3["populations"] = pops
    conf3["UNK_priors"] = "MR"
    tab3 = synth.Table(hpf3, "AAA")
    races = synth.race_fields(pops)
    s3 = RefSession(conf3, hpf3, cnt3)
    ok &= compare("pop3 typed races MR", s3, hpf3, cnt3, synth.typed_subjects(tab3, n, seed + 10, races))
    ok &= compare("pop3 messy races MR", s3, hpf3, cnt3, synth.messy_subjects(tab3, n, seed + 11, races=races))
    ok &= compare("pop3 messy races SR", s3, hpf3, cnt3, synth.messy_subjects(tab3, n, seed + 12, races=races),
                  UNK_priors="SR")
    ok &= compare("pop3 messy races thr=40", s3, hpf3, cnt3,
                  synth.messy_subjects(tab3, n, seed + 13, max_amb=5, races=races),
                  number_of_options_threshold=40)
    ok &= compare("pop3 messy eta>0", s3, hpf3, cnt3, synth.messy_subjects(tab3, n, seed + 14, races=races),
                  priority={"alpha": 0.4, "eta": 0.01, "beta": 1e-3, "gamma": 1e-2, "delta": 0.3})
    # EM-facing modes: hap_pop_pair rows and per-subject phase masks
    import numpy as np
    rng = np.random.RandomState(seed + 20)

    def masks_for(lines):
        return {ln.split(",")[0]: [int(x) for x in rng.randint(0, 2, size=4)] for ln in lines}

    lines = synth.typed_subjects(tab3, n, seed + 15, races) + synth.messy_subjects(tab3, n, seed + 16, races=races)
    ok &= compare("pop3 hap_pop_pair", s3, hpf3, cnt3, lines, hap_pop_pair=True)
    ok &= compare("pop3 hap_pop_pair nres=4", s3, hpf3, cnt3, lines, hap_pop_pair=True, number_of_results=4)
    ok &= compare("pop3 phase masks", s3, hpf3, cnt3, lines, phase_masks=masks_for(lines))
    lines = synth.messy_subjects(tab3, n, seed + 17, max_amb=3, p_missing=0.4, races=races)
    ok &= compare("pop3 phase masks missing loci", s3, hpf3, cnt3, lines, phase_masks=masks_for(lines))
    ok &= compare("pop3 phase masks + hap_pop_pair", s3, hpf3, cnt3, lines, hap_pop_pair=True, phase_masks=masks_for(lines))
    s3.close()
    print("ALL OK" if ok else "SOME MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
