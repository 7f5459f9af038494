"""Generates tests/golden/cases/* by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Each case directory holds: case.json (table name, config overrides), subjects.csv and the six
expected output files exp.umug / exp.umug_pops / exp.pmug / exp.pmug_pops / exp.miss /
exp.problem exactly as the reference wrote them.  Tables live in tests/golden/data/ (CAU hpf
from the reference's README flow: produce_hpf on data/freqs/CAU.freqs.gz; pop3 = perturbed
CAU, see tests/synth.py).  G-numbers follow SURVEY.md section 8(c)."""
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)

import synth  # noqa: E402
from refrun import RefSession, ensure_reference, REF_COPY  # noqa: E402

DATA = os.path.join(HERE, "data")
CASES = os.path.join(HERE, "cases")
POPS3 = ["AAA", "BBB", "CCC"]


def make_tables():
    """CAU hpf via the reference's own produce_hpf; pop3 via synth."""
    ensure_reference()
    import contextlib
    import io
    from graph_generation import generate_hpf
    conf = json.load(open(os.path.join(REF_COPY, "conf", "minimal-configuration.json")))
    json.dump(conf, open(os.path.join(DATA, "base_conf.json"), "w"), indent=1)
    cwd = os.getcwd()
    os.chdir(REF_COPY)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            generate_hpf.produce_hpf(conf_file="conf/minimal-configuration.json")
        shutil.copy("output/hpf.csv", os.path.join(DATA, "cau_hpf.csv"))
        shutil.copy("output/pop_counts_file.txt", os.path.join(DATA, "cau_pop_counts.txt"))
        shutil.copy("data/subjects/donor.csv", os.path.join(DATA, "donor.csv"))
    finally:
        os.chdir(cwd)
    cau = open(os.path.join(DATA, "cau_hpf.csv")).read()
    hpf3, cnt3 = synth.multipop_hpf(cau, POPS3, 7)
    open(os.path.join(DATA, "pop3_hpf.csv"), "w").write(hpf3)
    open(os.path.join(DATA, "pop3_pop_counts.txt"), "w").write(cnt3)


def edge_lines(tab):
    h = tab.haps
    a, b = h[0], h[5]

    def gl(pairs):
        return "^".join(x + "+" + y for x, y in pairs)

    full = list(zip(a, b))
    amb = [(a[0] + "/" + h[9][0], b[0])] + full[1:]
    miss = [full[0], full[1], full[4]]
    unk = [("A*99:99", b[0])] + full[1:]
    sfx = [(a[0] + "g", b[0] + "L")] + [(h[17][i], h[431][i]) for i in range(1, 5)]
    allunk = [("%s*99:01" % l, "%s*99:02" % l) for l in tab.loci]
    return [
        "E1," + gl(full) + ",CAU,CAU\n",
        "E2," + gl(amb) + ",CAU,CAU\n",
        "E3," + gl(miss) + ",CAU,CAU\n",
        "E4," + gl(unk) + ",CAU,CAU\n",
        "E5," + a[0] + "^" + gl(full[1:]) + ",CAU,CAU\n",
        "E6,,CAU,CAU\n",
        "E7," + gl(full[:2]) + "^C*UUUU+C*UUUU^" + gl(full[3:]) + ",XXX,CAU\n",
        "E8," + gl(sfx) + "\n",
        "E9," + gl(allunk) + ",CAU,CAU\n",
        "E10%" + gl(full) + "%CAU%CAU\n",
        "E11," + gl([(x, x) for x in a]) + ",CAU,CAU\n",
        "E12," + gl(full[:4]) + "^" + a[4] + "+" + a[4] + ",CAU,CAU\n",
        "E13,+" + gl(full) + ",CAU,CAU\n",
        "E14\n",
        "E15," + gl(full) + ",CAU\n",
    ]


def t1_lines(tab):
    last = tab_last_allele(tab)
    other = "DRB1*15:01"
    return [
        "T1a,%s+%s\n" % (last, other),
        "T1b,%s+%s\n" % (other, other),
        "T1c,%s+%s\n" % (last, last),
        "T1d,DQB1*06:02+DQB1*06:02^%s+%s\n" % (last, other),
    ]


def tab_last_allele(tab):
    # last first-seen allele of the last locus in hpf order (the node hit by SURVEY trap T1)
    seen = []
    for hp in tab.haps:
        if hp[-1] not in seen:
            seen.append(hp[-1])
    return seen[-1]


def ambiguous_lines(tab, per_locus, n_missing, seed, sid):
    import numpy as np
    rng = np.random.RandomState(seed)
    pairs = []
    for l in range(len(tab.loci)):
        if l < n_missing:
            continue
        al = tab.alleles[l]
        s1 = [al[i] for i in rng.choice(len(al), size=min(per_locus, len(al)), replace=False)]
        s2 = [al[i] for i in rng.choice(len(al), size=min(per_locus, len(al)), replace=False)]
        pairs.append(("/".join(s1), "/".join(s2)))
    return "%s,%s,CAU,CAU\n" % (sid, "^".join(x + "+" + y for x, y in pairs))


LOCI9 = ["A", "B", "C", "DPA1", "DPB1", "DQA1", "DQB1", "DRB1", "DRBX"]
NINE_OVER = {
    "populations": ["AAA", "BBB"], "UNK_priors": "MR", "freq_trim_threshold": 1e-9,
    "loci_map": {l: i + 1 for i, l in enumerate(LOCI9)},
    "Plan_B_Matrix": [[[1, 2, 3, 4, 5, 6, 7, 8, 9]], [[1, 2, 3], [4, 5], [6, 7, 8, 9]],
                      [[1], [2, 3], [4, 5], [6, 7], [8, 9]], [[1], [2], [3], [4], [5], [6], [7], [8], [9]]],
}


def nine_table(n_full, seed):
    """Small synthetic 9-locus table (BASELINE config 5 shape: 256 phases, 510 marginal labels)."""
    import numpy as np
    rng = np.random.RandomState(seed)
    na = [12, 20, 10, 4, 9, 5, 7, 11, 3]
    seen, haps = set(), []
    while len(haps) < n_full:
        h = tuple(min(int(rng.zipf(1.6)) - 1, n - 1) for n in na)
        if h not in seen:
            seen.add(h)
            haps.append(h)
    out = ["hap,pop,freq\n"]
    for p in NINE_OVER["populations"]:
        f = rng.rand(n_full) ** 3
        f /= f.sum()
        for h, x in zip(haps, f):
            if rng.rand() < 0.85:
                name = "~".join("%s*%02d:01" % (l, a + 1) for l, a in zip(LOCI9, h))
                out.append("%s,%s,%s\n" % (name, p, repr(float(x))))
    return "".join(out), "AAA,100.0,0.5\nBBB,100.0,0.5\n"


def main():
    os.makedirs(DATA, exist_ok=True)
    only = sys.argv[1:]
    if not only:
        make_tables()
        shutil.rmtree(CASES, ignore_errors=True)
    if not os.path.exists(os.path.join(DATA, "nine_hpf.csv")):
        h9, c9 = nine_table(150, 5)
        open(os.path.join(DATA, "nine_hpf.csv"), "w").write(h9)
        open(os.path.join(DATA, "nine_pop_counts.txt"), "w").write(c9)
    base = json.load(open(os.path.join(DATA, "base_conf.json")))
    cau = open(os.path.join(DATA, "cau_hpf.csv")).read()
    cau_cnt = open(os.path.join(DATA, "cau_pop_counts.txt")).read()
    hpf3 = open(os.path.join(DATA, "pop3_hpf.csv")).read()
    cnt3 = open(os.path.join(DATA, "pop3_pop_counts.txt")).read()
    tab = synth.Table(cau)
    tab3 = synth.Table(hpf3, "AAA")
    races3 = synth.race_fields(POPS3)

    conf3 = dict(base)
    conf3["populations"] = POPS3
    conf3["UNK_priors"] = "MR"
    hpf9 = open(os.path.join(DATA, "nine_hpf.csv")).read()
    cnt9 = open(os.path.join(DATA, "nine_pop_counts.txt")).read()
    conf9 = dict(base)
    conf9.update(NINE_OVER)
    tab9 = synth.Table(hpf9, "AAA", LOCI9)
    sessions = {"cau": RefSession(base, cau, cau_cnt), "pop3": RefSession(conf3, hpf3, cnt3)}
    if not only or any(o.startswith("g6") for o in only):
        sessions["nine"] = RefSession(conf9, hpf9, cnt9)
    base_over = {"cau": {}, "pop3": {"populations": POPS3, "UNK_priors": "MR"}, "nine": NINE_OVER}

    cases = [
        ("g1_readme_donor", "cau", {}, open(os.path.join(DATA, "donor.csv")).readlines()),
        ("g2_edges", "cau", {}, edge_lines(tab)),
        ("g2_edges_umug_only", "cau", {"output_haplotypes": False}, edge_lines(tab)),
        ("g2_edges_pmug_only", "cau", {"output_MUUG": False}, edge_lines(tab)),
        ("g2_edges_planb_off", "cau", {"planb": False}, edge_lines(tab)),
        ("g2_t1_last_node", "cau", {}, t1_lines(tab)),
        ("g3_pop3_typed", "pop3", {}, synth.typed_subjects(tab3, 60, 21, races3)),
        ("g3_pop3_messy", "pop3", {}, synth.messy_subjects(tab3, 60, 22, races=races3)),
        ("g3_pop3_messy_sr", "pop3", {"UNK_priors": "SR"}, synth.messy_subjects(tab3, 40, 23, races=races3)),
        ("g3_pop3_priority", "pop3",
         {"priority": {"alpha": 0.4, "eta": 0.01, "beta": 1e-3, "gamma": 1e-2, "delta": 0.3}},
         synth.messy_subjects(tab3, 40, 24, races=races3)),
        ("g4_amb6", "cau", {}, [ambiguous_lines(tab, 6, 0, 31, "A6")]),
        ("g4_amb12_over_threshold", "cau", {}, [ambiguous_lines(tab, 12, 0, 32, "A12")]),
        ("g4_amb14_missing_locus", "cau", {}, [ambiguous_lines(tab, 14, 1, 33, "A14")]),
        ("g4_low_threshold", "cau", {"number_of_options_threshold": 40},
         synth.messy_subjects(tab, 80, 34, max_amb=5)),
        ("g4_low_threshold_pop3", "pop3", {"number_of_options_threshold": 40},
         synth.messy_subjects(tab3, 50, 35, max_amb=5, races=races3)),
        ("g4_topk5", "cau", {"max_haplotypes_number_in_phase": 5, "number_of_results": 3},
         synth.messy_subjects(tab, 50, 36)),
        ("g4_save_space", "cau", {"save_space_mode": True}, synth.messy_subjects(tab, 50, 37)),
        ("g4_unknown_heavy", "cau", {}, synth.messy_subjects(tab, 60, 38, p_unknown=0.4, p_random=0.4)),
        ("g5_typed_cau", "cau", {}, synth.typed_subjects(tab, 400, 41, ["CAU,CAU"])),
        ("g5_messy_cau", "cau", {}, synth.messy_subjects(tab, 120, 42)),
        ("g5_typed_nores1000", "cau", {"number_of_results": 1000, "epsilon": 1e-2},
         synth.typed_subjects(tab, 60, 43)),
        ("g6_nine_loci", "nine", {},
         synth.typed_subjects(tab9, 8, 1, ["AAA,BBB", ",", "AAA;BBB,XXX"])
         + synth.messy_subjects(tab9, 14, 2, max_amb=2, p_missing=0.3, races=["AAA,BBB", ","])),
    ]
    # per-subject phase masks (bin_imputation_in_file) and the hap_pop_pair output mode
    import numpy as np
    rng = np.random.RandomState(77)
    mask_lines = synth.typed_subjects(tab3, 40, 71, races3) + synth.messy_subjects(tab3, 40, 72, max_amb=3, races=races3)
    masks = {}
    for ln in mask_lines:
        sid = ln.split(",")[0]
        masks[sid] = [int(x) for x in rng.randint(0, 2, size=4)]
    masks[mask_lines[0].split(",")[0]] = [1, 1, 1, 1]
    masks[mask_lines[1].split(",")[0]] = [0, 0, 0, 0]
    del masks[mask_lines[2].split(",")[0]]          # unknown subject id: KeyError -> raw line in .problem
    extra = {
        "g7_phase_masks": {"phase_masks": masks},
        "g7_hap_pop_pair": {"hap_pop_pair": True},
        "g7_hap_pop_pair_cau": {"hap_pop_pair": True},
    }
    cases += [
        ("g7_phase_masks", "pop3", {}, mask_lines),
        ("g7_hap_pop_pair", "pop3", {"number_of_results": 7},
         synth.typed_subjects(tab3, 40, 73, races3) + synth.messy_subjects(tab3, 50, 74, races=races3)),
        ("g7_hap_pop_pair_cau", "cau", {}, edge_lines(tab) + synth.messy_subjects(tab, 40, 75)),
    ]
    for name, table, over, lines in cases:
        if only and name not in only:
            continue
        run_kw = extra.get(name, {})
        res = sessions[table].run(lines, **run_kw, **over)
        d = os.path.join(CASES, name)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        o = dict(base_over[table])
        o.update(over)
        meta = {"table": table, "overrides": o}
        if run_kw.get("hap_pop_pair"):
            meta["hap_pop_pair"] = True
        if run_kw.get("phase_masks") is not None:
            json.dump(run_kw["phase_masks"], open(os.path.join(d, "phase_masks.json"), "w"))
            meta["phase_masks"] = "phase_masks.json"
        json.dump(meta, open(os.path.join(d, "case.json"), "w"), indent=1)
        open(os.path.join(d, "subjects.csv"), "w").writelines(lines)
        for k, v in res.items():
            open(os.path.join(d, "exp." + k), "w").write(v)
        print("%-28s %4d subjects  umug=%d pmug=%d miss=%d problem=%d" % (
            name, len(lines), res["umug"].count("\n"), res["pmug"].count("\n"),
            res["miss"].count("\n"), res["problem"].count("\n")))
    for s in sessions.values():
        s.close()


if __name__ == "__main__":
    main()
