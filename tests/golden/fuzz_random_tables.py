"""Three-way differential fuzz on RANDOM small frequency tables (build container only): the
unmodified reference, oracle/grim_oracle.py and the emulated kernel source (tests/emu/) must
write byte-identical files.  fuzz_vs_reference.py / fuzz_emu_vs_oracle.py use the README table
(3,380 haplotypes, most alleles rare); the tables here are small and dense -- a few alleles per
locus shared by most haplotypes -- so recombinant phases hit, top-link lists are long, the
top-K cap and the geno_seen rule bind, and the last-node quirk lands on common alleles.

    python tests/golden/fuzz_random_tables.py [n_tables] [n_subjects_per_case] [seed]
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")


def reference_worker(path_in, path_out):
    """Separate process: the reference's package is also called `grim`, so it cannot share an
    interpreter with the product's host code."""
    sys.path.insert(0, HERE)
    from refrun import RefSession
    job = json.load(open(path_in))
    sess = RefSession(job["conf"], job["hpf"], job["counts"])
    res = []
    for case in job["cases"]:
        lines, over = case[0], case[1]
        hpp = bool(case[2]) if len(case) > 2 else False
        masks = case[3] if len(case) > 3 else None
        t0 = time.time()
        out = sess.run(lines, hap_pop_pair=hpp, phase_masks=masks, **over)
        out["_seconds"] = time.time() - t0
        res.append(out)
    sess.close()
    json.dump(res, open(path_out, "w"))


if len(sys.argv) > 1 and sys.argv[1] == "--ref":
    reference_worker(sys.argv[2], sys.argv[3])
    sys.exit(0)

for p in (os.path.join(ROOT, "py-graph-imputation_b200"), os.path.join(ROOT, "oracle"), os.path.join(HERE, "..")):
    sys.path.insert(0, p)

import subprocess  # noqa: E402
import tempfile  # noqa: E402

import grim_oracle as go  # noqa: E402
import synth  # noqa: E402
from emu_backend import EmuGraph, emu_imputation  # noqa: E402
from grim.run_impute_def import load_config  # noqa: E402

BASE_CONF = json.load(open(os.path.join(HERE, "data", "base_conf.json")))
KEYS = ("umug", "umug_pops", "pmug", "pmug_pops", "miss", "problem")


def first_diff(a, b):
    a, b = a.splitlines(), b.splitlines()
    for i in range(max(len(a), len(b))):
        x = a[i] if i < len(a) else None
        y = b[i] if i < len(b) else None
        if x != y:
            return i, x, y
    return None


LOCI9 = ["A", "B", "C", "DPA1", "DPB1", "DQA1", "DQB1", "DRB1", "DRBX"]
NINE_OVER = {
    "populations": ["AAA", "BBB"], "UNK_priors": "MR", "freq_trim_threshold": 1e-9,
    "loci_map": {l: i + 1 for i, l in enumerate(LOCI9)},
    "Plan_B_Matrix": [[[1, 2, 3, 4, 5, 6, 7, 8, 9]], [[1, 2, 3], [4, 5], [6, 7, 8, 9]],
                      [[1], [2, 3], [4, 5], [6, 7], [8, 9]], [[1], [2], [3], [4], [5], [6], [7], [8], [9]]],
}


def main_nine(n_tables, n, seed):
    """Nine loci (BASELINE config 5 shape: 256 phases, 510 marginal labels, 9-block Plan-B matrix)."""
    rng = np.random.RandomState(seed)
    ok = True
    for t in range(n_tables):
        n_full = int(rng.choice([40, 100, 200]))
        n_alleles = [int(x) for x in (rng.randint(2, 5, size=9) if t % 2 == 0 else rng.randint(3, 15, size=9))]
        tseed = int(rng.randint(1, 1 << 30))
        pops = NINE_OVER["populations"]
        hpf = synth.zipf_table(n_full, n_alleles, tseed, loci=LOCI9, pops=tuple(pops))
        counts = "AAA,100.0,0.5\nBBB,100.0,0.5\n"
        conf = dict(BASE_CONF)
        conf.update(NINE_OVER)
        tab = synth.Table(hpf, pops[0], loci=LOCI9)
        print("== nine-locus table %d: %d haplotypes, alleles/locus %s (seed %d)" % (t, len(tab.haps), n_alleles, tseed), flush=True)
        og = go.OracleGraph(hpf.splitlines(True), pops, conf["loci_map"], conf["freq_trim_threshold"], counts.splitlines(True))
        eg = EmuGraph(og, conf["loci_map"])
        cbp = np.array([float(l.split(",")[2]) for l in counts.splitlines()])
        races = ["AAA,BBB", ",", "AAA;BBB,XXX"]
        cases = [
            ("typed", synth.typed_subjects(tab, n, tseed + 1, races), {}),
            ("messy", synth.messy_subjects(tab, n, tseed + 2, max_amb=2, p_missing=0.3, races=races[:2]), {}),
            ("messy save_space nres=3", synth.messy_subjects(tab, n, tseed + 3, max_amb=2, p_missing=0.4, races=races[:2]),
             {"save_space_mode": True, "number_of_results": 3}),
        ]
        ok &= run_cases(conf, hpf, counts, cases, og, eg, cbp)
    print("ALL OK" if ok else "SOME MISMATCH")
    return 0 if ok else 1


def run_cases(conf, hpf, counts, cases, og, eg, cbp, text_too=False):
    ok = True
    with tempfile.TemporaryDirectory() as td:
        json.dump({"conf": conf, "hpf": hpf, "counts": counts, "cases": [list(c[1:]) for c in cases]},
                  open(os.path.join(td, "in.json"), "w"))
        subprocess.run([sys.executable, os.path.abspath(__file__), "--ref", os.path.join(td, "in.json"),
                        os.path.join(td, "out.json")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        refs = json.load(open(os.path.join(td, "out.json")))
    for case, ref in zip(cases, refs):
        tag, lines, over = case[0], case[1], case[2]
        hpp = bool(case[3]) if len(case) > 3 else False
        masks = case[4] if len(case) > 4 else None
        c = dict(conf)
        c.update(over)
        mpath = None
        if masks is not None:
            fd, mpath = tempfile.mkstemp(suffix=".json")
            with os.fdopen(fd, "w") as f:
                json.dump(masks, f)
            c["bin_imputation_in_file"] = mpath
        t_r = ref.pop("_seconds")
        orc = go.OracleImputation(og, go.load_config(c), cbp).impute_lines(lines, em_mr=hpp)
        out = emu_imputation(eg, load_config(c), cbp, arena=1 << 30).impute_lines(lines, em_mr=hpp)   # the GPU's largest tier
        if mpath:
            os.unlink(mpath)
        txt = None
        if text_too:   # the C++ tokeniser / formatter (grimb_text.cpp) around the same emulated kernel source
            from emu_backend import emu_impute_text
            timp = emu_imputation(eg, load_config(c), cbp, arena=1 << 30)
            txt = emu_impute_text(timp, eg, "".join(lines).encode("utf8"), arena=1 << 30)
        emu = {k: "".join(v) for k, v in out.items()}
        bad_o = [k for k in KEYS if ref[k] != orc[k]]
        bad_e = [k for k in KEYS if ref[k] != emu[k]]
        print("   %-24s %4d subj  oracle %s  kernel %s  (reference %.1fs)  rows umug=%d pmug=%d miss=%d problem=%d" % (
            tag, len(lines), "OK" if not bad_o else "MISMATCH " + ",".join(bad_o),
            "OK" if not bad_e else "MISMATCH " + ",".join(bad_e), t_r, ref["umug"].count("\n"),
            ref["pmug"].count("\n"), ref["miss"].count("\n"), ref["problem"].count("\n")), flush=True)
        bad_t = [k for k in KEYS if txt is not None and ref[k] != txt[k]]
        if txt is not None:
            print("   %-24s        C++ text pipeline %s" % ("", "OK" if not bad_t else "MISMATCH " + ",".join(bad_t)), flush=True)
        for who, bad, mine in (("oracle", bad_o, orc), ("kernel", bad_e, emu), ("text", bad_t, txt)):
            for k in bad:
                i, x, y = first_diff(ref[k], mine[k])
                print("      first diff (%s) in %s line %d\n        ref : %r\n        mine: %r" % (who, k, i, x, y), flush=True)
        ok &= not bad_o and not bad_e and not bad_t
    return ok


def main_modes(n_tables, n, seed):
    """Output modes and configuration switches on random dense five-locus tables."""
    rng = np.random.RandomState(seed)
    ok = True
    for t in range(n_tables):
        n_full = int(rng.choice([30, 120, 500]))
        n_alleles = [int(x) for x in (rng.randint(2, 6, size=5) if t % 2 == 0 else rng.randint(3, 13, size=5))]
        pops = [["CAU"], ["AAA", "BBB"], ["AAA", "BBB", "CCC", "DDD"]][t % 3]
        tseed = int(rng.randint(1, 1 << 30))
        hpf = synth.zipf_table(n_full, n_alleles, tseed, pops=tuple(pops))
        cnt = 1000.0 / np.arange(1, len(pops) + 1) ** 1.1
        counts = "".join("%s,%s,%s\n" % (p, repr(float(c)), repr(float(c / cnt.sum()))) for p, c in zip(pops, cnt))
        conf = dict(BASE_CONF)
        conf["populations"] = pops
        conf["UNK_priors"] = "MR"
        tab = synth.Table(hpf, pops[0])
        races = synth.race_fields(pops) if len(pops) > 1 else None
        kw = {"races": races} if races else {}
        print("== modes table %d: %d haplotypes, alleles/locus %s, pops %s (seed %d)" % (t, len(tab.haps), n_alleles, pops, tseed),
              flush=True)
        og = go.OracleGraph(hpf.splitlines(True), pops, conf["loci_map"], conf["freq_trim_threshold"], counts.splitlines(True))
        eg = EmuGraph(og, conf["loci_map"])
        cbp = np.array([float(l.split(",")[2]) for l in counts.splitlines()])
        mixed = synth.typed_subjects(tab, n, tseed + 1, races) + synth.messy_subjects(tab, n, tseed + 2, **kw)
        mrng = np.random.RandomState(tseed % (1 << 31))

        def masks_for(lines):
            return {ln.split(",")[0]: [int(x) for x in mrng.randint(0, 2, size=4)] for ln in lines}

        cases = [
            ("planb off", mixed, {"planb": False}),
            ("umug only", mixed, {"output_haplotypes": False}),
            ("pmug only", mixed, {"output_MUUG": False}),
            ("save_space missing", synth.messy_subjects(tab, n, tseed + 3, p_missing=0.4, **kw), {"save_space_mode": True}),
            ("epsilon 1e-1", mixed, {"epsilon": 1e-1}),
            ("epsilon 1e-7 nres=1000", mixed, {"epsilon": 1e-7, "number_of_results": 1000}),
            ("SR priors", mixed, {"UNK_priors": "SR"}),
            ("priority eta>0", mixed, {"priority": {"alpha": 0.4, "eta": 0.01, "beta": 1e-3, "gamma": 1e-2, "delta": 0.3}}),
            ("hap_pop_pair", mixed, {}, True),
            ("hap_pop_pair nres=4", mixed, {"number_of_results": 4}, True),
            ("phase masks", mixed, {}, False, masks_for(mixed)),
            ("phase masks + hap_pop_pair", mixed, {}, True, masks_for(mixed)),
        ]
        ok &= run_cases(conf, hpf, counts, cases, og, eg, cbp)
    print("ALL OK" if ok else "SOME MISMATCH")
    return 0 if ok else 1


def main_matrix(n_tables, n, seed):
    """Plan_A_Matrix on random dense five-locus tables: random matrices (the full label first, a random
    set of marginal labels in random order), Plan B off -- everything the reference does is then defined -- plus a
    Plan-B-on case restricted to subjects that finish in Plan A (checked: the oracle refuses otherwise)."""
    import itertools
    rng = np.random.RandomState(seed)
    ok = True
    all_labels = [list(c) for r in range(4, 0, -1) for c in itertools.combinations([1, 2, 3, 4, 5], r)]
    for t in range(n_tables):
        n_full = int(rng.choice([30, 120, 500]))
        n_alleles = [int(x) for x in (rng.randint(2, 6, size=5) if t % 2 == 0 else rng.randint(3, 13, size=5))]
        pops = [["CAU"], ["AAA", "BBB"], ["AAA", "BBB", "CCC", "DDD"]][t % 3]
        tseed = int(rng.randint(1, 1 << 30))
        hpf = synth.zipf_table(n_full, n_alleles, tseed, pops=tuple(pops))
        cnt = 1000.0 / np.arange(1, len(pops) + 1) ** 1.1
        counts = "".join("%s,%s,%s\n" % (p, repr(float(c)), repr(float(c / cnt.sum()))) for p, c in zip(pops, cnt))
        k = int(rng.randint(1, 9))
        pick = [all_labels[i] for i in rng.choice(len(all_labels), size=k, replace=False)]
        matrix = [[1, 2, 3, 4, 5]] + pick
        conf = dict(BASE_CONF)
        conf.update({"populations": pops, "UNK_priors": "MR", "Plan_A_Matrix": matrix, "planb": False})
        tab = synth.Table(hpf, pops[0])
        races = synth.race_fields(pops) if len(pops) > 1 else None
        kw = {"races": races} if races else {}
        print("== matrix table %d: %d haplotypes, alleles/locus %s, pops %s, matrix %s (seed %d)" % (
            t, len(tab.haps), n_alleles, pops, matrix, tseed), flush=True)
        og = go.OracleGraph(hpf.splitlines(True), pops, conf["loci_map"], conf["freq_trim_threshold"], counts.splitlines(True),
                            plan_a_matrix=matrix)
        eg = EmuGraph(og, conf["loci_map"])
        cbp = np.array([float(l.split(",")[2]) for l in counts.splitlines()])
        subsets = [[x - 1 for x in row] for row in matrix] + [[0, 1], [2, 3, 4]]
        mixed = (synth.typed_subjects(tab, n, tseed + 1, races) + synth.messy_subjects(tab, n, tseed + 2, **kw)
                 + synth.subset_subjects(tab, 2 * n, tseed + 3, subsets, races, amb=2))
        # Plan B on: keep the subjects the oracle finishes in Plan A (it raises on the first that would not)
        plan_a_only = []
        oimp = go.OracleImputation(og, go.load_config(dict(conf, planb=True)), cbp)
        for ln in mixed:
            try:
                oimp.impute_lines([ln])
                plan_a_only.append(ln)
            except go.PlanBUnderMatrix:
                pass
        cases = [
            ("matrix planb off", mixed, {}),
            ("matrix low threshold", mixed, {"number_of_options_threshold": 40}),
            ("matrix pmug only nres=3", mixed, {"output_MUUG": False, "number_of_results": 3}),
            ("matrix planb on (Plan A only)", plan_a_only, {"planb": True}),
        ]
        ok &= run_cases(conf, hpf, counts, cases, og, eg, cbp, text_too=True)
    print("ALL OK" if ok else "SOME MISMATCH")
    return 0 if ok else 1


def main_heavy(n, seed, rounds):
    """Highly ambiguous subjects on the README table: 6-40 alleles per locus side, 0-3
    missing loci, products on both sides of number_of_options_threshold (default 100,000 and 2,000)."""
    cau = open(os.path.join(HERE, "data", "cau_hpf.csv")).read()
    counts = "CAU,3380.0,1.0\n"
    conf = dict(BASE_CONF)
    tab = synth.Table(cau)
    og = go.OracleGraph(cau.splitlines(True), conf["populations"], conf["loci_map"], conf["freq_trim_threshold"],
                        counts.splitlines(True))
    eg = EmuGraph(og, conf["loci_map"])
    cbp = np.array([3380.0 / 3380.0])
    ok = True
    for r in range(rounds):
        print("== heavy round %d (seed %d)" % (r, seed + 100 * r), flush=True)
        cases = [
            ("heavy thr=100000", synth.heavy_subjects(tab, n, seed + 100 * r), {}),
            ("heavy thr=2000", synth.heavy_subjects(tab, n, seed + 100 * r + 1, sizes=(3, 6, 12)),
             {"number_of_options_threshold": 2000}),
            ("heavy thr=2000 topk=10 nres=5", synth.heavy_subjects(tab, n, seed + 100 * r + 2, sizes=(3, 6, 12)),
             {"number_of_options_threshold": 2000, "max_haplotypes_number_in_phase": 10, "number_of_results": 5}),
        ]
        ok &= run_cases(conf, cau, counts, cases, og, eg, cbp)
    print("ALL OK" if ok else "SOME MISMATCH")
    return 0 if ok else 1


def dirty(lines, rng, loci, crlf):
    """Mutations of well-formed subject lines toward what real input files contain (SURVEY 8c edge set)."""
    out = []
    for ln in lines:
        f = ln.rstrip("\n").split(",")
        sid, gl, race = f[0], f[1], f[2:]
        loc = gl.split("^")
        for _ in range(int(rng.randint(0, 3))):
            m = int(rng.randint(0, 9))
            if not loc:
                break
            k = int(rng.randint(len(loc)))
            if m == 0:      # 'g' / 'L' suffixes (clean_up_gl deletes the characters)
                a = loc[k].split("+")
                a[int(rng.randint(len(a)))] += "gL"[int(rng.randint(2))]
                loc[k] = "+".join(a)
            elif m == 1:    # untyped locus written as UUUU
                name = loc[k].split("*")[0]
                loc[k] = "%s*UUUU+%s*UUUU" % (name, name)
            elif m == 2:    # a locus without '+': unparsable
                loc[k] = loc[k].split("+")[0]
            elif m == 3:    # loci in another order
                rng.shuffle(loc)
            elif m == 4:    # race field variants
                race = [["", ""], ["XXX", "CAU"], ["CAU"], [], ["CAU;XXX", "CAU"], ["CAU", "CAU", ""], ["AAA;BBB", "XXX"],
                        ["BBB", ""]][int(rng.randint(8))]
            elif m == 5:    # fully homozygous
                loc = ["+".join([x.split("+")[0]] * 2) for x in loc]
            elif m == 6:    # empty GL string
                loc = []
            elif m == 7:    # one side empty
                loc[k] = loc[k].split("+")[0] + "+"
            elif m == 8:    # locus dropped
                if len(loc) > 1:
                    del loc[k]
        out.append(",".join([sid, "^".join(loc)] + race) + ("\r\n" if crlf else "\n"))
    return out


def main_dirty(n_tables, n, seed, crlf=False):
    rng = np.random.RandomState(seed)
    ok = True
    for t in range(n_tables):
        n_full = int(rng.choice([30, 120, 500]))
        n_alleles = [int(x) for x in (rng.randint(2, 6, size=5) if t % 2 == 0 else rng.randint(3, 13, size=5))]
        pops = [["CAU"], ["AAA", "BBB"]][t % 2]
        tseed = int(rng.randint(1, 1 << 30))
        hpf = synth.zipf_table(n_full, n_alleles, tseed, pops=tuple(pops))
        cnt = 1000.0 / np.arange(1, len(pops) + 1) ** 1.1
        counts = "".join("%s,%s,%s\n" % (p, repr(float(c)), repr(float(c / cnt.sum()))) for p, c in zip(pops, cnt))
        conf = dict(BASE_CONF)
        conf["populations"] = pops
        conf["UNK_priors"] = "MR"
        tab = synth.Table(hpf, pops[0])
        races = synth.race_fields(pops) if len(pops) > 1 else ["CAU,CAU"]
        print("== dirty table %d: %d haplotypes, alleles/locus %s, pops %s (seed %d)" % (t, len(tab.haps), n_alleles, pops, tseed),
              flush=True)
        og = go.OracleGraph(hpf.splitlines(True), pops, conf["loci_map"], conf["freq_trim_threshold"], counts.splitlines(True))
        eg = EmuGraph(og, conf["loci_map"])
        cbp = np.array([float(l.split(",")[2]) for l in counts.splitlines()])
        base = synth.typed_subjects(tab, n, tseed + 1, races) + synth.messy_subjects(tab, n, tseed + 2, races=races)
        drng = np.random.RandomState(tseed % (1 << 31))
        cases = [("dirty", dirty(base, drng, synth.LOCI5, crlf), {}),
                 ("dirty umug only", dirty(base, drng, synth.LOCI5, crlf), {"output_haplotypes": False})]
        ok &= run_cases(conf, hpf, counts, cases, og, eg, cbp, text_too=True)
    print("ALL OK" if ok else "SOME MISMATCH")
    return 0 if ok else 1


def main():
    if len(sys.argv) > 1 and sys.argv[1] in ("--dirty", "--dirty-crlf"):
        a = sys.argv[2:]
        return main_dirty(int(a[0]) if a else 4, int(a[1]) if len(a) > 1 else 20, int(a[2]) if len(a) > 2 else 1,
                          crlf=sys.argv[1] == "--dirty-crlf")
    if len(sys.argv) > 1 and sys.argv[1] == "--matrix":
        a = sys.argv[2:]
        return main_matrix(int(a[0]) if a else 6, int(a[1]) if len(a) > 1 else 20, int(a[2]) if len(a) > 2 else 1)
    if len(sys.argv) > 1 and sys.argv[1] == "--heavy":
        a = sys.argv[2:]
        return main_heavy(int(a[0]) if a else 10, int(a[1]) if len(a) > 1 else 1, int(a[2]) if len(a) > 2 else 1)
    if len(sys.argv) > 1 and sys.argv[1] == "--modes":
        a = sys.argv[2:]
        return main_modes(int(a[0]) if a else 4, int(a[1]) if len(a) > 1 else 15, int(a[2]) if len(a) > 2 else 1)
    if len(sys.argv) > 1 and sys.argv[1] == "--nine":
        a = sys.argv[2:]
        return main_nine(int(a[0]) if a else 4, int(a[1]) if len(a) > 1 else 10, int(a[2]) if len(a) > 2 else 1)
    n_tables = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    rng = np.random.RandomState(seed)
    ok = True
    for t in range(n_tables):
        n_full = int(rng.choice([30, 120, 500]))
        # every other table is dense: 2-5 alleles per locus, so most recombinants exist in the table
        n_alleles = [int(x) for x in (rng.randint(2, 6, size=5) if t % 2 == 0 else rng.randint(3, 13, size=5))]
        pops = [["CAU"], ["AAA", "BBB"], ["AAA", "BBB", "CCC", "DDD"]][int(rng.randint(0, 3))]
        if os.environ.get("FUZZ_POPS"):   # e.g. FUZZ_POPS=21: BASELINE config 3 shape (21 populations, top-100 pop rows)
            pops = ["P%02d" % i for i in range(int(os.environ["FUZZ_POPS"]))]
        tseed = int(rng.randint(1, 1 << 30))
        hpf = synth.zipf_table(n_full, n_alleles, tseed, pops=tuple(pops))
        cnt = 1000.0 / np.arange(1, len(pops) + 1) ** 1.1
        counts = "".join("%s,%s,%s\n" % (p, repr(float(c)), repr(float(c / cnt.sum()))) for p, c in zip(pops, cnt))
        conf = dict(BASE_CONF)
        conf["populations"] = pops
        conf["UNK_priors"] = "MR" if len(pops) > 1 else conf.get("UNK_priors", "MR")
        tab = synth.Table(hpf, pops[0])
        races = synth.race_fields(pops) if len(pops) > 1 else None
        print("== table %d: %d haplotypes, alleles/locus %s, pops %s (seed %d)" % (t, len(tab.haps), n_alleles, pops, tseed),
              flush=True)
        og = go.OracleGraph(hpf.splitlines(True), pops, conf["loci_map"], conf["freq_trim_threshold"],
                            counts.splitlines(True))
        eg = EmuGraph(og, conf["loci_map"])
        cbp = np.array([float(l.split(",")[2]) for l in counts.splitlines()])
        cases = [
            ("typed", synth.typed_subjects(tab, n, tseed + 1, races), {}),
            ("messy", synth.messy_subjects(tab, n, tseed + 2, races=races) if races else synth.messy_subjects(tab, n, tseed + 2), {}),
            ("messy thr=40 topk=7", synth.messy_subjects(tab, n, tseed + 3, max_amb=5, races=races) if races
             else synth.messy_subjects(tab, n, tseed + 3, max_amb=5),
             {"number_of_options_threshold": 40, "max_haplotypes_number_in_phase": 7}),
            ("unknown heavy nres=3", synth.messy_subjects(tab, n, tseed + 4, p_unknown=0.3, p_random=0.3, races=races) if races
             else synth.messy_subjects(tab, n, tseed + 4, p_unknown=0.3, p_random=0.3),
             {"number_of_results": 3, "number_of_pop_results": 2}),
        ]
        with tempfile.TemporaryDirectory() as td:
            json.dump({"conf": conf, "hpf": hpf, "counts": counts, "cases": [[l, o] for _t, l, o in cases]},
                      open(os.path.join(td, "in.json"), "w"))
            subprocess.run([sys.executable, os.path.abspath(__file__), "--ref", os.path.join(td, "in.json"),
                            os.path.join(td, "out.json")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            refs = json.load(open(os.path.join(td, "out.json")))
        for (tag, lines, over), ref in zip(cases, refs):
            c = dict(conf)
            c.update(over)
            t_r = ref.pop("_seconds")
            orc = go.OracleImputation(og, go.load_config(c), cbp).impute_lines(lines)
            out = emu_imputation(eg, load_config(c), cbp).impute_lines(lines)
            emu = {k: "".join(v) for k, v in out.items()}
            bad_o = [k for k in KEYS if ref[k] != orc[k]]
            bad_e = [k for k in KEYS if ref[k] != emu[k]]
            print("   %-24s %4d subj  oracle %s  kernel %s  (reference %.1fs)  rows umug=%d pmug=%d miss=%d problem=%d" % (
                tag, len(lines), "OK" if not bad_o else "MISMATCH " + ",".join(bad_o),
                "OK" if not bad_e else "MISMATCH " + ",".join(bad_e), t_r, ref["umug"].count("\n"),
                ref["pmug"].count("\n"), ref["miss"].count("\n"), ref["problem"].count("\n")), flush=True)
            for who, bad, mine in (("oracle", bad_o, orc), ("kernel", bad_e, emu)):
                for k in bad:
                    i, x, y = first_diff(ref[k], mine[k])
                    print("      first diff (%s) in %s line %d\n        ref : %s\n        mine: %s" % (who, k, i, x, y), flush=True)
            ok &= not bad_o and not bad_e
    print("ALL OK" if ok else "SOME MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
