"""Generates the g8_plan_a_* cases (Plan_A_Matrix) by running the UNMODIFIED reference (build
container only):    python tests/golden/make_golden_plan_a.py

What is pinned: with a Plan_A_Matrix the reference generates the graph for the matrix labels only
(generate_neo4j_multi_hpf.py:101-192), imputes only subjects whose typed-locus pattern is a matrix row
(impute.py:1592-1596; the others go to .problem) and runs Plan A on that graph.  Cases either switch Plan B off
or hold only subjects that finish in Plan A: what the reference's Plan B does under a matrix is an accident of
vertex-list positions read as node ids (oracle/grim_oracle.py: PlanBUnderMatrix), and the product refuses it.
Every case is checked against the oracle before it is written: the oracle raises if a subject would have
reached Plan B, so a case that got through holds none."""
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))

import synth  # noqa: E402
from make_golden import LOCI9, NINE_OVER, POPS3, edge_lines, t1_lines  # noqa: E402
from refrun import RefSession  # noqa: E402

DATA = os.path.join(HERE, "data")
CASES = os.path.join(HERE, "cases")

M_BLOCKS = [[1, 2, 3, 4, 5], [1, 2, 3], [4, 5], [1], [2, 3], [4], [5], [2], [3]]
M_FOUR = [[1, 2, 3, 4, 5], [1, 2, 3, 4], [1, 2, 3, 5], [1, 2, 4, 5], [1, 3, 4, 5], [2, 3, 4, 5]]
M_NINE = [[1, 2, 3, 4, 5, 6, 7, 8, 9], [1, 2, 3, 7, 8], [1, 2, 3], [7, 8], [1, 2, 3, 4, 5, 6, 7, 8]]


def idx0(rows):
    return [[x - 1 for x in r] for r in rows]


def main():
    import grim_oracle as go
    base = json.load(open(os.path.join(DATA, "base_conf.json")))
    cau = open(os.path.join(DATA, "cau_hpf.csv")).read()
    cau_cnt = open(os.path.join(DATA, "cau_pop_counts.txt")).read()
    hpf3 = open(os.path.join(DATA, "pop3_hpf.csv")).read()
    cnt3 = open(os.path.join(DATA, "pop3_pop_counts.txt")).read()
    hpf9 = open(os.path.join(DATA, "nine_hpf.csv")).read()
    cnt9 = open(os.path.join(DATA, "nine_pop_counts.txt")).read()
    tab = synth.Table(cau)
    tab3 = synth.Table(hpf3, "AAA")
    tab9 = synth.Table(hpf9, "AAA", LOCI9)
    races3 = synth.race_fields(POPS3)
    over3 = {"populations": POPS3, "UNK_priors": "MR"}
    tables = {"cau": (cau, cau_cnt, {}), "pop3": (hpf3, cnt3, over3), "nine": (hpf9, cnt9, NINE_OVER)}
    in_and_out = idx0(M_BLOCKS) + [[0, 1], [0, 1, 2, 3], [2, 3, 4], [1, 2, 3, 4]]
    cases = [
        # Plan B on (the default): only subjects that finish in Plan A, or whose pattern is not a matrix row
        ("g8_plan_a_blocks", "cau", {"Plan_A_Matrix": M_BLOCKS},
         synth.typed_subjects(tab, 40, 81, ["CAU,CAU"]) + synth.subset_subjects(tab, 130, 82, in_and_out, ["CAU,CAU"])
         + [edge_lines(tab)[i] for i in (0, 5, 9, 10, 12, 13, 14)]),
        # Plan B off: everything is defined; messy subjects (ambiguity, missing loci, unknown alleles) incl. misses
        ("g8_plan_a_blocks_planb_off", "cau", {"Plan_A_Matrix": M_BLOCKS, "planb": False},
         synth.messy_subjects(tab, 120, 83) + synth.subset_subjects(tab, 60, 84, in_and_out, ["CAU,CAU"], amb=3)
         + edge_lines(tab) + t1_lines(tab)),
        ("g8_plan_a_four_pop3", "pop3", {"Plan_A_Matrix": M_FOUR, "planb": False},
         synth.messy_subjects(tab3, 100, 85, races=races3)
         + synth.subset_subjects(tab3, 60, 86, idx0(M_FOUR) + [[0, 1, 2]], races3, amb=2)),
        ("g8_plan_a_four_low_threshold", "cau", {"Plan_A_Matrix": M_FOUR, "planb": False, "number_of_options_threshold": 40},
         synth.messy_subjects(tab, 80, 87, max_amb=5) + synth.subset_subjects(tab, 40, 88, idx0(M_FOUR), ["CAU,CAU"], amb=4)),
        ("g8_plan_a_nine", "nine", {"Plan_A_Matrix": M_NINE, "planb": False},
         synth.typed_subjects(tab9, 8, 89, ["AAA,BBB", ","])
         + synth.subset_subjects(tab9, 40, 90, idx0(M_NINE) + [[0, 1]], ["AAA,BBB", ",", "AAA;BBB,XXX"], amb=1)
         + synth.messy_subjects(tab9, 12, 91, max_amb=2, p_missing=0.3, races=["AAA,BBB", ","])),
    ]
    # the last node of the last Plan-A label: its adjacency is cut by the CSR sentinel (nxg.py:195-196)
    last3 = None
    for hp in tab.haps:
        last3 = hp[2]                       # label "3" = locus index 2 is the last row of M_BLOCKS
    seen = []
    for hp in tab.haps:
        if hp[2] not in seen:
            seen.append(hp[2])
    last3 = seen[-1]
    other3 = seen[0]
    cases.append(("g8_plan_a_last_node", "cau", {"Plan_A_Matrix": M_BLOCKS, "planb": False},
                  ["L1,%s+%s\n" % (last3, other3), "L2,%s+%s\n" % (other3, other3), "L3,%s+%s\n" % (last3, last3),
                   "L4,%s+%s,CAU,CAU\n" % (seen[-2], last3)]))
    only = sys.argv[1:]
    for name, table, over, lines in cases:
        if only and name not in only:
            continue
        hpf, cnt, tover = tables[table]
        conf = dict(base)
        conf.update(tover)
        conf.update(over)
        s = RefSession(conf, hpf, cnt)
        res = s.run(lines)
        s.close()
        d = os.path.join(CASES, name)
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d)
        o = dict(tover)
        o.update(over)
        json.dump({"table": table, "overrides": o}, open(os.path.join(d, "case.json"), "w"), indent=1)
        open(os.path.join(d, "subjects.csv"), "w").writelines(lines)
        for k, v in res.items():
            open(os.path.join(d, "exp." + k), "w").write(v)
        print("%-30s %4d subjects  umug=%d pmug=%d miss=%d problem=%d" % (
            name, len(lines), res["umug"].count("\n"), res["pmug"].count("\n"),
            res["miss"].count("\n"), res["problem"].count("\n")))
        # the oracle must agree (and must not meet Plan B) before the case counts
        import goldenlib
        _t, gconf, glines, exp = goldenlib.load_case(name)
        mine, _e = go.impute_file(gconf, lines=glines)
        bad = [k for k in exp if mine[k] != exp[k]]
        print("    oracle:", "identical" if not bad else "DIFFERS in %s" % bad)


if __name__ == "__main__":
    main()
