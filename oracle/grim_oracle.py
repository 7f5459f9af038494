"""CPU oracle for the GRIM per-subject imputation hot path.

TEST INFRASTRUCTURE ONLY.  This module is a plain-Python restatement of the
reference algorithm (nmdp-bioinformatics/py-graph-imputation).  It exists so
that the CUDA path can be checked on a GPU box where /root/reference is absent.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import it; the product package never does.

Parity status: PINNED.  The reference's own tests hold no result-bearing
vectors (tests/unit/test_grim.py is a placeholder), so the oracle is pinned
against outputs of the reference itself, run in the build container:
tests/golden/make_golden.py imports /root/reference, generates the fixtures in
tests/golden/, and tests/test_oracle_golden.py checks this restatement against
every one of them (file-level equality).

Each function cites the reference lines it follows (paths relative to the
reference root; "impute.py" = grim/imputation/impute.py, "nxg.py" =
grim/imputation/networkx_graph.py, "gen.py" =
graph_generation/generate_neo4j_multi_hpf.py).  Nothing here is tuned for speed
beyond what the reference itself does; it is the CPU baseline ("port") in
bench.py.
"""
from __future__ import annotations

import itertools
import json
import os
from collections import defaultdict

import numpy as np

BLOCK_FACTOR = 0.0001  # impute.py:196 (self.factor)
NEVER = 10             # impute.py:1410 "row never found" sentinel


# ---------------------------------------------------------------------------------------------
# Frequency store: hpf.csv -> nodes / top-links / connectors
# ---------------------------------------------------------------------------------------------
class _Fault:
    """Marks an adjacency whose reference range runs past the edge array (IndexError)."""


def _sentinel_slice(own_adj, n_edges, n_vertices):
    # nxg.py:195-196: the closing CSR sentinel is len(Vertices), not len(Edges); the last
    # vertex therefore owns range(start_last, n_vertices) of the sorted edge array.
    start = n_edges - len(own_adj)
    if n_vertices <= start:
        return []
    if n_vertices > n_edges:
        return _Fault
    return own_adj[: n_vertices - start]


def _checked(adj):
    if adj is _Fault:
        raise IndexError("CSR sentinel range past the edge array")
    return adj


class PlanBUnderMatrix(NotImplementedError):
    """A subject leaves Plan A while a Plan_A_Matrix is in force.  The reference's Plan B then indexes its
    second vertex list (nxg.py:32-66: nodes of the Plan-B labels only) with node ids of the CSV files
    (nxg.py:91-130,285-306): it reads the adjacency of unrelated nodes, and the ids involved depend on
    PYTHONHASHSEED (gen.py:181).  Nothing to be exact against: the oracle refuses instead of guessing."""


def plan_a_labels(matrix, full_label):
    """gen.py:101-164 (labels_for_grap), the Plan-A part: label strings of the matrix rows in matrix order, the
    full label appended when absent.  Only what has a defined behaviour downstream is accepted: rows strictly
    ascending (input_type, impute.py:1574-1579, yields ascending lists, and a permuted row would also permute the
    alleles inside the node names, gen.py:367-368) and the full label FIRST (full haplotypes always get ids
    0..N-1, gen.py:341-358, while nodes.csv lists labels in matrix order: with the full label anywhere else the
    position of a vertex in nxg.py's filtered list no longer equals its id, nxg.py:49-57 vs :262-270)."""
    labels = []
    for row in matrix:
        row = [int(x) for x in row]
        if row != sorted(set(row)) or not row or any(str(x) not in full_label for x in row):
            raise NotImplementedError("Plan_A_Matrix rows must be strictly ascending lists of loci_map indices")
        lab = "".join(str(x) for x in row)
        if lab in labels:
            raise NotImplementedError("Plan_A_Matrix lists a label twice")
        labels.append(lab)
    if labels[0] != full_label:
        raise NotImplementedError("Plan_A_Matrix must list the full label first (see plan_a_labels)")
    if len(labels) < 2:
        # nxg.py:134: np.vstack of an empty edge list
        raise ValueError("need at least one array to concatenate")
    return labels


class OracleGraph:
    """Restates gen.py:209-486 (trim, node ids, sequential marginal sums, top links, parent
    edges) fused with nxg.py:42-213 (dict + CSR load, including the sentinel quirk at
    :195-196), without going through the CSV files."""

    def __init__(self, hpf_lines, pops, loci_map, freq_trim, pop_count_lines=None, marginals=True, plan_a_matrix=None):
        # marginals=False (checker speed only, not a reference feature): keep just the full
        # haplotype label.  Exact for subjects typed at every locus whose Plan A succeeds -- the
        # only queries are full-label lookups (nxg.py:262-266) -- which is how the big 9-locus
        # configuration is sampled; callers must check that every sampled subject used Plan A.
        self.pops = list(pops)
        self.loci_map = {k: int(v) for k, v in loci_map.items()}
        nloc = len(self.loci_map)
        # gen.py:195-206 loci_order
        self.full_label = "".join(sorted({str(v) for v in self.loci_map.values()}))
        full = self.full_label
        # gen.py:105-110 label order: full, then combinations of decreasing size
        self.labels = [full]
        for r in range(len(full) - 1, 0, -1):
            self.labels.extend("".join(c) for c in itertools.combinations(full, r))
        if not marginals:
            self.labels = [full]
        # Plan_A_Matrix (gen.py:101-192, nxg.py:32-66): Plan A sees the labels of the matrix only, ids in matrix
        # order.  The single-locus labels always exist in the reference's second vertex list (gen.py:166-170), which
        # is where allele-existence checks look (nxg.py:309-316); they follow the Plan-A labels here and take no part
        # in the CSR.  No connectors: Plan B is refused under a matrix (PlanBUnderMatrix).
        self.plan_a = None
        if plan_a_matrix:
            self.plan_a = plan_a_labels(plan_a_matrix, full)
            self.labels = self.plan_a + [ch for ch in full if ch not in self.plan_a]

        # gen.py:259-266 trim threshold per population
        trim = {}
        if pop_count_lines is None:
            for p in self.pops:
                trim[p] = freq_trim
        else:
            for line in pop_count_lines:
                p, cnt, _ratio = line.strip().split(",")
                trim[p] = freq_trim / float(cnt)

        # gen.py:320-339 read hpf rows
        seen = {}
        pop_hap = {}
        for line in hpf_lines:
            if not line:
                continue
            hap, pop, freq = line.split(",")
            if hap == "hap":
                continue
            freq = float(freq)
            if freq == 0.0:
                continue
            if freq < trim[pop]:
                continue
            alleles = self._canon(hap, nloc)
            name = "~".join(alleles)
            seen[name] = alleles
            pop_hap[(pop, name)] = freq

        # gen.py:341-358 full nodes, ids in first-appearance order
        self.names = []            # id -> name
        self.node = {}             # name -> (label, vec, id)
        self.by_label = {lab: [] for lab in self.labels}
        full_alleles = []
        full_vecs = []
        for name, alleles in seen.items():
            vec = [pop_hap.get((p, name), 0.0) for p in self.pops]
            self._add(name, full, vec)
            full_alleles.append(alleles)
            full_vecs.append(vec)
        self.n_full = len(full_alleles)

        # gen.py:364-415 marginal nodes: sequential sums in full-haplotype order; top links
        tl = {}
        for lab in self.labels[1:]:
            pos = [full.index(ch) for ch in lab]
            acc = {}
            for fid in range(self.n_full):
                al = full_alleles[fid]
                name = "~".join([al[i] for i in pos])
                cur = acc.get(name)
                if cur is None:
                    acc[name] = [0.0 + x for x in full_vecs[fid]]
                    tl[name] = [fid]
                else:
                    acc[name] = [a + b for a, b in zip(cur, full_vecs[fid])]
                    tl[name].append(fid)
            for name, vec in acc.items():
                self._add(name, lab, vec)
        self.n_nodes = len(self.names)

        # nxg.py:71-88,149-201 CSR of top links (partial -> full, ascending id)
        self.toplinks = tl
        if self.plan_a is not None:
            # the CSR holds the Plan-A vertices only (nxg.py:49-57,71-88)
            n_vertices = sum(len(self.by_label[lab]) for lab in self.plan_a)
            n_edges = sum(len(tl[n]) for lab in self.plan_a[1:] for n in self.by_label[lab])
            last = self.names[n_vertices - 1]
            self.toplinks[last] = _sentinel_slice(tl[last], n_edges, n_vertices)
            self.n_plan_a_nodes = n_vertices
            self.conn = {}
            return
        n_edges = sum(len(v) for v in tl.values())
        if tl:
            last = self.names[-1]
            self.toplinks[last] = _sentinel_slice(tl[last], n_edges, self.n_nodes)

        # nxg.py:91-130 connectors: (parent label, child name) -> parents (ascending id)
        conn = defaultdict(list)
        for lab in self.labels:
            if len(lab) < 2 or not marginals:
                continue
            for name in self.by_label[lab]:
                al = name.split("~")
                nid = self.node[name][2]
                for drop in range(len(al)):
                    child = "~".join(al[:drop] + al[drop + 1:])
                    conn[(lab, child)].append(nid)
        self.conn = dict(conn)
        if self.conn:
            # The last connector created while scanning edges.csv: rows are grouped by child
            # label (gen.py:439-455) and, per child, by the order parents were appended
            # (gen.py:82-98: `list(set difference)`), so it belongs to the last node of the
            # last single-locus label.  Its CSR range is cut by the same sentinel quirk.
            last_child = self.by_label[self.labels[-1]][-1]
            own = full.index(self.labels[-1])
            others = list(set(range(len(full))).difference([own]))
            plab = "".join(full[i] for i in sorted([own, others[-1]]))
            n_conn = len(self.conn)
            n_whole_edges = n_conn + sum(len(v) for v in self.conn.values())
            key = (plab, last_child)
            self.conn[key] = _sentinel_slice(self.conn[key], n_whole_edges, self.n_nodes + n_conn)

    def _canon(self, hap, nloc):
        # gen.py:59-68 make_allele_list: strip one trailing 'g', order by loci_map index
        out = ["0"] * nloc
        for a in hap.split("~"):
            if a[-1] == "g":
                a = a[:-1]
            out[self.loci_map[a.split("*")[0]] - 1] = a
        return out

    def _add(self, name, label, vec):
        self.node[name] = (label, vec, len(self.names))
        self.names.append(name)
        self.by_label[label].append(name)

    # ---- queries (nxg.py:215-321) ----
    def haps_by_label(self, label):
        return self.by_label.get(label, [])

    def haps_with_probs_by_label(self, label):
        return {n: self.node[n][1] for n in self.by_label.get(label, [])}

    def adjs_query(self, names):
        # nxg.py:253-278
        out = {}
        for n in names:
            hit = self.node.get(n)
            if hit is None:
                continue
            if hit[0] == self.full_label:
                out[n] = hit[1]
            else:
                for fid in _checked(self.toplinks[n]):
                    fn = self.names[fid]
                    out[fn] = self.node[fn][1]
        return out

    def adjs_query_by_color(self, names, label_a, label_b):
        # nxg.py:280-307
        if label_a == label_b:
            return self.node_probs(names)
        out = {}
        for n in names:
            if n in self.node:
                for pid in _checked(self.conn.get((label_b, n), [])):
                    pn = self.names[pid]
                    out[pn] = self.node[pn][1]
        return out

    def node_probs(self, names):
        # nxg.py:309-321
        return {n: self.node[n][1] for n in names if n in self.node}


# ---------------------------------------------------------------------------------------------
# Configuration (grim/run_impute_def.py:54-129)
# ---------------------------------------------------------------------------------------------
def load_config(json_conf):
    """Defaults as in run_impute_def.py:63-129."""
    c = json_conf
    return {
        "planb": c.get("planb", True),
        "pops": c.get("populations"),
        "priority": c.get("priority"),
        "epsilon": c.get("epsilon", 1e-3),
        "number_of_results": c.get("number_of_results", 1000),
        "number_of_pop_results": c.get("number_of_pop_results", 100),
        "output_MUUG": c.get("output_MUUG", True),
        "output_haplotypes": c.get("output_haplotypes", False),
        "factor_missing_data": c.get("factor_missing_data", 0.01),
        "loci_map": c.get("loci_map", {"A": 1, "B": 3, "C": 2, "DQB1": 4, "DRB1": 5}),
        "matrix_planb": c.get(
            "Plan_B_Matrix",
            [
                [[1, 2, 3, 4, 5]],
                [[1, 2, 3], [4, 5]],
                [[1], [2, 3], [4, 5]],
                [[1, 2, 3], [4], [5]],
                [[1], [2, 3], [4], [5]],
                [[1], [2], [3], [4], [5]],
            ],
        ),
        "number_of_options_threshold": c.get("number_of_options_threshold", 100000),
        "max_haplotypes_number_in_phase": c.get("max_haplotypes_number_in_phase", 100),
        "save_mode": c.get("save_space_mode", False),
        "UNK_priors": c.get("UNK_priors", "MR"),
        # run_impute_def.py:124-125: optional JSON {subject id: [0/1 per typed position]} phase masks
        "bin_imputation_input_file": c.get("bin_imputation_in_file", "None"),
        "nodes_for_plan_A": c.get("Plan_A_Matrix", []),      # run_impute_def.py:126
    }


def graph_from_config(json_conf, base_dir="", marginals=True):
    """hpf.csv (+ pop counts file) -> OracleGraph, following gen.py:240-266."""
    with open(os.path.join(base_dir, json_conf["freq_file"])) as f:
        hpf = f.readlines()
    pc_path = os.path.join(base_dir, json_conf.get("pops_count_file", ""))
    pc = None
    if os.path.isfile(pc_path):
        with open(pc_path) as f:
            pc = f.readlines()
    return OracleGraph(
        hpf,
        json_conf["populations"],
        json_conf.get("loci_map"),
        json_conf["freq_trim_threshold"],
        pc,
        marginals,
        json_conf.get("Plan_A_Matrix") or None,
    )


def count_by_prob_from_file(n_pops, path):
    # impute.py:205-210
    out = np.ones(n_pops)
    with open(path) as f:
        for i, line in enumerate(f):
            out[i] = float(line.strip().split(",")[2])
    return out


# ---------------------------------------------------------------------------------------------
# Per-subject imputation
# ---------------------------------------------------------------------------------------------
def clean_gl(gl):
    # impute.py:105-118
    gl = gl.replace("g", "").replace("L", "")
    loci = gl.split("^")
    for bad in [x for x in loci if x.strip("UUUU") != x]:
        loci.remove(bad)
    return "^".join(loci)


class _Acc:
    """Accumulators shared by all phases of one evaluation (impute.py:717-723,663-671)."""

    __slots__ = ("seen", "geno", "pops", "maxp", "pairs", "pair_pops", "pair_probs")

    def __init__(self):
        self.seen = set()
        self.geno = {}
        self.pops = {}
        self.maxp = 0
        self.pairs = []
        self.pair_pops = []
        self.pair_probs = []


class OracleImputation:
    def __init__(self, graph, config, count_by_prob=None):
        # impute.py:147-221
        self.g = graph
        self.cfg = config
        self.pops = config["pops"]
        npop = len(self.pops)
        self.M = np.ones((npop, npop))
        self.unk_priors = config["UNK_priors"]
        self.loci = list(config["loci_map"].keys())
        self.index = {k: int(v) for k, v in config["loci_map"].items()}
        self.index_str = {k: str(v) for k, v in self.index.items()}   # cypher_query.py:19-21
        self.full_label = "".join(sorted(set(self.index_str.values())))
        self.all_indices = list(set(self.index.values()))              # impute.py:1118
        self.fmd = config["factor_missing_data"]
        self.matrix = config["matrix_planb"]
        self.count_by_prob = np.ones(npop) if count_by_prob is None else count_by_prob
        self.threshold = config["number_of_options_threshold"]
        self.top_k = config["max_haplotypes_number_in_phase"]
        self.save_space = config["save_mode"]
        self.plan = "a"
        self.pair_evals = 0  # instrumentation: iterations reaching impute.py:464 / :573
        self.binary = None   # per-subject phase mask of the subject in flight (impute.py:2030-2032)
        self.em = False      # impute_file(em=True): no Plan C for the haplotype output (impute.py:1648)

    # ---- GL string -> phases ----
    def gl2haps(self, gl):
        # impute.py:246-272
        if gl == "" or gl == " ":
            return []
        parts = gl.split("^")
        n = len(parts)
        t1, t2, empty = [], [], 0
        for p in parts:
            if p[0] == "+":
                p = p[1:]
            sides = p.split("+")
            if len(sides) == 1:
                if sides == [""]:
                    empty += 1
                    continue
                return []
            t1.append(sides[0])
            t2.append(sides[1])
        return {"Genotype": [sorted(t1), sorted(t2)], "N_Loc": n - empty}

    def gen_phases(self, gen, n_loci, b_phases=None):
        # impute.py:274-303; b_phases: positions whose flip is allowed (impute.py:277-290)
        out, seen = [], set()
        allowed = None if b_phases is None else [i for i, e in enumerate(b_phases) if e == 1]
        for i in range(2 ** (n_loci - 1)):
            pick = [(i >> m) & 1 for m in range(n_loci)]
            if allowed is not None:
                pick = [v if (v == 0 or m in allowed) else 0 for m, v in enumerate(pick)]
            h1 = [gen[pick[k]][k] for k in range(n_loci)]
            h2 = [gen[1 - pick[k]][k] for k in range(n_loci)]
            a = "~".join(h1) + "^" + "~".join(h2)
            b = "~".join(h2) + "^" + "~".join(h1)
            if a not in seen or b not in seen:
                seen.add(a)
                seen.add(b)
                out.append([h1, h2])
        return out

    def _typed_label(self, side):
        # impute.py:952-963
        present = []
        for locus, name in self.index_str.items():
            if any(s.split("*", 1)[0] == locus for s in side):
                present.append(name)
        return "".join(sorted(present))

    def open_phases(self, pmags, n_loci):
        # impute.py:914-989 with cutils.pyx:6-31 (Cartesian product, last locus fastest) and
        # cutils.pyx:35-51 (filter of all haplotypes of the typed label)
        out = []
        for ph in pmags:
            opened = []
            for side in ph:
                splits = [s.split("/") for s in side]
                options = 1
                for i in range(n_loci):
                    options *= len(splits[i])
                if options < self.threshold:
                    cands = [list(t) for t in itertools.product(*splits)]
                else:
                    allowed = set()
                    for s in splits:
                        allowed.update(s)
                    cands = []
                    for name in self.g.haps_by_label(self._typed_label(side)):
                        al = name.split("~")
                        if len(al) == n_loci and all(a in allowed for a in al):
                            cands.append(al)
                opened.append(cands)
            if opened[0] and opened[1]:
                out.append(opened)
        return out

    def _alleles_exist(self, alleles):
        # impute.py:1218-1222 -> 1207-1216 -> nxg.py:309-321 (single-locus label == label)
        _ = self.index[alleles[0].split("*")[0]]
        return self.g.node_probs(alleles)

    def _reduce_valid(self, pmags, n_loci, planc=False):
        # impute.py:864-879
        for ph in pmags:
            for side in ph:
                options = 1
                for i in range(n_loci):
                    options *= len(side[i].split("/"))
                if options >= self.threshold or planc:
                    for i, g in enumerate(side):
                        found = self._alleles_exist(g.split("/"))
                        if found:
                            side[i] = "/".join(found.keys())

    def _reduce_common(self, pmags, n_loci, keep, planc=False):
        # impute.py:881-912
        for ph in pmags:
            for side in ph:
                options = 1
                for i in range(n_loci):
                    options *= len(side[i].split("/"))
                if options >= self.threshold or planc:
                    for i, g in enumerate(side):
                        found = self._alleles_exist(g.split("/"))
                        if found:
                            score = {}
                            for a, vec in found.items():
                                s = 0
                                for p, f in enumerate(vec):
                                    s += f * self.M[p, p]
                                score[a] = s
                            best = sorted(score.items(), key=lambda kv: kv[1], reverse=True)[:keep]
                            side[i] = "/".join(a for a, _ in best)

    # ---- flatten + pair evaluation ----
    def _flatten(self, probs):
        # impute.py:424-442
        items = []
        M = self.M
        for k in range(len(probs)):
            vec = probs[k]
            for j in range(len(vec)):
                if vec[j] > 0:
                    items.append((vec[j] * M[j][j], vec[j], k, j))
        items.sort(key=lambda t: t[0], reverse=True)
        return items[: self.top_k]

    def _pairs(self, haps1, haps2, top1, top2, eps, acc, muug):
        # impute.py:444-548 (UMUG) and :550-658 (PMUG)
        M = self.M
        pops = self.pops
        for _w1, f1, k1, p1 in top1:
            x = eps / f1
            x2 = x * 2
            for _w2, f2, k2, p2 in top2:
                self.pair_evals += 1
                if f2 >= x:
                    m = M[p1][p2]
                    if m > 0:
                        hap1 = haps1[k1]
                        hap2 = haps2[k2]
                        if (hap1 != hap2 and m * f2 >= x) or (hap1 == hap2 and m * f2 >= x2):
                            race1 = pops[p1]
                            race2 = pops[p2]
                            gid = "-".join(sorted([hap1 + "," + race1, hap2 + "," + race2]))
                            if gid in acc.seen:
                                continue
                            acc.seen.add(gid)
                            if muug:
                                key = "^".join(
                                    "+".join(sorted(pr))
                                    for pr in zip(sorted(hap1.split("~")), sorted(hap2.split("~")))
                                )
                            else:
                                key = "~".join(sorted(hap1.split("~") + hap2.split("~")))
                            prob = f1 * f2 * m
                            if hap1 != hap2:
                                prob = prob * 2
                            if prob > acc.maxp:
                                acc.maxp = prob
                            if key in acc.geno:
                                acc.geno[key] = acc.geno[key] + prob
                            else:
                                acc.geno[key] = prob
                            rk = ",".join(sorted([race1, race2]))
                            if rk in acc.pops:
                                acc.pops[rk] = acc.pops[rk] + prob
                            else:
                                acc.pops[rk] = prob
                            if not muug:
                                acc.pairs.append([hap1, hap2])
                                acc.pair_pops.append([race1, race2])
                                acc.pair_probs.append(prob)
                else:
                    break

    @staticmethod
    def _result(acc, muug):
        if muug:
            return {"MaxProb": acc.maxp, "Haps": acc.geno, "Pops": acc.pops}
        return {"MaxProb": acc.maxp, "Haps": acc.pairs, "Probs": acc.pair_probs, "Pops": acc.pair_pops}

    # ---- Plan A ----
    def plan_a(self, phases, eps, muug):
        # impute.py:714-754 / :660-712; probe = nxg.py:253-278
        acc = _Acc()
        haps2, probs2 = [], []
        for c1, c2 in phases:
            d1 = self.g.adjs_query(["~".join(c) for c in c1])
            haps1, probs1 = list(d1.keys()), list(d1.values())
            if probs1:
                d2 = self.g.adjs_query(["~".join(c) for c in c2])
                haps2, probs2 = list(d2.keys()), list(d2.values())
            self._pairs(haps1, haps2, self._flatten(probs1), self._flatten(probs2), eps, acc, muug)
        return self._result(acc, muug)

    # ---- Plan B ----
    def _label_of(self, indices):
        # cypher_plan_b.py:14-33
        return "".join(str(i) for i in sorted(indices))

    def _lookup(self, strings, division):
        # impute.py:1207-1216 + cypher_plan_b.py:39-42 + nxg.py:280-307
        if not strings:
            return {}
        want = self._label_of(division)
        have = "".join(sorted(self.index_str[a.split("*")[0]] for a in strings[0].split("~")))
        return self.g.adjs_query_by_color(strings, have, want)

    def _block_strings(self, cands, division, missing):
        # impute.py:1015-1039
        out = []
        for hap in cands:
            parts = []
            for d in division:
                if d not in missing:
                    place = d - sum(1 for m in missing if d > m)
                    parts.append(str(hap[place - 1]))
            if parts:
                out.append("~".join(parts))
        return out

    def _combine(self, new, acc, planc=False):
        # impute.py:1041-1069 open_option_(dict2=new, dict1=acc)
        size = 1 if planc else len(self.pops)
        if self.save_space:
            for d in (acc, new):
                if len(d) > 10:
                    order = sorted(((h, sum(d[h])) for h in d), key=lambda kv: kv[1])
                    while len(d) > 10:
                        del d[order[0][0]]
                        del order[0]
        out = {}
        for k1 in acc:
            for k2 in new:
                v1 = acc[k1]
                v2 = new[k2]
                vec = [v1[i] * v2[i] * BLOCK_FACTOR for i in range(size)]
                if max(vec) > 0:
                    out["~".join(sorted(k1.split("~") + k2.split("~")))] = vec
        return out

    def _row_freqs(self, row, cands, missing):
        # impute.py:1072-1115
        acc = self._lookup(self._block_strings(cands, row[0], missing), row[0])
        if acc != {}:
            for blk in row[1:]:
                d = self._lookup(self._block_strings(cands, blk, missing), blk)
                if d == {}:
                    if all(e in missing for e in blk):
                        d = self.g.haps_with_probs_by_label(self._label_of(blk))
                    else:
                        acc = {}
                        break
                acc = self._combine(d, acc)
        return acc

    def _side_plan_b(self, cands, row, missing):
        # impute.py:1117-1123
        if row[0] == self.all_indices:
            return self.g.adjs_query(["~".join(c) for c in cands])
        return self._row_freqs(row, cands, missing)

    def _side_missing_data(self, cands, nid):
        # impute.py:1125-1172
        drop = list(set(nid))
        keep = [x for x in set(self.index.values()) if x not in drop]
        out = {}
        for hap in cands:
            kept = [a for a in hap if self.index[a.split("*")[0]] not in nid]
            gone = list(set(a for a in hap if self.index[a.split("*")[0]] in nid))
            s = "~".join(kept)
            if s != "":
                d = self._lookup([s], keep)
                for key in d:
                    parts = key.split("~")
                    parts = parts[: nid[0] - 1] + gone + parts[nid[0] - 1:]
                    out["~".join(sorted(parts))] = [x * (self.fmd ** len(drop)) for x in d[key]]
        return out

    def _not_in_data(self, phases, side):
        # impute.py:1224-1241
        nid = []
        for t in range(len(phases[0][0][0])):
            alleles = list(set(c[t] for ph in phases for c in ph[side]))
            if self._alleles_exist(alleles) == {}:
                nid.append(self.index[alleles[0].split("*")[0]])
        return nid

    def _not_in_data_one(self, cands):
        # impute.py:1243-1258
        nid = []
        for t in range(len(cands[0])):
            alleles = list(set(c[t] for c in cands))
            if self._alleles_exist(alleles) == {}:
                nid.append(self.index[alleles[0].split("*")[0]])
        return nid

    def _untyped(self, phases):
        # impute.py:1193-1200, 994-1006
        first = phases[0][0][0]
        if len(first) < len(self.full_label):
            typed = [self.index[a.split("*")[0]] for a in first]
            out = []
            for locus in self.loci:
                i = self.index[locus]
                if i not in typed and i not in out:
                    out.append(i)
            return out
        return []

    def plan_b(self, phases, eps, muug):
        # impute.py:1392-1570
        acc = _Acc()
        first_row = [[NEVER, NEVER] for _ in phases]
        nid1 = self._not_in_data(phases, 0)
        nid2 = self._not_in_data(phases, 1)
        haps2, probs2 = [], []
        d1 = d2 = None
        row_i = 0
        missing = None
        while acc.geno == {}:
            if row_i >= len(self.matrix) or self.matrix[row_i] == []:
                break
            missing = self._untyped(phases)
            for i, (c1, c2) in enumerate(phases):
                if nid1 == []:
                    idx = min(row_i, first_row[i][0])
                    d1 = self._side_plan_b(c1, self.matrix[idx], missing)
                    if len(d1):
                        first_row[i][0] = idx
                else:
                    d1 = self._side_missing_data(c1, nid1)
                haps1, probs1 = list(d1.keys()), list(d1.values())
                if nid2 == []:
                    idx = min(row_i, first_row[i][1])
                    d2 = self._side_plan_b(c2, self.matrix[idx], missing)
                    if len(d2):
                        first_row[i][1] = idx
                    haps2, probs2 = list(d2.keys()), list(d2.values())
                elif len(probs1) > 0:
                    d2 = self._side_missing_data(c2, nid2)
                    haps2, probs2 = list(d2.keys()), list(d2.values())
                self._pairs(haps1, haps2, self._flatten(probs1), self._flatten(probs2), eps, acc, muug)
            row_i += 1

        # second stage, impute.py:1490-1558
        cur = 0
        while acc.geno == {} and cur < 6:
            for i, (c1, c2) in enumerate(phases):
                i1 = min(NEVER, first_row[i][0])
                i2 = min(NEVER, first_row[i][1])
                if not (i1 == NEVER and i2 == NEVER):
                    if i1 == NEVER and len(c1) > 0:
                        d1 = self._side_missing_data(c1, self._not_in_data_one(c1))
                        d2 = self._side_plan_b(c2, self.matrix[i2], missing)
                    if i2 == NEVER and len(c2) > 0:
                        d1 = self._side_plan_b(c1, self.matrix[i1], missing)
                        d2 = self._side_missing_data(c2, self._not_in_data_one(c2))
                    if d1 is None or d2 is None:
                        raise UnboundLocalError("P1/P2 unbound (impute.py:1521-1524)")
                    self._pairs(
                        list(d1.keys()), list(d2.keys()),
                        self._flatten(list(d1.values())), self._flatten(list(d2.values())),
                        eps, acc, muug,
                    )
            cur += 1
        return self._result(acc, muug)

    # ---- Plan C ----
    def _side_plan_c(self, cands, missing):
        # impute.py:1264-1311
        out = {}
        for hap in cands:
            tmp = {}
            miss = []
            for allele in hap:
                d = self._lookup([allele], [self.index[allele.split("*")[0]]])
                d = {a: [sum(v)] for a, v in d.items()}
                if d == {}:
                    miss.append(allele)
                elif tmp == {}:
                    tmp = d
                else:
                    tmp = self._combine(d, tmp, True)
                    if not tmp:
                        break
            if miss:
                for key in tmp:
                    parts = key.split("~") + miss
                    out["~".join(sorted(parts))] = [x * (self.fmd ** len(miss)) for x in tmp[key]]
            else:
                for key in tmp:
                    out[key] = tmp[key]
        d = self.g.haps_with_probs_by_label(self._label_of(missing))
        d = {a: [sum(v)] for a, v in d.items()}
        if out:
            if d:
                out = self._combine(d, out, True)
            else:
                for m in missing:
                    d = self.g.haps_with_probs_by_label(self._label_of([m]))
                    d = {a: [sum(v)] for a, v in d.items()}
                    if d:
                        out = self._combine(d, out, True)
        return out

    def plan_c(self, phases, muug):
        # impute.py:1313-1389
        acc = _Acc()
        haps2, probs2 = [], []
        missing = self._untyped(phases)
        for c1, c2 in phases:
            d1 = self._side_plan_c(c1, missing)
            haps1, probs1 = list(d1.keys()), list(d1.values())
            if probs1:
                d2 = self._side_plan_c(c2, missing)
                haps2, probs2 = list(d2.keys()), list(d2.values())
            self._pairs(haps1, haps2, self._flatten(probs1), self._flatten(probs2), 0, acc, muug)
        if muug:
            return {"MaxProb": acc.maxp, "Haps": acc.geno,
                    "Pops": {"all_pops,all_pops": sum(acc.pops.values())}}
        return {"MaxProb": acc.maxp, "Haps": acc.pairs, "Probs": acc.pair_probs,
                "Pops": [["all_pops", "all_pops"] for _ in acc.pair_pops]}

    # ---- epsilon schedule + fallbacks ----
    def evaluate(self, phases, muug):
        # impute.py:1658-1724
        planb = self.cfg["planb"]
        eps = self.cfg["epsilon"]
        res = {"Haps": "NaN", "Probs": 0}
        last_round = False
        while eps > 0:
            eps /= 10
            if eps < 1.0e-9:
                eps = 0.0
            res = self.plan_a(phases, eps, muug)
            if len(res["Haps"]) > 0 and eps > 0:
                eps = res["MaxProb"] / 100000
                last_round = True
                break
        if last_round:
            res = self.plan_a(phases, eps, muug)
        npop = len(self.pops)
        for level in range(2):
            if level == 1:
                self.M = np.ones((npop, npop))
            if planb and len(res["Haps"]) == 0:
                if self.cfg.get("nodes_for_plan_A"):
                    raise PlanBUnderMatrix("a subject leaves Plan A under a Plan_A_Matrix")
                self.plan = "b"
                eps = 1e-14
                n_res = 0
                while eps > 0 and n_res < 10:
                    eps /= 10
                    if eps < 1.0e-3:
                        eps = 0.0
                    res = self.plan_b(phases, eps, muug)
                    n_res = len(res["Haps"])
        return res

    def comp_cand(self, gl):
        # impute.py:1584-1656
        chrom = self.gl2haps(gl)
        if chrom == []:
            return None, None
        if self.cfg.get("nodes_for_plan_A"):
            # impute.py:1592-1596 with input_type (:1574-1579): the typed loci, in the order of the sorted side
            geno_type = [self.index[x.split("*")[0]] for x in chrom["Genotype"][0]]
            if geno_type not in self.cfg["nodes_for_plan_A"]:
                return None, None
        n_loci = chrom["N_Loc"]
        pmags = self.gen_phases(chrom["Genotype"], n_loci, self.binary)
        if pmags == []:
            return None, None
        res_muugs = {"MaxProb": 0, "Haps": {}, "Pops": {}}
        res_haps = {"Haps": "Nan", "Probs": 0, "Pops": {}}
        planb = self.cfg["planb"]
        phases = self.open_phases(pmags, n_loci)
        if not phases:
            self._reduce_valid(pmags, n_loci)
            phases = self.open_phases(pmags, n_loci)
        if not phases:
            self._reduce_common(pmags, n_loci, 10)
            phases = self.open_phases(pmags, n_loci)
        if phases:
            if self.cfg["output_MUUG"]:
                saved = np.array(self.M, order="K", copy=True)
                res_muugs = self.evaluate(phases, True)
                if planb and len(res_muugs["Haps"]) == 0:
                    self.plan = "c"
                    self._reduce_common(pmags, n_loci, 1, True)
                    phases = self.open_phases(pmags, n_loci)
                    res_muugs = self.plan_c(phases, True)
                self.M = saved
            if self.cfg["output_haplotypes"]:
                res_haps = self.evaluate(phases, False)
                if planb and len(res_haps["Haps"]) == 0 and not self.em:
                    self._reduce_common(pmags, n_loci, 1, True)
                    phases = self.open_phases(pmags, n_loci)
                    res_haps = self.plan_c(phases, False)
        return res_muugs, res_haps

    # ---- prior matrix ----
    def prior_matrix(self, races1, races2):
        # impute.py:1844-1924
        pr = self.cfg["priority"]
        n = len(self.pops)
        M = np.zeros((n, n))
        eye = np.identity(n)
        for a in races1:
            for b in races2:
                if a == "" and b == "":
                    continue
                T = np.zeros((n, n))
                if a == "" or b == "":
                    r = self.pops.index(b) if a == "" else self.pops.index(a)
                    for i in range(n):
                        T[r, i] = T[r, i] + pr["gamma"] * 2
                    T = T + T.transpose()
                    T[r, r] -= pr["gamma"] * 2
                else:
                    r1 = self.pops.index(a)
                    r2 = self.pops.index(b)
                    for i in range(n):
                        T[r1, i] = T[r1, i] + pr["gamma"]
                        T[i, r2] = T[i, r2] + pr["gamma"]
                    T[r1, r2] -= pr["gamma"]
                    T[r1, r2] = T[r1, r2] + pr["alpha"]
                    if r1 != r2:
                        T = T + T.transpose()
                        T[r1, r1] -= pr["gamma"]
                        T[r2, r2] -= pr["gamma"]
                    T[r1, r1] += pr["delta"]
                    if r1 != r2:
                        T[r2, r2] += pr["delta"]
                T = pr["eta"] * np.ones((n, n)) + T + pr["beta"] * eye
                M += T
        total = 0
        for i in range(n):
            for j in range(n):
                M[i][j] = M[i][j] * self.count_by_prob[i] * self.count_by_prob[j]
                total += M[i][j]
        return M / total

    def impute_one(self, gl, race1, race2):
        # impute.py:1940-1983
        cleaned = clean_gl(gl)
        n = len(self.pops)
        self.M = np.ones((n, n)) if self.unk_priors == "MR" else np.identity(n)
        if race1 or race2:
            known = False
            r1 = race1.split(";")
            for i, r in enumerate(r1):
                if r not in self.pops:
                    r1[i] = ""
                else:
                    known = True
            r2 = race2.split(";")
            for i, r in enumerate(r2):
                if r not in self.pops:
                    r2[i] = ""
                else:
                    known = True
            if known:
                self.M = self.prior_matrix(r1, r2)
        if gl:
            return self.comp_cand(cleaned)
        return None, None

    # ---- subject loop + writers (impute.py:1985-2155, :24-76) ----
    def impute_lines(self, lines, em_mr=False, em=False, first_index=0):
        """em_mr: grim.grim.impute(hap_pop_pair=True) (impute.py:2079-2088); em: impute_file(em=True);
        first_index: line number of lines[0] in the whole input (checker convenience: lets a test split one
        input over forked workers; the reference's enumerate() starts at 0)."""
        out = {k: [] for k in ("umug", "umug_pops", "pmug", "pmug_pops", "miss", "problem")}
        n_res = self.cfg["number_of_results"]
        n_pop = self.cfg["number_of_pop_results"]
        self.em = em
        f_bin = None
        bin_path = self.cfg.get("bin_imputation_input_file", "None")
        if os.path.isfile(bin_path):          # impute.py:2001-2005
            with open(bin_path) as f:
                f_bin = json.load(f)
        for i, raw in enumerate(lines, first_index):
            try:
                raw = raw.rstrip()
                fields = raw.split(",") if "," in raw else raw.split("%")
                sid = fields[0]
                gl = fields[1]
                self.binary = [1] * (len(self.full_label) - 1)     # impute.py:2030-2032
                if f_bin is not None:
                    self.binary = f_bin[sid]
                race1 = race2 = None
                if len(fields) > 2:
                    race1 = fields[2]
                    race2 = fields[3]
                self.plan = "a"
                res_muugs, res_haps = self.impute_one(gl, race1, race2)
                if res_muugs is None:
                    out["problem"].append(str(i) + "," + str(sid) + "\n")
                    continue
                if (len(res_haps["Haps"]) == 0 or res_haps["Haps"] == "NaN") and len(res_muugs["Haps"]) == 0:
                    out["miss"].append(str(i) + "," + str(sid) + "\n")
                if self.cfg["output_haplotypes"] and em_mr:
                    _write_hap_race_pairs(sid, res_haps["Haps"], res_haps["Pops"], res_haps["Probs"], n_res, out["pmug"])
                    _write_pairs(sid, res_haps["Pops"], res_haps["Probs"], 1, out["pmug_pops"], ",")
                elif self.cfg["output_haplotypes"]:
                    _write_pairs(sid, res_haps["Haps"], res_haps["Probs"], n_res, out["pmug"], "+")
                    _write_pairs(sid, res_haps["Pops"], res_haps["Probs"], n_pop, out["pmug_pops"], ",")
                if self.cfg["output_MUUG"]:
                    _write_dict(sid, res_muugs["Haps"], n_res, out["umug"])
                    _write_dict(sid, res_muugs["Pops"], n_pop, out["umug_pops"])
            except PlanBUnderMatrix:
                raise
            except Exception:
                out["problem"].append(str(raw) + "\n")
                continue
        return {k: "".join(v) for k, v in out.items()}


def _write_pairs(sid, res, probs, limit, rows, sign):
    # impute.py:24-58 write_best_prob
    sums = {}
    for k in range(len(res)):
        key = res[k][0] + sign + res[k][1]
        if key in sums:
            sums[key] = probs[k] + sums[key]
        else:
            key2 = res[k][1] + sign + res[k][0]
            if key2 in sums:
                sums[key2] = probs[k] + sums[key2]
            else:
                sums[key] = probs[k]
    ranked = sorted(sums.items(), key=lambda kv: kv[1], reverse=True)
    for k in range(min(limit, len(ranked))):
        rows.append(sid + "," + str(ranked[k][0]) + "," + str(ranked[k][1]) + "," + str(k) + "\n")


def _write_hap_race_pairs(sid, haps, pops, probs, limit, rows):
    # impute.py:79-99 write_best_hap_race_pairs: every (haplotype;population) pair on its own, no merging
    allr = []
    for k in range(len(probs)):
        allr.append([probs[k], haps[k][0] + ";" + pops[k][0] + "," + haps[k][1] + ";" + pops[k][1]])
    allr.sort(key=lambda x: x[0], reverse=True)
    for k in range(min(limit, len(allr))):
        rows.append(sid + "," + str(allr[k][1]) + "," + str(allr[k][0]) + "," + str(k) + "\n")


def _write_dict(sid, res, limit, rows):
    # impute.py:61-76 write_best_prob_genotype
    ranked = sorted(res.items(), key=lambda kv: kv[1], reverse=True)
    for k in range(min(limit, len(ranked))):
        rows.append(sid + "," + str(ranked[k][0]) + "," + str(ranked[k][1]) + "," + str(k) + "\n")


# ---------------------------------------------------------------------------------------------
# File-level driver (grim/grim.py:57-74 + run_impute_def.py:41-211), used by tests and bench
# ---------------------------------------------------------------------------------------------
def impute_file(json_conf, base_dir="", graph=None, lines=None, em_mr=False, em=False):
    """Returns (dict of the six file texts, graph)."""
    if graph is None:
        graph = graph_from_config(json_conf, base_dir)
    cfg = load_config(json_conf)
    cbp = None
    pc = json_conf.get("pops_count_file", False)
    if pc:
        cbp = count_by_prob_from_file(len(cfg["pops"]), os.path.join(base_dir, pc))
    imp = OracleImputation(graph, cfg, cbp)
    if lines is None:
        with open(os.path.join(base_dir, json_conf["imputation_in_file"])) as f:
            lines = f.readlines()
    return imp.impute_lines(lines, em_mr=em_mr, em=em), graph


def impute_conf_file(conf_path, base_dir=""):
    with open(conf_path) as f:
        return impute_file(json.load(f), base_dir)
