"""Array-backed twin of grim_oracle.OracleGraph for tables too large for per-node Python dicts.

TEST INFRASTRUCTURE ONLY (same rules as grim_oracle.py: imported by tests/, smoke() and
bench.py's CPU legs, never by the product package).

Same five queries, same answers, same ordering as OracleGraph -- i.e. as the reference's
Graph (grim/imputation/networkx_graph.py:215-321 = "nxg.py") loaded from the CSV files that
graph_generation/generate_neo4j_multi_hpf.py:209-486 ("gen.py") writes -- but the marginal
labels are materialised lazily with numpy, one label at a time, the first time a query touches
them.  A 1M-haplotype x 21-population table (11.4M nodes) then costs a few seconds and a few
hundred MB instead of minutes and tens of GB, which is what lets `pytest -m gpu` check the CUDA
path against the oracle on BASELINE configs 2/3/5 at their own table shape.

Parity status: PINNED through OracleGraph -- tests/test_oracle_graph_np.py checks every query
of this class against OracleGraph (itself pinned against the unmodified reference by the golden
fixtures) on the README table, the 3-population table and the nine-locus table, including the
CSR sentinel quirk (nxg.py:195-196) on the last node and the last connector.

What is restated, with the reference lines:
  node ids      full haplotypes in first-appearance order, then per label (gen.py:105-110 order)
                marginal nodes in first-appearance order (gen.py:296-309,368-371)
  sums          sequential FP64 adds in full-haplotype order (gen.py:405): np.add.at is an
                unbuffered in-order accumulation, so every (node, population) sum sees its
                addends in exactly that order
  top links     partial node -> full ids ascending (nxg.py:149-201)
  connectors    (parent label, child) -> parent ids ascending (nxg.py:91-130)
  quirk         the closing CSR sentinel is len(Vertices) (nxg.py:195-196)
"""
from __future__ import annotations

import itertools

import numpy as np


class _Fault:
    """An adjacency whose reference range runs past the edge array (IndexError in the reference)."""


def _sentinel_count(own, n_edges, n_vertices):
    # nxg.py:195-196, as grim_oracle._sentinel_slice: how many of its own `own` edges the last
    # vertex keeps (None = the reference raises)
    start = n_edges - own
    if n_vertices <= start:
        return 0
    if n_vertices > n_edges:
        return None
    return n_vertices - start


class _Label:
    __slots__ = ("mask", "n", "first", "keys", "index", "vec", "tl_perm", "tl_start", "names", "vecs")


class NumpyOracleGraph:
    def __init__(self, allele_names, full_alleles, full_freqs, pops, loci_map, plan_a_matrix=None):
        """allele_names[l]: names of locus l (position = loci_map index - 1), id = position + 1;
        full_alleles uint16 [N][L] ids of the full haplotypes in first-appearance order (after the
        trim of gen.py:320-339); full_freqs float64 [N][P] in `pops` order (gen.py:341-358)."""
        self.pops = list(pops)
        self.loci_map = {k: int(v) for k, v in loci_map.items()}
        L = len(self.loci_map)
        self.L = L
        self.loci = [k for k, _v in sorted(self.loci_map.items(), key=lambda kv: kv[1])]
        self.locus_pos = {k: v - 1 for k, v in self.loci_map.items()}
        self.full_label = "".join(sorted({str(v) for v in self.loci_map.values()}))
        full = self.full_label
        self.labels = [full]
        for r in range(len(full) - 1, 0, -1):
            self.labels.extend("".join(c) for c in itertools.combinations(full, r))
        # Plan_A_Matrix: as grim_oracle.OracleGraph -- the matrix labels in matrix order, then the single-locus
        # labels (look-ups only), no connectors, the CSR sentinel on the last Plan-A node
        self.plan_a = None
        if plan_a_matrix:
            from grim_oracle import plan_a_labels
            self.plan_a = plan_a_labels(plan_a_matrix, full)
            self.labels = self.plan_a + [ch for ch in full if ch not in self.plan_a]
        self.allele_names = [list(a) for a in allele_names]
        self.allele_id = [{a: i + 1 for i, a in enumerate(al)} for al in self.allele_names]
        self.fa = np.ascontiguousarray(full_alleles, dtype=np.uint16)
        self.ff = np.ascontiguousarray(full_freqs, dtype=np.float64)
        self.n_full = int(self.fa.shape[0])
        self.P = int(self.ff.shape[1])
        # packed key: one bit field per locus, as many 64-bit words as needed
        self.bits = [max(1, int(len(a) + 1).bit_length()) for a in self.allele_names]
        self.word, self.shift = [], []
        w, used = 0, 0
        for b in self.bits:
            if used + b > 64:
                w, used = w + 1, 0
            self.word.append(w)
            self.shift.append(used)
            used += b
        self.n_words = w + 1
        self.full_mask = (1 << L) - 1
        # per-label node counts (every label: the CSR sentinel quirk depends on the totals)
        self._count = {}
        self._first_id = {}
        nid = 0
        for lab in self.labels:
            m = self._mask_of(lab)
            if m == self.full_mask:
                c = self.n_full
            else:
                c = int(len(self._unique(self._project(m))[0]))
            self._count[m] = c
            self._first_id[m] = nid
            nid += c
        self.n_nodes = nid
        self.n_edges = self.n_full * (len(self.labels) - 1)
        if self.plan_a is not None:
            self.n_edges = self.n_full * (len(self.plan_a) - 1)
            self.n_plan_a_nodes = sum(self._count[self._mask_of(lab)] for lab in self.plan_a)
            self.n_conn = self.n_whole_edges = 0
            self._lab, self._conn, self._fname, self._fvec = {}, {}, {}, {}
            self.last_mask = self._mask_of(self.plan_a[-1])
            return
        n_conn = 0
        whole = 0
        for lab in self.labels:
            m = self._mask_of(lab)
            if bin(m).count("1") < 2:
                continue
            for l in range(L):
                if m >> l & 1:
                    n_conn += self._count[m & ~(1 << l)]
            whole += bin(m).count("1") * self._count[m]
        self.n_conn = n_conn
        self.n_whole_edges = n_conn + whole
        self._lab = {}
        self._conn = {}
        self._fname = {}   # full id -> name (memo)
        self._fvec = {}    # full id -> list of P floats (memo)
        self.last_mask = self._mask_of(self.labels[-1]) if len(self.labels) > 1 else None

    # ---- construction helpers ----
    @classmethod
    def from_hpf(cls, hpf_lines, pops, loci_map, freq_trim, pop_count_lines=None, plan_a_matrix=None):
        """hpf.csv rows -> arrays, following gen.py:259-266 (trim), :320-339 (rows), :341-358."""
        lm = {k: int(v) for k, v in loci_map.items()}
        L = len(lm)
        trim = {}
        if pop_count_lines is None:
            for p in pops:
                trim[p] = freq_trim
        else:
            for line in pop_count_lines:
                p, cnt, _ratio = line.strip().split(",")
                trim[p] = freq_trim / float(cnt)
        seen = {}
        pop_hap = {}
        for line in hpf_lines:
            if not line:
                continue
            hap, pop, freq = line.split(",")
            if hap == "hap":
                continue
            freq = float(freq)
            if freq == 0.0 or freq < trim[pop]:
                continue
            al = ["0"] * L
            for a in hap.split("~"):
                if a[-1] == "g":
                    a = a[:-1]
                al[lm[a.split("*")[0]] - 1] = a
            al = tuple(al)
            seen[al] = True
            pop_hap[(pop, al)] = freq
        rows = list(seen)
        names = [sorted({r[l] for r in rows}) for l in range(L)]
        ids = [{a: i + 1 for i, a in enumerate(n)} for n in names]
        fa = np.zeros((len(rows), L), np.uint16)
        for l in range(L):
            fa[:, l] = [ids[l][r[l]] for r in rows]
        ff = np.zeros((len(rows), len(pops)), np.float64)
        for j, p in enumerate(pops):
            ff[:, j] = [pop_hap.get((p, r), 0.0) for r in rows]
        return cls(names, fa, ff, pops, loci_map, plan_a_matrix)

    def _mask_of(self, label):
        m = 0
        for ch in label:
            m |= 1 << (int(ch) - 1)
        return m

    def _project(self, mask):
        """Packed key words of every full haplotype restricted to the loci of `mask`."""
        words = [np.zeros(self.n_full, np.uint64) for _ in range(self.n_words)]
        for l in range(self.L):
            if mask >> l & 1:
                words[self.word[l]] |= self.fa[:, l].astype(np.uint64) << np.uint64(self.shift[l])
        return words

    @staticmethod
    def _unique(words):
        """-> (position of the first member of each group, group index of every element); groups
        are numbered by first appearance."""
        n = len(words[0])
        if n == 0:
            return np.zeros(0, np.int64), np.zeros(0, np.int64)
        if len(words) == 1:
            _u, first, inv = np.unique(words[0], return_index=True, return_inverse=True)
        else:
            order = np.lexsort(tuple(words))
            diff = np.zeros(n, bool)
            diff[0] = True
            for w in words:
                ws = w[order]
                diff[1:] |= ws[1:] != ws[:-1]
            gid_sorted = np.cumsum(diff) - 1
            inv = np.empty(n, np.int64)
            inv[order] = gid_sorted
            first = np.full(int(gid_sorted[-1]) + 1, n, np.int64)
            np.minimum.at(first, inv, np.arange(n))
        rank = np.empty(len(first), np.int64)
        rank[np.argsort(first, kind="stable")] = np.arange(len(first))
        return np.sort(first), rank[inv.reshape(-1)]

    def _pykeys(self, words, idx):
        if len(words) == 1:
            return words[0][idx].tolist()
        out = [0] * len(idx)
        for w, arr in enumerate(words):
            col = arr[idx].tolist()
            sh = 64 * w
            out = [o | (c << sh) for o, c in zip(out, col)]
        return out

    def _label(self, mask):
        lab = self._lab.get(mask)
        if lab is not None:
            return lab
        lab = _Label()
        lab.mask = mask
        lab.first = self._first_id[mask]
        words = self._project(mask)
        if mask == self.full_mask:
            lab.n = self.n_full
            first = np.arange(self.n_full)
            lab.vec = self.ff
            lab.tl_perm = lab.tl_start = None
        else:
            first, gid = self._unique(words)
            lab.n = len(first)
            acc = np.zeros((lab.n, self.P), np.float64)
            np.add.at(acc, gid, self.ff)        # in-order, unbuffered: gen.py:405's sequential +=
            lab.vec = acc
            lab.tl_perm = np.argsort(gid, kind="stable")
            lab.tl_start = np.concatenate([[0], np.cumsum(np.bincount(gid, minlength=lab.n))])
        lab.keys = self._pykeys(words, first)
        lab.index = {k: i for i, k in enumerate(lab.keys)}
        lab.names = None
        lab.vecs = None
        self._lab[mask] = lab
        return lab

    # ---- name <-> key ----
    def _parse(self, name):
        """-> (label mask, packed key) or None when some allele is not in the dictionaries."""
        mask, key = 0, 0
        for a in name.split("~"):
            l = self.locus_pos.get(a.split("*")[0])
            if l is None or mask >> l & 1:
                return None
            i = self.allele_id[l].get(a)
            if i is None:
                return None
            mask |= 1 << l
            key |= i << (self.shift[l] + 64 * self.word[l])
        return mask, key

    def _name(self, mask, key):
        out = []
        for l in range(self.L):
            if mask >> l & 1:
                i = (key >> (self.shift[l] + 64 * self.word[l])) & ((1 << self.bits[l]) - 1)
                out.append(self.allele_names[l][i - 1])
        return "~".join(out)

    def _find(self, name):
        pk = self._parse(name)
        if pk is None:
            return None
        mask, key = pk
        if mask not in self._count:
            return None
        lab = self._label(mask)
        i = lab.index.get(key)
        if i is None:
            return None
        return lab, i

    def _full_name(self, fid):
        n = self._fname.get(fid)
        if n is None:
            row = self.fa[fid].tolist()
            n = self._fname[fid] = "~".join(self.allele_names[l][row[l] - 1] for l in range(self.L))
        return n

    def _full_vec(self, fid):
        v = self._fvec.get(fid)
        if v is None:
            v = self._fvec[fid] = self.ff[fid].tolist()
        return v

    # ---- queries (nxg.py:215-321) ----
    def _names_of(self, lab):
        if lab.names is None:
            lab.names = [self._name(lab.mask, k) for k in lab.keys]
        return lab.names

    def haps_by_label(self, label):
        m = self._mask_of(label)
        if m not in self._count:
            return []
        return self._names_of(self._label(m))

    def haps_with_probs_by_label(self, label):
        m = self._mask_of(label)
        if m not in self._count:
            return {}
        lab = self._label(m)
        if lab.vecs is None:
            lab.vecs = lab.vec.tolist()
        return dict(zip(self._names_of(lab), lab.vecs))

    def _toplinks(self, lab, i):
        own = lab.tl_perm[lab.tl_start[i]:lab.tl_start[i + 1]]
        if lab.mask == self.last_mask and i == lab.n - 1:
            c = _sentinel_count(len(own), self.n_edges, self.n_nodes if self.plan_a is None else self.n_plan_a_nodes)
            if c is None:
                raise IndexError("CSR sentinel range past the edge array")
            own = own[:c]
        return own

    def adjs_query(self, names):
        # nxg.py:253-278
        out = {}
        for n in names:
            hit = self._find(n)
            if hit is None:
                continue
            lab, i = hit
            if lab.mask == self.full_mask:
                out[n] = self._full_vec(i)
            else:
                for fid in self._toplinks(lab, i).tolist():
                    out[self._full_name(fid)] = self._full_vec(fid)
        return out

    def node_probs(self, names):
        # nxg.py:309-321
        out = {}
        for n in names:
            hit = self._find(n)
            if hit is not None:
                lab, i = hit
                out[n] = self._full_vec(i) if lab.mask == self.full_mask else lab.vec[i].tolist()
        return out

    def _connectors(self, mask_b, mask_a):
        key = (mask_b, mask_a)
        c = self._conn.get(key)
        if c is not None:
            return c
        lab_b = self._label(mask_b)
        fm = 0
        for l in range(self.L):
            if mask_a >> l & 1:
                fm |= ((1 << self.bits[l]) - 1) << (self.shift[l] + 64 * self.word[l])
        groups = {}
        for j, k in enumerate(lab_b.keys):     # ascending parent id
            groups.setdefault(k & fm, []).append(j)
        self._conn[key] = groups
        return groups

    def adjs_query_by_color(self, names, label_a, label_b):
        # nxg.py:280-307
        if label_a == label_b:
            return self.node_probs(names)
        if self.plan_a is not None:
            return {}       # Plan B is refused under a matrix (grim_oracle.PlanBUnderMatrix) before it gets here
        mb = self._mask_of(label_b)
        out = {}
        for n in names:
            hit = self._find(n)
            if hit is None:
                continue
            lab, i = hit
            ma = lab.mask
            if mb not in self._count or (ma & ~mb) or bin(mb & ~ma).count("1") != 1:
                continue
            parents = self._connectors(mb, ma).get(lab.keys[i], [])
            # the connector created last (see grim_oracle.OracleGraph: last node of the last
            # single-locus label under the parent label that adds locus L-2) is cut by the quirk
            if ma == self.last_mask and i == lab.n - 1 and self.L >= 2:
                own = self.L - 1
                others = list(set(range(self.L)).difference([own]))
                if mb == (1 << own) | (1 << others[-1]):
                    c = _sentinel_count(len(parents), self.n_whole_edges, self.n_nodes + self.n_conn)
                    if c is None:
                        raise IndexError("CSR sentinel range past the edge array")
                    parents = parents[:c]
            lab_b = self._label(mb)
            for j in parents:
                out[self._name(mb, lab_b.keys[j])] = lab_b.vec[j].tolist()
        return out
