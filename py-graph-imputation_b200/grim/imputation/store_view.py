"""Host-side view of the frequency store with the reference's query methods (SURVEY 8b "store
seam"): adjs_query / adjs_query_by_color / node_probs / haps_by_label / haps_with_probs_by_label
of grim/imputation/networkx_graph.py:215-321, answered from the table arrays the device holds
(`Graph.export()`: packed node keys, per-population frequency vectors, top-link and connector CSR,
node-id range per label).  The per-subject engine never calls these -- the kernels probe the
device tables directly -- they exist for callers that inspect the graph the way they did with
the reference's `Graph`.  Results are ordered dicts / lists in the reference's order (node-id
order; insertion order of the traversal).

No frequency arithmetic happens here: vectors are the FP64 values of the table, as Python floats.
"""
import numpy as np

ADJ_FAULT = 0xFFFFFFFF   # CSR count patched to "the reference raises IndexError here" 


class StoreView(object):
    def __init__(self, arrays, loci, loci_map, alleles):
        """arrays: Graph.export() (or the same layout built by the tests from the oracle graph);
        loci: locus names in key order; loci_map: config loci_map; alleles: per-locus allele names
        (dictionary order: id = index + 1)."""
        self.a = arrays
        self.loci = list(loci)
        self.L = len(self.loci)
        self.alleles = alleles
        self.allele_id = [{n: i + 1 for i, n in enumerate(al)} for al in alleles]
        self.locus_of = {n: l for l, n in enumerate(self.loci)}
        self.digit = [str(loci_map[n]) for n in self.loci]            # label strings concatenate these
        bits = arrays.get("bits")
        if bits is None:
            raise ValueError("store view needs the key layout ('bits')")
        self.bits = [int(b) for b in bits]
        self.shift = [sum(self.bits[:l]) for l in range(self.L)]
        self.freq = arrays["node_freq"] if "node_freq" in arrays else arrays["freq"]
        self.key = arrays["node_key"]
        self.label_first = np.asarray(arrays["label_first"]).astype(np.int64)
        self.label_count = np.asarray(arrays["label_count"]).astype(np.int64)
        self.full = (1 << self.L) - 1
        self._index = {}          # label mask -> {packed key: node id}
        self._names = {}          # node id -> name (cache)

    # ---- labels, names, keys
    def label_mask(self, label):
        """'135' -> bit mask over the loci in key order; None if the string is not a label."""
        m, rest = 0, str(label)
        for l in range(self.L):                       # digits appear in locus order
            d = self.digit[l]
            if rest.startswith(d):
                m |= 1 << l
                rest = rest[len(d):]
        return m if not rest and m else None

    def label_of_node(self, node):
        for m in range(1, self.full + 1):
            f, c = self.label_first[m], self.label_count[m]
            if c and f <= node < f + c:
                return m
        raise IndexError(node)

    def name_of(self, node):
        node = int(node)
        nm = self._names.get(node)
        if nm is None:
            m = self.label_of_node(node)
            k = int(self.key[node])
            nm = "~".join(self.alleles[l][((k >> self.shift[l]) & ((1 << self.bits[l]) - 1)) - 1]
                          for l in range(self.L) if m >> l & 1)
            self._names[node] = nm
        return nm

    def _parse(self, name):
        """name -> (label mask, packed key) or None when a locus / allele is not in the table."""
        m, k = 0, 0
        for part in name.split("~"):
            l = self.locus_of.get(part.split("*")[0])
            if l is None or m >> l & 1:
                return None
            i = self.allele_id[l].get(part)
            if i is None:
                return None
            m |= 1 << l
            k |= i << self.shift[l]
        return (m, k) if m else None

    def find(self, name):
        """Node id of a (full or partial) haplotype name, or None.  Names list their loci in key
        order, as the reference's node names do."""
        p = self._parse(name)
        if p is None:
            return None
        m, k = p
        idx = self._index.get(m)
        if idx is None:
            f, c = int(self.label_first[m]), int(self.label_count[m])
            idx = {int(self.key[f + j]): f + j for j in range(c)}
            self._index[m] = idx
        node = idx.get(k)
        if node is None or self.name_of(node) != name:     # loci out of order: not a node name
            return None
        return node

    def vector(self, node):
        return [float(x) for x in np.atleast_1d(self.freq[int(node)])]

    # ---- the reference's queries
    def haps_by_label(self, label):
        """nxg.py:215-236: names of the label's nodes, node-file order."""
        m = self.label_mask(label)
        if m is None:
            return []
        f, c = int(self.label_first[m]), int(self.label_count[m])
        return [self.name_of(f + j) for j in range(c)]

    def haps_with_probs_by_label(self, label):
        """nxg.py:238-251."""
        m = self.label_mask(label)
        if m is None:
            return {}
        f, c = int(self.label_first[m]), int(self.label_count[m])
        return {self.name_of(f + j): self.vector(f + j) for j in range(c)}

    def _adj(self, start, cnt, adj):
        if int(cnt) == ADJ_FAULT:
            raise IndexError("CSR range past the end of the edge array (reference quirk, networkx_graph.py:195-196)")
        s = int(start)
        return [int(x) for x in adj[s:s + int(cnt)]]

    def adjs_query(self, alleleList):
        """nxg.py:253-278: a full-label name gives its own vector; a partial name gives the vectors of
        the full haplotypes that contain it (ascending node id); ordered, de-duplicated by name."""
        out = {}
        for name in alleleList:
            node = self.find(name)
            if node is None:
                continue
            if self.label_first[self.full] <= node < self.label_first[self.full] + self.label_count[self.full]:
                out[name] = self.vector(node)
            else:
                for fid in self._adj(self.a["tl_start"][node], self.a["tl_cnt"][node], self.a["tl_adj"]):
                    out[self.name_of(fid)] = self.vector(fid)
        return out

    def node_probs(self, nodes, label=None):
        """nxg.py:309-321 (the label only selects a vertex set there; one set exists without Plan_A_Matrix)."""
        out = {}
        for name in nodes:
            node = self.find(name)
            if node is not None:
                out[name] = self.vector(node)
        return out

    def adjs_query_by_color(self, alleleList, labelA, labelB):
        """nxg.py:280-307: parents of each node inside label B (one locus longer), through the connector
        `labelB + name`."""
        if labelA == labelB:
            return self.node_probs(alleleList, labelA)
        mb = self.label_mask(labelB)
        out = {}
        for name in alleleList:
            node = self.find(name)
            if node is None or mb is None:
                continue
            ma = self.label_of_node(node)
            extra = mb & ~ma
            if (ma & ~mb) or extra == 0 or extra & (extra - 1):
                continue                                   # no such connector in the reference's node file
            l = extra.bit_length() - 1
            for pid in self._adj(self.a["cn_start"][node][l], self.a["cn_cnt"][node][l], self.a["cn_adj"]):
                out[self.name_of(pid)] = self.vector(pid)
        return out
