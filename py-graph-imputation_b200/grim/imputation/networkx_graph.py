"""Frequency store of the B200 build: class Graph keeps the reference's name and role
(grim/imputation/networkx_graph.py:14 in the reference) but holds device-resident hash tables
and CSR arrays built straight from hpf.csv by libgrimb200.so, instead of dicts loaded from the
nodes/edges/top_links CSV files.

Host side here = the data-format work only: read hpf.csv (trim, canonical allele order,
first-appearance order: generate_neo4j_multi_hpf.py:259-358 of the reference), assign allele
ids, hand the packed rows to grimb_tables_build.  Marginal sums, node ids, top links and
connectors are computed on the GPU."""
import ctypes as C
import os

import numpy as np

from . import _lib


def loci_in_order(loci_map):
    """Locus names by loci_map index (1..L).  The reference silently mis-looks-up unless index
    order equals the alphabetical order of the allele prefixes (SURVEY trap T8), so anything
    else is rejected."""
    items = sorted(((int(v), k) for k, v in loci_map.items()))
    idx = [i for i, _ in items]
    if idx != list(range(1, len(items) + 1)):
        raise ValueError("loci_map indices must be 1..L")
    names = [k for _, k in items]
    if sorted(n + "*" for n in names) != [n + "*" for n in names]:
        raise NotImplementedError(
            "loci_map index order must equal the alphabetical order of the locus names "
            "(the reference's lookups silently miss otherwise)")
    if len(names) > _lib.MAX_LOCI:
        raise NotImplementedError("more than %d loci" % _lib.MAX_LOCI)
    return names


def read_hpf(freq_file, pops, loci, freq_trim, pops_count_file=None):
    """-> (alleles per locus [sorted list of str], full_alleles uint16 [N][L], full_freqs [N][P]).
    Follows generate_neo4j_multi_hpf.py:259-266 (trim), :320-339 (rows), :341-358 (vectors)."""
    L = len(loci)
    lidx = {n: i for i, n in enumerate(loci)}
    trim = {}
    if pops_count_file and os.path.isfile(pops_count_file):
        with open(pops_count_file) as f:
            for line in f:
                p, cnt, _ = line.strip().split(",")
                trim[p] = freq_trim / float(cnt)
    else:
        for p in pops:
            trim[p] = freq_trim
    pidx = {p: i for i, p in enumerate(pops)}
    order = {}       # canonical allele tuple -> row
    rows = []        # canonical allele tuples
    vals = {}        # (row, pop index) -> freq (later lines overwrite)
    with open(freq_file) as f:
        for line in f:
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
            al = ["0"] * L
            for a in hap.split("~"):
                if a[-1] == "g":
                    a = a[:-1]
                al[lidx[a.split("*")[0]]] = a
            key = tuple(al)
            r = order.get(key)
            if r is None:
                r = len(rows)
                order[key] = r
                rows.append(key)
            if pop in pidx:
                vals[(r, pidx[pop])] = freq
    alleles = [sorted({r[l] for r in rows}) for l in range(L)]
    ids = [{a: i + 1 for i, a in enumerate(alleles[l])} for l in range(L)]
    n = len(rows)
    fa = np.zeros((n, L), dtype=np.uint16)
    for l in range(L):
        d = ids[l]
        fa[:, l] = [d[r[l]] for r in rows]
    ff = np.zeros((n, len(pops)), dtype=np.float64)
    for (r, p), v in vals.items():
        ff[r, p] = v
    return alleles, fa, ff


def key_layout(n_alleles):
    """Field width per locus: enough for the table ids plus the subject-local ids of alleles
    absent from the table; spare bits of the 63 (64-bit keys) or 127 (128-bit keys, when the
    ids alone need more than 63) are shared out evenly (max 16 per locus)."""
    L = len(n_alleles)
    base = [max(1, int(n + 1).bit_length()) for n in n_alleles]
    wide = os.environ.get("GRIMB_KEY_WORDS", "") == "2"   # force the 128-bit build (tests)
    total = 63 if (sum(base) <= 63 and not wide) else 127
    if sum(base) > total:
        raise NotImplementedError("packed key needs more than 127 bits")
    spare = (total - sum(base)) // L
    return [min(16, b + spare) for b in base]


def key_words(key_bits):
    """1: libgrimb200.so (64-bit packed keys), 2: libgrimb200w.so (128-bit)."""
    return 1 if sum(key_bits) <= 63 else 2


class Graph(object):
    """Device tables for one GPU.  `Graph(config)` + `build_graph()` mirrors the reference's
    two-step construction (networkx_graph.py:15,42); the CSV file arguments are accepted and
    ignored because the tables come from config["freq_file"]."""

    def __init__(self, config, device=0):
        self.config = config
        self.device = device
        self.handle = None
        self.full_loci = config.get("full_loci")
        self.loci = loci_in_order(config["loci_map"])
        self.pops = list(config["pops"])
        self.alleles = None
        self.allele_id = None
        self.key_bits = None
        self.shift = None
        self.kw = 1           # key words: which of the two library builds serves this table
        # Plan_A_Matrix rows as locus bit masks (None: every label); the tables are built for it
        self.plan_a_masks = config.get("plan_a_masks") or None
        self.lib = None
        self._engines = {}

    def build_graph(self, nodesFile=None, edgesFile=None, allEdgesFile=None):
        """Tables from config["freq_file"] (hpf.csv); if that file is absent but a GRIM graph
        directory exists (nodesFile = nodes.csv written by the reference or by write_csv), from
        its full-haplotype rows."""
        cfg = self.config
        if os.path.isfile(cfg.get("freq_file", "")):
            alleles, fa, ff = read_hpf(cfg["freq_file"], self.pops, self.loci, cfg["freq_trim_threshold"],
                                       cfg.get("pops_count_file") if cfg.get("use_pops_count_file") else None)
        elif nodesFile and os.path.isfile(nodesFile):
            from .graph_io import read_nodes_csv
            alleles, fa, ff = read_nodes_csv(nodesFile, self.loci, self.full_loci)
        else:
            raise FileNotFoundError("neither freq_file (%s) nor a nodes csv (%s) exists"
                                    % (cfg.get("freq_file"), nodesFile))
        self.from_arrays(alleles, fa, ff)
        return self

    def _set_dictionaries(self, alleles):
        self.alleles = [list(a) for a in alleles]
        self.allele_id = [{a: i + 1 for i, a in enumerate(al)} for al in self.alleles]
        self.key_bits = key_layout([len(a) for a in self.alleles])
        self.shift = [int(sum(self.key_bits[:l])) for l in range(len(self.alleles))]
        self.kw = key_words(self.key_bits)
        self.lib = _lib.load(self.kw)

    def image_to_host(self):
        """The device image of the tables as a host uint8 array (binary cache / replication)."""
        lib = self.lib
        n = C.c_int64()
        _lib.check(lib.grimb_tables_image_size(self.handle, C.byref(n)), "grimb_tables_image_size", lib)
        img = np.empty(n.value, dtype=np.uint8)
        _lib.check(lib.grimb_tables_image_copy(self.handle, img.ctypes.data), "grimb_tables_image_copy", lib)
        return img

    def from_image_host(self, img, alleles):
        self._set_dictionaries(alleles)
        lib = self.lib
        h = C.c_void_p()
        _lib.check(lib.grimb_tables_from_image(img.ctypes.data, img.nbytes, self.device, C.byref(h)),
                   "grimb_tables_from_image", lib)
        self.handle = h
        return self

    def write_csv(self, out_dir, **names):
        from .graph_io import write_csv
        write_csv(self, out_dir, **names)

    def save_cache(self, path):
        from .graph_io import save_cache
        save_cache(self, path)

    def load_cache(self, path):
        from .graph_io import load_cache
        return load_cache(self, path)

    def from_arrays(self, alleles, full_alleles, full_freqs):
        L, P = len(self.loci), len(self.pops)
        self._set_dictionaries(alleles)
        self._store = None
        lib = self.lib
        fa = np.ascontiguousarray(full_alleles, dtype=np.uint16)
        ff = np.ascontiguousarray(full_freqs, dtype=np.float64)
        d = _lib.TableDesc()
        d.n_loci, d.n_pops, d.n_full = L, P, fa.shape[0]
        d.full_alleles = fa.ctypes.data
        d.full_freqs = ff.ctypes.data
        for l in range(L):
            d.n_alleles[l] = len(alleles[l])
            d.key_bits[l] = self.key_bits[l]
        # the connector created last while the reference scans edges.csv (SURVEY trap T1):
        # parents are appended in `list(set difference)` order (generate_neo4j_multi_hpf.py:82-98)
        others = list(set(range(L)).difference([L - 1]))
        d.last_parent_locus = others[-1] if others else -1
        d.device = self.device
        # Plan_A_Matrix: the store holds the matrix labels (+ the single-locus labels), no connectors
        if self.plan_a_masks:
            store = np.asarray(self.config["store_label_masks"], dtype=np.uint32)
            d.label_masks = store.ctypes.data
            d.n_labels = len(store)
            d.n_plan_a_labels = len(self.plan_a_masks)
        h = C.c_void_p()
        _lib.check(lib.grimb_tables_build(C.byref(d), C.byref(h)), "grimb_tables_build", lib)
        self.handle = h
        return self

    def info(self):
        i = _lib.TableInfo()
        _lib.check(self.lib.grimb_tables_info(self.handle, C.byref(i)), "grimb_tables_info", self.lib)
        return {k: getattr(i, k) for k, _ in _lib.TableInfo._fields_}

    def export(self):
        """Copies the device arrays back (tests / CSV export)."""
        i = self.info()
        L, P, n = i["n_loci"], i["n_pops"], i["n_nodes"]
        out = {
            "node_key": np.zeros(n * self.kw, np.uint64), "node_freq": np.zeros((n, P), np.float64),
            "tl_start": np.zeros(n, np.uint32), "tl_cnt": np.zeros(n, np.uint32),
            "tl_adj": np.zeros(max(1, i["n_toplinks"]), np.uint32),
            "cn_start": np.zeros((n, L), np.uint32), "cn_cnt": np.zeros((n, L), np.uint32),
            "cn_adj": np.zeros(max(1, i["n_conn_edges"]), np.uint32),
            "label_first": np.zeros(1 << L, np.uint32), "label_count": np.zeros(1 << L, np.uint32),
        }
        order = ["node_key", "node_freq", "tl_start", "tl_cnt", "tl_adj", "cn_start", "cn_cnt", "cn_adj",
                 "label_first", "label_count"]
        _lib.check(self.lib.grimb_tables_export(self.handle, *[out[k].ctypes.data for k in order]),
                   "grimb_tables_export", self.lib)
        if self.kw == 2:   # 128-bit keys: Python ints (little-endian word pairs)
            w = out["node_key"].reshape(n, 2)
            out["node_key"] = np.array([int(lo) | (int(hi) << 64) for lo, hi in w], dtype=object)
        return out

    # ---- store seam: the reference's Graph queries, answered on the host from a copy of the
    # device arrays (networkx_graph.py:215-321).  The engine never calls these; they serve callers that
    # inspect the graph.  The copy is made on first use (export(): the whole table image comes back, so
    # this is for tables of README / NEMO size, not for an HBM-filling nine-locus table).
    def store_view(self):
        if getattr(self, "_store", None) is None:
            from .store_view import StoreView
            arr = self.export()
            arr["bits"] = list(self.key_bits)
            self._store = StoreView(arr, self.loci, self.config["loci_map"], self.alleles)
        return self._store

    def haps_by_label(self, label):
        return self.store_view().haps_by_label(label)

    def haps_with_probs_by_label(self, label):
        return self.store_view().haps_with_probs_by_label(label)

    def adjs_query(self, alleleList):
        return self.store_view().adjs_query(alleleList)

    def adjs_query_by_color(self, alleleList, labelA, labelB):
        return self.store_view().adjs_query_by_color(alleleList, labelA, labelB)

    def node_probs(self, nodes, label):
        return self.store_view().node_probs(nodes, label)

    def engine(self, workspace_bytes):
        e = self._engines.get(workspace_bytes)
        if e is None:
            e = C.c_void_p()
            _lib.check(self.lib.grimb_engine_create(self.handle, workspace_bytes, C.byref(e)),
                       "grimb_engine_create", self.lib)
            self._engines[workspace_bytes] = e
        return e

    def close(self):
        lib = self.lib
        if lib is None:
            return
        for e in self._engines.values():
            lib.grimb_engine_free(e)
        self._engines = {}
        self._store = None
        if self.handle is not None:
            lib.grimb_tables_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
