"""TEST-ONLY: runs the product's host code (tokeniser / formatter in grim.imputation.impute)
against the single-thread emulation build of the kernel source (tests/emu/), with tables
laid out by plain numpy from the oracle graph.  This checks the integer/FP64 logic of the CUDA
source on a machine without a GPU; it is not a product path (see tests/emu/grimb_emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from grim.imputation import _lib
from grim.imputation.impute import Imputation
from grim.imputation.networkx_graph import key_layout, loci_in_order

HERE = os.path.dirname(os.path.abspath(__file__))
# GRIMB_EMU_SO: an alternative build of the same source (e.g. -fsanitize=address,undefined; tests/emu/build.sh asan)
EMU_SO = os.environ.get("GRIMB_EMU_SO") or os.path.join(HERE, "emu", "libgrimb_emu.so")

M64 = (1 << 64) - 1


def hash_key(k):
    """grimb_group.h hash_key() for 64-bit keys."""
    M32 = 0xFFFFFFFF
    h = (k & M32) ^ (((k >> 32) * 0x9E3779B1) & M32)
    h ^= h >> 16
    h = (h * 0x85ebca6b) & M32
    h ^= h >> 13
    h = (h * 0xc2b2ae35) & M32
    h ^= h >> 16
    return h


class EmuTables(C.Structure):
    _fields_ = [
        ("L", C.c_int32), ("P", C.c_int32), ("n_nodes", C.c_uint32), ("n_full", C.c_uint32),
        ("shift", C.c_uint8 * 9), ("width", C.c_uint8 * 9), ("n_alleles", C.c_uint32 * 9),
        ("label_first", C.c_void_p), ("label_count", C.c_void_p), ("ht_off", C.c_void_p),
        ("ht_mask", C.c_void_p), ("slots", C.c_void_p), ("node_key", C.c_void_p), ("freq", C.c_void_p),
        ("tl_start", C.c_void_p), ("tl_cnt", C.c_void_p), ("tl_adj", C.c_void_p),
        ("cn_start", C.c_void_p), ("cn_cnt", C.c_void_p), ("cn_adj", C.c_void_p),
    ]


ADJ_FAULT = 0xFFFFFFFF


def arrays_from_oracle(g, loci):
    """OracleGraph -> the table arrays of csrc/grimb_tables.h (reference node ids)."""
    L, P = len(loci), len(g.pops)
    full = g.full_label
    per_locus = [set() for _ in range(L)]
    for name in g.by_label[full]:
        for l, a in enumerate(name.split("~")):
            per_locus[l].add(a)
    alleles = [sorted(s) for s in per_locus]
    aid = [{a: i + 1 for i, a in enumerate(al)} for al in alleles]
    bits = key_layout([len(a) for a in alleles])
    shift = [sum(bits[:l]) for l in range(L)]
    n = g.n_nodes
    node_key = np.zeros(n, np.uint64)
    freq = np.zeros((n, P), np.float64)
    label_first = np.zeros(1 << L, np.uint32)
    label_count = np.zeros(1 << L, np.uint32)
    label_mask = {}
    for lab in g.labels:
        m = 0
        for ch in lab:
            m |= 1 << full.index(ch)
        label_mask[lab] = m
        names = g.by_label[lab]
        if names:
            label_first[m] = g.node[names[0]][2]
        label_count[m] = len(names)
    node_label = np.zeros(n, np.uint32)
    for name, (lab, vec, i) in g.node.items():
        m = label_mask[lab]
        k = 0
        pos = [l for l in range(L) if m >> l & 1]
        for l, a in zip(pos, name.split("~")):
            k |= aid[l][a] << shift[l]
        node_key[i] = k
        freq[i] = vec
        node_label[i] = m
    # top links
    tl_start = np.zeros(n, np.uint32)
    tl_cnt = np.zeros(n, np.uint32)
    adj = []
    for name, lst in g.toplinks.items():
        i = g.node[name][2]
        tl_start[i] = len(adj)
        if isinstance(lst, list):
            tl_cnt[i] = len(lst)
            adj.extend(lst)
        else:
            tl_cnt[i] = ADJ_FAULT
    tl_adj = np.array(adj if adj else [0], np.uint32)
    # connectors
    cn_start = np.zeros((n, L), np.uint32)
    cn_cnt = np.zeros((n, L), np.uint32)
    cadj = []
    for (plab, child), lst in g.conn.items():
        ci = g.node[child][2]
        added = label_mask[plab] & ~int(node_label[ci])
        l = added.bit_length() - 1
        cn_start[ci, l] = len(cadj)
        if isinstance(lst, list):
            cn_cnt[ci, l] = len(lst)
            cadj.extend(lst)
        else:
            cn_cnt[ci, l] = ADJ_FAULT
    cn_adj = np.array(cadj if cadj else [0], np.uint32)
    # hash regions
    ht_off = np.zeros(1 << L, np.uint64)
    ht_mask = np.zeros(1 << L, np.uint32)
    so = 0
    for m in range(1 << L):
        sz = 2
        while sz < (4 if m == (1 << L) - 1 else 2) * int(label_count[m]):
            sz <<= 1
        ht_off[m] = so
        ht_mask[m] = sz - 1
        so += sz
    slots = np.zeros(so, dtype=[("key", np.uint64), ("node", np.uint32), ("pad", np.uint32)])
    slots["key"] = M64
    slots["node"] = 0xFFFFFFFF
    for i in range(n):
        m = int(node_label[i])
        k = int(node_key[i])
        h = hash_key(k) & int(ht_mask[m]) & ~1      # ht_home(): sector-aligned
        base = int(ht_off[m])
        while slots["node"][base + h] != 0xFFFFFFFF:
            h = (h + 1) & int(ht_mask[m])
        slots["key"][base + h] = k
        slots["node"][base + h] = i
    return {
        "alleles": alleles, "bits": bits, "shift": shift, "node_key": node_key, "freq": freq,
        "label_first": label_first, "label_count": label_count, "tl_start": tl_start, "tl_cnt": tl_cnt,
        "tl_adj": tl_adj, "cn_start": cn_start, "cn_cnt": cn_cnt, "cn_adj": cn_adj, "ht_off": ht_off,
        "ht_mask": ht_mask, "slots": slots, "n_full": g.n_full,
    }


def build_emu():
    src = os.path.join(HERE, "emu", "grimb_emu.cpp")
    hdr_dir = os.path.join(HERE, "..", "py-graph-imputation_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(hdr_dir, f)) for f in os.listdir(hdr_dir) if f.endswith(".h"))
    newest = max(newest, os.path.getmtime(src))
    if not os.path.exists(EMU_SO) or os.path.getmtime(EMU_SO) < newest:
        subprocess.run(["sh", os.path.join(HERE, "emu", "build.sh")], check=True)
    lib = C.CDLL(EMU_SO)
    lib.grimb_emu_impute.argtypes = [C.POINTER(EmuTables), C.POINTER(_lib.Config), C.POINTER(_lib.Batch),
                                     C.POINTER(_lib.Results), C.c_uint64]
    return lib


class EmuGraph(object):
    """Stands in for grim.imputation.networkx_graph.Graph in emulation tests."""

    def __init__(self, oracle_graph, loci_map):
        self.loci = loci_in_order(loci_map)
        self.pops = list(oracle_graph.pops)
        self.arr = arrays_from_oracle(oracle_graph, self.loci)
        self.alleles = self.arr["alleles"]
        self.allele_id = [{a: i + 1 for i, a in enumerate(al)} for al in self.alleles]
        self.key_bits = self.arr["bits"]
        self.shift = self.arr["shift"]
        a = self.arr
        t = EmuTables()
        t.L, t.P, t.n_nodes, t.n_full = len(self.loci), len(self.pops), len(a["node_key"]), a["n_full"]
        for l in range(len(self.loci)):
            t.shift[l] = a["shift"][l]
            t.width[l] = a["bits"][l]
            t.n_alleles[l] = len(self.alleles[l])
        for k in ("label_first", "label_count", "ht_off", "ht_mask", "slots", "node_key", "freq", "tl_start",
                  "tl_cnt", "tl_adj", "cn_start", "cn_cnt", "cn_adj"):
            a[k] = np.ascontiguousarray(a[k])
            setattr(t, k, a[k].ctypes.data)
        self.tables = t
        self.emu = build_emu()
        self.kw = 1
        self.lib = _lib.load(1)   # host-only halves of the text pipeline (no CUDA calls)
        # Plan_A_Matrix: the rows as locus bit masks (what networkx_graph.Graph records)
        self.plan_a_masks = None
        if getattr(oracle_graph, "plan_a", None):
            full = oracle_graph.full_label
            self.plan_a_masks = [sum(1 << full.index(ch) for ch in lab) for lab in oracle_graph.plan_a]


def emu_imputation(emu_graph, config, count_by_prob=None, arena=64 << 20):
    imp = Imputation(emu_graph, config, count_by_prob)

    def backend(cfg, batch, res, workspace):
        return emu_graph.emu.grimb_emu_impute(C.byref(emu_graph.tables), C.byref(cfg), C.byref(batch),
                                              C.byref(res), arena)

    imp._backend = backend
    return imp


def emu_impute_text(imp, emu_graph, data, first_index=0, arena=64 << 20):
    """The C++ text pipeline's two host halves (grimb_text_tokenise / grimb_text_format from
    libgrimb200.so -- no CUDA calls) around the emulated kernel source.  CPU test aid."""
    import numpy as np
    lib = _lib.load()
    t = imp._text_handle()
    b = _lib.Batch()
    _lib.check(lib.grimb_text_tokenise(t, C.byref(imp.cfg), data, len(data), first_index, C.byref(b)), "tokenise")
    S = b.n_subjects
    res = _lib.ResultArrays(S, 1, general=S + 16, hap=4096, pop=4096)
    while True:
        r = res.struct
        rc = emu_graph.emu.grimb_emu_impute(C.byref(emu_graph.tables), C.byref(imp.cfg), C.byref(b), C.byref(r), arena)
        if rc == _lib.E_CAPACITY:
            res.grow()
            continue
        assert rc == 0
        break
    out = _lib.TextOut()
    _lib.check(lib.grimb_text_format(t, C.byref(imp.cfg), C.byref(r), C.byref(out)), "format")
    return {k: C.string_at(out.data[i], out.size[i]).decode("utf8") for i, k in enumerate(_lib.OUT_KEYS)}
