"""On-disk graph formats.

* write_csv: nodes.csv / edges.csv / top_links.csv / info_node.csv in the reference's layout
  (graph_generation/generate_neo4j_multi_hpf.py:419-486 of the reference) from the device
  tables, so tools that consume a GRIM graph directory keep working.  nodes.csv, edges.csv and
  info_node.csv are byte-identical to the reference's; top_links.csv has the same rows, ordered
  by full-haplotype id inside each partial node (the reference iterates a Python `set`, whose
  order is not reproducible between runs).
* read_nodes_csv: an existing GRIM graph directory -> the full-haplotype rows the device build
  needs (marginal rows are recomputed on the GPU and equal the file's).
* save_cache / load_cache: binary image of the device tables + the allele dictionaries, for
  fast start-up of impute(graph=None).
"""
import csv
import json
import os

import numpy as np


def _fmt(v):
    # absent populations are the int 0 in the reference ("0"); sums of present ones are floats
    return "0" if v == 0.0 else repr(float(v))


def _label_string(mask, L):
    return "".join(str(l + 1) for l in range(L) if mask >> l & 1)


def _label_order(L):
    """(mask, label string) in the reference's order: full, then combinations of decreasing size."""
    import itertools
    out = [((1 << L) - 1)]
    for r in range(L - 1, 0, -1):
        for c in itertools.combinations(range(L), r):
            m = 0
            for l in c:
                m |= 1 << l
            out.append(m)
    return out


def _kfield(keys, shift, bits):
    """Allele-id field of packed keys: uint64 array (64-bit keys) or object array of ints (128-bit)."""
    if keys.dtype == object:
        m = (1 << bits) - 1
        return np.array([(int(k) >> shift) & m for k in keys], dtype=np.int64)
    return ((keys >> np.uint64(shift)) & np.uint64((1 << bits) - 1)).astype(np.int64)


def _kand(keys, mask):
    if keys.dtype == object:
        return np.array([int(k) & mask for k in keys], dtype=object)
    return keys & np.uint64(mask)


def node_names(graph, arrays):
    """Name of every node id, from the packed keys."""
    L = len(graph.loci)
    keys = arrays["node_key"]
    cols = []
    for l in range(L):
        ids = _kfield(keys, graph.shift[l], graph.key_bits[l])
        names = np.array([""] + list(graph.alleles[l]), dtype=object)
        cols.append(names[ids])
    out = []
    for i in range(len(keys)):
        out.append("~".join(c[i] for c in cols if c[i] != ""))
    return out


def write_csv(graph, out_dir, node_csv="nodes.csv", edges_csv="edges.csv", top_links_csv="top_links.csv",
              info_node_csv="info_node.csv"):
    if graph.plan_a_masks:
        # the reference's generator also writes Plan-B labels whose order (hence node ids) follows the hash seed of
        # the Python process (generate_neo4j_multi_hpf.py:181): there is no file to be equal to
        raise NotImplementedError("graph CSV export under a Plan_A_Matrix")
    os.makedirs(out_dir, exist_ok=True)
    a = graph.export()
    L, P = len(graph.loci), len(graph.pops)
    n = len(a["node_key"])
    names = node_names(graph, a)
    order = _label_order(L)
    label_of = np.zeros(n, dtype=np.int64)
    for m in order:
        f, c = int(a["label_first"][m]), int(a["label_count"][m])
        label_of[f:f + c] = m
    freq = a["node_freq"]
    n_full = int(a["label_count"][(1 << L) - 1])
    with open(os.path.join(out_dir, node_csv), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["haplotypeId:ID(HAPLOTYPE)", "name", "loci:LABEL", "frequency:DOUBLE[]"])
        for i in range(n):
            w.writerow([i, names[i], _label_string(int(label_of[i]), L), ";".join(_fmt(v) for v in freq[i])])
    # top links of a node list its full haplotypes in hpf order.  The reference's sentinel quirk
    #  truncates the LAST node's list only in the loaded graph, not in the files, so
    # the true degree is recomputed for it.
    tl_cnt = a["tl_cnt"].astype(np.int64).copy()
    if n > n_full:
        lm = int(label_of[n - 1])
        km = 0
        for l in range(L):
            if lm >> l & 1:
                km |= ((1 << graph.key_bits[l]) - 1) << graph.shift[l]
        tl_cnt[n - 1] = int(np.count_nonzero(_kand(a["node_key"][:n_full], km) == a["node_key"][n - 1]))
    with open(os.path.join(out_dir, top_links_csv), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([":START_ID(HAPLOTYPE)", ":END_ID(HAPLOTYPE)", ":TYPE"])
        for i in range(n_full, n):
            s = int(a["tl_start"][i])
            for fid in a["tl_adj"][s:s + int(tl_cnt[i])]:
                w.writerow([i, int(fid), "TOP"])
    # edges: for every partial node, for every full haplotype containing it (hpf order), one row
    # per locus that can be added (ascending), CP = full frequency / child frequency
    key_of = {(int(label_of[i]), int(a["node_key"][i])): i for i in range(n)}
    masks = [((1 << graph.key_bits[l]) - 1) << graph.shift[l] for l in range(L)]
    with open(os.path.join(out_dir, edges_csv), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([":START_ID(HAPLOTYPE)", ":END_ID(HAPLOTYPE)", "CP:DOUBLE[]", ":TYPE"])
        for i in range(n_full, n):
            m = int(label_of[i])
            add = list(set(range(L)).difference([l for l in range(L) if m >> l & 1]))
            s = int(a["tl_start"][i])
            for fid in a["tl_adj"][s:s + int(tl_cnt[i])]:
                fkey = int(a["node_key"][fid])
                for l in add:
                    pm = m | (1 << l)
                    pkey = 0
                    for q in range(L):
                        if pm >> q & 1:
                            pkey |= fkey & masks[q]
                    parent = key_of[(pm, pkey)]
                    cp = []
                    for p in range(P):
                        child = freq[i, p]
                        full_v = freq[fid, p]
                        if child == 0.0:
                            cp.append("0")
                        else:
                            cp.append(repr(float(full_v / child)))
                    w.writerow([i, parent, ";".join(cp), "CP"])
    with open(os.path.join(out_dir, info_node_csv), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["INFO_NODE_ID:ID(INFO_NODE)", "populations:STRING[]", "INFO_NODE:LABEL"])
        w.writerow([1, ";".join(graph.pops), "INFO_NODE"])


def read_nodes_csv(nodes_file, loci, full_label):
    """-> (alleles per locus, full_alleles uint16 [N][L], full_freqs [N][P]) from the full-label
    rows of a reference nodes.csv (networkx_graph.py:45-68 reads the same columns)."""
    L = len(loci)
    rows = []
    vecs = []
    with open(nodes_file) as f:
        r = csv.reader(f)
        next(r)
        for row in r:
            if len(row) > 0 and row[2] == full_label:
                rows.append(row[1].split("~"))
                vecs.append([float(x) for x in row[3].split(";")])
    alleles = [sorted({h[l] for h in rows}) for l in range(L)]
    ids = [{a: i + 1 for i, a in enumerate(alleles[l])} for l in range(L)]
    fa = np.zeros((len(rows), L), dtype=np.uint16)
    for l in range(L):
        fa[:, l] = [ids[l][h[l]] for h in rows]
    return alleles, fa, np.array(vecs, dtype=np.float64).reshape(len(rows), -1)


def save_cache(graph, path):
    """Binary table cache: <path> = device image, <path>.json = dictionaries + layout."""
    img = graph.image_to_host()
    img.tofile(path)
    meta = {"loci": graph.loci, "pops": graph.pops, "alleles": graph.alleles, "bytes": int(img.nbytes),
            "plan_a_masks": graph.plan_a_masks}
    with open(path + ".json", "w") as f:
        json.dump(meta, f)


def load_cache(graph, path):
    with open(path + ".json") as f:
        meta = json.load(f)
    if meta["loci"] != graph.loci or meta["pops"] != graph.pops:
        raise ValueError("table cache was built for other loci / populations")
    if meta.get("plan_a_masks") != graph.plan_a_masks:
        raise ValueError("table cache was built for another Plan_A_Matrix")
    img = np.fromfile(path, dtype=np.uint8)
    if img.nbytes != meta["bytes"]:
        raise ValueError("table cache is truncated")
    graph.from_image_host(img, meta["alleles"])
    return graph
