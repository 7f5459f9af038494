"""Per-configuration records for bench.py (BASELINE.json configs[2..4]: C3, C4, C5) and for the table build
(K0): every kernel of the path gets a CUDA-event time, its algorithmic bytes by SURVEY 8(d) and a roofline
fraction in the same driver run, next to an oracle-checked sample.  Imported by bench.py only.

Per configuration:
  text    text in -> six texts out through grimb_impute_text (tokeniser + C ABI with host buffers + formatter)
  device  the tokenised batch resident in HBM, grimb_impute_device, CUDA events (whole call and per kernel)
  bytes   per SURVEY 8(d): the subject's input + 32 B per probe issued + one frequency vector (32 * ceil(8P / 32) B)
          per vector read + 4 B per adjacency entry followed (bounded by the vectors read) + the result bytes
  parity  the first `sample` lines against the CPU oracle (forked workers)
"""
import ctypes as C
import json
import os
import time

import numpy as np


def _arr(ptr, n, ct):
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(max(1, n),))[:n].copy()


def device_batch(lib, imp, data, torch, dev):
    """Tokenises `data` (C++ text pipeline) and places the batch arrays in HBM.  -> (Batch with device pointers,
    keep-alive list, number of subjects, input bytes per subject)."""
    from grim.imputation import _lib
    t = imp._text_handle()
    hb = _lib.Batch()
    _lib.check(lib.grimb_text_tokenise(t, C.byref(imp.cfg), data, len(data), 0, C.byref(hb)), "tokenise", lib)
    S, L, P = hb.n_subjects, imp.L, imp.P
    keep = []

    def put(ptr, n, ct, tdt):
        if not ptr:
            return None
        a = _arr(ptr, n, ct)
        tt = torch.from_numpy(a.view(tdt)).to(dev)
        keep.append(tt)
        return tt.data_ptr()

    db = _lib.Batch()
    db.n_subjects = S
    n_al = hb.n_alleles_total
    db.typed_mask = put(hb.typed_mask, S, C.c_uint16, np.int16)
    db.counts = put(hb.counts, S * L * 2, C.c_uint16, np.int16)
    db.allele_off = put(hb.allele_off, S + 1, C.c_uint32, np.int32)
    db.alleles = put(hb.alleles, max(1, n_al), C.c_uint16, np.int16)
    db.n_alleles_total = n_al
    db.prior_index = put(hb.prior_index, S, C.c_uint32, np.int32)
    db.priors = put(hb.priors, hb.n_priors * P * P, C.c_double, np.float64)
    db.n_priors = hb.n_priors
    db.phase_mask = None
    in_bytes = S * 2 + (S + 1) * 4 + n_al * 2 + (S * L * 4 if hb.counts else 0) + (S * 4 if hb.prior_index else 0)
    return db, keep, S, in_bytes


def device_results(torch, dev, S, kw, caps):
    from grim.imputation import _lib
    sz = {"compact": 16, "words": 8, "general": 48, "hap_rows": 24 if kw == 1 else 40, "pop_rows": 16}
    cap = dict(caps, compact=S)
    bufs = {k: torch.zeros(max(16, cap[k]) * sz[k], dtype=torch.uint8, device=dev) for k in sz}
    totals = np.zeros(9, np.int64)
    r = _lib.Results()
    r.compact = bufs["compact"].data_ptr()
    r.words, r.word_capacity = bufs["words"].data_ptr(), max(16, cap["words"])
    r.general, r.general_capacity = bufs["general"].data_ptr(), max(16, cap["general"])
    r.hap_rows, r.hap_capacity = bufs["hap_rows"].data_ptr(), max(16, cap["hap_rows"])
    r.pop_rows, r.pop_capacity = bufs["pop_rows"].data_ptr(), max(16, cap["pop_rows"])
    r.totals = totals.ctypes.data
    return r, bufs, totals, sz


def measure(name, g, cfg, cbp, lines, oracle_graph, conf, sample, torch, dev, stream, peak, steps=5, workspace=None,
            note=None):
    """One configuration on a built Graph `g`.  -> record (dict)."""
    import oracle_par
    from grim.imputation import _lib
    from grim.imputation.impute import Imputation
    lib = g.lib
    data = "".join(lines).encode("utf8")
    imp = Imputation(g, cfg, cbp)
    if workspace:
        imp.workspaces = workspace
    imp.impute_text(data)                       # warm-up: engines, pinned staging buffers
    warm = imp
    imp = Imputation(g, cfg, cbp)
    if workspace:
        imp.workspaces = workspace
    imp._text = warm._text                      # same GrimbText: its staging is sized by the warm-up pass
    # three calls, the median reported (single calls of the nine-locus configurations showed sporadic stalls of
    # 0.1-1 s inside the ABI call that are not understood yet: all three times are kept in the record)
    text_runs = []
    for _rep in range(3):
        imp = Imputation(g, cfg, cbp)
        if workspace:
            imp.workspaces = workspace
        imp._text = warm._text
        t0 = time.time()
        texts = imp.impute_text(data)
        text_runs.append((time.time() - t0, imp))
    text_runs.sort(key=lambda x: x[0])
    t_text, imp = text_runs[1]
    # device-resident leg
    db, keep, S, in_bytes = device_batch(lib, imp, data, torch, dev)
    caps = {"words": 0, "general": 0, "hap_rows": 0, "pop_rows": 0}
    eng = g.engine(imp.workspaces[0])
    P, L = imp.P, imp.L
    while True:
        r, bufs, totals, sz = device_results(torch, dev, S, g.kw, caps)
        rc = lib.grimb_impute_device(eng, C.byref(imp.cfg), C.byref(db), C.byref(r), C.c_void_p(stream.cuda_stream))
        if rc == _lib.E_CAPACITY:
            caps = {"words": int(totals[0]) + 64, "general": int(totals[1]) + 64, "hap_rows": int(totals[2]) + 64,
                    "pop_rows": int(totals[3]) + 64}
            continue
        _lib.check(rc, "grimb_impute_device", lib)
        break
    comp = bufs["compact"].cpu().numpy().view(_lib.COMPACT_DTYPE)
    ws = int((comp["status"] == _lib.ST_WORKSPACE).sum())     # served by a bigger workspace tier in the text leg
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kms = {"k_impute": [], "k_impute_typed": [], "k_impute_slots": []}
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        ev[0].record(stream)
        _lib.check(lib.grimb_impute_device(eng, C.byref(imp.cfg), C.byref(db), C.byref(r), C.c_void_p(stream.cuda_stream)),
                   "grimb_impute_device", lib)
        ev[1].record(stream)
        torch.cuda.synchronize()
        ms.append(ev[0].elapsed_time(ev[1]))
        kms["k_impute"].append(lib.grimb_engine_kernel_ms(eng, 1))
        kms["k_impute_typed"].append(lib.grimb_engine_kernel_ms(eng, 2))
        kms["k_impute_slots"].append(lib.grimb_engine_kernel_ms(eng, 5))
    ms_call = float(np.median(ms))
    k_typed = float(np.median(kms["k_impute_typed"]))
    k_gen = float(np.median(kms["k_impute"]))
    out_bytes = S * 16 + int(totals[0]) * 8 + int(totals[1]) * 48 + int(totals[2]) * sz["hap_rows"] + int(totals[3]) * 16
    vec_bytes = 32 * ((8 * P + 31) // 32)
    probes, hits, vecs = int(totals[6]), int(totals[7]), int(totals[8])
    algo = in_bytes + probes * 32 + vecs * (vec_bytes + 4) + out_bytes
    k_slots = float(np.median(kms["k_impute_slots"]))
    # the kernel that finishes the bulk of the subjects; the other one is the tail of the call
    handed = int(totals[5])
    dominant = "k_impute_typed" if (k_typed > 0 and handed < S // 2) else "k_impute"
    k_ms = k_typed if dominant == "k_impute_typed" else k_gen + max(0.0, k_slots)
    # oracle sample
    ns = min(sample, len(lines))
    mine = {k: v.decode("utf8") for k, v in Imputation(g, cfg, cbp).impute_text("".join(lines[:ns]).encode("utf8")).items()} \
        if not workspace else None
    if mine is None:
        imp_s = Imputation(g, cfg, cbp)
        imp_s.workspaces = workspace
        mine = {k: v.decode("utf8") for k, v in imp_s.impute_text("".join(lines[:ns]).encode("utf8")).items()}
    t0 = time.time()
    ref, _e = oracle_par.oracle_texts(oracle_graph, conf, lines[:ns], cbp)
    t_cpu = time.time() - t0
    rec = {
        "subjects": S, "populations": P, "loci": L, "key_words": g.kw,
        "subjects_per_s": S / (ms_call * 1e-3), "ms_per_call": ms_call,
        "pair_evals_per_s": int(totals[4]) / (ms_call * 1e-3),
        "kernel": dominant, "kernel_ms": k_ms, "k_impute_typed_ms": k_typed, "k_impute_ms": k_gen,
        "k_impute_slots_ms": k_slots,
        "handed_to_general_kernel": int(totals[5]), "workspace_overflow_in_tier0": ws,
        "probes": probes, "probe_hits": hits, "vectors_read": vecs,
        "algorithmic_bytes": algo, "algorithmic_bytes_per_subject": algo / max(1, S),
        "roofline": {"bound": "hbm", "achieved": algo / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None, "peak": peak, "unit": "GB/s",
                     "frac": (algo / (k_ms * 1e-3) / 1e9 / peak) if k_ms > 0 else None},
        "e2e_text": {"subjects_per_s": S / t_text, "out_bytes": sum(len(v) for v in texts.values()),
                     "tokenise_s": imp.stats.get("tokenise_seconds"), "abi_s": imp.stats.get("abi_seconds"),
                     "format_s": imp.stats.get("format_seconds"),
                     # subjects re-issued on a bigger workspace tier in this call (the first use of a tier creates its
                     # engine: a device allocation of tens of GB, 0.2-0.6 s, inside the call)
                     "workspace_retries": imp.stats.get("workspace_retries"),
                     "seconds_of_three_calls": [round(x[0], 4) for x in text_runs]},
        "plans": imp.stats["plan"],
        "oracle": {"sample": ns, "identical": all(mine[k] == ref[k] for k in ref), "subjects_per_s_all_cores": ns / t_cpu,
                   "cores": os.cpu_count()},
    }
    if note:
        rec["note"] = note
    return rec


def table_record(g, n_full, P, L, build_s, peak, labels=None):
    """K0 as a streaming kernel chain: bytes = per marginal label the (key, index) pairs read and written by the
    sort-by-key (12 B each way) and the members' frequency vectors read by the ordered segmented sum, plus the
    table image written once."""
    info = g.info()
    labels = (1 << L) - 2 if labels is None else labels
    algo = labels * n_full * (24 + 8 * P) + info["device_bytes"]
    dev_s = float(g.lib.grimb_tables_build_ms(g.handle)) * 1e-3     # CUDA events inside grimb_tables_build
    t = dev_s if dev_s > 0 else build_s
    return {"n_full": n_full, "n_nodes": info["n_nodes"], "populations": P, "loci": L, "labels_built": labels + 1,
            "device_bytes": info["device_bytes"],
            "build_s": build_s, "build_device_s": dev_s, "kernel_launches": int(g.lib.grimb_tables_build_launches(g.handle)),
            "algorithmic_bytes": algo,
            "note": "build_s = wall time of Graph.from_arrays (host staging + copies + kernels); the roofline uses the "
                    "device time of the kernel chain",
            "roofline": {"bound": "hbm", "achieved": algo / t / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": algo / t / 1e9 / peak}}


def run_all(args, names, fa, ff1, torch, dev, stream, peak):
    """C3 on bench.py's own 1M-haplotype table x 21 populations, C4 (messy + heavy) on the README table,
    C5 on a nine-locus table with 128-bit keys.  -> dict of records."""
    import goldenlib
    import grim_oracle as go
    import synth
    from grim.imputation.networkx_graph import Graph
    from grim.run_impute_def import load_config
    from oracle_graph_np import NumpyOracleGraph
    out = {}
    base = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    lm5 = {"A": 1, "B": 2, "C": 3, "DQB1": 4, "DRB1": 5}
    # ---- C3
    pops = ["P%02d" % i for i in range(21)]
    ff = synth.multipop_freqs(ff1[:, 0], 21, 21)
    _ct, ratio = synth.pop_counts(pops)
    conf = dict(base)
    conf.update({"populations": pops, "loci_map": lm5, "UNK_priors": "MR", "number_of_pop_results": 100, "number_of_results": 10})
    cfg = load_config(conf)
    t0 = time.time()
    g = Graph(cfg, device=dev.index).from_arrays(names, fa, ff)
    torch.cuda.synchronize()
    t_build = time.time() - t0
    p = ff.mean(axis=1)
    p = p / p.sum()
    lines = synth.array_subject_lines(names, fa, p, args.c3_subjects, 3, synth.race_fields(pops), variants=False)
    og = NumpyOracleGraph(names, fa, ff, pops, lm5)
    out["C3"] = measure("C3", g, cfg, ratio, lines, og, conf, args.config_sample, torch, dev, stream, peak,
                        note="%d fully typed subjects, race-field shapes cycled, on the C2 table x 21 populations "
                             "(168-byte frequency vectors), top-100 population rows" % len(lines))
    out["K0"] = {"C3_table": table_record(g, len(fa), 21, 5, t_build, peak)}
    g.close()
    del og, ff
    # ---- C4 on the README table
    _t, conf4, _l, _e = goldenlib.load_case("g1_readme_donor")
    cfg4 = load_config(conf4)
    t0 = time.time()
    g4 = Graph(cfg4, device=dev.index).build_graph()
    torch.cuda.synchronize()
    tab = synth.Table(open(conf4["freq_file"]).read())
    og4 = go.graph_from_config(conf4)
    cbp4 = go.count_by_prob_from_file(1, conf4["pops_count_file"])
    messy = synth.messy_subjects(tab, args.c4_subjects, 4, max_amb=6)
    out["C4_messy"] = measure("C4_messy", g4, cfg4, cbp4, messy, og4, conf4, min(60, args.config_sample), torch, dev, stream,
                              peak, steps=3, note="<= 6 alleles per side, missing loci, unknown alleles; README table")
    heavy = synth.heavy_subjects(tab, args.c4_heavy_subjects, 44, races=["CAU,CAU"])
    out["C4_heavy"] = measure("C4_heavy", g4, cfg4, cbp4, heavy, og4, conf4, min(16, args.config_sample), torch, dev, stream,
                              peak, steps=3, workspace=[512 << 20, 4 << 30],
                              note="6-40 alleles per side, products on both sides of the 100,000-option threshold, "
                                   "0-3 missing loci; README table; 512 MB workspace tier")
    wide = synth.wide_subjects(tab, args.c4_wide_subjects, 45, races=["CAU,CAU"])
    out["C4_wide"] = measure("C4_wide", g4, cfg4, cbp4, wide, og4, conf4, min(8, args.config_sample), torch, dev, stream, peak,
                             steps=3, workspace=[128 << 20, 4 << 30],
                             note="fully typed, 7-9 alleles per side: Cartesian products of 16,807-59,049 candidates per "
                                  "phase and side, under the threshold (cooperative slot pass); README table")
    g4.close()
    # ---- C5
    loci = ["A", "B", "C", "DPA1", "DPB1", "DQA1", "DQB1", "DRB1", "DRBX"]
    n_all = [700, 1200, 600, 40, 300, 60, 250, 700, 100]
    pops5 = ["Q%d" % i for i in range(5)]
    lm9 = {l: i + 1 for i, l in enumerate(loci)}
    names9, fa9, base_f = synth.zipf_arrays(args.c5_haps, n_all, 20261018, loci)
    ff9 = synth.multipop_freqs(base_f, 5, 9, zero_frac=0.2)
    _ct, ratio5 = synth.pop_counts(pops5)
    conf5 = dict(base)
    conf5.update({"populations": pops5, "UNK_priors": "MR", "number_of_pop_results": 100, "loci_map": lm9,
                  "Plan_B_Matrix": [[[1, 2, 3, 4, 5, 6, 7, 8, 9]], [[1, 2, 3], [4, 5], [6, 7, 8, 9]],
                                    [[1], [2, 3], [4, 5], [6, 7], [8, 9]], [[1], [2], [3], [4], [5], [6], [7], [8], [9]]]})
    cfg5 = load_config(conf5)
    t0 = time.time()
    g5 = Graph(cfg5, device=dev.index).from_arrays(names9, fa9, ff9)
    torch.cuda.synchronize()
    t_build5 = time.time() - t0
    p5 = ff9.mean(axis=1)
    p5 = p5 / p5.sum()
    lines5 = synth.array_subject_lines(names9, fa9, p5, args.c5_subjects, 9, synth.race_fields(pops5), variants=False)
    og5 = NumpyOracleGraph(names9, fa9, ff9, pops5, lm9)
    out["C5"] = measure("C5", g5, cfg5, ratio5, lines5, og5, conf5, min(200, args.config_sample), torch, dev, stream, peak,
                        note="nine loci (256 phases), 5 populations, %d haplotypes, 128-bit keys" % len(fa9))
    out["K0"]["C5_table"] = table_record(g5, len(fa9), 5, 9, t_build5, peak)
    g5.close()
    del og5, names9, fa9, ff9
    # ---- C5 at N_full = 2M: only a Plan_A_Matrix makes a nine-locus table of that size fit (510 marginal labels per
    # haplotype otherwise); the store holds the matrix labels + the single-locus labels, Plan A only
    if args.c5m_haps > 0:
        matrix = [[1, 2, 3, 4, 5, 6, 7, 8, 9], [1, 2, 3, 7, 8], [1, 2, 3], [7, 8], [1, 2, 3, 4, 5, 6, 7, 8]]
        names9, fa9, base_f = synth.zipf_arrays(args.c5m_haps, n_all, 20261019, loci)
        ff9 = synth.multipop_freqs(base_f, 5, 9, zero_frac=0.2)
        conf5m = dict(conf5)
        conf5m.update({"Plan_A_Matrix": matrix, "planb": False})
        cfg5m = load_config(conf5m)
        t0 = time.time()
        g5 = Graph(cfg5m, device=dev.index).from_arrays(names9, fa9, ff9)
        torch.cuda.synchronize()
        t_build5 = time.time() - t0
        p5 = ff9.mean(axis=1)
        p5 = p5 / p5.sum()
        lines5 = synth.array_subject_lines(names9, fa9, p5, args.c5_subjects, 10, synth.race_fields(pops5), variants=False)
        og5 = NumpyOracleGraph(names9, fa9, ff9, pops5, lm9, plan_a_matrix=matrix)
        out["C5_matrix"] = measure("C5_matrix", g5, cfg5m, ratio5, lines5, og5, conf5m, min(200, args.config_sample), torch, dev,
                                   stream, peak, note="nine loci, 5 populations, %d haplotypes, 128-bit keys, Plan_A_Matrix of %d "
                                   "labels (Plan B off)" % (len(fa9), len(matrix)))
        out["K0"]["C5_matrix_table"] = table_record(g5, len(fa9), 5, 9, t_build5, peak, labels=len(matrix) + 9 - 1 - 1)
        g5.close()
    return out
