"""TEST-ONLY helper: the CPU oracle over forked workers (contiguous line ranges, global line indices),
so that oracle samples of a few thousand subjects fit the GPU suite's time budget.  The store is built in
the parent and inherited copy-on-write."""
import multiprocessing as mp
import os

import grim_oracle as go

KEYS = ("umug", "umug_pops", "pmug", "pmug_pops", "miss", "problem")


def oracle_texts(graph, conf, lines, count_by_prob=None, procs=None, warm=8):
    """-> (dict of the six texts, total pair evaluations)."""
    cfg = go.load_config(conf)
    procs = procs or min(len(lines) // 8 + 1, os.cpu_count() or 1)
    if warm:   # materialise the labels the first subjects touch before forking
        go.OracleImputation(graph, cfg, count_by_prob).impute_lines(lines[:warm])
    if procs <= 1:
        imp = go.OracleImputation(graph, cfg, count_by_prob)
        return imp.impute_lines(lines), imp.pair_evals
    n = len(lines)
    # interleave small blocks so that slow subjects spread over the workers
    blk = max(1, min(64, n // (procs * 4) or 1))
    bounds = list(range(0, n, blk)) + [n]

    def work(w, q):
        imp = go.OracleImputation(graph, cfg, count_by_prob)
        out = []
        for b in range(w, len(bounds) - 1, procs):
            out.append((b, imp.impute_lines(lines[bounds[b]:bounds[b + 1]], first_index=bounds[b])))
        q.put((out, imp.pair_evals))

    ctx = mp.get_context("fork")
    q = ctx.Queue()
    ps = [ctx.Process(target=work, args=(w, q)) for w in range(procs)]
    for p in ps:
        p.start()
    got = [q.get() for _ in ps]
    for p in ps:
        p.join()
    parts = sorted((b, t) for out, _e in got for b, t in out)
    texts = {k: "".join(t[k] for _b, t in parts) for k in KEYS}
    return texts, sum(e for _o, e in got)
