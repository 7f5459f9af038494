#!/usr/bin/env python
"""The north-star job through the PUBLIC API: grim.grim.impute(conf_file) on >= 1M synthetic 5-locus
subjects against a 1M-haplotype table, single process or one process per GPU under torchrun
(tables built on rank 0, one NCCL broadcast, subjects sharded, rank 0 writes the six files).  Reports
wall times and checks the first `--sample` subjects of the written files against the CPU oracle.

    python tests/tools/run_1m_job.py [--subjects N] [--haps H]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tests/tools/run_1m_job.py
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "py-graph-imputation_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--subjects", type=int, default=1 << 20)
    ap.add_argument("--haps", type=int, default=1000000)
    ap.add_argument("--sample", type=int, default=3000)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    from grim import grim
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    work = [None]
    names = fa = ff = alleles = None
    if rank == 0:
        t = time.time()
        d = tempfile.mkdtemp(prefix="grim1m_")
        names, fa, ff = bench.make_table(args.haps)
        cols = [np.array(names[l], dtype=object)[fa[:, l] - 1] for l in range(5)]
        with open(d + "/hpf.csv", "w") as f:
            f.write("hap,pop,freq\n")
            f.writelines("%s~%s~%s~%s~%s,CAU,%r\n" % (cols[0][i], cols[1][i], cols[2][i], cols[3][i], cols[4][i], float(ff[i, 0]))
                         for i in range(len(fa)))
        open(d + "/cnt.txt", "w").write("CAU,%d.0,1.0\n" % len(fa))
        _batch, alleles = bench.make_subjects(fa, ff, args.subjects, bench.SUBJECT_SEED)
        with open(d + "/subjects.csv", "w") as f:
            for lo in range(0, args.subjects, 65536):
                f.writelines(bench.subject_lines(names, alleles, lo, min(args.subjects, lo + 65536)))
        conf = bench.base_conf()
        conf.update({"freq_file": d + "/hpf.csv", "pops_count_file": d + "/cnt.txt", "freq_trim_threshold": 1e-30,
                     "imputation_in_file": d + "/subjects.csv", "imputation_out_path": d + "/out",
                     "graph_files_path": d + "/csv/"})
        json.dump(conf, open(d + "/conf.json", "w"))
        work = [d]
        print("inputs written in %.1f s: %s" % (time.time() - t, d), flush=True)
    if world > 1:
        dist.broadcast_object_list(work, src=0)
        dist.barrier()
    d = work[0]
    import contextlib
    import io
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        g = grim.impute(conf_file=d + "/conf.json", device=local)
    if world > 1:
        dist.barrier()
    t_first = time.time() - t0          # includes reading hpf.csv, the table build and its broadcast
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        grim.impute(conf_file=d + "/conf.json", graph=g, device=local)
    if world > 1:
        dist.barrier()
    t_again = time.time() - t0          # tables reused: read input, impute, gather, write the six files
    if rank == 0:
        import grim_oracle as go
        conf = json.load(open(d + "/conf.json"))
        og = bench._FullOnlyGraph(names, fa, ff)
        oimp = go.OracleImputation(og, go.load_config(conf), np.ones(1))
        lines = bench.subject_lines(names, alleles, 0, args.sample)
        ref = oimp.impute_lines(lines)
        out_names = {"umug": "imputation_out_umug_freq_filename", "umug_pops": "imputation_out_umug_pops_filename",
                     "pmug": "imputation_out_hap_freq_filename", "pmug_pops": "imputation_out_hap_pops_filename"}
        same = True
        rows = {}
        for k, ck in out_names.items():
            text = open(os.path.join(d, "out", conf[ck])).read()
            rows[k] = text.count("\n")
            same = same and text.startswith(ref[k]) and (len(ref[k]) > 0)
        miss = open(os.path.join(d, "out", conf["imputation_out_miss_filename"])).read().count("\n")
        prob = open(os.path.join(d, "out", conf["imputation_out_problem_filename"])).read().count("\n")
        print(json.dumps({
            "job": "grim.grim.impute(conf_file) on %d subjects, %d-haplotype table" % (args.subjects, args.haps),
            "n_gpus": world, "wall_s_first_call_incl_table_build": t_first, "wall_s_with_tables_resident": t_again,
            "subjects_per_s_with_tables_resident": args.subjects / t_again, "rows": rows, "miss": miss, "problem": prob,
            "first_%d_subjects_identical_to_oracle" % args.sample: bool(same)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
