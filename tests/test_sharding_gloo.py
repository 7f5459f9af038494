"""CPU, world_size 2 over gloo: the multi-rank host path (contiguous shards, global line indices,
ordered concatenation on rank 0) reproduces the single-process output files.  The per-subject
work runs on the emulation backend here; on the GPU box the same code path runs with NCCL and
the CUDA backend (bench.py --gpus N, tests/test_gpu_parity.py)."""
import os
import sys
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, case, outdir):
    for p in (HERE, os.path.join(HERE, "..", "oracle"), os.path.join(HERE, "..", "py-graph-imputation_b200")):
        sys.path.insert(0, p)
    import torch.distributed as dist

    import goldenlib
    import grim_oracle as go
    from emu_backend import EmuGraph, emu_imputation
    from grim.imputation.multi_gpu import impute_lines_sharded
    from grim.run_impute_def import load_config
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    table, conf, lines, exp = goldenlib.load_case(case)
    eg = EmuGraph(go.graph_from_config(conf), conf["loci_map"])
    imp = emu_imputation(eg, load_config(conf))

    def gather(obj):
        parts = [None] * world if rank == 0 else None
        dist.gather_object(obj, parts, dst=0)
        return parts

    out = impute_lines_sharded(imp, lines, rank, world, gather)
    if rank == 0:
        for k, v in out.items():
            open(os.path.join(outdir, k), "w").write("".join(v))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["g2_edges", "g3_pop3_typed"])
def test_two_rank_sharding_matches_golden(case):
    import torch.multiprocessing as mp

    import goldenlib
    _, _, _, exp = goldenlib.load_case(case)
    port = 29500 + (os.getpid() % 2000)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, port, case, d), nprocs=2, join=True)
        for k in goldenlib.KEYS:
            assert open(os.path.join(d, k)).read() == exp[k], k
