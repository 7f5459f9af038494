"""Run under torchrun (one process per GPU): grim.grim.impute(conf_file) with an initialised NCCL
process group must write the same six files as the unmodified reference (golden case).
Usage: torchrun --nproc-per-node N tests/multi_gpu_worker.py <case> <outdir>"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.join(HERE, "..", "py-graph-imputation_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import goldenlib  # noqa: E402


def main():
    case, outdir = sys.argv[1], sys.argv[2]
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from grim import grim
    table, conf, lines, exp = goldenlib.load_case(case)
    conf = dict(conf)
    if dist.get_rank() == 0:
        open(outdir + "/subjects.csv", "w").writelines(lines)
    dist.barrier()
    conf["imputation_in_file"] = outdir + "/subjects.csv"
    conf["imputation_out_path"] = outdir + "/out"
    names = {"umug": "imputation_out_umug_freq_filename", "umug_pops": "imputation_out_umug_pops_filename",
             "pmug": "imputation_out_hap_freq_filename", "pmug_pops": "imputation_out_hap_pops_filename",
             "miss": "imputation_out_miss_filename", "problem": "imputation_out_problem_filename"}
    for k, ck in names.items():
        conf[ck] = "x." + k
    cpath = outdir + "/conf_%d.json" % dist.get_rank()
    json.dump(conf, open(cpath, "w"))
    grim.impute(conf_file=cpath)
    ok = True
    if dist.get_rank() == 0:
        for k in names:
            same = open(outdir + "/out/x." + k).read() == exp[k]
            ok = ok and same
            print("multi-gpu", case, k, "identical" if same else "DIFFERENT", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
