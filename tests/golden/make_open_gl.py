"""Fixture for Imputation.open_gl_string / open_phases_for_em, from the UNMODIFIED
reference (build container only): tests/golden/data/open_gl_string.json."""
import contextlib
import io
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
from refrun import RefSession  # noqa: E402

DATA = os.path.join(HERE, "data")
base = json.load(open(os.path.join(DATA, "base_conf.json")))
cau = open(os.path.join(DATA, "cau_hpf.csv")).read()
cnt = open(os.path.join(DATA, "cau_pop_counts.txt")).read()
s = RefSession(base, cau, cnt)
s.run(open(os.path.join(DATA, "donor.csv")).readlines())
from grim.imputation.impute import Imputation  # noqa: E402  (the reference)
from grim.run_impute_def import run_impute  # noqa: F401,E402
conf = json.load(open(s.conf_path))
cfg = {"pops": conf["populations"], "loci_map": conf["loci_map"], "matrix_planb": conf.get("Plan_B_Matrix", [[[1, 2, 3, 4, 5]]]),
       "factor_missing_data": 0.01, "number_of_options_threshold": 100000, "max_haplotypes_number_in_phase": 100,
       "save_mode": False, "UNK_priors": "MR", "nodes_for_plan_A": [], "full_loci": "12345", "use_pops_count_file": False}
with contextlib.redirect_stdout(io.StringIO()):
    imp = Imputation(s.graph, cfg)
cases = [
    ("A*01:01+A*02:01^B*08:01+B*07:02", 100),
    ("A*01:01/A*01:02+A*02:01^B*08:01+B*07:02/B*15:01^C*07:01+C*07:02", 100),
    ("A*01:01/A*01:02+A*02:01^B*08:01+B*07:02/B*15:01^C*07:01+C*07:02", 2),
    ("A*01:01+A*01:01^B*08:01+B*08:01", 10),
    ("A*01:01/A*03:01/A*11:01+A*02:01/A*24:02^DRB1*15:01+DRB1*03:01/DRB1*04:01", 7),
    ("A*01:01", 10),
    ("", 10),
]
out = []
for gl, cutoff in cases:
    out.append({"gl": gl, "cutoff": cutoff, "phases": imp.open_gl_string(gl, cutoff)})
json.dump(out, open(os.path.join(DATA, "open_gl_string.json"), "w"), indent=0)
print("wrote", len(out), "cases")
s.close()
