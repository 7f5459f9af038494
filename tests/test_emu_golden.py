"""CPU check of the kernel source: the per-subject algorithm (csrc/grimb_plan.h), compiled for
the host in single-thread emulation, driven by the product's tokeniser/formatter, must
reproduce every golden file written by the unmodified reference.  The GPU parity tests
(test_gpu_parity.py) run the same cases through libgrimb200.so on the B200."""
import numpy as np
import pytest

import goldenlib
import grim_oracle as go
from emu_backend import EmuGraph, emu_imputation
from grim.run_impute_def import load_config

_cache = {}


def _setup(table, conf):
    if table not in _cache:
        og = go.graph_from_config(conf)
        _cache[table] = (og, EmuGraph(og, conf["loci_map"]))
    return _cache[table]


def run_case(name):
    table, conf, lines, exp = goldenlib.load_case(name)
    og, eg = _setup(table, conf)
    cfg = load_config(conf)
    imp = emu_imputation(eg, cfg)
    files = imp.impute_lines(lines, em_mr=conf["_hap_pop_pair"])
    return {k: "".join(v) for k, v in files.items()}, exp


@pytest.mark.parametrize("name", goldenlib.case_names())
def test_emulated_kernel_matches_reference_files(name):
    out, exp = run_case(name)
    for k in goldenlib.KEYS:
        assert out[k] == exp[k], "%s: %s differs" % (name, k)


@pytest.mark.parametrize("name", ["g1_readme_donor", "g2_edges", "g4_amb6", "g4_amb12_over_threshold", "g4_low_threshold",
                                  "g4_unknown_heavy", "g5_messy_cau", "g3_pop3_messy", "g6_nine_loci"])
def test_cooperative_slot_pass_gives_the_same_files(name, monkeypatch):
    """The cooperative CTA group of the heaviest subjects (k_impute mode 1: every (phase, side) slot opened,
    probed and reduced to its top-K list by a group of its own; mode 0 then starts from the lists) -- here with
    EVERY subject sent through it, on the emulation build: same files as the reference."""
    monkeypatch.setenv("GRIMB_EMU_GROUP", "1")
    out, exp = run_case(name)
    for k in goldenlib.KEYS:
        assert out[k] == exp[k], "%s: %s differs" % (name, k)
