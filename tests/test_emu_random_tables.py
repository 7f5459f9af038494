"""CPU check of the kernel source on RANDOM small dense tables: csrc/grimb_plan.h compiled for the
host (tests/emu/), driven by the product's tokeniser / formatter, against the oracle -- the same
generator as tests/test_gpu_random_tables.py (GPU) and tests/golden/fuzz_random_tables.py
(three-way with the unmodified reference, build container only)."""
import numpy as np
import pytest

import grim_oracle as go
import synth
from emu_backend import EmuGraph, emu_imputation
from grim.run_impute_def import load_config
from test_gpu_random_tables import cases_for, random_table

import goldenlib
import json
import os

KEYS = goldenlib.KEYS


@pytest.mark.parametrize("t", range(4))
def test_emulated_kernel_matches_oracle_on_random_tables(t):
    rng = np.random.RandomState(977 + t)
    hpf, counts, pops, tseed, _na = random_table(rng, t)
    conf = json.load(open(os.path.join(goldenlib.GOLD, "data", "base_conf.json")))
    conf.update({"populations": pops, "UNK_priors": "MR"})
    og = go.OracleGraph(hpf.splitlines(True), pops, conf["loci_map"], conf["freq_trim_threshold"],
                        counts.splitlines(True))
    eg = EmuGraph(og, conf["loci_map"])
    cbp = np.array([float(l.split(",")[2]) for l in counts.splitlines()])
    tab = synth.Table(hpf, pops[0])
    races = synth.race_fields(pops) if len(pops) > 1 else None
    for tag, lines, over in cases_for(tab, 6, tseed, races):
        c = dict(conf)
        c.update(over)
        ref = go.OracleImputation(og, go.load_config(c), cbp).impute_lines(lines)
        out = emu_imputation(eg, load_config(c), cbp).impute_lines(lines)
        for k in KEYS:
            assert "".join(out[k]) == ref[k], "table %d, %s: %s differs" % (t, tag, k)
