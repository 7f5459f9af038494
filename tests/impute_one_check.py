"""Shared by the CPU (emulation) and GPU tests: Imputation.impute_one against the oracle's impute_one
(impute.py:1940-1983 of the reference): same dict keys in the same ORDER with the same sums, same un-merged lists."""
import pytest

import grim_oracle as go
from grim.run_impute_def import load_config


def check_impute_one(conf, lines, binary=None, imp=None, og=None):
    """imp: the product's Imputation (default: over the emulated kernel source, CPU)."""
    og = og or go.graph_from_config(conf)
    cfg = load_config(conf)
    if imp is None:
        imp = _emu_imputation(_emu_graph(og, conf["loci_map"]), cfg)
    cbp = go.count_by_prob_from_file(len(conf["populations"]), conf["pops_count_file"])
    oimp = go.OracleImputation(og, go.load_config(conf), cbp)
    n_checked = 0
    for raw in lines:
        f = raw.rstrip().split(",")
        if len(f) < 2 or len(f) == 3:
            continue
        sid, gl = f[0], f[1]
        r1, r2 = (f[2], f[3]) if len(f) > 2 else (None, None)
        oimp.binary = binary if binary is not None else [1] * (len(conf["loci_map"]) - 1)
        oimp.em = False
        oimp.plan = "a"
        try:
            want = oimp.impute_one(gl, r1, r2)
        except Exception:
            with pytest.raises(Exception):
                imp.impute_one(sid, gl, binary, r1, r2, cfg["priority"], cfg["epsilon"], 1000, True, True, cfg["planb"], False)
            continue
        got = imp.impute_one(sid, gl, binary, r1, r2, cfg["priority"], cfg["epsilon"], 1000, True, True, cfg["planb"], False)
        assert got[0] == sid
        if want[0] is None:
            assert got[1] is None and got[2] is None
            continue
        wm, wh = want
        gm, gh = got[1], got[2]
        assert list(gm["Haps"].items()) == list(wm["Haps"].items()), sid      # same keys, same ORDER, same sums
        assert list(gm["Pops"].items()) == list(wm["Pops"].items()), sid
        if wh["Haps"] in ("Nan", "NaN"):
            assert gh["Haps"] == "Nan"
        else:
            assert gh["Haps"] == wh["Haps"] and gh["Probs"] == wh["Probs"] and gh["Pops"] == wh["Pops"], sid
            if wh["Haps"]:
                assert gh["MaxProb"] == wh["MaxProb"] == gm["MaxProb"]
        n_checked += 1
    return n_checked




def _emu_graph(og, loci_map):
    from emu_backend import EmuGraph
    return EmuGraph(og, loci_map)


def _emu_imputation(eg, cfg):
    from emu_backend import emu_imputation
    return emu_imputation(eg, cfg)
