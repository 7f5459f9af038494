"""Runs the UNMODIFIED reference (imported from a scratch copy of /root/reference) on a
configuration and returns its six output files.  Build-container only: the reference cannot
travel to the GPU box, which is why its outputs are committed as fixtures (make_golden.py).

The copy is needed because the reference's Cython helper must be compiled in-tree and
/root/reference is read-only.
"""
import contextlib
import io
import json
import os
import shutil
import subprocess
import sys
import tempfile

REF_SRC = "/root/reference"
REF_COPY = os.environ.get("GRIM_REF_COPY", "/tmp/grim_ref_copy")

OUT_KEYS = {
    "umug": "imputation_out_umug_freq_filename",
    "umug_pops": "imputation_out_umug_pops_filename",
    "pmug": "imputation_out_hap_freq_filename",
    "pmug_pops": "imputation_out_hap_pops_filename",
    "miss": "imputation_out_miss_filename",
    "problem": "imputation_out_problem_filename",
}


def ensure_reference():
    if not os.path.isdir(REF_SRC):
        raise RuntimeError("reference tree not present (this script only runs in the build container)")
    so = [f for f in os.listdir(os.path.join(REF_COPY, "grim", "imputation"))
          if f.startswith("cutils") and f.endswith(".so")] if os.path.isdir(REF_COPY) else []
    if not so:
        shutil.rmtree(REF_COPY, ignore_errors=True)
        shutil.copytree(REF_SRC, REF_COPY)
        subprocess.run(["chmod", "-R", "u+w", REF_COPY], check=True)
        subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=REF_COPY,
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if REF_COPY not in sys.path:
        sys.path.insert(0, REF_COPY)
    sys.argv = [sys.argv[0]]  # generate_graph() parses sys.argv


class RefSession:
    """One frequency table (hpf.csv [+ pop counts]) -> reference graph, reused across runs."""

    def __init__(self, conf, hpf_text, pop_counts_text=None):
        ensure_reference()
        from grim import grim as ref_grim  # noqa: the reference package
        self.ref_grim = ref_grim
        self.dir = tempfile.mkdtemp(prefix="grimref_")
        self.conf = dict(conf)
        self.conf["freq_file"] = os.path.join(self.dir, "hpf.csv")
        self.conf["graph_files_path"] = os.path.join(self.dir, "csv") + "/"
        self.conf["imputation_out_path"] = os.path.join(self.dir, "out")
        self.conf["imputation_in_file"] = os.path.join(self.dir, "subjects.csv")
        for k, v in OUT_KEYS.items():
            self.conf[v] = "o." + k
        with open(self.conf["freq_file"], "w") as f:
            f.write(hpf_text)
        if pop_counts_text is not None:
            self.conf["pops_count_file"] = os.path.join(self.dir, "pop_counts_file.txt")
            with open(self.conf["pops_count_file"], "w") as f:
                f.write(pop_counts_text)
        else:
            self.conf.pop("pops_count_file", None)
        self.conf_path = os.path.join(self.dir, "conf.json")
        self._write_conf()
        with contextlib.redirect_stdout(io.StringIO()):
            ref_grim.graph_freqs(conf_file=self.conf_path)
        self.graph = None

    def _write_conf(self):
        with open(self.conf_path, "w") as f:
            json.dump(self.conf, f)

    def run(self, subject_lines, hap_pop_pair=False, phase_masks=None, **overrides):
        """subject_lines: list of str (with newlines).  overrides: config keys for this run.
        hap_pop_pair: grim.impute(hap_pop_pair=True); phase_masks: {subject id: [0/1, ...]} written
        as the bin_imputation_in_file JSON."""
        saved = dict(self.conf)
        self.conf.update(overrides)
        if phase_masks is not None:
            self.conf["bin_imputation_in_file"] = os.path.join(self.dir, "phase_masks.json")
            with open(self.conf["bin_imputation_in_file"], "w") as f:
                json.dump(phase_masks, f)
        self._write_conf()
        with open(self.conf["imputation_in_file"], "w") as f:
            f.writelines(subject_lines)
        out_dir = self.conf["imputation_out_path"]
        shutil.rmtree(out_dir, ignore_errors=True)
        cwd = os.getcwd()
        os.chdir(self.dir)  # full_path() makes outputs relative to cwd
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                self.graph = self.ref_grim.impute(conf_file=self.conf_path, hap_pop_pair=hap_pop_pair, graph=self.graph)
        finally:
            os.chdir(cwd)
        res = {}
        for k in OUT_KEYS:
            p = os.path.join(self.dir, os.path.basename(out_dir.rstrip("/")), "o." + k)
            res[k] = open(p).read() if os.path.exists(p) else ""
        self.conf = saved
        return res

    def close(self):
        shutil.rmtree(self.dir, ignore_errors=True)
