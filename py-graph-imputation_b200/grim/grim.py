"""Public API of the B200 build; signatures as in grim/grim.py:40-87 of the reference."""
import json
import os

from .imputation.impute import Imputation
from .imputation.networkx_graph import Graph
from .run_impute_def import load_config, run_impute

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
_DEFAULT_CONF = os.path.join(_PKG_ROOT, "conf", "minimal-configuration.json")


def graph_freqs(conf_file="", for_em=False, em_pop=None, device=0):
    """Builds the frequency store for `conf_file` on the GPU and returns it (the reference writes
    nodes/edges/top_links CSV files here; this build keeps the tables in HBM and `impute(...,
    graph=g)` reuses them)."""
    project = ""
    if conf_file == "":
        conf_file = _DEFAULT_CONF
        project = _PKG_ROOT + "/"
    with open(conf_file) as f:
        config = load_config(json.load(f), project, project)
    # EM variants (generate_neo4j_multi_hpf.py:249-261 of the reference): em_pop replaces the
    # population list; for_em ignores the population counts file when trimming
    if em_pop:
        config["pops"] = list(em_pop)
    if for_em:
        config["use_pops_count_file"] = False
    return graph_instance(config, device=device)


def impute(conf_file="", hap_pop_pair=False, graph=None, device=0):
    project_dir_in_file, project_dir_graph = "", ""
    if conf_file == "":
        conf_file = _DEFAULT_CONF
        project_dir_graph = _PKG_ROOT + "/"
        project_dir_in_file = _PKG_ROOT + "/"
    return run_impute(conf_file, project_dir_graph, project_dir_in_file, hap_pop_pair, graph, device=device)


def impute_instance(config, graph, count_by_prob=None):
    return Imputation(graph, config, count_by_prob)


def graph_instance(config, device=0):
    graph = Graph(config, device=device)
    graph.build_graph(config.get("node_file"), config.get("top_links_file"), config.get("edges_file"))
    return graph
