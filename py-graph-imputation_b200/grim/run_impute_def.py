"""Configuration + driver of the B200 build: same JSON keys, defaults and output file layout
as grim/run_impute_def.py:41-211 of the reference; the Graph is built on the GPU from
`freq_file` (hpf.csv) instead of being loaded from the nodes/edges/top_links CSV files."""
import json
import pathlib
from pathlib import Path

from .imputation.impute import Imputation
from .imputation.networkx_graph import Graph


def full_path(output, original_path):
    # <dir of original_path>/<output>/<file name>, as the reference places its outputs
    p = Path(original_path)
    return str(p.parent / output / p.name)


def load_config(json_conf, project_dir_graph="", project_dir_in_file=""):
    """JSON -> config dict (keys and defaults of the reference, run_impute_def.py:63-129, plus the
    graph-side keys freq_file / freq_trim_threshold the table build needs)."""
    graph_files_path = json_conf.get("graph_files_path", "output/csv/")
    if graph_files_path[-1] != "/":
        graph_files_path += "/"
    output_dir = json_conf.get("imputation_out_path", "output")
    if output_dir[-1] != "/":
        output_dir += "/"
    g = json_conf.get
    config = {
        "planb": g("planb", True),
        "pops": g("populations"),
        "priority": g("priority"),
        "epsilon": g("epsilon", 1e-3),
        "number_of_results": g("number_of_results", 1000),
        "number_of_pop_results": g("number_of_pop_results", 100),
        "output_MUUG": g("output_MUUG", True),
        "output_haplotypes": g("output_haplotypes", False),
        "node_file": project_dir_graph + graph_files_path + g("node_csv_file", "nodes.csv"),
        "top_links_file": project_dir_graph + graph_files_path + g("top_links_csv_file", "top_links.csv"),
        "edges_file": project_dir_graph + graph_files_path + g("edges_csv_file", "edges.csv"),
        "imputation_input_file": project_dir_in_file + g("imputation_in_file", ""),
        "imputation_out_umug_freq_file": full_path(output_dir, g("imputation_out_umug_freq_filename", "out.umug")),
        "imputation_out_umug_pops_file": full_path(output_dir, g("imputation_out_umug_pops_filename", "out.umug.pops")),
        "imputation_out_hap_freq_file": full_path(output_dir, g("imputation_out_hap_freq_filename", "out.pmug")),
        "imputation_out_hap_pops_file": full_path(output_dir, g("imputation_out_hap_pops_filename", "out.pmug.pops")),
        "imputation_out_miss_file": full_path(output_dir, g("imputation_out_miss_filename", "out.miss")),
        "imputation_out_problem_file": full_path(output_dir, g("imputation_out_problem_filename", "out.problem")),
        "factor_missing_data": g("factor_missing_data", 0.01),
        "loci_map": g("loci_map", {"A": 1, "B": 3, "C": 2, "DQB1": 4, "DRB1": 5}),
        "matrix_planb": g(
            "Plan_B_Matrix",
            [
                [[1, 2, 3, 4, 5]],
                [[1, 2, 3], [4, 5]],
                [[1], [2, 3], [4, 5]],
                [[1, 2, 3], [4], [5]],
                [[1], [2, 3], [4], [5]],
                [[1], [2], [3], [4], [5]],
            ],
        ),
        "pops_count_file": project_dir_graph + g("pops_count_file", ""),
        "use_pops_count_file": g("pops_count_file", False),
        "number_of_options_threshold": g("number_of_options_threshold", 100000),
        "max_haplotypes_number_in_phase": g("max_haplotypes_number_in_phase", 100),
        "bin_imputation_input_file": project_dir_in_file + g("bin_imputation_in_file", "None"),
        "nodes_for_plan_A": g("Plan_A_Matrix", []),
        "save_mode": g("save_space_mode", False),
        "UNK_priors": g("UNK_priors", "MR"),
        # graph-side keys (generate_neo4j_multi_hpf.py:243-256 of the reference)
        "freq_file": project_dir_graph + g("freq_file", "output/hpf.csv"),
        "freq_trim_threshold": g("freq_trim_threshold", 1e-5),
        "imputation_out_path": output_dir,
    }
    all_loci = {str(v) for v in config["loci_map"].values()}
    config["full_loci"] = "".join(sorted(all_loci))
    if config["nodes_for_plan_A"]:
        config["plan_a_masks"], config["store_label_masks"] = plan_a_label_masks(
            config["nodes_for_plan_A"], len(config["loci_map"]))
    return config


def plan_a_label_masks(matrix, n_loci):
    """"Plan_A_Matrix" -> (locus bit masks of the matrix rows, masks of the labels the store builds).

    The reference generates graph nodes for the matrix labels only (generate_neo4j_multi_hpf.py:101-192), loads
    them as the vertex set of Plan A (networkx_graph.py:32-66) and imputes only subjects whose typed-locus
    pattern is a matrix row (impute.py:1592-1596).  What is accepted here is what has a defined behaviour there:
    rows strictly ascending (input_type yields ascending lists; a permuted row would also permute the alleles
    inside node names), no repeated row, the full label FIRST (full haplotypes always take ids 0..N-1 while
    nodes.csv lists labels in matrix order: with the full label elsewhere a vertex's list position no longer
    equals its id and every adjacency is read from another node), at least one marginal label (the reference's
    np.vstack of an empty edge list raises otherwise).  The store also builds the single-locus labels (the
    reference always generates them, generate_neo4j_multi_hpf.py:166-170; the allele-existence checks of the
    reduce steps read them)."""
    full = list(range(1, n_loci + 1))
    masks = []
    for row in matrix:
        row = [int(x) for x in row]
        if not row or row != sorted(set(row)) or row[0] < 1 or row[-1] > n_loci:
            raise NotImplementedError("Plan_A_Matrix rows must be strictly ascending lists of loci_map indices")
        m = 0
        for i in row:
            m |= 1 << (i - 1)
        if m in masks:
            raise NotImplementedError("Plan_A_Matrix lists a label twice")
        masks.append(m)
    if [int(x) for x in matrix[0]] != full:
        raise NotImplementedError(
            "Plan_A_Matrix must list the full label first: otherwise the reference reads every adjacency from "
            "another node (vertex list positions vs. node ids)")
    if len(masks) < 2:
        raise ValueError("need at least one array to concatenate")   # what the reference's graph load raises
    store = list(masks) + [1 << l for l in range(n_loci) if (1 << l) not in masks]
    return masks, store


def run_impute(conf_file="../conf/minimal-configuration.json", project_dir_graph="", project_dir_in_file="",
               hap_pop_pair=False, graph=None, device=0):
    with open(conf_file) as f:
        json_conf = json.load(f)
    config = load_config(json_conf, project_dir_graph, project_dir_in_file)
    print("*" * 100)
    print("Performing imputation based on:")
    for label, key in (("Population", "pops"), ("Priority", "priority"), ("UNK priority", "UNK_priors"),
                       ("Epsilon", "epsilon"), ("Plan B", "planb"), ("Number of Results", "number_of_results"),
                       ("Number of Population Results", "number_of_pop_results"), ("Freq File", "freq_file"),
                       ("Input File", "imputation_input_file"), ("Output UMUG Format", "output_MUUG"),
                       ("Output Haplotype Format", "output_haplotypes"), ("Factor Missing Data", "factor_missing_data"),
                       ("Loci Map", "loci_map"), ("Plan B Matrix", "matrix_planb"),
                       ("Pops Count File", "pops_count_file"),
                       ("Number of Options Threshold", "number_of_options_threshold"),
                       ("Max Number of haplotypes in phase", "max_haplotypes_number_in_phase"),
                       ("Save space mode", "save_mode")):
        print("\t{}: {}".format(label, config[key]))
    print("*" * 100)
    dist = _process_group()
    if dist is not None and dist.get_world_size() > 1:
        return _run_impute_sharded(dist, config, hap_pop_pair, graph)
    if graph is None:
        graph = Graph(config, device=device)
        graph.build_graph(config["node_file"], config["top_links_file"], config["edges_file"])
    imputation = Imputation(graph, config)
    pathlib.Path(config["imputation_out_path"]).mkdir(parents=False, exist_ok=True)
    imputation.impute_file(config, em_mr=hap_pop_pair)
    return graph


def _process_group():
    """torch.distributed, if the caller started one process per GPU (torchrun); else None."""
    import sys
    if "torch" not in sys.modules:
        return None
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _run_impute_sharded(dist, config, hap_pop_pair, graph):
    """One process per GPU: tables built on rank 0 and replicated by one NCCL broadcast, input
    lines sharded by contiguous ranges, rank 0 writes the six files in input order."""
    import os
    from .imputation import multi_gpu
    rank, world = dist.get_rank(), dist.get_world_size()
    device = int(os.environ.get("LOCAL_RANK", rank))
    multi_gpu.bind_to_device_numa_node(device)
    if graph is None:
        if rank == 0:
            graph = Graph(config, device=device)
            graph.build_graph(config["node_file"], config["top_links_file"], config["edges_file"])
        graph = multi_gpu.broadcast_graph(graph, config, device, src=0)
    # the ranks share the host's cores: split them for the C++ tokeniser / formatter threads
    os.environ.setdefault("GRIMB_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // world)))
    imputation = Imputation(graph, config)
    targets = {"miss": "imputation_out_miss_file", "problem": "imputation_out_problem_file"}
    if config["output_MUUG"]:
        targets.update(umug="imputation_out_umug_freq_file", umug_pops="imputation_out_umug_pops_file")
    if config["output_haplotypes"]:
        targets.update(pmug="imputation_out_hap_freq_file", pmug_pops="imputation_out_hap_pops_file")
    in_path = config["imputation_input_file"]
    size = os.path.getsize(in_path)
    lib = graph.lib
    import ctypes as C
    import threading
    from .imputation import _lib
    # outputs of an earlier run: moved aside now and deleted in the background while the subjects are imputed
    # (truncating gigabytes in place frees their page-cache pages synchronously, on the critical path)
    cleaner = None
    if rank == 0:
        old = []
        for ck in targets.values():
            try:
                if os.path.isfile(config[ck]) and os.path.getsize(config[ck]) > (1 << 20):
                    aside = "%s.old.%d" % (config[ck], os.getpid())
                    os.rename(config[ck], aside)
                    old.append(aside)
            except OSError:
                pass
        if old:
            cleaner = threading.Thread(target=lambda: [os.unlink(f) for f in old])
            cleaner.start()
    if not (hap_pop_pair or imputation.phase_masks is not None) and os.environ.get("GRIMB_SHARD_BOARD", "1") != "0":
        # default mode: the ranks take the input's chunks round-robin and stream their rows straight into the six
        # final files (grimb_impute_file_sharded: line counts and piece sizes meet on a small memory-mapped board)
        chunk = int(os.environ.get("GRIMB_FILE_CHUNK", "0")) or (16 << 20)
        board = os.path.join(config["imputation_out_path"], ".grimb_board.%s" % os.environ.get("MASTER_PORT", "0"))
        n_tiers = 1
        while True:
            if rank == 0:
                pathlib.Path(config["imputation_out_path"]).mkdir(parents=False, exist_ok=True)
                for ck in targets.values():
                    open(config[ck], "wb").close()
                with open(board, "wb") as f:
                    f.truncate(int(lib.grimb_file_board_bytes(size, chunk)))
            dist.barrier()
            err = None
            try:
                imputation.impute_file_sharded(in_path, {k: config[ck] for k, ck in targets.items()}, chunk, rank, world,
                                               board, n_tiers)
            except (RuntimeError, NotImplementedError) as e:      # every rank must reach the exchange below
                err = (type(e).__name__, str(e))
            errs = [None] * world
            dist.all_gather_object(errs, err)
            if rank == 0:
                os.unlink(board)
            if all(e is None for e in errs):
                break
            # a subject overflowed the last workspace tier on some rank: all ranks start over with one more
            if any(e is not None and "(-3)" in e[1] for e in errs) and n_tiers < len(imputation.workspaces):
                n_tiers += 1
                continue
            # the rank that hit the problem names it; the others only report that a rank failed
            found = [e for e in errs if e is not None]
            first = next((e for e in found if "another rank failed" not in e[1]), found[0])
            if cleaner is not None:
                cleaner.join()
            raise (NotImplementedError if first[0] == "NotImplementedError" else RuntimeError)(first[1])
        if cleaner is not None:
            cleaner.join()
        return graph
    # every rank takes the lines that START in its byte range of the input (memory-mapped by the library: no
    # rank reads the whole file); the ranks count their own lines and exchange the counts, because .miss /
    # .problem rows carry global line indices
    lo_b, hi_b = size * rank // world, size * (rank + 1) // world
    n_lines, lo_adj, hi_adj = C.c_int64(), C.c_int64(), C.c_int64()
    _lib.check(lib.grimb_file_count_lines(in_path.encode("utf8"), lo_b, hi_b if rank + 1 < world else -1, 0,
                                          C.byref(n_lines), C.byref(lo_adj), C.byref(hi_adj)), "grimb_file_count_lines", lib)
    counts = [None] * world
    dist.all_gather_object(counts, int(n_lines.value))
    first = sum(counts[:rank])
    if hap_pop_pair or imputation.phase_masks is not None:
        # EM-facing modes go through the numpy host front end (see Imputation.impute_file)
        with open(in_path, "rb") as f:
            f.seek(lo_adj.value)
            data = f.read(hi_adj.value - lo_adj.value)
        rows = imputation.impute_lines(data.decode("utf8").splitlines(True), first_index=first, em_mr=hap_pop_pair)
        mine = {k: "".join(v).encode("utf8") for k, v in rows.items()}
        ptr = {k: (mine[k], len(mine[k])) for k in targets}
    else:
        out, _st = imputation.impute_file_native(in_path, None, lo_b, hi_b if rank + 1 < world else -1, first)
        ptr = {k: (out.data[i], int(out.size[i])) for i, k in enumerate(_lib.OUT_KEYS) if k in targets}
    # every rank writes its rows straight into the final files at its offset (rank order == input order):
    # only the sizes are exchanged, not the gigabytes of text
    sizes = [None] * world
    dist.all_gather_object(sizes, {k: ptr[k][1] for k in targets})
    if rank == 0:
        pathlib.Path(config["imputation_out_path"]).mkdir(parents=False, exist_ok=True)
        for k, ck in targets.items():
            with open(config[ck], "wb") as f:
                f.truncate(sum(sz[k] for sz in sizes))
    dist.barrier()
    errs = []

    def put(k, ck):
        data, n = ptr[k]
        if n:
            rc = lib.grimb_file_write_at(config[ck].encode("utf8"), sum(sz[k] for sz in sizes[:rank]), data, n)
            if rc != 0:
                errs.append((k, rc))

    ths = [threading.Thread(target=put, args=kv) for kv in targets.items()]   # ctypes releases the GIL in the call
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    if cleaner is not None:
        cleaner.join()
    if errs:
        raise RuntimeError("writing the output files failed: %s" % errs)
    dist.barrier()
    return graph
